#!/usr/bin/env python
"""Attributes executed warp instructions / stall samples of one kernel in an ncu report to CUDA source lines.
usage: sass_by_line.py <nvdisasm -g -c output> <ncu --page source --csv output> <mangled-name-substring> [top] [demangled-name-substring]
The report may hold several kernels: the CSV is a sequence of sections, each introduced by a "Kernel Name" row; the
optional fifth argument selects the sections of one kernel (e.g. 'traceClosestKernel<(int)2, (bool)0>')."""
import collections
import csv
import re
import sys

dis, src_csv, name = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = open(dis).read().splitlines()
start = next(i for i, l in enumerate(lines) if ".text." in l and name in l and ".section" in l)
off2line, cur, n = {}, None, 0
for l in lines[start + 1:]:
    if ".section" in l and n > 50:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/", l)
    if m:
        off2line[int(m.group(1), 16)] = cur
        n += 1
rows = list(csv.reader(open(src_csv)))
if len(sys.argv) > 5:
    wanted, keep, selected = sys.argv[5], False, []
    for r in rows:
        if r and r[0] == "Kernel Name":
            keep = len(r) > 1 and wanted in r[1]
            continue
        if keep:
            selected.append(r)
    rows = selected
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
ai, ii, si, ti = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


seen, uniq = set(), []
for r in rows[hi + 1:]:
    if len(r) <= ii or r[ai] in seen or r[ai] == "Address":
        continue
    seen.add(r[ai])
    uniq.append(r)
addrs = [int(r[ai], 16) if r[ai].startswith("0x") else int(r[ai]) for r in uniq]
base = min(addrs)
inst, samp, thr = collections.Counter(), collections.Counter(), collections.Counter()
for r, a in zip(uniq, addrs):
    ln = off2line.get(a - base)
    inst[ln] += num(r[ii])
    samp[ln] += num(r[si])
    thr[ln] += num(r[ti])
tot, tots = sum(inst.values()), sum(samp.values())
print(f"total warp instructions {tot:.0f}, avg active threads {sum(thr.values()) / tot:.1f}")
src = {}
for k in inst:
    if k and k[0] not in src:
        try:
            src[k[0]] = open("/root/repo/cpupathtrace_b200/csrc/" + k[0]).read().splitlines()
        except OSError:
            src[k[0]] = []
for k, v in inst.most_common(top):
    text = src[k[0]][k[1] - 1].strip()[:80] if k and src.get(k[0]) and k[1] - 1 < len(src[k[0]]) else ""
    print(f"{v / tot:6.2%} inst {samp[k] / tots:6.2%} samp thr {thr[k] / max(v, 1):5.1f}  {str(k):28s} {text}")
