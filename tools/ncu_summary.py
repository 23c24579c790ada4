#!/usr/bin/env python
"""Summarises ncu output brought back in gpurun_out/ into small tracked text files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches_<tag>.md
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/r01_ncu_<tag>.md
  python tools/ncu_summary.py traffic gpurun_out/prof.ncu-rep profiles/traffic.json "<how the capture was made>"
      DRAM bytes per launch of the captured kernels (what bench.py quotes as roofline.traffic): written from the SAME
      report as the full summary, so the two cannot drift apart
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "launch__grid_size",
    "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
]


def launches(src, dst):
    rows = list(csv.reader(open(src, errors="replace")))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        c = agg.setdefault(name, [0, 0.0])
        c[0] += 1
        c[1] += v
    total = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src}): gpu__time_duration.sum, --clock-control none (cold-cache, serialised: compare shares)\n\n")
        f.write("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k[:80]}` | {n} | {t / 1e3:.3f} | {t / total:.1%} | {t / n:.1f} |\n")
        f.write(f"\ntotal {total / 1e3:.3f} ms over {sum(v[0] for v in agg.values())} launches\n")


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src}); one column per captured launch\n\n")
        names = [r[idx["Kernel Name"]].split("(")[0].replace("void ", "") for r in rows[2:]]
        f.write("| metric | unit | " + " | ".join(f"`{n[:28]}`" for n in names) + " |\n")
        f.write("|---|---|" + "---:|" * len(names) + "\n")
        for m in METRICS:
            if m not in idx:
                continue
            f.write(f"| {m} | {units[idx[m]]} | " + " | ".join(r[idx[m]] for r in rows[2:]) + " |\n")


def traffic(src, dst, how):
    import json

    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    time_scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}

    def val(row, metric, table):
        return float(row[idx[metric]].replace(",", "")) * table[units[idx[metric]]]

    launches_out = []
    for r in rows[2:]:
        launches_out.append({
            "kernel": r[idx["Kernel Name"]].split("(")[0].replace("void ", ""),
            "dram_read_bytes": val(r, "dram__bytes_read.sum", scale),
            "dram_write_bytes": val(r, "dram__bytes_write.sum", scale),
            "duration_ms": val(r, "gpu__time_duration.sum", time_scale),
            "issue_active_pct": float(r[idx["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
            "threads_per_instruction": float(r[idx["smsp__thread_inst_executed_per_inst_executed.ratio"]]),
            "dram_throughput_pct": float(r[idx["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]),
        })
    top = next((k for k in launches_out if "traceClosestKernel<2" in k["kernel"]), launches_out[0])
    out = {
        "source": f"{src} -> this file and the matching profiles/*.md summary come from the same report; {how}",
        "kernel": top["kernel"],
        "dram_bytes_per_launch": top["dram_read_bytes"] + top["dram_write_bytes"],
        "launches": launches_out,
    }
    json.dump(out, open(dst, "w"), indent=1)


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
