#!/bin/bash
PTB_LOG_BATCHES=1 timeout -s KILL 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-adaptive-line > gpurun_out/var.json 2> gpurun_out/var.err
python - <<'PY'
import json, re
d = json.load(open("gpurun_out/var.json"))
print("steps", d["ms_steps"])
frames, cur = [], []
for line in open("gpurun_out/var.err"):
    m = re.search(r"batch of (\d+) iterations done ([0-9.]+) ms after the call started \((\d+) paths", line)
    if not m:
        continue
    t = float(m.group(2))
    if cur and t < cur[-1][0]:
        frames.append(cur); cur = []
    cur.append((t, int(m.group(3))))
if cur:
    frames.append(cur)
for f in frames:
    if len(f) >= 3 and f[0][1] > 100_000_000:
        ts = [round(t) for t, _ in f]
        print("frame batches end at", ts, "deltas", [ts[0]] + [ts[i] - ts[i - 1] for i in range(1, len(ts))])
PY
