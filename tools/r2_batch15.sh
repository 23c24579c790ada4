#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout -s KILL 600 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err || tail -5 gpurun_out/bench_c2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_c2.json")); r = d["roofline"]
print("C2:", round(d["value"],1), "Msamples/s", d["ms_steps"], "e2e", round(d["e2e"]["value"],1), d["e2e"].get("ms_steps"), "frac", round(r["frac"],3), "breakdown", {k: (round(v,1) if isinstance(v,float) else v) for k,v in d["breakdown"].items() if k!="source"})
print("scene:", {k: v for k, v in d["scene"].items() if k != "note"}); print("adaptive:", d["adaptive"]); print("cpu:", d["cpu_baseline"]["value"])
PY
