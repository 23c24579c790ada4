#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout -s KILL 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err || tail -5 gpurun_out/bench_c2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_c2.json")); r = d["roofline"]
print("C2:", round(d["value"],1), "Msamples/s e2e", round(d["e2e"]["value"],1), d["e2e"].get("ms_steps"), "ms/step", round(d["ms_per_step"],1), "frac", round(r["frac"],3), "breakdown", {k: (round(v,1) if isinstance(v,float) else v) for k,v in d["breakdown"].items() if k!="source"})
print("scene:", d["scene"]); print("adaptive:", d["adaptive"]); print("cpu:", d["cpu_baseline"])
PY
timeout -s KILL 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 600 gpurun_out/bench_reference.json
