#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
rm -f gpurun_out/r02_soup_bench.jsonl
for cfg in "4 incoherent" "4 coherent" "4 shadow" "16 incoherent"; do
  set -- $cfg
  timeout -s KILL 900 python bench.py --workload soup --tris $1 --rays $2 --steps 5 --warmup 3 --cpu-seconds 5 >> gpurun_out/r02_soup_bench.jsonl 2> gpurun_out/soup.err || tail -5 gpurun_out/soup.err
done
python - <<'PY'
import json
for line in open("gpurun_out/r02_soup_bench.jsonl"):
    s = json.loads(line); q = s["roofline"]
    print("soup", s["config"]["tris"] >> 20, s["config"]["rays"], round(s["value"],1), "Mrays/s e2e", round(s["e2e"]["value"],1), "frac", round(q["frac"],3), "identical", s["identical_to_reference_walk"], "cpu", (s.get("cpu_baseline") or {}).get("value"), "build_s", round(s["scene"]["build_s"],3))
PY
