#!/bin/bash
# final validation of the round: GPU tests, smoke, default bench, reference arm, C3 soup lines
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -1
timeout -s KILL 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err || tail -5 gpurun_out/bench_c2.err
timeout -s KILL 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
rm -f gpurun_out/r02_soup_bench.jsonl
for cfg in "4 incoherent" "4 coherent" "4 shadow" "16 incoherent"; do
  set -- $cfg
  timeout -s KILL 900 python bench.py --workload soup --tris $1 --rays $2 --steps 5 --warmup 3 --cpu-seconds 5 >> gpurun_out/r02_soup_bench.jsonl 2> gpurun_out/soup.err || tail -5 gpurun_out/soup.err
done
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_c2.json")); r = d["roofline"]
print("C2:", round(d["value"],1), d["ms_steps"], "e2e", round(d["e2e"]["value"],1), d["e2e"]["ms_steps"], "frac", round(r["frac"],3), {k: (round(v,1) if isinstance(v,float) else v) for k,v in d["breakdown"].items() if k!="source"})
print("scene:", {k: (round(v,3) if isinstance(v,float) else v) for k, v in d["scene"].items() if k != "note"})
print("adaptive:", d["adaptive"]["frame_ms"], d["adaptive"]["samples_traced"])
ref = json.load(open("gpurun_out/bench_reference.json")); print("reference arm:", ref["value"], "-> e2e ratio", round(d["e2e"]["value"] / ref["value"], 1))
for line in open("gpurun_out/r02_soup_bench.jsonl"):
    s = json.loads(line); q = s["roofline"]
    print("soup", s["config"]["tris"] >> 20, s["config"]["rays"], round(s["value"],1), "Mrays/s e2e", round(s["e2e"]["value"],1), "frac", round(q["frac"],3), "identical", s["identical_to_reference_walk"], "cpu", (s.get("cpu_baseline") or {}).get("value"), "build_s", round(s["scene"]["build_s"],3))
PY
