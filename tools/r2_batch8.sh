#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
for q in sweep host; do
  PTB_QUERY_TREE=$q timeout -s KILL 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-adaptive-line > gpurun_out/bench_c2_$q.json 2> gpurun_out/bench_c2_$q.err || tail -5 gpurun_out/bench_c2_$q.err
done
python - <<'PY'
import json
for q in ("sweep", "host"):
    d = json.load(open(f"gpurun_out/bench_c2_{q}.json")); r = d["roofline"]
    print(q, round(d["value"],1), "Msamples/s e2e", round(d["e2e"]["value"],1), "closest", round(r["closest_inner_per_ray"],2), round(r["closest_leaf_per_ray"],2), "shadow", round(r["shadow_inner_per_ray"],2), round(r["shadow_leaf_per_ray"],2), "breakdown", {k: (round(v,1) if isinstance(v,float) else v) for k,v in d["breakdown"].items() if k!="source"}, "scene", d["scene"])
PY
