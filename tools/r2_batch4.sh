#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -v "^$" | tail -12
for rays in incoherent coherent shadow; do
  timeout -s KILL 600 python bench.py --workload soup --tris 4 --rays $rays --steps 5 --warmup 3 --cpu-seconds 5 > gpurun_out/soup_4_$rays.json 2> gpurun_out/soup.err || tail -5 gpurun_out/soup.err
done
PTB_SORT_RAYS=0 timeout -s KILL 600 python bench.py --workload soup --tris 4 --rays incoherent --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/soup_4_incoherent_nosort.json 2> gpurun_out/soup.err || tail -5 gpurun_out/soup.err
timeout -s KILL 900 python bench.py --workload soup --tris 16 --rays incoherent --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/soup_16_incoherent.json 2> gpurun_out/soup.err || tail -5 gpurun_out/soup.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/soup_*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f"{f:45s} {d['value']:8.1f} Mrays/s e2e {d['e2e']['value']:7.1f} frac {r['frac']:.3f} ({r['inner_fetches_per_ray']:.1f}+{r['leaf_fetches_per_ray']:.1f} fetches, {r['bytes_per_ray']:.0f} B/ray) trace {r['ms_per_launch']:.1f} ms sort {r['sort_ms_per_launch']:.1f} ms identical {d['identical_to_reference_walk']} cpu {(d.get('cpu_baseline') or {}).get('value')}")
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout -s KILL 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err || tail -5 gpurun_out/bench_c2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_c2.json")); r = d["roofline"]
print("C2:", round(d["value"],1), "Msamples/s e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],1), "frac", round(r["frac"],3), "breakdown", {k: (round(v,1) if isinstance(v,float) else v) for k,v in d["breakdown"].items() if k!="source"}, "cpu", d["cpu_baseline"]["value"] if d["cpu_baseline"] else None, "setup", d["scene"]["scene_ctor_s"])
PY
