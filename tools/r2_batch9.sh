#!/bin/bash
mkdir -p gpurun_out
for q in sweep host; do
  PTB_QUERY_TREE=$q timeout -s KILL 900 python bench.py --workload soup --tris 16 --rays incoherent --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/soup_16_incoherent_$q.json 2> gpurun_out/soup.err || tail -5 gpurun_out/soup.err
  PTB_QUERY_TREE=$q timeout -s KILL 900 python bench.py --workload soup --tris 4 --rays shadow --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/soup_4_shadow_$q.json 2> gpurun_out/soup.err || tail -5 gpurun_out/soup.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/soup_*_sweep.json") + glob.glob("gpurun_out/soup_*_host.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f"{f:45s} {d['value']:8.1f} Mrays/s e2e {d['e2e']['value']:7.1f} frac {r['frac']:.3f} ({r['inner_fetches_per_ray']:.1f}+{r['leaf_fetches_per_ray']:.1f} fetches, {r['bytes_per_ray']:.0f} B/ray) trace {r['ms_per_launch']:.1f} ms identical {d['identical_to_reference_walk']} scene {d['scene']}")
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout -s KILL 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-adaptive-line > gpurun_out/bench_c2_quick.json 2> gpurun_out/bench_c2_quick.err || tail -5 gpurun_out/bench_c2_quick.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c2_quick.json')); print(round(d['value'],1), d['scene'])"
