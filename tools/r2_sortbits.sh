#!/bin/bash
mkdir -p gpurun_out
for b in 1 2 3 4; do
  for cfg in "16 incoherent" "4 incoherent" "4 shadow"; do
    set -- $cfg
    PTB_SORT_DIR_BITS=$b timeout -s KILL 900 python bench.py --workload soup --tris $1 --rays $2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/soup_v.json 2> gpurun_out/soup.err || tail -3 gpurun_out/soup.err
    python -c "
import json; s=json.load(open('gpurun_out/soup_v.json')); print('dir bits $b soup $1 $2:', round(s['value'],1), 'Mrays/s frac', round(s['roofline']['frac'],3), 'trace ms', round(s['roofline']['ms_per_launch'],1), 'identical', s['identical_to_reference_walk'])"
  done
done
