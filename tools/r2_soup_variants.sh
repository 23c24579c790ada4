#!/bin/bash
mkdir -p gpurun_out
cp cpupathtrace_b200/lib/libptb.so /tmp/libptb_base.so
for v in "$@"; do
  if [ $v == base ]; then cp /tmp/libptb_base.so cpupathtrace_b200/lib/libptb.so; else cp variants/libptb_$v.so cpupathtrace_b200/lib/libptb.so; fi
  for cfg in "16 incoherent" "4 incoherent" "4 shadow"; do
    set -- $cfg
    timeout -s KILL 900 python bench.py --workload soup --tris $1 --rays $2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/soup_v.json 2> gpurun_out/soup.err || tail -3 gpurun_out/soup.err
    python -c "
import json; s=json.load(open('gpurun_out/soup_v.json')); print('$v soup $1 $2:', round(s['value'],1), 'Mrays/s frac', round(s['roofline']['frac'],3), 'trace ms', round(s['roofline']['ms_per_launch'],1))"
  done
done
cp /tmp/libptb_base.so cpupathtrace_b200/lib/libptb.so
