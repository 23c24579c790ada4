#!/bin/bash
mkdir -p gpurun_out
cp cpupathtrace_b200/lib/libptb.so /tmp/libptb_base.so
for rep in 1 2; do
  for v in base evict1 evict2; do
    if [ $v == base ]; then cp /tmp/libptb_base.so cpupathtrace_b200/lib/libptb.so; else cp variants/libptb_$v.so cpupathtrace_b200/lib/libptb.so; fi
    bash tools/r2_sweep_env.sh PTB_VARIANT $v
  done
done
cp /tmp/libptb_base.so cpupathtrace_b200/lib/libptb.so
