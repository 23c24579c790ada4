#!/bin/bash
# ncu --set full of one steady-state shade + accumulate launch, exact build vs production-math build
mkdir -p gpurun_out
for pm in 0 1; do
  PTB_PRODUCTION_MATH=$pm timeout -s KILL 400 ncu --set full --clock-control none --import-source on --kernel-name-base mangled \
    -k regex:"shadeKernel|accumulateKernel" -s 10 -c 2 -f -o gpurun_out/r2_shade_pm$pm \
    python bench.py --spp 64 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_shade_pm$pm.log 2>&1; echo "pm=$pm rc=$?"
done
ls -la gpurun_out/*.ncu-rep
