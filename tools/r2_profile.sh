#!/bin/bash
# final round-2 profile set: ncu launch list of the default command (render kernels; scene setup launches ~1500 short
# kernels of its own first) + ONE full capture of a steady-state bounce iteration
mkdir -p gpurun_out
timeout -s KILL 500 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:"traceClosestKernelILi2ELb0|traceShadowKernelILb0|shadeKernel|accumulateKernel" -s 28 -c 4 -f -o gpurun_out/r02_iteration python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-adaptive-line > gpurun_out/ncu_full.log 2>&1; echo "full capture rc=$?"
timeout -s KILL 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"trace|shade|accumulate|generate|resolve|adaptive" -c 600 --csv --log-file gpurun_out/r02_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-adaptive-line > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
