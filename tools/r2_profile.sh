#!/bin/bash
# final round-2 profile set: launch list of the default command + ONE full capture of a steady-state bounce iteration
mkdir -p gpurun_out
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_default.csv python bench.py --no-cpu-baseline --no-adaptive-line > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:"traceClosestKernelILi2ELb0|traceShadowKernelILb0|shadeKernel|accumulateKernel" -s 28 -c 4 -f -o gpurun_out/r02_iteration python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-adaptive-line > gpurun_out/ncu_full.log 2>&1; echo "full capture rc=$?"
ls -la gpurun_out/*.ncu-rep
