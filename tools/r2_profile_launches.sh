#!/bin/bash
# ncu launch list of the default command, render kernels only (scene setup launches ~1500 short kernels of its own first)
mkdir -p gpurun_out
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"trace|shade|accumulate|generate|resolve|adaptive" -c 900 --csv --log-file gpurun_out/r02_launches_default.csv python bench.py --no-cpu-baseline --no-adaptive-line > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"Kernel" -c 1300 --csv --log-file gpurun_out/r02_launches_setup.csv python -c "
import sys, os
sys.path.insert(0, os.getcwd())
from cpupathtrace_b200 import pth, scenes
b = pth.load_b200()
v, n = scenes.standin_triangles(1000, 500, scenes.DEMO_DRAGON_TRANSFORM)
sc = scenes.cornell_demo(('triangles', v, n)).build(b)
" > gpurun_out/ncu_setup.log 2>&1; echo "setup list rc=$?"
