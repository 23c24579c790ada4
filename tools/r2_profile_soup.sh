#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:"intersectKernelILi2ELb0" -s 3 -c 1 -f -o gpurun_out/r02_soup16 python bench.py --workload soup --tris 16 --rays incoherent --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_soup.log 2>&1; echo "soup capture rc=$?"
ls -la gpurun_out/r02_soup16.ncu-rep
