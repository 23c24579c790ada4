#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
bash tools/r2_sweep_env.sh PTB_PIPELINED_BATCHES 1 0 1 0
cp cpupathtrace_b200/lib/libptb.so /tmp/libptb_base.so
cp variants/libptb_prefetch.so cpupathtrace_b200/lib/libptb.so
bash tools/r2_sweep_env.sh PTB_VARIANT_PREFETCH 1 1
cp /tmp/libptb_base.so cpupathtrace_b200/lib/libptb.so
python tools/obj_load_bench.py 1000x500
PTB_MESH_THREADS=1 python tools/obj_load_bench.py 1000x500
