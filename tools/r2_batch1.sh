#!/bin/bash
# round 2, GPU batch 1: parity tests, then A/B of the first perf changes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader
timeout -s KILL 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
export SPP=128 T=200
bash tools/sweep.sh "VARIANT=base" "PTB_PRODUCTION_MATH=0" 2>&1
bash tools/variant_sweep.sh stack0 stack12 stack8_tb10 shade7 shade8 2>&1 | grep -v "^$"
