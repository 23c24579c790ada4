#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -v "^$" | tail -25
export SPP=128 T=200 CUDA_VISIBLE_DEVICES=0
bash tools/sweep.sh "VARIANT=base" "VARIANT=base -- --guarded" "PTB_ITERATIONS_PER_SYNC=8" "PTB_ITERATIONS_PER_SYNC=16" "PTB_PROFILE=0" 2>&1
