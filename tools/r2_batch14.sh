#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
bash tools/r2_sweep_env.sh PTB_STREAMS 2 1 3 2 1
