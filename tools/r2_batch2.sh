#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
export SPP=128 T=200
bash tools/sweep.sh "VARIANT=base" "PTB_PRODUCTION_MATH=0" "VARIANT=base -- --guarded" 2>&1
bash tools/variant_sweep.sh closest12 flat256 shade7 shade8 flat256s4 2>&1 | grep -v "^$"
