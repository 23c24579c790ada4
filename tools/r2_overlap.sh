#!/bin/bash
mkdir -p gpurun_out
python tools/r2_overlap_probe.py 1
for st in 0 8 16 24; do
for cfg in "16 6" "8 3" "5 3"; do
  set -- $cfg
  PTB_MULTI_STAGGER_MS=$st PTB_TRACE_BLOCKS_PER_SM=$1 PTB_SHADE_BLOCKS_PER_SM=$2 python tools/r2_overlap_probe.py 2 | sed "s/^/stagger $st: /"
done
done
