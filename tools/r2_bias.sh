#!/bin/bash
cp cpupathtrace_b200/lib/libptb.so /tmp/libptb_base.so
for name in "$@"; do
  cp variants/libptb_$name.so cpupathtrace_b200/lib/libptb.so
  echo "== $name"; python -m pytest tests/test_parity_gpu.py -m gpu -q -x -s -k "bench_scene_image" 2>&1 | grep -E "^trimmed" | sed -E 's/trimmed RMSE.*clipped/clipped/'
done
cp /tmp/libptb_base.so cpupathtrace_b200/lib/libptb.so
