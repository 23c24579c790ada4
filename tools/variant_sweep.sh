#!/bin/bash
# usage: tools/variant_sweep.sh name1 name2 ...  — swaps variants/libptb_<name>.so in and runs one bench each ("base" = the built library)
cp cpupathtrace_b200/lib/libptb.so /tmp/libptb_base.so
for name in "$@"; do
  if [ "$name" == "base" ]; then cp /tmp/libptb_base.so cpupathtrace_b200/lib/libptb.so; else cp variants/libptb_$name.so cpupathtrace_b200/lib/libptb.so; fi
  SPP=${SPP:-256} T=${T:-200} bash tools/sweep.sh "VARIANT=$name"
done
cp /tmp/libptb_base.so cpupathtrace_b200/lib/libptb.so
