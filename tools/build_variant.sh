#!/bin/bash
# builds variants/libptb_<name>.so with extra nvcc flags (A/B experiments on launch bounds etc.); run from the repo root
# usage: tools/build_variant.sh name "-DPTB_SHADE_MIN_BLOCKS=8"
set -e
name=$1; flags=$2
mkdir -p variants/obj_$name
cd cpupathtrace_b200/csrc
nvcc $flags -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-O2 --expt-relaxed-constexpr -Xptxas -v -c ptb.cu -o ../../variants/obj_$name/ptb.o 2> ../../variants/obj_$name/ptxas.log
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/libptb_$name.so ../../variants/obj_$name/ptb.o ../lib/obj/bvh_build.o ../lib/obj/host_math.o -Xlinker -soname,libptb.so -lpthread
grep -A2 "shadeKernelINS_4RngTILb0\|traceClosestKernelILi2ELb0\|traceShadowKernelILb0" ../../variants/obj_$name/ptxas.log | grep "registers\|spill" | tr '\n' ' '; echo
