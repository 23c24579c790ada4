#!/bin/bash
# builds variants/libptb_<name>.so with extra nvcc flags (A/B experiments on launch bounds etc.); run from the repo root
# usage: tools/build_variant.sh name "<flags for both CUDA translation units>" ["<extra flags for ptb_fast.cu only>"]
set -e
name=$1; flags=$2; fast_flags=$3
mkdir -p variants/obj_$name
cd cpupathtrace_b200/csrc
nvcc $flags -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-O2 --expt-relaxed-constexpr -Xptxas -v -c ptb.cu -o ../../variants/obj_$name/ptb.o 2> ../../variants/obj_$name/ptxas.log &
nvcc $flags $fast_flags -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=true -use_fast_math -Xcompiler -fPIC,-O2 --expt-relaxed-constexpr -Xptxas -v -c ptb_fast.cu -o ../../variants/obj_$name/ptb_fast.o 2> ../../variants/obj_$name/ptxas_fast.log &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/libptb_$name.so ../../variants/obj_$name/ptb.o ../../variants/obj_$name/ptb_fast.o ../lib/obj/bvh_build.o ../lib/obj/host_math.o ../lib/obj/cert_guard.o -Xlinker -soname,libptb.so -lpthread
echo -n "$name: "
grep -A2 "shadeKernelINS_4RngTILb0\|traceClosestKernelILi2ELb0\|traceShadowKernelILb0" ../../variants/obj_$name/ptxas.log ../../variants/obj_$name/ptxas_fast.log | grep "registers\|spill" | sed -E 's/.*ptxas info    : //' | tr '\n' ' '; echo
