#!/bin/bash
# Diagnostic build of libhrt.so with extra -D switches, kept beside the product build:
#   tools/build_variant.sh lb4 -DHRT_LOGIC_BLOCKS=4      -> build/libhrt_lb4.so   (select with HRT_LIB=build/libhrt_lb4.so)
# The product library (hyper-ray-tracer_b200/csrc/libhrt.so) is rebuilt without switches afterwards.
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
cd "$root/hyper-ray-tracer_b200/csrc"
mkdir -p "$root/build"
touch hrt_kernels.cu
make -s EXTRA="$*" libhrt.so
grep -A3 "wave_logic\|wave_tree\|wave_trace" ptxas_fast.log | grep "Used" || true
cp libhrt.so "$root/build/libhrt_$name.so"
touch hrt_kernels.cu
make -s libhrt.so
