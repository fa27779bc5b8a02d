#!/bin/bash
# `ncu --set full` of ONE launch of a kernel, skipping the first SKIP matching launches (a representative iteration of the
# wavefront render).  usage (under gpurun): tools/ncu_full_at.sh TAG CONFIG SPP KERNEL-REGEX SKIP
TAG=$1; CFG=${2:-C5}; SPP=${3:-64}; K=${4:-wave_trace_kernel}; SKIP=${5:-200}
CMD="python bench.py --config $CFG --steps 1 --warmup 0 --samples $SPP --no-cpu-baseline --no-e2e"
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:$K -s $SKIP -c 1 -o gpurun_out/$TAG -f $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/$TAG.ncu-rep
