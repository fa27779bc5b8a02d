#!/bin/bash
# One `ncu --set full` capture of the dominant render kernel (after the same command exited 0 without ncu), plus a light
# launch list.  usage (under gpurun): tools/ncu_full.sh TAG CONFIG SPP [kernel-regex]
TAG=$1; CFG=${2:-C5}; SPP=${3:-64}; K=${4:-render_interp_kernel}
CMD="python bench.py --config $CFG --steps 1 --warmup 1 --samples $SPP --no-cpu-baseline --no-e2e"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
tail -1 gpurun_out/plain_$TAG.log | cut -c1-300
ncu --set full --import-source on --clock-control none -k regex:$K -s 1 -c 1 -o gpurun_out/$TAG -f $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/$TAG.ncu-rep
