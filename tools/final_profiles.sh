#!/bin/bash
# Round-end ncu evidence on one GPU (run under gpurun, after the same commands exited 0 without ncu): the launch list of a
# bench command, the per-kernel window and the DRAM-traffic pass of the wavefront render, one `ncu --set full` capture
# of a steady-state launch of each wave kernel.  TAG names the files (gpurun_out/TAG_*).
TAG=${1:-r02b}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --samples 256 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
tools/wave_window.sh $TAG C5 512 800 | tee gpurun_out/${TAG}_wave_window.txt
tools/traffic_pass.sh C5 64 | tail -1 > gpurun_out/${TAG}_traffic.json
for k in wave_trace_kernel wave_logic_kernel wave_tree_kernel; do tools/ncu_full_at.sh ${TAG}_$k C5 512 $k 100 | tail -1; done
