#!/bin/bash
# A/B of diagnostic (-DHRT_SCHED_STATS) builds: throughput + scheduler statistics.  Usage: tools/ab_stats.sh CONFIG SPP lib...
cfg=$1; spp=$2; shift 2
mkdir -p gpurun_out
for lib in "$@"; do
  echo "== $lib"
  HRT_LIB=$PWD/hyper-ray-tracer_b200/csrc/$lib HRT_SCHED_STATS=1 python bench.py --config $cfg --steps 1 --warmup 1 --samples $spp --no-cpu-baseline --no-e2e 2> gpurun_out/stats_$lib.log | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   %.1f Mpaths/s %.1f ms' % (d['value'], d['ms_per_step']))"
  grep sched gpurun_out/stats_$lib.log | tail -14 | sed "s/^/   /"
done
