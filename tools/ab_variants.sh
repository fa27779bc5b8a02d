#!/bin/bash
# A/B of render-kernel build variants on one GPU (run under gpurun).  Usage: tools/ab_variants.sh SPP "name:lib[:ENV=VAL]" ...
# Each variant: python bench.py --samples SPP (CUDA-event timed, no profiler); lines -> gpurun_out/ab_variants.jsonl
spp=$1; shift
out=gpurun_out/ab_variants.jsonl
mkdir -p gpurun_out; : > $out
for v in "$@"; do
  IFS=: read -r name lib envkv <<< "$v"
  line=$(env HRT_LIB=$PWD/hyper-ray-tracer_b200/csrc/$lib ${envkv:-_X=1} python bench.py --steps 2 --warmup 3 --samples $spp --no-cpu-baseline --no-e2e 2>/dev/null | tail -1)
  echo "{\"variant\": \"$name\", \"line\": $line}" >> $out
  python - "$name" <<PY
import json,sys
d=json.loads('''$line''')
print(sys.argv[1], '%.1f Mpaths/s  %.1f ms  clocks %s grid %s' % (d['value'], d['ms_per_step'], d.get('clocks',{}).get('sm_mhz'), d['config'].get('grid')))
PY
done
