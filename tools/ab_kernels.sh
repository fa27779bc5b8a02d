#!/bin/bash
# A/B of the render-kernel variants of the CURRENT build on one GPU (run under gpurun): tools/ab_kernels.sh "C5:1024 C3:512 ..." "pool uniform ..."
# Lines -> gpurun_out/ab_kernels.jsonl
out=gpurun_out/ab_kernels.jsonl
mkdir -p gpurun_out
for cs in $1; do
  IFS=: read -r cfg spp <<< "$cs"
  for k in $2; do
    line=$(HRT_KERNEL=$k python bench.py --config $cfg --steps 2 --warmup 3 --samples $spp --no-cpu-baseline --no-e2e 2>/dev/null | tail -1)
    echo "{\"config\": \"$cfg\", \"spp\": $spp, \"kernel\": \"$k\", \"line\": $line}" >> $out
    python - "$cfg" "$spp" "$k" <<PY
import json,sys
try:
    d=json.loads('''$line''')
    print(sys.argv[1], sys.argv[2], sys.argv[3], '%.1f Mpaths/s  %.1f ms  rays/path %.3f clocks %s grid %s' % (d['value'], d['ms_per_step'], d.get('rays_per_path') or 0, d.get('clocks',{}).get('sm_mhz'), d['config'].get('grid')))
except Exception as e:
    print(sys.argv[1:], 'FAILED', e)
PY
  done
done
