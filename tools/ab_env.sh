#!/bin/bash
# A/B of environment-selected variants of the CURRENT build on one GPU (run under gpurun):
#   tools/ab_env.sh "C5:512 C3:512" "HRT_SHAPE=0 HRT_SHAPE=1 ..."      lines -> gpurun_out/ab_env.jsonl
out=gpurun_out/ab_env.jsonl
mkdir -p gpurun_out
for cs in $1; do
  IFS=: read -r cfg spp <<< "$cs"
  for kv in $2; do
    line=$(env $kv python bench.py --config $cfg --steps 2 --warmup 3 --samples $spp --no-cpu-baseline --no-e2e 2>/dev/null | tail -1)
    echo "{\"config\": \"$cfg\", \"spp\": $spp, \"env\": \"$kv\", \"line\": $line}" >> $out
    python - "$cfg" "$spp" "$kv" <<PY
import json,sys
try:
    d=json.loads('''$line''')
    print(sys.argv[1], sys.argv[2], sys.argv[3], '%.1f Mpaths/s  %.1f ms  clocks %s grid %s block %s' % (d['value'], d['ms_per_step'], d.get('clocks',{}).get('sm_mhz'), d['config'].get('grid'), d['config'].get('block')))
except Exception as e:
    print(sys.argv[1:], 'FAILED', e)
PY
  done
done
