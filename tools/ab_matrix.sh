#!/bin/bash
# A/B of builds and environment switches on one GPU (run under gpurun):
#   tools/ab_matrix.sh "C5:2048" "HRT_LIB=build/libhrt_base.so HRT_WAVE_PARTS=2,HRT_NO_CHAIN=1 ..."
# every second argument word is one variant: comma-separated VAR=value assignments ("-" = none).  Lines -> gpurun_out/ab_matrix.jsonl
out=gpurun_out/ab_matrix.jsonl
mkdir -p gpurun_out
for cs in $1; do
  IFS=: read -r cfg spp <<< "$cs"
  for kv in $2; do
    envs=$(echo "$kv" | tr ',' ' ')
    [ "$kv" = "-" ] && envs=""
    line=$(env $envs python bench.py --config $cfg --steps 2 --warmup 3 --samples $spp --no-cpu-baseline --no-e2e 2>gpurun_out/ab_matrix.err | tail -1)
    echo "{\"config\": \"$cfg\", \"spp\": $spp, \"env\": \"$kv\", \"line\": $line}" >> $out
    python - "$cfg" "$spp" "$kv" <<PY
import json,sys
try:
    d=json.loads('''$line''')
    print(sys.argv[1], sys.argv[2], sys.argv[3], '%.1f Mpaths/s  %.1f ms  rays/path %.4f launches %s clocks %s grid %s' % (d['value'], d['ms_per_step'], d.get('rays_per_path') or 0, d.get('gpu_launches'), d.get('clocks',{}).get('sm_mhz'), d['config'].get('grid')))
except Exception as e:
    print(sys.argv[1:], 'FAILED', e)
PY
  done
done
