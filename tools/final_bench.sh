#!/bin/bash
# Round-end numbers on one GPU (run under gpurun): both bench arms at the default settings and one line per BASELINE config.
# -> gpurun_out/bench_ref_final.json, bench_final.json, bench_all.jsonl
mkdir -p gpurun_out
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref_final.json
python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_final.json
python - <<'PY'
import json
r=json.load(open('gpurun_out/bench_ref_final.json')); b=json.load(open('gpurun_out/bench_final.json'))
print('reference arm: %.3f %s' % (r['value'], r['unit']))
print('b200 arm: value %.1f e2e %.1f %s  ms/step %.1f  roofline frac %.4f  cpu_baseline %.3f (%d cores)  clocks %s launches %s parity_vs_n1 %s' % (
    b['value'], b['e2e']['value'], b['unit'], b['ms_per_step'], b['roofline']['frac'], b['cpu_baseline']['value'], b['cpu_baseline']['cores'], b['clocks'], b['gpu_launches'], b['parity_vs_n1']))
PY
bash tools/bench_all.sh
