#!/bin/bash
# All BASELINE configs on one GPU, CPU oracle timed beside each (run under gpurun); JSON lines -> gpurun_out/bench_all.jsonl
out=gpurun_out/bench_all.jsonl
: > $out
for cfg in C1 C2a C2b C3 C4; do  # (C5 is the default bench line)
  python bench.py --config $cfg --steps 2 --warmup 3 --cpu-seconds 10 2>/dev/null | tail -1 >> $out
done
python - <<'PY'
import json
for l in open('gpurun_out/bench_all.jsonl'):
    d=json.loads(l)
    print(d['config']['workload'], '| GPU %.1f Mpaths/s (%.0f Mrays/s) e2e %.1f | CPU %.3f Mpaths/s (%d cores) | x%.0f | roofline fp32 %.3f l2 %.3f' % (
        d['value'], d['mrays_per_s'], d['e2e']['value'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['e2e']['value']/d['cpu_baseline']['value'], d['roofline']['frac'], d['roofline']['l2']['frac']))
PY
