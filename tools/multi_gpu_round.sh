#!/bin/bash
# Multi-GPU evidence on one 8-GPU box (run under `gpurun --gpus 8`): BASELINE config 4 (Cornell-smoke) at 2 / 4 / 8 ranks and
# the headline config C5 at 8 ranks, each through bench.py's torchrun path (spp sharding + one NCCL all-reduce).
mkdir -p gpurun_out
run() {  # run CONFIG N
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29600 + $2)) bench.py --config $1 --gpus $2 --steps 2 --warmup 3 --no-cpu-baseline 2> gpurun_out/multi_$1_n$2.err | tail -1 > gpurun_out/multi_$1_n$2.json
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/multi_%s_n%s.json' % (sys.argv[1], sys.argv[2])))
    print('%s N=%s: %.1f Mpaths/s (e2e %.1f), %.1f ms/step, parity_vs_n1 %s' % (sys.argv[1], sys.argv[2], d['value'], d['e2e']['value'], d['ms_per_step'], d['parity_vs_n1']))
except Exception as e:
    print(sys.argv[1:], 'FAILED', e)
PY
}
run C5 8
run C4 8
run C4 4
run C4 2
python bench.py --config C4 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/multi_C4_n1.json
python -c "
import json; d=json.load(open('gpurun_out/multi_C4_n1.json')); print('C4 N=1: %.1f Mpaths/s (e2e %.1f)' % (d['value'], d['e2e']['value']))"
