#!/bin/bash
# quick per-config throughput of the current build (no CPU arm, no e2e): tools/cfg_quick.sh C3 C4 ...
for cfg in "$@"; do
  python bench.py --config $cfg --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$cfg %.1f Mpaths/s grid %s block %s' % (d['value'], d['config']['grid'], d['config']['block']))"
done
