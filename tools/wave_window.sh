#!/bin/bash
# Per-kernel figures of a steady-state window of the wavefront render (64 consecutive launches): tools/wave_window.sh TAG CONFIG SPP SKIP
TAG=$1; CFG=${2:-C5}; SPP=${3:-128}; SKIP=${4:-400}
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --cache-control none -s $SKIP -c 64 --csv --log-file gpurun_out/window_$TAG.csv python bench.py --config $CFG --steps 1 --warmup 0 --samples $SPP --no-cpu-baseline --no-e2e > gpurun_out/window_$TAG.log 2>&1
python - "$TAG" <<'PY'
import csv, collections, sys
rows=[r for r in csv.reader(l for l in open('gpurun_out/window_%s.csv' % sys.argv[1]) if not l.startswith('=='))]
h=rows[0]; ki=h.index('Kernel Name'); ni=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID')
d=collections.defaultdict(dict)
for r in rows[1:]:
    d[(int(r[ii]), r[ki].split('(')[0])][r[ni]]=float(r[vi].replace(',',''))
agg=collections.defaultdict(lambda: collections.Counter())
for (i,k),m in sorted(d.items()):
    a=agg[k]; a['n']+=1; a['us']+=m['gpu__time_duration.sum']/1e3; a['inst']+=m['smsp__inst_executed.sum']; a['lanes']+=m['smsp__thread_inst_executed_per_inst_executed.ratio']; a['issue']+=m['smsp__issue_active.avg.pct_of_peak_sustained_active']
tot=sum(a['us'] for a in agg.values())
for k,a in agg.items(): print('%-20s n %3d  avg %7.1f us  share %4.1f%%  Minst %6.1f  lanes %4.1f  issue %4.1f%%'%(k, a['n'], a['us']/a['n'], 100*a['us']/tot, a['inst']/a['n']/1e6, a['lanes']/a['n'], a['issue']/a['n']))
PY
