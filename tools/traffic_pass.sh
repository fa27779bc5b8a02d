#!/bin/bash
# DRAM traffic + issue utilisation of EVERY launch of a short frame (ncu metric pass), summed per kernel:
#   tools/traffic_pass.sh CONFIG SPP   -> gpurun_out/traffic_CONFIG.csv and a JSON line for profiles/r02_traffic.json
CFG=${1:-C5}; SPP=${2:-16}
CMD="python bench.py --config $CFG --steps 1 --warmup 0 --samples $SPP --no-cpu-baseline --no-e2e"
mkdir -p gpurun_out
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none --cache-control none -k regex:"wave_|render_" --csv --log-file gpurun_out/traffic_$CFG.csv $CMD > gpurun_out/traffic_$CFG.log 2>&1
python - "$CFG" "$SPP" <<'PY'
import csv, collections, json, sys
cfg, spp = sys.argv[1], int(sys.argv[2])
rows=[r for r in csv.reader(l for l in open('gpurun_out/traffic_%s.csv' % cfg) if not l.startswith('=='))]
h=rows[0]; ki=h.index('Kernel Name'); ni=h.index('Metric Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit'); ii=h.index('ID')
mult={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}
d=collections.defaultdict(dict)
for r in rows[1:]:
    v=float(r[vi].replace(',',''))
    if r[ni].startswith('dram__'): v*=mult.get(r[ui],1)
    d[(int(r[ii]), r[ki].split('(')[0])][r[ni]]=v
b=json.loads([l for l in open('gpurun_out/traffic_%s.log' % cfg) if l.startswith('{')][-1])
paths=b['value']*1e6*b['ms_per_step']*1e-3
# the frame itself: when it ran as a wavefront, only the wave kernels (bench.py's parity check renders a few samples
# with the persistent kernel after the timed step)
if any(k.startswith('wave_') for (_, k) in d):
    d={ik: m for ik, m in d.items() if ik[1].startswith('wave_')}
agg=collections.defaultdict(lambda: collections.Counter())
for (i,k),m in d.items():
    a=agg[k]; a['n']+=1; a['bytes']+=m['dram__bytes_read.sum']+m['dram__bytes_write.sum']; a['us']+=m['gpu__time_duration.sum']/1e3
    a['issue_x_us']+=m['smsp__issue_active.avg.pct_of_peak_sustained_active']*m['gpu__time_duration.sum']/1e3; a['inst']+=m['smsp__inst_executed.sum']
tot_b=sum(a['bytes'] for a in agg.values()); tot_us=sum(a['us'] for a in agg.values())
out={"dram_bytes_per_path": tot_b/paths, "paths_measured": paths, "spp": spp, "issue_active_pct": sum(a['issue_x_us'] for a in agg.values())/tot_us,
     "issue_source": "profiles/r02b_traffic_%s.csv.gz (time-weighted over all launches of a %d-spp frame)" % (cfg, spp),
     "kernels": {k: {"launches": a['n'], "time_share": a['us']/tot_us, "dram_bytes": a['bytes'], "warp_inst_per_path": a['inst']/paths} for k,a in agg.items()}}
print(json.dumps({cfg: out}))
PY
