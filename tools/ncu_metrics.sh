#!/bin/bash
# Light ncu metric pass (stall reasons, pipes, caches, instruction counts) on one render kernel launch.
# usage (under gpurun): tools/ncu_metrics.sh TAG CONFIG SPP [kernel-regex]   -> gpurun_out/m_TAG.csv + a printed summary
TAG=$1; CFG=${2:-C5}; SPP=${3:-256}; K=${4:-render_interp_kernel}
S=smsp__average_warps_issue_stalled
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,sass__inst_executed_local_loads,sass__inst_executed_local_stores,smsp__warps_eligible.avg.per_cycle_active
for r in long_scoreboard wait not_selected short_scoreboard no_instruction math_pipe_throttle branch_resolving mio_throttle lg_throttle dispatch_stall imc_miss tex_throttle barrier; do M=$M,${S}_${r}_per_issue_active.ratio; done
CMD="python bench.py --config $CFG --steps 1 --warmup 1 --samples $SPP --no-cpu-baseline --no-e2e"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics $M --clock-control none -k regex:$K -s 1 -c 1 --csv --log-file gpurun_out/m_$TAG.csv $CMD > gpurun_out/ncu_$TAG.log 2>&1
python - "$TAG" <<'PY'
import csv,json,sys
tag=sys.argv[1]
rows=[r for r in csv.reader(l for l in open('gpurun_out/m_%s.csv'%tag) if not l.startswith('=='))]
h=rows[0]; ni=h.index('Metric Name'); vi=h.index('Metric Value')
d={r[ni]:float(r[vi].replace(',','')) for r in rows[1:] if len(r)>vi and r[vi] not in ('', 'n/a')}
b=json.loads([l for l in open('gpurun_out/plain_%s.log'%tag) if l.startswith('{')][-1])
iss=d['smsp__issue_active.avg.pct_of_peak_sustained_active']/100
rays=b['rays_per_path']*b['value']*1e6*b['ms_per_step']*1e-3
print('%s: plain %.1f Mpaths/s | warp-instr/ray %.0f | issue %.1f%% warps %.1f%% lanes %.1f | L1 hit %.1f%% L2 hit %.1f%% | local ld/st %.0fM/%.0fM | dram %.0f MB' % (tag, b['value'], d['smsp__inst_executed.sum']/rays, iss*100, d['sm__warps_active.avg.pct_of_peak_sustained_active'], d['smsp__thread_inst_executed.sum']/d['smsp__inst_executed.sum'], d.get('l1tex__t_sector_hit_rate.pct',-1), d.get('lts__t_sector_hit_rate.pct',-1), d.get('sass__inst_executed_local_loads',0)/1e6, d.get('sass__inst_executed_local_stores',0)/1e6, (d.get('dram__bytes_read.sum',0)+d.get('dram__bytes_write.sum',0))))
print('   pipes alu %.0f fma %.0f lsu %.0f xu %.0f | stalls per issue: ' % tuple(d['sm__inst_executed_pipe_%s.avg.pct_of_peak_sustained_active'%p] for p in ('alu','fma','lsu','xu')) + ' '.join('%s %.2f' % (k.split('stalled_')[1].replace('_per_issue_active.ratio',''), v) for k,v in sorted(d.items(), key=lambda kv:-kv[1]) if 'stalled' in k and v>=0.05))
PY
