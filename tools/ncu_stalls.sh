#!/bin/bash
# usage: tools/ncu_stalls.sh SPP lib...   — metric pass (stall reasons, pipes, caches) on render_pool_kernel for each build
SPP=$1; shift
S=smsp__average_warps_issue_stalled
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sass__inst_executed_local_loads,sass__inst_executed_local_stores
for r in long_scoreboard wait not_selected short_scoreboard no_instruction math_pipe_throttle branch_resolving mio_throttle lg_throttle dispatch_stall imc_miss tex_throttle barrier; do M=$M,${S}_${r}_per_issue_active.ratio; done
for lib in "$@"; do
  export HRT_LIB=$PWD/hyper-ray-tracer_b200/csrc/$lib
  CMD="python bench.py --steps 1 --warmup 2 --samples $SPP --no-cpu-baseline --no-e2e"
  $CMD > gpurun_out/plain_$lib.log 2>&1 && ncu --metrics $M --clock-control none -k regex:render_pool_kernel -s 2 -c 1 --csv --log-file gpurun_out/m_$lib.csv $CMD > gpurun_out/ncu_$lib.log 2>&1
  python - "$lib" <<'PY'
import csv,json,sys
lib=sys.argv[1]
rows=[r for r in csv.reader(l for l in open('gpurun_out/m_%s.csv'%lib) if not l.startswith('=='))]
h=rows[0]; ni=h.index('Metric Name'); vi=h.index('Metric Value')
d={r[ni]:float(r[vi].replace(',','')) for r in rows[1:] if len(r)>vi and r[vi] not in ('', 'n/a')}
b=json.loads([l for l in open('gpurun_out/plain_%s.log'%lib) if l.startswith('{')][-1])
iss=d['smsp__issue_active.avg.pct_of_peak_sustained_active']/100
print('%s: plain %.1f Mpaths/s | issue %.1f%% warps %.1f%% lanes %.1f | L1 hit %.1f%% lsu-wf %.1f%% (shared %.1f%%) local ld/st %.0fM/%.0fM' % (lib, b['value'], iss*100, d['sm__warps_active.avg.pct_of_peak_sustained_active'], d['smsp__thread_inst_executed.sum']/d['smsp__inst_executed.sum'], d.get('l1tex__t_sector_hit_rate.pct',-1), d.get('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',-1), d.get('l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed',-1), d.get('sass__inst_executed_local_loads',0)/1e6, d.get('sass__inst_executed_local_stores',0)/1e6))
print('   pipes alu %.0f fma %.0f lsu %.0f xu %.0f | warps per scheduler-cycle in state: ' % tuple(d['sm__inst_executed_pipe_%s.avg.pct_of_peak_sustained_active'%p] for p in ('alu','fma','lsu','xu')) + ' '.join('%s %.2f' % (k.split('stalled_')[1].replace('_per_issue_active.ratio',''), v*iss) for k,v in sorted(d.items(), key=lambda kv:-kv[1]) if 'stalled' in k and v*iss>=0.01))
PY
done
