#!/bin/bash
# Light ncu pass on render_kernel (C5, 64 spp): instruction counts, issue utilisation, top stall reasons.
# usage (under gpurun): bash tools/ncu_light.sh <tag>
TAG=${1:-x}
CMD="python bench.py --steps 1 --warmup 3 --samples 64 --no-cpu-baseline --no-e2e"
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.sum,sm__cycles_active.avg
$CMD > gpurun_out/plain_$TAG.log 2>&1 && ncu --metrics $M --clock-control none -k regex:render_kernel -s 3 -c 1 --csv --log-file gpurun_out/m_$TAG.csv $CMD > gpurun_out/ncu_$TAG.log 2>&1
python - <<PY
import csv,json
rows=[r for r in csv.reader(l for l in open('gpurun_out/m_$TAG.csv') if not l.startswith('=='))]
h=rows[0]; ni=h.index('Metric Name'); vi=h.index('Metric Value')
d={r[ni]:float(r[vi].replace(',','')) for r in rows[1:] if len(r)>vi}
b=json.loads([l for l in open('gpurun_out/plain_$TAG.log') if l.startswith('{')][-1])
rays=b['rays_per_path']*800*800*64
print('plain run: %.1f Mpaths/s kernel %.1f ms'%(b['value'],b['kernel_ms']))
print('lanes/instr %.2f  warp-instr/ray %.0f  thread-instr/ray %.0f  issue_active %.1f%%  warps_active %.1f%%'%(d['smsp__thread_inst_executed.sum']/d['smsp__inst_executed.sum'], d['smsp__inst_executed.sum']/rays, d['smsp__thread_inst_executed.sum']/rays, d['smsp__issue_active.avg.pct_of_peak_sustained_active'], d['sm__warps_active.avg.pct_of_peak_sustained_active']))
for k,v in sorted(d.items()):
    if 'stalled' in k: print('  %-28s %.2f'%(k.split('stalled_')[1].split('_per')[0],v))
print('pipes fma %.1f alu %.1f lsu %.1f xu %.1f | L1 hit %.1f L2 hit %.1f | lsu wavefronts/cycle/SM %.3f'%(d['sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'],d['sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'],d['sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'],d['sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'],d['l1tex__t_sector_hit_rate.pct'],d['lts__t_sector_hit_rate.pct'], d['l1tex__data_pipe_lsu_wavefronts.sum']/148/d['sm__cycles_active.avg']))
PY
