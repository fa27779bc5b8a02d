#!/bin/bash
# usage: bash tools/ncu_light2.sh <tag> <lib> <spp>   (render_pool_kernel, light metric pass)
TAG=$1; export HRT_LIB=$PWD/$2; SPP=${3:-512}
CMD="python bench.py --steps 1 --warmup 2 --samples $SPP --no-cpu-baseline --no-e2e"
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
$CMD > gpurun_out/plain_$TAG.log 2>&1 && ncu --metrics $M --clock-control none -k regex:render_pool_kernel -s 2 -c 1 --csv --log-file gpurun_out/m_$TAG.csv $CMD > gpurun_out/ncu_$TAG.log 2>&1
python - <<PY
import csv,json
rows=[r for r in csv.reader(l for l in open('gpurun_out/m_$TAG.csv') if not l.startswith('=='))]
h=rows[0]; ni=h.index('Metric Name'); vi=h.index('Metric Value')
d={r[ni]:float(r[vi].replace(',','')) for r in rows[1:] if len(r)>vi}
b=json.loads([l for l in open('gpurun_out/plain_$TAG.log') if l.startswith('{')][-1])
rays=b['rays_per_path']*800*800*$SPP
print('$TAG: plain %.1f Mpaths/s | warp-instr/ray %.0f thread-instr/ray %.0f lanes %.1f | issue %.1f%% warps %.1f%% | alu %.0f fma %.0f lsu %.0f xu %.0f | stalls long_sb %.2f wait %.2f notsel %.2f short_sb %.2f noinst %.2f math %.2f'%(b['value'], d['smsp__inst_executed.sum']/rays, d['smsp__thread_inst_executed.sum']/rays, d['smsp__thread_inst_executed.sum']/d['smsp__inst_executed.sum'], d['smsp__issue_active.avg.pct_of_peak_sustained_active'], d['sm__warps_active.avg.pct_of_peak_sustained_active'], d['sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'], d['sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'], d['sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'], d['sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'], d['smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio'], d['smsp__average_warps_issue_stalled_wait_per_issue_active.ratio'], d['smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio'], d['smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio'], d['smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio'], d['smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio']))
PY
