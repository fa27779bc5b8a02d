"""Key figures of one `ncu --page raw --csv` export: tools/ncu_raw_summary.py <raw.csv>"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]
d = dict(zip(h, rows[2] if len(rows) > 2 else rows[1]))
keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'smsp__warps_eligible.avg.per_cycle_active', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'sass__inst_executed_shared_loads', 'sass__inst_executed_global_loads']
for k in keys:
    for kk in d:
        if kk == k:
            print(f"{kk:70s} {d[kk]}")
st = []
for kk in d:
    if 'issue_stalled' in kk and kk.endswith('per_issue_active.ratio') and 'not_issued' not in kk:
        try:
            st.append((float(d[kk]), kk.split('stalled_')[1].replace('_per_issue_active.ratio', '')))
        except ValueError:
            pass
print('stalls per issue: ' + ', '.join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True) if v >= 0.05))
