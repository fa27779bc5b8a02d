#!/bin/bash
# Round-end evidence on one GPU (run under gpurun): full GPU test suite, both bench arms, all configs, the launch list of
# the bench command and the DRAM-traffic pass of the dominant kernel.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref_final.json
python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_final.json
python - <<'PY'
import json
r=json.load(open('gpurun_out/bench_ref_final.json')); b=json.load(open('gpurun_out/bench_final.json'))
print('reference arm: %.3f %s' % (r['value'], r['unit']))
print('b200 arm: value %.1f e2e %.1f %s  ms/step %.1f  roofline frac %.4f  cpu_baseline %.3f (%d cores)  clocks %s launches %s' % (
    b['value'], b['e2e']['value'], b['unit'], b['ms_per_step'], b['roofline']['frac'], b['cpu_baseline']['value'], b['cpu_baseline']['cores'], b['clocks'], b['gpu_launches']))
PY
bash tools/bench_all.sh
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:render_pool_kernel -s 1 -c 1 --csv --log-file gpurun_out/traffic_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_traffic.log 2>&1
tail -3 gpurun_out/traffic_final.csv | cut -c1-300
