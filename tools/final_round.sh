#!/bin/bash
# (Superseded by tools/final_bench.sh + tools/final_profiles.sh, which fit a smaller GPU budget.)
# Round-end evidence on one GPU (run under gpurun): full GPU test suite, smoke, both bench arms, all configs, the launch
# list of the bench command, the per-kernel window and DRAM-traffic pass of the wavefront render, one `ncu --set full`
# capture per wave kernel (steady-state iteration) and of the persistent kernel on Cornell.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref_final.json
python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_final.json
python - <<'PY'
import json
r=json.load(open('gpurun_out/bench_ref_final.json')); b=json.load(open('gpurun_out/bench_final.json'))
print('reference arm: %.3f %s' % (r['value'], r['unit']))
print('b200 arm: value %.1f e2e %.1f %s  ms/step %.1f  roofline frac %.4f  cpu_baseline %.3f (%d cores)  clocks %s launches %s parity_vs_n1 %s' % (
    b['value'], b['e2e']['value'], b['unit'], b['ms_per_step'], b['roofline']['frac'], b['cpu_baseline']['value'], b['cpu_baseline']['cores'], b['clocks'], b['gpu_launches'], b['parity_vs_n1']))
PY
bash tools/bench_all.sh
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --samples 256 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
tools/wave_window.sh final C5 128 600 | tee gpurun_out/wave_window_final.txt
HRT_KERNEL=w tools/traffic_pass.sh C5 64 | tail -1 > gpurun_out/traffic_final.json
for k in wave_trace_kernel wave_logic_kernel wave_tree_kernel; do tools/ncu_full_at.sh r02_$k C5 64 $k 150 | tail -1; done
tools/ncu_full.sh r02_uniform_c3 C3 64 render_interp_kernel | tail -1
