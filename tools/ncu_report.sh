#!/bin/bash
# Summaries of an ncu report for profiles/: tools/ncu_report.sh gpurun_out/TAG.ncu-rep KERNEL-SUBSTRING profiles/PREFIX
# -> PREFIX_raw.csv (all metrics), PREFIX_summary.txt (key figures + stalls per issue), PREFIX_by_function.txt
# (the SASS of the CURRENT production build must be the one that was profiled)
rep=$1; k=$2; out=$3
ncu -i $rep --page raw --csv > ${out}_raw.csv 2>/dev/null
python tools/ncu_raw_summary.py ${out}_raw.csv > ${out}_summary.txt
ncu -i $rep --page source --csv > /tmp/ncu_src_$$.csv 2>/dev/null
nvdisasm --print-line-info hyper-ray-tracer_b200/csrc/hrt_kernels_fast.cubin > /tmp/ncu_dis_$$.txt 2>/dev/null || {
  cuobjdump -xelf all hyper-ray-tracer_b200/csrc/hrt_kernels_fast.o >/dev/null 2>&1; mv *.cubin /tmp/ 2>/dev/null; nvdisasm --print-line-info /tmp/*sm_100a*.cubin > /tmp/ncu_dis_$$.txt; }
python tools/ncu_by_function.py /tmp/ncu_src_$$.csv /tmp/ncu_dis_$$.txt $k > ${out}_by_function.txt
