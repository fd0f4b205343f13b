#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:ffpa_bwd -s 3 -c 3 --csv --log-file gpurun_out/g_bwd_c2_stash.csv python tools/prof_bwd_c2.py > gpurun_out/g_ncu.log 2>&1
FFPA_BWD_STASH=0 timeout 600 ncu --metrics $M --clock-control none -k regex:ffpa_bwd -s 3 -c 3 --csv --log-file gpurun_out/g_bwd_c2_recompute.csv python tools/prof_bwd_c2.py >> gpurun_out/g_ncu.log 2>&1
for f in gpurun_out/g_bwd_c2_stash.csv gpurun_out/g_bwd_c2_recompute.csv; do echo $f; grep -E "gpu__time|tensor" $f | awk -F'","' '{print substr($5,1,38), $(NF-2), $NF}'; done
