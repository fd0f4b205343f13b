#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:ffpa_fwd --csv --log-file gpurun_out/n_fwd_d1024.csv python tools/prof_fwd_d1024.py > gpurun_out/n_ncu.log 2>&1
grep -E "gpu__time|tensor|dram" gpurun_out/n_fwd_d1024.csv | awk -F'","' '{print $1, substr($5,1,34), $(NF-2), $NF}'
