#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
for d in 1 2 3; do
FFPA_DBG=$d PROF_H=16 timeout 300 ncu --metrics $M --clock-control none -k regex:ffpa_bwd_kernel -s 1 -c 1 --csv --log-file gpurun_out/i_dbg$d.csv python tools/prof_bwd_c2.py > gpurun_out/i_ncu.log 2>&1
echo "dbg=$d"; grep -E "gpu__time|tensor" gpurun_out/i_dbg$d.csv | awk -F'","' '{print $(NF-2), $NF}'
done
FFPA_DBG=0 PROF_H=16 timeout 300 ncu --metrics $M --clock-control none -k regex:ffpa_bwd_kernel -s 1 -c 1 --csv --log-file gpurun_out/i_dbg0.csv python tools/prof_bwd_c2.py > gpurun_out/i_ncu.log 2>&1
echo "dbg=0"; grep -E "gpu__time|tensor" gpurun_out/i_dbg0.csv | awk -F'","' '{print $(NF-2), $NF}'
FFPA_BWD_STASH=0 PROF_H=16 timeout 300 ncu --metrics $M --clock-control none -k regex:ffpa_bwd_kernel -s 3 -c 1 --csv --log-file gpurun_out/i_nostash.csv python tools/prof_bwd_c2.py > gpurun_out/i_ncu.log 2>&1
echo "nostash"; grep -E "gpu__time|tensor" gpurun_out/i_nostash.csv | awk -F'","' '{print $(NF-2), $NF}'
