#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,lts__t_sectors_srcunit_tex_op_read.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_elapsed.avg,lts__t_bytes.sum.per_second
timeout 600 ncu --metrics $M --clock-control none -k regex:ffpa_bwd -s 3 -c 3 --csv --log-file gpurun_out/c_bwd_c2_metrics.csv python tools/prof_bwd_c2.py > gpurun_out/c_ncu.log 2>&1
tail -5 gpurun_out/c_ncu.log
cat gpurun_out/c_bwd_c2_metrics.csv | tail -40
