#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_bwd_gpu.py -q -m gpu -k "stash" --timeout 150 -p no:cacheprovider > gpurun_out/h_new_tests.log 2>&1
echo "new tests exit $?" >> gpurun_out/h_new_tests.log
tail -5 gpurun_out/h_new_tests.log
timeout 300 python tools/bench_more.py c2_self_d512 c3_gqa_causal_n4096_d512 > gpurun_out/h_bench_more.log 2>&1
cut -c1-420 gpurun_out/h_bench_more.log
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:ffpa_bwd -s 3 -c 3 --csv --log-file gpurun_out/h_bwd_c2_metrics.csv python tools/prof_bwd_c2.py > gpurun_out/h_ncu.log 2>&1
grep -E "gpu__time|tensor|dram" gpurun_out/h_bwd_c2_metrics.csv | awk -F'","' '{print substr($5,1,38), $(NF-2), $NF}'
