#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_bwd_gpu.py -q -m gpu -k "chunked" --timeout 200 -p no:cacheprovider > gpurun_out/k_chunk_test.log 2>&1
echo "chunk test exit $?" >> gpurun_out/k_chunk_test.log
tail -12 gpurun_out/k_chunk_test.log
timeout 900 python -m pytest tests -q -m gpu --timeout 200 -p no:cacheprovider > gpurun_out/k_all_tests.log 2>&1
echo "all tests exit $?" >> gpurun_out/k_all_tests.log
tail -6 gpurun_out/k_all_tests.log
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread
timeout 900 ncu --metrics $M --clock-control none -k regex:"ffpa|quantize|merge|colsum|kmean|preprocess" --csv --log-file gpurun_out/k_variants.csv python tools/prof_variants.py > gpurun_out/k_variants.log 2>&1
tail -3 gpurun_out/k_variants.log; wc -l gpurun_out/k_variants.csv
