#!/bin/bash
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_bwd_gpu.py -q -m gpu -k "stash" --timeout 150 -p no:cacheprovider > gpurun_out/b_new_tests.log 2>&1
echo "new tests exit $?" >> gpurun_out/b_new_tests.log
tail -40 gpurun_out/b_new_tests.log
timeout 600 python -m pytest tests/test_bwd_gpu.py tests/test_varlen_gpu.py -q -m gpu --timeout 150 -p no:cacheprovider > gpurun_out/b_bwd_tests.log 2>&1
echo "bwd tests exit $?" >> gpurun_out/b_bwd_tests.log
tail -15 gpurun_out/b_bwd_tests.log
timeout 300 python tools/bench_more.py c2_self_d512 c2_causal_d512 c3_gqa_causal_n4096_d512 d384 > gpurun_out/b_bench_more.log 2>&1
cat gpurun_out/b_bench_more.log
