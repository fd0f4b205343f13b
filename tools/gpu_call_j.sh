#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_bwd_gpu.py -q -m gpu -k "stash or large" --timeout 150 -p no:cacheprovider > gpurun_out/j_new_tests.log 2>&1
echo "new tests exit $?" >> gpurun_out/j_new_tests.log
tail -8 gpurun_out/j_new_tests.log
timeout 300 python tools/bench_more.py c2_self_d512 d768 d1024 > gpurun_out/j_bench_more.log 2>&1
cut -c1-420 gpurun_out/j_bench_more.log
