#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fwd_gpu.py tests/test_fp8_gpu.py -q -m gpu --timeout 150 -p no:cacheprovider > gpurun_out/l_fwd_tests.log 2>&1
echo "fwd tests exit $?" >> gpurun_out/l_fwd_tests.log
tail -6 gpurun_out/l_fwd_tests.log
timeout 300 python tools/bench_more.py d128 d256 c4_b4_d256 c2_self_d512 > gpurun_out/l_bench_more.log 2>&1
cut -c1-420 gpurun_out/l_bench_more.log
