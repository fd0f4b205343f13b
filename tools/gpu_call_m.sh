#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_fwd_gpu.py -q -m gpu -k "replay or large_headdims or dispatch_smoke" --timeout 120 -x -p no:cacheprovider > gpurun_out/m_new_tests.log 2>&1
echo "new tests exit $?" >> gpurun_out/m_new_tests.log
tail -25 gpurun_out/m_new_tests.log
timeout 300 python tools/bench_more.py d1024 d768 > gpurun_out/m_bench_more.log 2>&1
cut -c1-500 gpurun_out/m_bench_more.log
