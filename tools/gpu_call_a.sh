#!/bin/bash
# dev validation run: new tests first (fail fast on hangs), then the whole GPU suite, then timings
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 420 python -m pytest tests/test_bwd_gpu.py tests/test_varlen_gpu.py -q -m gpu -k "large or varlen" --timeout 150 -p no:cacheprovider > gpurun_out/a_new_tests.log 2>&1
echo "new tests exit $?" >> gpurun_out/a_new_tests.log
tail -30 gpurun_out/a_new_tests.log
timeout 600 python -m pytest tests -q -m gpu --timeout 150 -p no:cacheprovider > gpurun_out/a_all_tests.log 2>&1
echo "all tests exit $?" >> gpurun_out/a_all_tests.log
tail -15 gpurun_out/a_all_tests.log
timeout 300 python tools/bench_more.py c2_self_d512 c3_gqa_causal_n4096_d512 d768 d1024 d128 varlen > gpurun_out/a_bench_more.log 2>&1
cat gpurun_out/a_bench_more.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
cat gpurun_out/a_bench.json
