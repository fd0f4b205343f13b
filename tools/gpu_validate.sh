#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/final_smi.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu -x --timeout 200 -p no:cacheprovider > gpurun_out/final_all_tests.log 2>&1
echo "all tests exit $?" >> gpurun_out/final_all_tests.log
tail -4 gpurun_out/final_all_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
cat gpurun_out/final_bench.json | cut -c1-2500
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2>> gpurun_out/final_bench.err
cut -c1-600 gpurun_out/final_bench_ref.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/final_ncu_bench.log 2>&1
grep -c ffpa gpurun_out/final_launches.csv
timeout 300 python tools/bench_more.py c2_causal_d512 gqa_d512 self_n16384_d512 d320 d256 d128 c4_b4_d256 decode_nq1_n8192_d512 varlen > gpurun_out/final_bench_more.log 2>&1
cut -c1-330 gpurun_out/final_bench_more.log
