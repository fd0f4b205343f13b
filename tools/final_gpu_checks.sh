set -x
timeout 420 python -m pytest tests -m gpu -x -q -rs > gpurun_out/r02_pytest_gpu_v4.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_v4.log
B="python bench.py --steps 2 --warmup 3 --no-also --no-e2e --no-cpu-baseline --sustain-seconds 0"
SRC="ffpa_fwd_sm100.cuh sm100_ptx.cuh ffpa_fwd_launch.cu ffpa_fwd_bf16.cu"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:ffpa_fwd_kernel -s 1 -c 1 -f -o gpurun_out/r02_fwd_c2_v3 $B > /dev/null 2>&1
ncu -i gpurun_out/r02_fwd_c2_v3.ncu-rep --page raw --csv > gpurun_out/r02_fwd_c2_v3_raw.csv && python tools/update_roofline_traffic.py c2_self_fwd_b1h32n8192d512 gpurun_out/r02_fwd_c2_v3_raw.csv profiles/r02_fwd_ncu.md $SRC
timeout 150 ncu --set full --clock-control none --import-source on -k regex:ffpa_fwd_kernel -s 1 -c 1 -f -o gpurun_out/r02_fwd_c3_v3 $B --workload c3_gqa_causal_fwd_hq32hkv8n4096d512 > /dev/null 2>&1
ncu -i gpurun_out/r02_fwd_c3_v3.ncu-rep --page raw --csv > gpurun_out/r02_fwd_c3_v3_raw.csv && python tools/update_roofline_traffic.py c3_gqa_causal_fwd_hq32hkv8n4096d512 gpurun_out/r02_fwd_c3_v3_raw.csv profiles/r02_fwd_ncu.md $SRC
cp profiles/roofline_traffic.json gpurun_out/roofline_traffic.json
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_v3.csv $B > /dev/null 2>&1
timeout 420 python bench.py --no-ab > gpurun_out/r02_bench_v3.json 2> gpurun_out/r02_bench_v3.err; tail -c 600 gpurun_out/r02_bench_v3.json
