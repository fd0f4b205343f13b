#!/bin/bash
mkdir -p gpurun_out
PROF_H=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ffpa_bwd_kernel -s 1 -c 1 -o gpurun_out/d_bwd_dq_stash -f python tools/prof_bwd_c2.py > gpurun_out/d_ncu.log 2>&1
tail -3 gpurun_out/d_ncu.log
ls -la gpurun_out/
