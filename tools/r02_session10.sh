#!/bin/bash
# round-2 GPU session 10: TC epilogue v3, L sweep v3 vs v2 (crossover), tests
mkdir -p gpurun_out/s10
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/s10/gpu_tests.txt
timeout 400 python tools/bench_flat.py 2>&1 | tail -12 | tee gpurun_out/s10/bench_flat.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flat_tc_kernel -s 1 -c 1 -o gpurun_out/s10/prof_flat_tc python tools/bench_flat.py 1000000 > gpurun_out/s10/ncu_tc.log 2>&1; tail -2 gpurun_out/s10/ncu_tc.log
timeout 600 python tools/sweep_l.py c2_1Mx128_f32_l2 2>&1 | tail -16 | tee gpurun_out/s10/sweep_l_c2.txt
timeout 900 python tools/sweep_l.py c3_1Mx768_f16_ip 2>&1 | tail -16 | tee gpurun_out/s10/sweep_l_c3.txt
