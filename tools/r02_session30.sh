#!/bin/bash
# round-2 GPU session 30: MinMax quantizer kernels (parity tests, racecheck / memcheck), fused LUT kernel with two query groups,
# ncu --set full of the two-level (batches in flight) variant of search_kernel_v2, selected by its template arguments
O=gpurun_out/s30; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/gpu_tests.txt
for tool in memcheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_check.py pq > $O/sanitize_pq_$tool.log 2>&1; echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_check pq done|hazard" $O/sanitize_pq_$tool.log | tail -4
done
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:search_kernel_v2<float, 0, 0, 4, 1>" -s 2 -c 1 -o $O/prof_search_c2_inflight python bench.py --steps 2 --warmup 3 --profile-range --no-cpu-baseline --no-parity > $O/ncu_search_c2.log 2>&1; tail -1 $O/ncu_search_c2.log
ls -la $O
