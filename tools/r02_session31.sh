#!/bin/bash
# round-2 GPU session 31: ncu --set full of the two-level (batches in flight) variant of search_kernel_v2 — with --steps 1
# --warmup 3 the timed region launches 9 global-table kernels first (4 + 1 overflow re-run + 4), the 11th launch is in the
# in-flight loop; MinMax kernels under ncu (tools/kernel_zoo.py minmax)
O=gpurun_out/s31; mkdir -p $O
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:search_kernel -s 10 -c 1 -o $O/prof_search_c2_inflight python bench.py --steps 1 --warmup 3 --profile-range --no-cpu-baseline --no-parity > $O/ncu_search_c2.log 2>&1; tail -2 $O/ncu_search_c2.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/kernel_zoo_minmax_launches.csv python tools/kernel_zoo.py minmax > $O/kernel_zoo_minmax.log 2>&1; tail -1 $O/kernel_zoo_minmax.log | cut -c1-400
ls -la $O
