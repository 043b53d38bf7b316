#!/bin/bash
# round-2 GPU session 32: full GPU suite on the final tree (wide-adjacency PQ test, 8-lane MinMax distance kernel), MinMax kernels
# under ncu, C3 line on the final build
O=gpurun_out/s32; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/gpu_tests.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.txt
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/kernel_zoo_minmax_launches.csv python tools/kernel_zoo.py minmax > $O/kernel_zoo_minmax.log 2>&1; tail -1 $O/kernel_zoo_minmax.log | cut -c1-200
timeout 600 python bench.py --workload c3_1Mx768_f16_ip --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err; cut -c1-330 $O/bench_c3.json; tail -2 $O/bench_c3.err
ls -la $O
