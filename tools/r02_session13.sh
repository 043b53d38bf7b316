#!/bin/bash
# round-2 GPU session 13 (evidence): full GPU suite, smoke, C2/C3/C4 bench lines, ncu launch list + full capture of the
# search kernel, tcgen05 scan timing + capture, one launch of every distance kernel family under ncu
O=gpurun_out/s13; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/gpu_tests.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.txt
timeout 600 python bench.py > $O/bench_c2.json 2> $O/bench_c2.err; cut -c1-1200 $O/bench_c2.json
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c2.csv python bench.py --steps 2 --warmup 3 --profile-range --no-cpu-baseline --no-parity > $O/launch_bench.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:search_kernel -c 1 -o $O/prof_search_c2 python bench.py --steps 1 --warmup 3 --profile-range --no-cpu-baseline --no-parity > $O/ncu_search.log 2>&1; tail -1 $O/ncu_search.log
timeout 900 python bench.py --workload c3_1Mx768_f16_ip --steps 10 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err; cut -c1-1200 $O/bench_c3.json
timeout 300 ncu --profile-from-start off --set full --clock-control none -k regex:search_kernel -c 1 -o $O/prof_search_c3 python bench.py --workload c3_1Mx768_f16_ip --steps 1 --warmup 3 --profile-range --no-cpu-baseline --no-parity --l-search 100 > $O/ncu_search_c3.log 2>&1; tail -1 $O/ncu_search_c3.log
timeout 900 python bench.py --workload c4_10Mx128_i8_pq32 --steps 10 --warmup 3 > $O/bench_c4.json 2> $O/bench_c4.err; cut -c1-1200 $O/bench_c4.json
timeout 400 python tools/bench_flat.py 2>&1 | tail -8 | tee $O/bench_flat.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flat_tc_kernel -s 1 -c 1 -o $O/prof_flat_tc python tools/bench_flat.py 1000000 > $O/ncu_tc.log 2>&1; tail -1 $O/ncu_tc.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/kernel_zoo_launches.csv python tools/kernel_zoo.py > $O/kernel_zoo.log 2>&1; tail -3 $O/kernel_zoo.log
ls -la $O
