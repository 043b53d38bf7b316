#!/bin/bash
# round-2 GPU session 24: full GPU suite (wide integer frontier kernel, level 1 only for batches in flight), kernel zoo under ncu,
# full ncu capture of the in-flight search kernel, headline bench line
O=gpurun_out/s24; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/gpu_tests.txt
DAB_V2_T1_BYTES=1024 timeout 600 python -m pytest tests -m gpu -x -q -k "search_batch_identical or in_flight" 2>&1 | tail -2 | tee $O/gpu_tests_t1_1024.txt
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/kernel_zoo_launches.csv python tools/kernel_zoo.py > $O/kernel_zoo.log 2>&1; tail -2 $O/kernel_zoo.log | cut -c1-300
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:search_kernel -s 8 -c 1 -o $O/prof_search_c2 python bench.py --steps 1 --warmup 3 --profile-range --no-cpu-baseline --no-parity > $O/ncu_search.log 2>&1; tail -1 $O/ncu_search.log
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c2.csv python bench.py --steps 2 --warmup 3 --profile-range --no-cpu-baseline --no-parity > $O/launch_bench.log 2>&1
timeout 600 python bench.py --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err; cut -c1-300 $O/bench_c2.json
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.txt
