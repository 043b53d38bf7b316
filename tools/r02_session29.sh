#!/bin/bash
# round-2 GPU session 29 (round-end validation, one B200): GPU tests, smoke, racecheck / synccheck of the PQ kernels, headline bench
# (C2) with the CPU arm, ncu launch list of the bench command, ncu --set full of the kernel the headline number is timed on (the
# in-flight search_kernel_v2, 9th launch of the timed region) and of search_kernel_pqs at C4, four batches in flight
O=gpurun_out/s29; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/gpu_tests.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.txt
for tool in racecheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_check.py pq > $O/sanitize_pq_$tool.log 2>&1; echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_check pq done|hazard" $O/sanitize_pq_$tool.log | tail -4
done
timeout 600 python bench.py > $O/bench_c2.json 2> $O/bench_c2.err; cut -c1-400 $O/bench_c2.json; tail -2 $O/bench_c2.err
timeout 300 python bench.py --in-flight 4 --steps 12 --warmup 4 --no-cpu-baseline > $O/bench_c2_inflight4.json 2> $O/bench_c2_inflight4.err; cut -c1-300 $O/bench_c2_inflight4.json; tail -2 $O/bench_c2_inflight4.err
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c2.csv python bench.py --steps 2 --warmup 3 --profile-range --no-cpu-baseline --no-parity > $O/launch_bench.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:search_kernel -s 8 -c 1 -o $O/prof_search_c2_inflight python bench.py --steps 1 --warmup 3 --profile-range --no-cpu-baseline --no-parity > $O/ncu_search_c2.log 2>&1; tail -1 $O/ncu_search_c2.log
timeout 700 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:search_kernel_pqs -s 3 -c 1 -o $O/prof_pqs_c4 python bench.py --workload c4_10Mx128_i8_pq32 --steps 1 --warmup 3 --profile-range --no-cpu-baseline --no-parity --l-search 800 > $O/ncu_pqs.log 2>&1; tail -1 $O/ncu_pqs.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/kernel_zoo_pq_launches.csv python tools/kernel_zoo.py pq > $O/kernel_zoo_pq.log 2>&1; tail -1 $O/kernel_zoo_pq.log | cut -c1-200
ls -la $O
