#!/bin/bash
# round-2 GPU session 23 (one GPU): default bench lines with batches in flight (C2 incl. CPU arm, C3), C5-shaped run at 20M points,
# reference arm, ncu launch list of the pipelined bench command
O=gpurun_out/s23; mkdir -p $O
timeout 600 python bench.py > $O/bench_c2.json 2> $O/bench_c2.err; cut -c1-400 $O/bench_c2.json; tail -2 $O/bench_c2.err
timeout 600 python bench.py --workload c3_1Mx768_f16_ip --steps 10 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err; cut -c1-300 $O/bench_c3.json
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c2.csv python bench.py --steps 2 --warmup 3 --profile-range --no-cpu-baseline --no-parity > $O/launch_bench.log 2>&1
timeout 900 python bench.py --workload c5_100Mx96_f32_l2 --n-points 20000000 --steps 10 --warmup 3 > $O/bench_c5_20M.json 2> $O/bench_c5_20M.err; cut -c1-300 $O/bench_c5_20M.json; tail -3 $O/bench_c5_20M.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-600 $O/bench_reference.json; tail -2 $O/bench_reference.err
