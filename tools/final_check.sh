#!/bin/bash
# round-end validation on one B200: GPU tests, smoke, ncu launch list of the bench command, headline bench, reference arm
mkdir -p gpurun_out/final
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/final/gpu_tests.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/final/smoke.txt
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/launches.csv python bench.py --steps 2 --warmup 3 --profile-range --no-cpu-baseline --no-parity > gpurun_out/final/launch_bench.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:search_kernel -c 1 -o gpurun_out/final/prof_search python bench.py --steps 1 --warmup 3 --profile-range --no-cpu-baseline --no-parity > gpurun_out/final/ncu_search.log 2>&1
timeout 600 python bench.py > gpurun_out/final/bench_default.json 2> gpurun_out/final/bench_default.err; cat gpurun_out/final/bench_default.json | cut -c1-1500
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/final/bench_reference.json 2> gpurun_out/final/bench_reference.err; cat gpurun_out/final/bench_reference.json | cut -c1-1500; tail -3 gpurun_out/final/bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final/kernel_zoo_launches.csv python tools/kernel_zoo.py > gpurun_out/final/kernel_zoo.log 2>&1; tail -3 gpurun_out/final/kernel_zoo.log
