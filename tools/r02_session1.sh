#!/bin/bash
# round-2 GPU session 1: A/B of the compiled-out search_kernel_v2 experiments + sanitizer passes
mkdir -p gpurun_out/s1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tail -1
for l in build/lib_*.so; do tools/quick_check.sh $l; done 2>&1 | tee gpurun_out/s1/quick.txt
for n in tag16 int wide split defer; do
  DAB_LIB_PATH=build/lib_$n.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2 | sed "s/^/$n: /"
done | tee gpurun_out/s1/parity.txt
tools/sweep_libs.sh build/lib_*.so 2>&1 | tee gpurun_out/s1/sweep.txt
tools/sweep_libs.sh build/lib_*.so 2>&1 | tee gpurun_out/s1/sweep2.txt
timeout 400 compute-sanitizer --tool racecheck python tools/sanitize_check.py > gpurun_out/s1/racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/s1/racecheck.txt
timeout 300 compute-sanitizer --tool synccheck python tools/sanitize_check.py > gpurun_out/s1/synccheck.txt 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/s1/synccheck.txt
