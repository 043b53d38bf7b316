#!/bin/bash
# round-2 GPU session 20: per-phase cycles of search_kernel_v2 — HEAD vs working tree (T1 off / on)
O=gpurun_out/s20; mkdir -p $O
ph() { name=$1; shift; env DAB_PHASE_PROFILE=1 "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --in-flight 1 2>&1 >/dev/null | grep "phase profile" | tail -1 | sed "s/^/$name /" | tee -a $O/phases.txt; }
ph head DAB_LIB_PATH=build/lib_head_phase.so
ph cur_t1_0 DAB_LIB_PATH=build/lib_cur_phase.so DAB_V2_T1_BYTES=0
ph cur_t1_4096 DAB_LIB_PATH=build/lib_cur_phase.so
