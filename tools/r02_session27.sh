#!/bin/bash
# round-2 GPU session 27: A/B of the search_kernel_pqs variants on one resident C4 index (tools/pq_ab.py), new GPU tests
O=gpurun_out/s27; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "device_memory or pq_traversal" 2>&1 | tail -3 | tee $O/gpu_tests_subset.txt
timeout 900 python tools/pq_ab.py c4_10Mx128_i8_pq32 500 800 2>&1 | tee $O/pq_ab_c4.txt | tail -16
timeout 300 python tools/pq_ab.py small_200Kx128_i8_pq32 350 2>&1 | tee $O/pq_ab_small.txt | tail -8
