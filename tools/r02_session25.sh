#!/bin/bash
# round-2 GPU session 25: PQ traversal with the pivots in shared memory (search_kernel_pqs) and the fused LUT + ADC kernel
# (pq_fused_kernel): full GPU suite, A/B against the global-table kernels on the small PQ workload, C4 at full size,
# ncu --set full of the new traversal kernel at C4, PQ kernel zoo under ncu
O=gpurun_out/s25; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/gpu_tests.txt
S="--workload small_200Kx128_i8_pq32 --steps 5 --warmup 3 --no-cpu-baseline --l-search 350"
timeout 300 python bench.py $S > $O/pq_small_pqs.json 2> $O/pq_small_pqs.err
DAB_PQ_GLOBAL_LUT=1 timeout 300 python bench.py $S > $O/pq_small_global.json 2> $O/pq_small_global.err
DAB_PQ_WARPS=8 timeout 300 python bench.py $S --no-parity > $O/pq_small_pqs_w8.json 2> $O/pq_small_pqs_w8.err
for f in pqs global pqs_w8; do python - <<PY
import json
try:
    d = json.loads(open("$O/pq_small_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["ms_per_step"], 3), "ms/step", round(d["value"]), "QPS recall", d["config"]["recall_at_10"], d["config"].get("parity_gate"))
except Exception as e:
    print("$f failed", e); print(open("$O/pq_small_$f.err").read()[-1500:])
PY
done
timeout 900 python bench.py --workload c4_10Mx128_i8_pq32 --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err; cut -c1-400 $O/bench_c4.json; tail -3 $O/bench_c4.err
timeout 700 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:search_kernel_pqs -s 3 -c 1 -o $O/prof_pqs_c4 python bench.py --workload c4_10Mx128_i8_pq32 --steps 1 --warmup 3 --profile-range --no-cpu-baseline --no-parity --l-search 500 > $O/ncu_pqs.log 2>&1; tail -2 $O/ncu_pqs.log
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/kernel_zoo_pq_launches.csv python tools/kernel_zoo.py pq > $O/kernel_zoo_pq.log 2>&1; tail -1 $O/kernel_zoo_pq.log | cut -c1-400
ls -la $O
