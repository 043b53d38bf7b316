#!/bin/bash
# round-2 GPU session 26: search_kernel_pqs with the next hop's row copied ahead + bucket prefetch, lists up to 1024 entries
# (two-tile merge), f16 / float-cosine rerank; full GPU suite, small PQ A/B, C4 with the L sweep up to 1000, strong-scaling
# shares of C2 on one GPU (tools/nq_sweep.py)
O=gpurun_out/s26; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/gpu_tests.txt
S="--workload small_200Kx128_i8_pq32 --steps 5 --warmup 3 --no-cpu-baseline --l-search 350"
timeout 300 python bench.py $S > $O/pq_small_pqs.json 2> $O/pq_small_pqs.err
for f in pqs; do python - <<PY
import json
try:
    d = json.loads(open("$O/pq_small_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["ms_per_step"], 3), "ms/step", round(d["value"]), "QPS recall", d["config"]["recall_at_10"], d["config"].get("parity_gate"))
except Exception as e:
    print("$f failed", e); print(open("$O/pq_small_$f.err").read()[-1500:])
PY
done
timeout 1200 python bench.py --workload c4_10Mx128_i8_pq32 > $O/bench_c4.json 2> $O/bench_c4.err; cut -c1-330 $O/bench_c4.json; tail -3 $O/bench_c4.err
timeout 600 python tools/nq_sweep.py c2_1Mx128_f32_l2 2>&1 | tee $O/nq_sweep_c2.txt | tail -12
ls -la $O
