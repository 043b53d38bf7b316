#!/bin/bash
# round-2 GPU session 4: search_kernel_v3 with the register-query fast path + linear-probing table; new bench.py; PQ training
mkdir -p gpurun_out/s5
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/s5/gpu_tests.txt
b() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/s5/$name.err > gpurun_out/s5/$name.json; python - $name <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/s5/{sys.argv[1]}.json"))
    m = d["config"]["at_min_l"]
    print(sys.argv[1], "ms/step %.3f" % d["ms_per_step"], "e2e ms %.3f" % d["e2e"]["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "recall", d["config"]["recall_at_10"], "minL ms %.3f" % m["ms_per_step"], "parity", (d["config"]["parity_gate"] or {}).get("result"))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(f"gpurun_out/s5/{sys.argv[1]}.err").read()[-600:])
PY
}
b v3_default A=1
b v3_generic DAB_V3_GENERIC=1
b v2_only DAB_DISABLE_V3=1
for l in build/lib_v3_*.so; do [ -f $l ] && b $(basename $l .so) DAB_LIB_PATH=$l; done
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:search_kernel_v3 -c 1 -o gpurun_out/s5/prof_v3 python bench.py --steps 1 --warmup 3 --profile-range --no-cpu-baseline --no-parity > gpurun_out/s5/ncu.log 2>&1; tail -2 gpurun_out/s5/ncu.log
timeout 600 python bench.py --workload small_200Kx128_i8_pq32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s5/pq_small.json 2> gpurun_out/s5/pq_small.err; tail -c 1500 gpurun_out/s5/pq_small.json; tail -5 gpurun_out/s5/pq_small.err
timeout 900 python bench.py --workload c3_1Mx768_f16_ip --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s5/c3.json 2> gpurun_out/s5/c3.err; tail -c 2500 gpurun_out/s5/c3.json; tail -5 gpurun_out/s5/c3.err
