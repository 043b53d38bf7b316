#!/bin/bash
# round-2 GPU session 22: compact two-level kernel (L1 template parameter) vs HEAD on one box; exact-next / tail-prefetch variants
O=gpurun_out/s22; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/gpu_tests.txt
b() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline $BARGS 2>$O/$name.err > $O/$name.json; python - $O/$name <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1] + ".json"))
    print(sys.argv[1].split("/")[-1], "in flight", d["config"]["batches_in_flight"], "ms/step %.3f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "e2e ms %.3f" % d["e2e"]["ms_per_step"],
          "serial ms %.3f" % d["config"]["serial"]["ms_per_step"], "recall", d["config"]["recall_at_10"], "parity", (d["config"]["parity_gate"] or {}).get("result"))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(sys.argv[1] + ".err").read()[-600:])
PY
}
BARGS=""
b c2_head DAB_LIB_PATH=build/lib_head.so
b c2_cur A=1
b c2_cur_t1_0 DAB_V2_T1_BYTES=0
b c2_exact DAB_LIB_PATH=build/lib_exact.so
b c2_tail DAB_LIB_PATH=build/lib_tail.so
b c2_cur_t1_3072 DAB_V2_T1_BYTES=3072
b c2_cur_t1_5120 DAB_V2_T1_BYTES=5120
BARGS="--workload c3_1Mx768_f16_ip --steps 10 --warmup 3"
b c3_head DAB_LIB_PATH=build/lib_head.so
b c3_cur A=1
b c3_tail DAB_LIB_PATH=build/lib_tail.so
b c3_exact DAB_LIB_PATH=build/lib_exact.so
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:search_kernel -c 1 -o $O/prof_search_c2 python bench.py --steps 1 --warmup 3 --profile-range --no-cpu-baseline --no-parity --in-flight 1 > $O/ncu_search.log 2>&1; tail -1 $O/ncu_search.log
