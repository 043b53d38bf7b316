#!/bin/bash
# round-2 GPU session 18: two-level visited set of search_kernel_v2 (shared-memory tags first, global table for the overflow)
O=gpurun_out/s18; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/gpu_tests.txt
DAB_V2_T1_BYTES=1024 timeout 900 python -m pytest tests -m gpu -x -q -k "search or build or smoke or flight or overflow" 2>&1 | tail -3 | tee $O/gpu_tests_t1_1024.txt
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
for t in 4096 0 2048 3072 6144 8192; do b c2_t1_$t DAB_V2_T1_BYTES=$t; done
BARGS="--workload c3_1Mx768_f16_ip --steps 10 --warmup 3"
for t in 4096 0 6144; do b c3_t1_$t DAB_V2_T1_BYTES=$t; done
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:search_kernel -c 1 -o $O/prof_search_c2 python bench.py --steps 1 --warmup 3 --profile-range --no-cpu-baseline --no-parity --in-flight 1 > $O/ncu_search.log 2>&1; tail -1 $O/ncu_search.log
