#!/bin/bash
# round-2 GPU session 19: same-box A/B — HEAD (atomic inserts, global table only) vs the working tree without / with the shared-memory level
O=gpurun_out/s19; mkdir -p $O
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
b c2_t1_0 DAB_V2_T1_BYTES=0
b c2_t1_4096 A=1
b c2_t1_4096_full DAB_V2_FULL_GRID=1
b c2_t1_5120_full DAB_V2_FULL_GRID=1 DAB_V2_T1_BYTES=5120
b c2_head_again DAB_LIB_PATH=build/lib_head.so
BARGS="--in-flight 3"
b c2_t1_4096_full_f3 DAB_V2_FULL_GRID=1
BARGS="--workload c3_1Mx768_f16_ip --steps 10 --warmup 3"
b c3_head DAB_LIB_PATH=build/lib_head.so
b c3_t1_0 DAB_V2_T1_BYTES=0
b c3_t1_4096_full DAB_V2_FULL_GRID=1
