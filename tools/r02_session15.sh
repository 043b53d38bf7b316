#!/bin/bash
# round-2 GPU session 15: with two batches in flight, which kernel wins at the headline L? v2 (global tables) vs v3 (shared-memory tables)
O=gpurun_out/s15; mkdir -p $O
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
b c2_v2 A=1
b c2_v3 DAB_V3_MAX_CAP=200
b c2_v3_c5 DAB_V3_MAX_CAP=200 DAB_LIB_PATH=build/lib_v3_c5.so
BARGS="--in-flight 3"
b c2_v2_f3 A=1
b c2_v3_f3 DAB_V3_MAX_CAP=200
BARGS="--workload c3_1Mx768_f16_ip --steps 10 --warmup 3"
b c3_v2 A=1
b c3_v3 DAB_V3_MAX_CAP=200
