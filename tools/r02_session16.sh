#!/bin/bash
# round-2 GPU session 16: atomic-free visited-table inserts (A/B against DAB_V2_CAS_ONLY), smaller tables (DAB_V2_SLOTS)
O=gpurun_out/s16; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/gpu_tests.txt
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
b c2_nocas A=1
b c2_cas DAB_V2_CAS_ONLY=1
b c2_nocas_s2400 DAB_V2_SLOTS=2400
b c2_nocas_s3200 DAB_V2_SLOTS=3200
BARGS="--workload c3_1Mx768_f16_ip --steps 10 --warmup 3"
b c3_nocas A=1
b c3_cas DAB_V2_CAS_ONLY=1
