#!/bin/bash
# round-2 GPU session 17: resident warps of search_kernel_v2 (launch bounds -> 72 / 64 registers, two-warp CTAs), ncu of the atomic-free kernel
O=gpurun_out/s17; mkdir -p $O
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
b c2_default A=1
for v in v2_c26 v2_c30 v2_w2c13 v2_w2c16; do b c2_$v DAB_LIB_PATH=build/lib_$v.so; done
BARGS="--workload c3_1Mx768_f16_ip --steps 10 --warmup 3"
for v in v2_c26 v2_c30 v2_w2c16; do b c3_$v DAB_LIB_PATH=build/lib_$v.so; done
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:search_kernel -c 1 -o $O/prof_search_c2 python bench.py --steps 1 --warmup 3 --profile-range --no-cpu-baseline --no-parity --in-flight 1 > $O/ncu_search.log 2>&1; tail -1 $O/ncu_search.log
