#!/bin/bash
# round-2 GPU session 21: exact next-node adjacency fetch and L2 prefetch of the later row groups, A/B on one box
O=gpurun_out/s21; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "search or build or smoke or flight or overflow" 2>&1 | tail -3 | tee $O/gpu_tests.txt
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
b c2_noexact DAB_LIB_PATH=build/lib_noexact.so
b c2_notail DAB_LIB_PATH=build/lib_notail.so
b c2_cur_t1_0 DAB_V2_T1_BYTES=0
b c2_cur_t1_3072 DAB_V2_T1_BYTES=3072
BARGS="--workload c3_1Mx768_f16_ip --steps 10 --warmup 3"
b c3_head DAB_LIB_PATH=build/lib_head.so
b c3_cur A=1
b c3_cur_t1_0 DAB_V2_T1_BYTES=0
b c3_notail DAB_LIB_PATH=build/lib_notail.so
