#!/bin/bash
# round-2 GPU session 9: full GPU suite; tcgen05 flat scan with the resident query tile; v3 prefetch A/B
mkdir -p gpurun_out/s9
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/s9/gpu_tests.txt
timeout 400 python tools/bench_flat.py 2>&1 | tail -12 | tee gpurun_out/s9/bench_flat.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flat_tc_kernel -s 1 -c 1 -o gpurun_out/s9/prof_flat_tc python tools/bench_flat.py 1000000 > gpurun_out/s9/ncu_tc.log 2>&1; tail -2 gpurun_out/s9/ncu_tc.log
b() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/s9/$name.err > gpurun_out/s9/$name.json; python - $name <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/s9/{sys.argv[1]}.json"))
    m = d["config"]["at_min_l"]
    print(sys.argv[1], "ms/step %.3f" % d["ms_per_step"], "e2e ms %.3f" % d["e2e"]["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "recall", d["config"]["recall_at_10"], "minL ms %.3f" % m["ms_per_step"], "parity", (d["config"]["parity_gate"] or {}).get("result"), d["config"]["setup_s"])
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(f"gpurun_out/s9/{sys.argv[1]}.err").read()[-600:])
PY
}
b v3_default A=1
for l in build/lib_v3_*.so; do [ -f $l ] && b $(basename $l .so) DAB_LIB_PATH=$l; done
b v2_only DAB_DISABLE_V3=1
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_check.py > gpurun_out/s9/memcheck.txt 2>&1; tail -2 gpurun_out/s9/memcheck.txt
