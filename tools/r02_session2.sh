#!/bin/bash
# round-2 GPU session 2: first run of search_kernel_v3 (shared-memory visited sets)
mkdir -p gpurun_out/s2
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/s2/gpu_tests.txt
b() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/s2/$name.err > gpurun_out/s2/$name.json; python - $name <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/s2/{sys.argv[1]}.json"))
    print(sys.argv[1], "ms/step %.3f" % d["ms_per_step"], "e2e ms %.3f" % d["e2e"]["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "recall", d["config"]["recall_at_10"], "minL", d["config"]["at_min_l"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
b v3_default A=1
b v2_only DAB_DISABLE_V3=1
b v3_ctas3 DAB_V3_CTAS_PER_SM=3
b v3_ctas2 DAB_V3_CTAS_PER_SM=2
b v3_t8k DAB_V3_TABLE_BYTES=8192
b v3_t12k DAB_V3_TABLE_BYTES=12288
for l in build/lib_v3_*.so; do b $(basename $l .so) DAB_LIB_PATH=$l; done
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:search_kernel_v3 -c 1 -o gpurun_out/s2/prof_v3 python bench.py --steps 1 --warmup 3 --profile-range --no-cpu-baseline > gpurun_out/s2/ncu.log 2>&1; tail -2 gpurun_out/s2/ncu.log
