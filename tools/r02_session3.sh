#!/bin/bash
# round-2 GPU session 3: first run of search_kernel_v4 (one CTA per query)
mkdir -p gpurun_out/s3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/s3/gpu_tests.txt
b() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/s3/$name.err > gpurun_out/s3/$name.json; python - $name <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/s3/{sys.argv[1]}.json"))
    m = d["config"]["at_min_l"]
    print(sys.argv[1], "ms/step %.3f" % d["ms_per_step"], "e2e ms %.3f" % d["e2e"]["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "recall", d["config"]["recall_at_10"], "minL ms %.3f" % m["ms_per_step"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
b v4_default A=1
b v4_c6 DAB_V4_CTAS_PER_SM=6
b v4_c10 DAB_V4_CTAS_PER_SM=10
b v4_c12 DAB_V4_CTAS_PER_SM=12
b v3_only DAB_DISABLE_V4=1 DAB_V3_TABLE_BYTES=11776
for l in build/lib_v4_*.so; do [ -f $l ] && b $(basename $l .so) DAB_LIB_PATH=$l; done
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:search_kernel_v4 -c 1 -o gpurun_out/s3/prof_v4 python bench.py --steps 1 --warmup 3 --profile-range --no-cpu-baseline > gpurun_out/s3/ncu.log 2>&1; tail -2 gpurun_out/s3/ncu.log
timeout 400 compute-sanitizer --tool racecheck python tools/sanitize_check.py > gpurun_out/s3/racecheck.txt 2>&1; echo "racecheck rc=$?"; grep -c "Race reported" gpurun_out/s3/racecheck.txt; tail -2 gpurun_out/s3/racecheck.txt
