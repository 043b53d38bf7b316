#!/bin/bash
# round-2 GPU session 6: tcgen05 flat scan first run; v3 occupancy variants; C3 A/B
mkdir -p gpurun_out/s6
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k tensor_core 2>&1 | tail -15 | tee gpurun_out/s6/tc_tests.txt
timeout 300 python tools/bench_flat.py 2>&1 | tail -12 | tee gpurun_out/s6/bench_flat.txt
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_tensor_core_flat_scan_equals_the_exact_scan 2>&1 | tail -5 | tee gpurun_out/s6/gpu_tests.txt
b() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/s6/$name.err > gpurun_out/s6/$name.json; python - $name <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/s6/{sys.argv[1]}.json"))
    m = d["config"]["at_min_l"]
    print(sys.argv[1], "ms/step %.3f" % d["ms_per_step"], "e2e ms %.3f" % d["e2e"]["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "recall", d["config"]["recall_at_10"], "minL ms %.3f" % m["ms_per_step"], "parity", (d["config"]["parity_gate"] or {}).get("result"))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(f"gpurun_out/s6/{sys.argv[1]}.err").read()[-600:])
PY
}
b v3_default A=1
for l in build/lib_v3_*.so; do [ -f $l ] && b $(basename $l .so) DAB_LIB_PATH=$l; done
b v2_only DAB_DISABLE_V3=1
c3() { name=$1; shift; env "$@" timeout 600 python bench.py --workload c3_1Mx768_f16_ip --steps 8 --warmup 3 --no-cpu-baseline --l-search 100 2>gpurun_out/s6/$name.err > gpurun_out/s6/$name.json; python -c "
import json; d=json.load(open('gpurun_out/s6/$name.json')); print('$name', 'ms/step %.3f' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'], 'recall', d['config']['recall_at_10'], d['config']['parity_gate']['result'] if d['config']['parity_gate'] else None)" || tail -3 gpurun_out/s6/$name.err; }
c3 c3_v3 A=1
c3 c3_v2 DAB_DISABLE_V3=1
