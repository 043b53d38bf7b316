#!/bin/bash
# round-2 GPU session 8: tcgen05 flat scan timing + ncu; v3 prefetch A/B; batched build parity; PQ recall sweep; C4
mkdir -p gpurun_out/s8
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "build or tensor_core or pq" 2>&1 | tail -5 | tee gpurun_out/s8/tests_subset.txt
timeout 400 python tools/bench_flat.py 2>&1 | tail -12 | tee gpurun_out/s8/bench_flat.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flat_tc_kernel -s 1 -c 1 -o gpurun_out/s8/prof_flat_tc python tools/bench_flat.py 1000000 > gpurun_out/s8/ncu_tc.log 2>&1; tail -2 gpurun_out/s8/ncu_tc.log
b() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/s8/$name.err > gpurun_out/s8/$name.json; python - $name <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/s8/{sys.argv[1]}.json"))
    m = d["config"]["at_min_l"]
    print(sys.argv[1], "ms/step %.3f" % d["ms_per_step"], "e2e ms %.3f" % d["e2e"]["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "recall", d["config"]["recall_at_10"], "minL ms %.3f" % m["ms_per_step"], "parity", (d["config"]["parity_gate"] or {}).get("result"))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(f"gpurun_out/s8/{sys.argv[1]}.err").read()[-600:])
PY
}
for l in build/lib_v3_*.so; do [ -f $l ] && b $(basename $l .so) DAB_LIB_PATH=$l; done
timeout 600 python bench.py --workload small_200Kx128_i8_pq32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s8/pq_small.json 2> gpurun_out/s8/pq_small.err; python -c "
import json; d=json.load(open('gpurun_out/s8/pq_small.json')); print('pq_small ms/step %.3f' % d['ms_per_step'], 'L', d['config']['l_search'], 'recall', d['config']['recall_at_10'], [(x['l'], x['recall']) for x in d['config']['l_sweep'][-6:]], d['config']['parity_gate'])" || tail -3 gpurun_out/s8/pq_small.err
timeout 1500 python bench.py --workload c4_10Mx128_i8_pq32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s8/c4.json 2> gpurun_out/s8/c4.err; python -c "
import json; d=json.load(open('gpurun_out/s8/c4.json')); print('c4 ms/step %.3f' % d['ms_per_step'], 'QPS %.0f' % d['value'], 'L', d['config']['l_search'], 'recall', d['config']['recall_at_10'], d['config']['setup_s'], d['config']['parity_gate'], 'frac %.3f' % d['roofline']['frac'])" || tail -5 gpurun_out/s8/c4.err
