#!/bin/bash
# round-2 GPU session 14: asynchronous batches — parity, tail share of a step, batches in flight; per-phase cycles of v2
O=gpurun_out/s14; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q -k "in_flight or overflow or grid_search" 2>&1 | tail -4 | tee $O/gpu_tests_async.txt
timeout 400 python tools/nq_sweep.py 2>&1 | tail -14 | tee $O/nq_sweep_c2.txt
DAB_PHASE_PROFILE=1 DAB_LIB_PATH=build/lib_phase.so timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --in-flight 1 2>&1 >/dev/null | grep "phase profile" | tail -3 | tee $O/phase_c2.txt
timeout 600 python bench.py --no-cpu-baseline > $O/bench_c2_inflight2.json 2> $O/bench_c2_inflight2.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/s14/bench_c2_inflight2.json"))
print("c2 in flight", d["config"]["batches_in_flight"], "ms/step %.3f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "e2e ms %.3f" % d["e2e"]["ms_per_step"], "serial", d["config"]["serial"], "parity", (d["config"]["parity_gate"] or {}).get("result"))
PY
tail -3 $O/bench_c2_inflight2.err
