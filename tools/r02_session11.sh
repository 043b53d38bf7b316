#!/bin/bash
# round-2 GPU session 11: TC epilogue overlap; PQ kernel residency / table sizing on C4; C2 default line
mkdir -p gpurun_out/s11
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_core or pq" 2>&1 | tail -3
timeout 400 python tools/bench_flat.py 2>&1 | tail -12 | tee gpurun_out/s11/bench_flat.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flat_tc_kernel -s 1 -c 1 -o gpurun_out/s11/prof_flat_tc python tools/bench_flat.py 1000000 > gpurun_out/s11/ncu_tc.log 2>&1; tail -1 gpurun_out/s11/ncu_tc.log
for c in 6 3 2; do
DAB_PQ_CTAS_PER_SM=$c timeout 600 python bench.py --workload small_200Kx128_i8_pq32 --steps 5 --warmup 3 --no-cpu-baseline --l-search 350 > gpurun_out/s11/pq_small_$c.json 2> gpurun_out/s11/pq_small_$c.err; python -c "
import json; d=json.load(open('gpurun_out/s11/pq_small_$c.json')); print('pq_small ctas=$c ms/step %.3f' % d['ms_per_step'], 'recall', d['config']['recall_at_10'], (d['config']['parity_gate'] or {}).get('result'))" || tail -3 gpurun_out/s11/pq_small_$c.err
done
timeout 1500 python bench.py --workload c4_10Mx128_i8_pq32 --steps 5 --warmup 3 --no-cpu-baseline --l-search 500 > gpurun_out/s11/c4.json 2> gpurun_out/s11/c4.err; python -c "
import json; d=json.load(open('gpurun_out/s11/c4.json')); print('c4 ms/step %.3f' % d['ms_per_step'], 'QPS %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'], 'L', d['config']['l_search'], 'recall', d['config']['recall_at_10'], d['config']['setup_s'], (d['config']['parity_gate'] or {}).get('result'), 'frac %.3f' % d['roofline']['frac'])" || tail -5 gpurun_out/s11/c4.err
timeout 400 python bench.py --steps 20 --warmup 4 > gpurun_out/s11/c2_default.json 2> gpurun_out/s11/c2_default.err; python -c "
import json; d=json.load(open('gpurun_out/s11/c2_default.json')); print('c2 ms/step %.3f' % d['ms_per_step'], 'QPS %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'], 'frac %.3f' % d['roofline']['frac'], d['cpu_baseline'])" || tail -5 gpurun_out/s11/c2_default.err
