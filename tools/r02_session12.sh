#!/bin/bash
# round-2 GPU session 12 (2 GPUs): replication through the library's NCCL path; weak and strong scaling; TC resident mode
mkdir -p gpurun_out/s12
timeout 300 python tools/bench_flat.py 2>&1 | tail -8 | tee gpurun_out/s12/bench_flat.txt
DAB_TC_STREAM=1 timeout 300 python tools/bench_flat.py 2>&1 | tail -8 | tee gpurun_out/s12/bench_flat_stream.txt
for mode in weak strong; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --scaling $mode > gpurun_out/s12/c2_2gpu_$mode.json 2> gpurun_out/s12/c2_2gpu_$mode.err; python -c "
import json; d=json.load(open('gpurun_out/s12/c2_2gpu_$mode.json')); print('2gpu $mode ms/step %.3f' % d['ms_per_step'], 'QPS %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'], 'recall', d['config']['recall_at_10'], d['config']['setup_s'], (d['config']['parity_gate'] or {}).get('result'))" || tail -8 gpurun_out/s12/c2_2gpu_$mode.err
done
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --scaling strong > gpurun_out/s12/c2_1gpu.json 2> gpurun_out/s12/c2_1gpu.err; python -c "
import json; d=json.load(open('gpurun_out/s12/c2_1gpu.json')); print('1gpu ms/step %.3f' % d['ms_per_step'], 'QPS %.0f' % d['value'])"
