#!/bin/bash
# round-2 GPU session 28 (2 GPUs): replication through the library's NCCL path, weak and strong scaling of C2 and the PQ workload
O=gpurun_out/s28; mkdir -p $O
run2() { # name, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 4 --no-cpu-baseline $2 > $O/$1.json 2> $O/$1.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/$1.json").read().strip().splitlines()[-1]); c = d["config"]
    print("$1", "ms/step %.3f" % d["ms_per_step"], "QPS %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"], "recall", c["recall_at_10"], "in flight", c["batches_in_flight"], c["setup_s"], (c["parity_gate"] or {}).get("result"))
except Exception as e:
    print("$1 failed", e); print(open("$O/$1.err").read()[-1500:])
PY
}
run2 c2_2gpu_weak "--scaling weak"
run2 c2_2gpu_strong "--scaling strong"
run2 pq_small_2gpu_weak "--scaling weak --workload small_200Kx128_i8_pq32"
