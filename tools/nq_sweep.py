"""tuning aid: ms per launch of the search kernel vs. queries per launch (how much of a 10K-query step is tail),
and batches in flight 1..4.  usage: python tools/nq_sweep.py [workload] [L]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench as B
import diskann_b200 as dab

wl = sys.argv[1] if len(sys.argv) > 1 else "c2_1Mx128_f32_l2"
cfg = dict(B.WORKLOADS[wl])
L = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["l_search"]
n, dim, md = cfg["n"], cfg["dim"], B.max_degree(cfg["R"])
centers = B.make_centers(cfg)
base = B.make_data(cfg, B.SEED_BASE, n, centers)
medoid = B.find_medoid(base)
dt, mt = B.dab_enums(dab, cfg)
g = dab.GpuIndex(dt, mt, dim, n, 1, md)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
g.set_stream(stream.cuda_stream)
g.upload_vectors(base)
g.upload_vectors(medoid[None, :], first=n)
g.build(cfg["R"], cfg["l_build"], B.ALPHA)
NQMAX = 40000
qs = [torch.from_numpy(B.make_data(cfg, B.SEED_QUERY + 97 * b, NQMAX, centers)).cuda() for b in range(4)]
out = [dict(ids=torch.empty((NQMAX, 10), dtype=torch.int32, device="cuda"), dists=torch.empty((NQMAX, 10), dtype=torch.float32, device="cuda"))
       for _ in range(4)]


def timed(nq, steps, slots):
    def go(i):
        s, b = i % slots, i % 4
        if slots == 1:
            g.search_batch_device(qs[b].data_ptr(), nq, 10, L, 1, out[0]["ids"].data_ptr(), out[0]["dists"].data_ptr())
        else:
            g.wait(s)
            g.search_batch_device_async(s, qs[b].data_ptr(), nq, 10, L, 1, out[s]["ids"].data_ptr(), out[s]["dists"].data_ptr())
    for i in range(4):
        go(i)
    for s in range(slots):
        g.wait(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        go(i)
    for s in range(slots):
        g.wait(s)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


print(f"{wl} L={L}")
for nq in (1667, 3334, 5000, 6668, 10000, 13336, 20000, 40000):
    ms = timed(nq, 12, 1)
    print(f"nq={nq:6d} in flight 1: {ms:7.3f} ms/launch  {ms / nq * 1e4:6.3f} ms per 10K queries", flush=True)
for slots in (1, 2, 3, 4):
    ms = timed(10000, 24, slots)
    print(f"nq= 10000 in flight {slots}: {ms:7.3f} ms/step   {1e4 / ms / 1e3:6.3f} M QPS", flush=True)
# what one rank of a strong-scaled 10K-query batch runs at N = 1, 2, 4, 8 GPUs (no collective on the path, so the
# per-rank step time IS the job's step time): efficiency(N) = t(10000) / (N * t(10000 / N))
t1 = {s_: timed(10000, 24, s_) for s_ in (1, 2, 4)}
for ngpu in (1, 2, 4, 8):
    nq = 10000 // ngpu
    for slots in (1, 2, 4):
        ms = timed(nq, 24, slots)
        print(f"strong-scaling share N={ngpu}: nq={nq:6d} in flight {slots}: {ms:7.3f} ms/step  job {1e4 / ms / 1e3:6.3f} M QPS  "
              f"efficiency {t1[slots] / (ngpu * ms):5.3f}", flush=True)
