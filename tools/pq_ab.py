"""tuning aid: A/B of the PQ traversal kernel variants on ONE resident index (dab_reload_tuning re-reads the DAB_* knobs).
usage: python tools/pq_ab.py [workload] [L ...]     e.g.  python tools/pq_ab.py c4_10Mx128_i8_pq32 500 800"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench as B
import diskann_b200 as dab

wl = sys.argv[1] if len(sys.argv) > 1 else "small_200Kx128_i8_pq32"
Ls = [int(a) for a in sys.argv[2:]] or [350]
cfg = dict(B.WORKLOADS[wl])
n, dim, md = cfg["n"], cfg["dim"], B.max_degree(cfg["R"])
centers = B.make_centers(cfg)
base = B.make_data(cfg, B.SEED_BASE, n, centers)
medoid = B.find_medoid(base)
dt, mt = B.dab_enums(dab, cfg)
g = dab.GpuIndex(dt, mt, dim, n, 1, md)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
g.set_stream(stream.cuda_stream)
t0 = time.time()
B.prepare_index(g, cfg, base, medoid, 0, 1, None, torch, lambda *a: None)
print(f"{wl}: index ready in {time.time() - t0:.1f}s", flush=True)
nq = cfg["nq"]
qs = [torch.from_numpy(B.make_data(cfg, B.SEED_QUERY + 97 * b, nq, centers)).cuda() for b in range(4)]
ids = torch.empty((nq, 10), dtype=torch.int32, device="cuda")
dists = torch.empty((nq, 10), dtype=torch.float32, device="cuda")
cmps = torch.empty(nq, dtype=torch.int32, device="cuda")

VARIANTS = [
    ("default", {}),
    ("no row copied ahead", {"DAB_PQ_NO_SPEC": "1"}),
    ("no code prefetch", {"DAB_PQ_NO_CODE_PREFETCH": "1"}),
    ("neither", {"DAB_PQ_NO_SPEC": "1", "DAB_PQ_NO_CODE_PREFETCH": "1"}),
    ("12 warps per SM", {"DAB_PQ_WARPS": "12"}),
    ("8 warps per SM", {"DAB_PQ_WARPS": "8"}),
    ("table in global memory (search_kernel_pq)", {"DAB_PQ_GLOBAL_LUT": "1"}),
]
KNOBS = sorted({k for _, env in VARIANTS for k in env})


def run(L, rerank):
    def go(i):
        g.search_batch_pq_device(qs[i % 4].data_ptr(), nq, 10, L, 1, ids.data_ptr(), dists.data_ptr(), 0, cmps.data_ptr(), 0, rerank=rerank)
    for i in range(4):
        go(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(12):
        go(i)
    e1.record(stream)
    torch.cuda.synchronize()
    go(0)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 12, ids.cpu().numpy().copy(), cmps.cpu().numpy().copy()


for L in Ls:
    ref = None
    for name, env in VARIANTS:
        for k in KNOBS:
            os.environ.pop(k, None)
        os.environ.update(env)
        g.reload_tuning()
        ms, i_, c_ = run(L, True)
        if ref is None:
            ref = (i_, c_)
        same = np.array_equal(ref[0], i_) and np.array_equal(ref[1], c_)
        print(f"L={L:4d} {name:45s} {ms:8.3f} ms per 10K-query step (traversal + rerank)  {nq / ms / 1e3:7.3f} M QPS  "
              f"mean cmps {c_.mean():8.1f}  {'same results' if same else 'RESULTS DIFFER'}", flush=True)
for k in KNOBS:
    os.environ.pop(k, None)
