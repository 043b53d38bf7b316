"""micro-benchmark of kernel (1), the batched frontier distance gather (dab_distances_device):
10K queries x C random candidate rows of a 1M x 128 f32 index.  Reports algorithmic GB/s
(dim*4 + 8 bytes per (query, candidate)) against the measured HBM peak."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import diskann_b200 as dab

n, dim, nq = 1_000_000, 128, 10_000
cands = [int(a) for a in sys.argv[1:]] or [64, 1205]
rng = np.random.default_rng(1)
base = rng.standard_normal((n + 1, dim), dtype=np.float32)
g = dab.GpuIndex(dab.DType.f32, dab.Metric.L2, dim, n, 1, 83)
g.upload_vectors(base)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
g.set_stream(stream.cuda_stream)
q = torch.from_numpy(rng.standard_normal((nq, dim), dtype=np.float32)).cuda()
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
lib = dab.lib()
for c in cands:
    ids = torch.from_numpy(rng.integers(0, n, (nq, c)).astype(np.int32)).cuda()
    out = torch.empty((nq, c), dtype=torch.float32, device="cuda")
    def step():
        dab._lib.check(lib.dab_distances_device(g._h, C.c_void_p(q.data_ptr()), nq, C.c_void_p(ids.data_ptr()), c, C.c_void_p(out.data_ptr())))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = nq * c * (dim * 4 + 8) / 1e9
    print(json.dumps({"kernel": "frontier_float_kernel<f32,L2>", "queries": nq, "candidates_per_query": c, "ms": ms,
                      "algorithmic_GB": gb, "achieved_GBps": gb / (ms / 1e3), "frac_of_measured_hbm_peak": gb / (ms / 1e3) / peak}))
