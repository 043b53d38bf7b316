"""tuning aid: search_kernel_v3 vs search_kernel_v2 over the list size L on the C2 (or C3) shape — where is the crossover?
usage: python tools/sweep_l.py [c2_1Mx128_f32_l2|c3_1Mx768_f16_ip]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import diskann_b200 as dab

wl = sys.argv[1] if len(sys.argv) > 1 else "c2_1Mx128_f32_l2"
cfg = bench.WORKLOADS[wl]
n, dim, md = cfg["n"], cfg["dim"], bench.max_degree(cfg["R"])
centers = bench.make_centers(cfg)
base = bench.make_data(cfg, bench.SEED_BASE, n, centers)
medoid = bench.find_medoid(base)
batches = [bench.make_data(cfg, bench.SEED_QUERY + 97 * b, cfg["nq"], centers) for b in range(4)]
dt, mt = bench.dab_enums(dab, cfg)
nq, K = cfg["nq"], 10


def make(env):
    for k in ("DAB_DISABLE_V3", "DAB_DISABLE_V2", "DAB_V3_MAX_CAP"):
        os.environ.pop(k, None)
    os.environ.update(env)
    g = dab.GpuIndex(dt, mt, dim, n, 1, md)
    g.upload_vectors(base)
    g.upload_vectors(medoid[None, :], first=n)
    return g


g3 = make({"DAB_V3_MAX_CAP": "512"})
g3.build(cfg["R"], cfg["l_build"], bench.ALPHA)
adj = g3.download_graph()
g2 = make({"DAB_DISABLE_V3": "1"})
g2.upload_graph(adj)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)  # events below are recorded on the stream the kernels run on
g3.set_stream(stream.cuda_stream)
g2.set_stream(stream.cuda_stream)
d_q = [torch.from_numpy(q).cuda() for q in batches]
d_ids = torch.empty((nq, K), dtype=torch.int32, device="cuda")
d_d = torch.empty((nq, K), dtype=torch.float32, device="cuda")
for L in (10, 15, 20, 30, 40, 50, 60, 70, 80, 100, 120, 150, 200):
    row = [f"L={L:4d}"]
    for name, g in (("v3", g3), ("v2", g2)):
        for i in range(4):
            g.search_batch_device(d_q[i % 4].data_ptr(), nq, K, L, 1, d_ids.data_ptr(), d_d.data_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(8):
            g.search_batch_device(d_q[i % 4].data_ptr(), nq, K, L, 1, d_ids.data_ptr(), d_d.data_ptr())
        e1.record(stream)
        torch.cuda.synchronize()
        row.append(f"{name} {e0.elapsed_time(e1) / 8:7.3f} ms")
    print("  ".join(row), flush=True)
