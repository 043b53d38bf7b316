"""debug aid: small device build + search (run under compute-sanitizer)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import diskann_b200 as dab
rng = np.random.default_rng(0)
n, d = 3000, 64
base = rng.normal(size=(n + 1, d)).astype(np.float32)
with dab.GpuIndex(dab.DType.f32, dab.Metric.L2, d, n, 1, 41) as g:
    g.upload_vectors(base)
    g.build(32, 64, 1.2)
    adj = g.download_graph()
    print("deg max", adj[:, 0].max(), "min", adj[:n, 0].min())
    print(g.search_batch(base[:5], 5, 64, 1)[0])
