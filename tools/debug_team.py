"""debug aid: one tiny search through the team kernel (run under compute-sanitizer)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import diskann_b200 as dab
from test_oracle_golden import grid
data, adj, n = grid(1, 100)
with dab.GpuIndex(dab.DType.f32, dab.Metric.L2, 1, n, 1, adj.shape[1] - 1) as g:
    g.upload_vectors(data)
    g.upload_graph(adj)
    print(g.search_batch(np.array([[-1.0]], np.float32), 10, 10, 1))
