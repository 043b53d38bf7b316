"""One call (repeated twice) of every distance kernel family at a production-like size; run under
`ncu --metrics gpu__time_duration.sum --csv` to get per-kernel durations, then tools/zoo_table.py
turns launches + the sizes printed here into achieved GB/s.  Not a benchmark of record."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import diskann_b200 as dab

rng = np.random.default_rng(7)
sizes = {}
ONLY = set(sys.argv[1:])  # e.g. `kernel_zoo.py pq`: only that family (f32, pq, minmax, i8, f16); default: all


def want(family):
    return not ONLY or family in ONLY


def twice(f):
    f()
    return f()


# ---- f32 1M x 128
n, dim, nq, c = 1_000_000, 128, 10_000, 256
base = rng.standard_normal((n + 1, dim), dtype=np.float32)
q = rng.standard_normal((nq, dim), dtype=np.float32)
ids = rng.integers(0, n, (nq, c)).astype(np.uint32)
with dab.GpuIndex(dab.DType.f32, dab.Metric.L2, dim, n, 1, 83) as g:
    g.upload_vectors(base)
    if want("f32"):
        twice(lambda: g.distances(q, ids))
        sizes["frontier_float_kernel f32 L2 128-d"] = {"pairs": nq * c, "bytes_per_pair": dim * 4 + 8}
        a = rng.integers(0, n, 2_000_000).astype(np.uint32)
        b = rng.integers(0, n, 2_000_000).astype(np.uint32)
        twice(lambda: g.row_pair_distances(a, b))
        sizes["rowpair_float_kernel f32 L2 128-d"] = {"pairs": 2_000_000, "bytes_per_pair": 2 * dim * 4 + 12}
        twice(lambda: g.flat_knn(q[:1000], 10))
        sizes["flat_knn f32 L2 (1000 queries x 1M rows)"] = {"pairs": 1000 * n, "bytes_per_pair": 0, "flop_per_pair": 3 * dim}
    if want("pq"):
        piv = base[rng.choice(n, 256, replace=False)].copy()
        off = np.arange(0, dim + 1, 4, dtype=np.uint64)  # 32 chunks of 4 dims
        codes = rng.integers(0, 256, (n + 1, 32)).astype(np.uint8)
        g.upload_pq(piv, off, codes)
        twice(lambda: g.pq_populate_lut(q[:2000]))
        sizes["pq_fused_kernel, table written out: 32 chunks x 256 centres (2000 queries)"] = {"pairs": 2000, "bytes_per_pair": 32 * 256 * 4}
        twice(lambda: g.pq_distances(q[:2000], ids[:2000]))
        sizes["pq_fused_kernel, table + ADC: 32-byte codes, 256 candidates per query"] = {"pairs": 2000 * c, "bytes_per_pair": 32 + 8}
        ids_wide = rng.integers(0, n, (2000, 4096)).astype(np.uint32)
        twice(lambda: g.pq_distances(q[:2000], ids_wide))
        sizes["pq_fused_kernel, table + ADC: 32-byte codes, 4096 candidates per query"] = {"pairs": 2000 * 4096, "bytes_per_pair": 32 + 8}
        twice(lambda: g.pq_encode(base[:100_000]))
        sizes["pq_encode_kernel (100K vectors)"] = {"pairs": 100_000, "bytes_per_pair": dim * 4 + 32}
del base

# ---- i8 1M x 128
if want("minmax"):
    vmm = rng.uniform(-1.0, 1.0, (1_000_000, dim)).astype(np.float32)
    for nb in (8, 4):
        rows_mm, _ = twice(lambda: dab.minmax_compress(vmm, nb))
        sizes[f"minmax_compress_kernel {nb}-bit, 1M x 128 f32"] = {"pairs": 1_000_000, "bytes_per_pair": dim * 4 + int(rows_mm.shape[1])}
        other = np.ascontiguousarray(rows_mm[::-1])
        twice(lambda: dab.minmax_distances(dab.Metric.L2, nb, nb, dim, rows_mm, other))
        sizes[f"minmax_distance_kernel {nb} x {nb} bits, 1M pairs"] = {"pairs": 1_000_000, "bytes_per_pair": 2 * int(rows_mm.shape[1]) + 4}
    del vmm

if want("i8"):
    base8 = rng.integers(-128, 128, (n + 1, dim)).astype(np.int8)
    q8 = rng.integers(-128, 128, (nq, dim)).astype(np.int8)
    with dab.GpuIndex(dab.DType.i8, dab.Metric.L2, dim, n, 1, 83) as g:
        g.upload_vectors(base8)
        twice(lambda: g.distances(q8, ids))
        sizes["frontier_int_wide_kernel i8 L2 128-d"] = {"pairs": nq * c, "bytes_per_pair": dim + 8}
    del base8

# ---- f16 200K x 768 (BASELINE configs[2] shape)
if want("f16"):
    n3, d3 = 200_000, 768
    base16 = rng.standard_normal((n3 + 1, d3), dtype=np.float32).astype(np.float16)
    q16 = rng.standard_normal((2000, d3), dtype=np.float32).astype(np.float16)
    ids3 = rng.integers(0, n3, (2000, c)).astype(np.uint32)
    with dab.GpuIndex(dab.DType.f16, dab.Metric.InnerProduct, d3, n3, 1, 83) as g:
        g.upload_vectors(base16)
        twice(lambda: g.distances(q16, ids3))
        sizes["frontier_float_kernel f16 IP 768-d"] = {"pairs": 2000 * c, "bytes_per_pair": d3 * 2 + 8}
print(json.dumps(sizes))
