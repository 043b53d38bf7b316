"""CPU checks of the oracle's scalar-quantized store and traversal (test infrastructure for the GPU parity tests):
the canonical-front row layout of the reference (diskann-quantization/src/meta/vector.rs:478-507, bits/slice.rs:261-323)
and the search through compressed rows (diskann-providers/src/model/graph/provider/async_/inmem/scalar.rs:449-570)."""
import os
import struct
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import oracle_lib as O


@pytest.mark.parametrize("nbits", [1, 2, 4, 8])
def test_rows_are_canonical_front_with_dense_codes(nbits):
    dim = 13
    maxv = (1 << nbits) - 1
    shift = np.linspace(-1.0, 0.5, dim).astype(np.float32)
    scale = 2.0
    # vectors that land exactly on code values: x = shift + code * scale / maxv
    codes = (np.arange(dim) * 5 + 3) % (maxv + 1)
    vec = (shift.astype(np.float64) + codes * (scale / maxv)).astype(np.float32)
    rows = O.sq_encode_rows(vec[None], shift, scale, nbits)
    assert rows.shape == (1, 4 + (dim * nbits + 7) // 8)
    got = [(int(rows[0, 4 + (i * nbits) // 8]) >> ((i * nbits) % 8)) & maxv for i in range(dim)]
    assert got == [int(c) for c in codes]
    # the unused high bits of the last byte stay zero
    used = dim * nbits % 8
    if used:
        assert int(rows[0, -1]) >> used == 0
    # compensation first: scale / maxv * sum(code * shift), accumulated as a sequential f32 FMA chain
    comp = struct.unpack("<f", rows[0, :4].tobytes())[0]
    want = np.float32(scale) * (np.float32(1.0) / np.float32(maxv))
    dot = np.float32(0)
    for c, s in zip(codes, shift):
        dot = np.float32(np.float64(np.float32(c)) * np.float64(s) + np.float64(dot))  # fma: one rounding
    assert comp == np.float32(want * dot)


def _index(metric, nbits, seed=0, n=3000, d=32):
    rng = np.random.default_rng(seed)
    centers = rng.standard_normal((16, d)).astype(np.float32)
    base = (centers[rng.integers(0, 16, n)] + 0.3 * rng.standard_normal((n, d))).astype(np.float32)
    if metric == O.COSINE_NORMALIZED:
        base /= np.linalg.norm(base, axis=1, keepdims=True)
    medoid = base[np.argmin(((base - base.mean(0)) ** 2).sum(1))]
    vecs = np.concatenate([base, medoid[None]])
    adj = O.build_graph(vecs, n, 1, O.L2 if metric == O.COSINE_NORMALIZED else metric, 16, 20, 30)
    std = float(vecs.std())
    shift = (vecs.mean(0) - 3 * std).astype(np.float32)
    scale = 6 * std
    ssn = float(-O.distance(shift, shift, O.INNER_PRODUCT))
    mean_norm = float(np.linalg.norm(vecs, axis=1).mean()) if metric == O.INNER_PRODUCT else 0.0
    rows = O.sq_encode_rows(vecs, shift, scale, nbits)
    return vecs, adj, n, (rows, nbits, shift, scale, ssn, mean_norm)


@pytest.mark.parametrize("metric", [O.L2, O.INNER_PRODUCT, O.COSINE_NORMALIZED])
def test_eight_bit_traversal_tracks_the_full_precision_search(metric):
    vecs, adj, n, sq = _index(metric, 8)
    rng = np.random.default_rng(1)
    queries = vecs[rng.integers(0, n, 100)] + 0.05 * rng.standard_normal((100, vecs.shape[1])).astype(np.float32)
    queries = queries.astype(np.float32)
    full = O.Index(vecs, adj, n, 1, metric)
    quant = O.Index(vecs, adj, n, 1, metric, sq=sq)
    want = full.search_batch(queries, 10, 60)[0]
    got = quant.search_batch(queries, 10, 60)[0]
    rer = quant.search_batch_rerank(queries, 10, 60)
    overlap = np.mean([len(set(a) & set(b)) / 10 for a, b in zip(got, want)])
    overlap_rerank = np.mean([len(set(a) & set(b)) / 10 for a, b in zip(rer[0], want)])
    assert overlap > 0.85, overlap
    assert overlap_rerank >= overlap and overlap_rerank > 0.95, (overlap, overlap_rerank)
    # Rerank reports full-precision distances in ascending order and never a start point
    assert (np.diff(rer[1], axis=1) >= 0).all()
    assert (rer[0] < n).all()
    d0 = O.distance(queries[0], vecs[rer[0][0, 0]], metric)
    assert d0 == rer[1][0, 0]


def test_one_bit_l2_distance_is_a_scaled_hamming_distance():
    vecs, adj, n, sq = _index(O.L2, 1, seed=3, n=500, d=40)
    rows, nbits, shift, scale, ssn, mean_norm = sq
    quant = O.Index(vecs, adj, n, 1, O.L2, sq=sq)
    q = vecs[7:8].copy()
    ids, dists, counts, _, _ = quant.search_batch(q, 5, 20)
    qrow = rows[7, 4:]
    for i, dist in zip(ids[0][:counts[0]], dists[0]):
        ham = int(np.unpackbits(qrow ^ rows[i, 4:]).sum())
        assert dist == np.float32(np.float32(scale) * np.float32(scale)) * np.float32(ham)
