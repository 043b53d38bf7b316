"""The oracle's MinMax quantizer (oracle/minmax.cpp) against the reference's own tests for it, restated with the
reference's tolerances (diskann-quantization/src/minmax/quantizer.rs:473-756, vectors.rs:520-700).  The reference draws
its inputs from Rust's StdRng; the same properties are checked here on numpy draws of the same distributions."""
import numpy as np
import pytest

import oracle_lib as O

NBITS = [1, 2, 4, 8]


def meta_of(row):
    dim = int(row[:4].view(np.uint32)[0])
    b, n, a, norm_squared = (float(x) for x in row[4:20].view(np.float32))
    return dim, b, n, a, norm_squared


def codes_of(row, nbits, dim):
    bits = np.unpackbits(row[20:], bitorder="little")[:dim * nbits].reshape(dim, nbits)
    return (bits * (1 << np.arange(nbits))).sum(1)


# quantizer.rs:586-613: (nbits, relative error per grid scale 1.0 / 1.1 / 0.9)
@pytest.mark.parametrize("nbits,errs", [(1, [0.5, 0.5, 0.5]), (2, [0.5, 0.5, 0.5]), (4, [1e-2, 1e-2, 3e-2]), (8, [2e-3, 2e-3, 7e-3])])
def test_encoding_random(nbits, errs):
    """test_quantizer_encoding_random (quantizer.rs:473-561)."""
    rng = np.random.default_rng(nbits)
    L = O.lib()
    for scale, err in zip([1.0, 1.1, 0.9], errs):
        for dim in range(10, 100, 3):
            v = rng.uniform(-1.0, 1.0, (4, dim)).astype(np.float32)
            rows, loss, nan = O.minmax_compress(v, nbits, scale)
            assert not nan.any()
            rec = O.minmax_decompress(rows, nbits, dim)
            for r in range(v.shape[0]):
                d, b, n, a, ns = meta_of(rows[r])
                assert d == dim
                codes = codes_of(rows[r], nbits, dim)
                assert np.array_equal(rec[r], (codes.astype(np.float32) * np.float32(a) + np.float32(b)).astype(np.float32))
                rerr = float(((v[r] - rec[r]) ** 2).sum(dtype=np.float32))
                norm = float((v[r] * v[r]).sum(dtype=np.float32))
                assert rerr / norm <= err, (nbits, scale, dim)
                assert loss[r] - rerr <= 1e-4
                assert abs(n / a - float(codes.sum())) <= 2e-5 * dim
                assert abs(ns - float((rec[r] * rec[r]).sum(dtype=np.float32))) <= 1e-3
            # FullQuery (quantizer.rs:525-560): the vector stays f32, meta = {sum, norm_squared}
            s, ns = np.zeros(1, np.float32), np.zeros(1, np.float32)
            assert L.orc_minmax_full_query_meta(O.ptr(v[0]), dim, O.ptr(s), O.ptr(ns)) == 0
            assert abs(float(ns[0]) - float((v[0] * v[0]).sum(dtype=np.float64))) < 1e-4
            assert abs(float(s[0]) - float(v[0].sum(dtype=np.float64))) < 1e-4


@pytest.mark.parametrize("nbits", NBITS)
def test_all_same_value_vector(nbits):
    """quantizer.rs:632-667: min == max != 0 compresses with (almost) no loss and reconstructs the constant."""
    v = np.full((1, 30), 42.5, np.float32)
    rows, loss, nan = O.minmax_compress(v, nbits, 1.0)
    assert not nan[0] and abs(loss[0]) <= 1e-6
    assert (np.abs(O.minmax_decompress(rows, nbits, 30) - 42.5) < 1e-3).all()


@pytest.mark.parametrize("nbits", NBITS)
def test_two_distinct_values(nbits):
    """quantizer.rs:670-724 as written there (`skip(dim)` leaves every element at the first value)."""
    v = np.full((1, 20), -10.0, np.float32)
    rows, loss, nan = O.minmax_compress(v, nbits, 1.0)
    assert not nan[0] and abs(loss[0]) <= 1e-6
    if nbits > 1:
        assert len(set(codes_of(rows[0], nbits, 20).tolist())) <= 2
    assert (np.abs(O.minmax_decompress(rows, nbits, 20) - v) < 1e-4).all()
    # and the case the comment of that test describes: two values, half each -> the two ends of the code range
    w = np.array([[-10.0] * 10 + [15.0] * 10], np.float32)
    rows, loss, nan = O.minmax_compress(w, nbits, 1.0)
    assert set(codes_of(rows[0], nbits, 20).tolist()) == {0, (1 << nbits) - 1}
    assert (np.abs(O.minmax_decompress(rows, nbits, 20) - w) < 1e-4).all() and abs(loss[0]) <= 1e-6


@pytest.mark.parametrize("nbits", NBITS)
def test_nan_input_is_an_error_but_the_meta_is_written(nbits):
    """quantizer.rs:728-750."""
    v = np.ones((1, 100), np.float32)
    v[0, 33] = np.nan
    rows, loss, nan = O.minmax_compress(v, nbits, 1.0)
    assert nan[0] and meta_of(rows[0])[0] == 100
    s, ns = np.zeros(1, np.float32), np.zeros(1, np.float32)
    assert O.lib().orc_minmax_full_query_meta(O.ptr(v[0]), 100, O.ptr(s), O.ptr(ns)) == 1


@pytest.mark.parametrize("nbits", NBITS)
def test_compensated_distances(nbits):
    """test_minmax_compensated_vectors (vectors.rs:520-640): random codes with random (a, b) against the f32 arithmetic
    on the reconstructed vectors, the reference's tolerances; plus the heterogeneous 8 x N pairing (vectors.rs:778-797)."""
    rng = np.random.default_rng(100 + nbits)
    for dim in list(range(1, 40)) + [64, 100, 128, 257]:
        def random_row(nb):
            codes = rng.integers(0, 1 << nb, dim)
            a, b = np.float32(rng.uniform(0.0, 2.0)), np.float32(rng.uniform(0.0, 2.0))
            orig = (a * codes.astype(np.float32) + b).astype(np.float32)
            row = np.zeros(20 + (dim * nb + 7) // 8, np.uint8)
            bits = ((codes[:, None] >> np.arange(nb)) & 1).astype(np.uint8).reshape(-1)
            row[20:] = np.packbits(np.pad(bits, (0, (-len(bits)) % 8)), bitorder="little")
            row[:4] = np.array([dim], np.uint32).view(np.uint8)
            meta = np.array([b, a * np.float32(codes.sum(dtype=np.float32)), a, np.float32((orig * orig).sum(dtype=np.float32))], np.float32)
            row[4:20] = meta.view(np.uint8)
            return row, orig
        for nbx in (nbits, 8):
            x, ox = random_row(nbx)
            y, oy = random_row(nbits)
            ip = float((ox * oy).sum(dtype=np.float32))
            l2 = float(((ox - oy) ** 2).sum(dtype=np.float32))
            nx, ny = float((ox * ox).sum(dtype=np.float32)), float((oy * oy).sum(dtype=np.float32))
            got = {m: float(O.lib().orc_minmax_distance(m, nbx, nbits, O.ptr(x), O.ptr(y))) for m in (O.L2, O.INNER_PRODUCT, O.COSINE, O.COSINE_NORMALIZED)}
            assert abs(ip - (-got[O.INNER_PRODUCT])) / abs(ip) < 1e-3
            assert l2 == 0 or abs(got[O.L2] - l2) / l2 < 1e-3 or abs(got[O.L2] - l2) < 1e-3 * max(nx, ny)
            cos = 1.0 - ip / (np.sqrt(nx) * np.sqrt(ny))
            assert abs(got[O.COSINE] - cos) < 1e-6 or abs(got[O.COSINE] - cos) / cos < 1e-3
            assert abs((1.0 - ip) - got[O.COSINE_NORMALIZED]) / abs(1.0 - ip) < 1e-5


def test_one_bit_range_uses_the_side_means():
    """quantizer.rs:119-136: for 1 bit the range is (mean of the values below the mean, mean of the others)."""
    v = np.array([[0.0, 0.0, 0.0, 1.0, 1.0, 10.0]], np.float32)  # mean 2.0: below {0,0,0,1,1} -> 0.4, above {10} -> 10
    rows, _, _ = O.minmax_compress(v, 1, 1.0)
    _, b, n, a, _ = meta_of(rows[0])
    assert abs(b - 0.4) < 1e-6 and abs(a - 9.6) < 1e-5
    assert codes_of(rows[0], 1, 6).tolist() == [0, 0, 0, 0, 0, 1] and abs(n - a * 1.0) < 1e-6


@pytest.mark.parametrize("nbits", NBITS)
def test_full_query_distances(nbits):
    """FullQuery x Data (vectors.rs:596-650 of the reference's test): the query stays f32; against the f32 arithmetic on the
    reconstructed data vector with the reference's tolerances, and against a float64 evaluation of the same formula (the
    lane order of the f32 x N-bit inner product must only change rounding) for every length 1..150."""
    rng = np.random.default_rng(200 + nbits)
    L = O.lib()
    for dim in list(range(1, 151)) + [256, 384, 1000]:
        v = rng.uniform(-1.0, 1.0, (2, dim)).astype(np.float32)
        rows, _, _ = O.minmax_compress(v[1:], nbits, 1.0)
        q = v[0]
        y = O.minmax_decompress(rows, nbits, dim)[0]
        got = O.minmax_query_distances(O.INNER_PRODUCT, nbits, q[None], rows)[0, 0]
        codes = codes_of(rows[0], nbits, dim).astype(np.float64)
        _, b, _, a, ns = meta_of(rows[0])
        exact_ip = float((q.astype(np.float64) * codes).sum()) * a + float(q.astype(np.float64).sum()) * b
        assert abs(-got - exact_ip) <= 1e-5 * max(1.0, float(np.abs(q).sum()) * max(abs(a) * ((1 << nbits) - 1), abs(b))), (dim, nbits)
        ip = float((q * y).sum(dtype=np.float32))
        l2 = float(((q - y) ** 2).sum(dtype=np.float32))
        d = {m: float(O.minmax_query_distances(m, nbits, q[None], rows)[0, 0]) for m in (O.L2, O.COSINE, O.COSINE_NORMALIZED)}
        assert abs(d[O.L2] - l2) <= 1e-3 * max(l2, 1e-3) + 1e-4
        nq, ny = float((q * q).sum(dtype=np.float32)), float((y * y).sum(dtype=np.float32))
        if nq > 0 and ny > 0:
            cos = 1.0 - ip / (np.sqrt(nq) * np.sqrt(ny))
            assert abs(d[O.COSINE] - cos) < 1e-4
        assert abs(d[O.COSINE_NORMALIZED] - (1.0 - ip)) < 1e-4 * max(1.0, abs(1.0 - ip))
