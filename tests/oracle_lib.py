"""ctypes loader for the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")

F32, F16, I8, U8 = 0, 1, 2, 3
COSINE, INNER_PRODUCT, L2, COSINE_NORMALIZED = 0, 1, 2, 3
SIMD, SCALAR, AVX2 = 0, 1, 2

NP_DTYPES = {F32: np.float32, F16: np.float16, I8: np.int8, U8: np.uint8}


def dtype_code(arr):
    return {np.dtype(np.float32): F32, np.dtype(np.float16): F16, np.dtype(np.int8): I8,
            np.dtype(np.uint8): U8}[arr.dtype]


class OrcIndex(C.Structure):
    _fields_ = [
        ("dtype", C.c_int), ("metric", C.c_int), ("dim", C.c_uint32), ("n_points", C.c_uint64),
        ("n_start", C.c_uint32), ("vectors", C.c_void_p), ("row_stride", C.c_uint64),
        ("adj", C.c_void_p), ("adj_stride", C.c_uint32),
        ("pq_pivots", C.c_void_p), ("pq_offsets", C.c_void_p), ("pq_chunks", C.c_uint32),
        ("pq_centers", C.c_uint32), ("pq_codes", C.c_void_p),
        ("sq_rows", C.c_void_p), ("sq_nbits", C.c_int), ("sq_shift", C.c_void_p), ("sq_scale", C.c_float),
        ("sq_shift_square_norm", C.c_float), ("sq_mean_norm", C.c_float),
    ]


_lib = None


def build():
    """(Re)build liboracle.so from source if it is missing or stale."""
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("distance.cpp", "pq.cpp", "graph.cpp", "oracle.h", "Makefile")]
    if os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return
    subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(LIB_PATH)
    vp, sz, f, i, u32, u64 = C.c_void_p, C.c_size_t, C.c_float, C.c_int, C.c_uint32, C.c_uint64
    L.orc_distance.restype = f
    L.orc_distance.argtypes = [i, i, i, i, vp, vp, sz, C.POINTER(C.c_int)]
    L.orc_distance_rows.restype = None
    L.orc_distance_rows.argtypes = [i, i, i, i, vp, vp, sz, sz, sz, vp]
    L.orc_f16_to_f32.restype = f
    L.orc_f16_to_f32.argtypes = [C.c_uint16]
    L.orc_f32_to_f16.restype = C.c_uint16
    L.orc_f32_to_f16.argtypes = [f]
    L.orc_pq_chunk_offsets.restype = None
    L.orc_pq_chunk_offsets.argtypes = [sz, sz, vp]
    L.orc_pq_populate_lut.restype = None
    L.orc_pq_populate_lut.argtypes = [vp, sz, sz, vp, sz, i, vp, vp]
    L.orc_pq_lookup.restype = f
    L.orc_pq_lookup.argtypes = [vp, sz, vp, sz]
    L.orc_pq_query_distances.restype = None
    L.orc_pq_query_distances.argtypes = [vp, sz, sz, vp, sz, i, vp, vp, sz, vp]
    L.orc_pq_direct_distance.restype = f
    L.orc_pq_direct_distance.argtypes = [vp, sz, vp, sz, i, vp, vp]
    L.orc_pq_self_distance.restype = f
    L.orc_pq_self_distance.argtypes = [vp, sz, vp, sz, i, vp, vp]
    L.orc_pq_encode.restype = i
    L.orc_pq_encode.argtypes = [vp, sz, sz, vp, sz, vp, vp]
    L.orc_sq_encode_row.restype = None
    L.orc_sq_encode_row.argtypes = [vp, f, sz, i, vp, vp]
    L.orc_pq_train.restype = i
    L.orc_pq_train.argtypes = [vp, u64, u32, u32, u32, u32, u64, vp, vp]
    L.orc_sq_compress.restype = f
    L.orc_sq_compress.argtypes = [vp, f, sz, i, vp, vp, C.POINTER(C.c_int)]
    L.orc_sq_distance.restype = f
    L.orc_sq_distance.argtypes = [i, i, f, f, vp, f, vp, f, sz]
    L.orc_minmax_row_bytes.restype = sz
    L.orc_minmax_row_bytes.argtypes = [sz, i]
    L.orc_minmax_compress.restype = i
    L.orc_minmax_compress.argtypes = [f, sz, i, vp, vp, vp]
    L.orc_minmax_full_query_meta.restype = i
    L.orc_minmax_full_query_meta.argtypes = [vp, sz, vp, vp]
    L.orc_minmax_distance.restype = f
    L.orc_minmax_distance.argtypes = [i, i, i, vp, vp]
    L.orc_minmax_query_distance.restype = f
    L.orc_minmax_query_distance.argtypes = [i, i, vp, f, f, vp]
    L.orc_minmax_decompress.restype = None
    L.orc_minmax_decompress.argtypes = [vp, i, vp]
    L.orc_search.restype = u32
    L.orc_search.argtypes = [C.POINTER(OrcIndex), vp, u32, u32, u32, i, vp, vp, vp, vp]
    L.orc_search_batch.restype = None
    L.orc_search_batch.argtypes = [C.POINTER(OrcIndex), vp, u64, u32, u32, u32, u32, i, i, vp, vp, vp, vp, vp]
    L.orc_search_batch_rerank.restype = None
    L.orc_search_batch_rerank.argtypes = [C.POINTER(OrcIndex), vp, u64, u32, u32, u32, u32, i, i, vp, vp, vp, vp, vp]
    L.orc_update_occlude_factor.restype = f
    L.orc_update_occlude_factor.argtypes = [i, f, f, f, f]
    L.orc_robust_prune.restype = u32
    L.orc_robust_prune.argtypes = [C.POINTER(OrcIndex), vp, vp, vp, u32, u32, f, i, vp, vp]
    L.orc_build.restype = None
    L.orc_build.argtypes = [i, i, u32, u64, u32, vp, u64, u32, u32, u32, f, vp, u32]
    L.orc_build_batched.restype = None
    L.orc_build_batched.argtypes = [i, i, u32, u64, u32, vp, u64, u32, u32, u32, f, u32, vp, u32]
    L.orc_set_pool_tie_mode.restype = None
    L.orc_set_pool_tie_mode.argtypes = [i]
    L.orc_last_build_counts.restype = None
    L.orc_last_build_counts.argtypes = [C.POINTER(u64), C.POINTER(u64)]
    L.orc_queue_new.restype = vp
    L.orc_queue_new.argtypes = [u32]
    L.orc_queue_free.restype = None
    L.orc_queue_free.argtypes = [vp]
    L.orc_queue_insert.restype = None
    L.orc_queue_insert.argtypes = [vp, u32, f]
    L.orc_queue_has_notvisited.restype = i
    L.orc_queue_has_notvisited.argtypes = [vp]
    L.orc_queue_closest_notvisited.restype = i
    L.orc_queue_closest_notvisited.argtypes = [vp, C.POINTER(u32), C.POINTER(f)]
    L.orc_queue_size.restype = u32
    L.orc_queue_size.argtypes = [vp]
    L.orc_queue_get.restype = None
    L.orc_queue_get.argtypes = [vp, u32, C.POINTER(u32), C.POINTER(f), C.POINTER(C.c_int)]
    L.orc_bruteforce_knn.restype = None
    L.orc_bruteforce_knn.argtypes = [i, i, u32, vp, u64, u64, vp, u64, u32, u32, i, vp, vp]
    L.orc_recall.restype = C.c_double
    L.orc_recall.argtypes = [vp, u32, vp, u32, vp, u32, u32, u32]
    L.orc_hardware_threads.restype = i
    _lib = L
    return L


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def distance(x, y, metric, flavour=SIMD):
    x = np.ascontiguousarray(x)
    y = np.ascontiguousarray(y)
    assert x.shape == y.shape and x.ndim == 1
    err = C.c_int(0)
    v = lib().orc_distance(flavour, dtype_code(x), dtype_code(y), metric, ptr(x), ptr(y), x.shape[0], C.byref(err))
    if err.value:
        raise ValueError("unsupported dtype pair / metric")
    return np.float32(v)


def distance_rows(query, rows, metric, flavour=SIMD):
    query = np.ascontiguousarray(query)
    rows = np.ascontiguousarray(rows)
    out = np.empty(rows.shape[0], np.float32)
    lib().orc_distance_rows(flavour, dtype_code(query), dtype_code(rows), metric, ptr(query), ptr(rows),
                            rows.strides[0], rows.shape[0], rows.shape[1], ptr(out))
    return out


def sq_encode_rows(vectors_f32, shift, scale, nbits):
    """SQStore::set_vector for every row: canonical-front rows (f32 compensation | dense N-bit codes)."""
    vectors_f32 = np.ascontiguousarray(vectors_f32, np.float32)
    shift = np.ascontiguousarray(shift, np.float32)
    n, dim = vectors_f32.shape
    rows = np.zeros((n, 4 + (dim * nbits + 7) // 8), np.uint8)
    L = lib()
    for i in range(n):
        L.orc_sq_encode_row(ptr(shift), scale, dim, nbits, ptr(vectors_f32[i]), ptr(rows[i]))
    return rows


def pq_offsets(dim, n_chunks):
    out = np.zeros(n_chunks + 1, np.uint64)
    lib().orc_pq_chunk_offsets(dim, n_chunks, ptr(out))
    return out


def pq_train(data, n_chunks, n_centers=256, lloyds_reps=5, seed=0):
    """(pivots [n_centers, dim] f32, offsets u64, status) — train_pq's per-chunk k-means++ + Lloyd."""
    data = np.ascontiguousarray(data, np.float32)
    pivots = np.zeros((n_centers, data.shape[1]), np.float32)
    offsets = np.zeros(n_chunks + 1, np.uint64)
    st = lib().orc_pq_train(ptr(data), data.shape[0], data.shape[1], n_chunks, n_centers, lloyds_reps, seed, ptr(pivots), ptr(offsets))
    return pivots, offsets, st


def lloyds(data, centers, reps):
    """lloyds(data, centers, max_reps): (assignments u32 [n], loss, centers after the update)."""
    data = np.ascontiguousarray(data, np.float32)
    centers = np.ascontiguousarray(centers, np.float32).copy()
    assign = np.zeros(data.shape[0], np.uint32)
    loss = C.c_float()
    lib().orc_lloyds.restype = None
    lib().orc_lloyds(ptr(data), C.c_uint64(data.shape[0]), C.c_uint32(data.shape[1]), ptr(centers), C.c_uint32(centers.shape[0]),
                     C.c_uint32(reps), ptr(assign), C.byref(loss))
    return assign, loss.value, centers


class Index:
    """Host-side view of an index for the oracle (keeps the numpy arrays alive)."""

    def __init__(self, vectors, adj, n_points, n_start, metric, pq=None, sq=None):
        self.vectors = np.ascontiguousarray(vectors)
        self.adj = np.ascontiguousarray(adj, dtype=np.uint32)
        assert self.vectors.shape[0] == n_points + n_start == self.adj.shape[0]
        self.n_points, self.n_start, self.metric = n_points, n_start, metric
        self.pq = pq  # (pivots f32 [centers, dim], offsets u64, codes u8 [n_total, chunks])
        s = OrcIndex()
        s.dtype = dtype_code(self.vectors)
        s.metric = metric
        s.dim = self.vectors.shape[1]
        s.n_points = n_points
        s.n_start = n_start
        s.vectors = self.vectors.ctypes.data
        s.row_stride = self.vectors.strides[0]
        s.adj = self.adj.ctypes.data
        s.adj_stride = self.adj.shape[1]
        if pq is not None:
            piv, off, codes = pq
            self._piv = np.ascontiguousarray(piv, np.float32)
            self._off = np.ascontiguousarray(off, np.uint64)
            self._codes = np.ascontiguousarray(codes, np.uint8)
            s.pq_pivots = self._piv.ctypes.data
            s.pq_offsets = self._off.ctypes.data
            s.pq_chunks = self._codes.shape[1]
            s.pq_centers = self._piv.shape[0]
            s.pq_codes = self._codes.ctypes.data
        if sq is not None:
            # (rows u8 [n_total, 4 + ceil(dim * nbits / 8)], nbits, shift f32 [dim], scale, shift_square_norm, mean_norm)
            rows, nbits, shift, scale, ssn, mean_norm = sq
            self._sq_rows = np.ascontiguousarray(rows, np.uint8)
            self._sq_shift = np.ascontiguousarray(shift, np.float32)
            assert self._sq_rows.shape == (n_points + n_start, 4 + (self.vectors.shape[1] * nbits + 7) // 8)
            s.sq_rows = self._sq_rows.ctypes.data
            s.sq_nbits = nbits
            s.sq_shift = self._sq_shift.ctypes.data
            s.sq_scale = scale
            s.sq_shift_square_norm = ssn
            s.sq_mean_norm = mean_norm
        self.c = s

    def search_batch(self, queries, k, l_search, beam=1, flavour=AVX2, threads=1):
        queries = np.ascontiguousarray(queries)
        nq = queries.shape[0]
        ids = np.empty((nq, k), np.uint32)
        dists = np.empty((nq, k), np.float32)
        counts = np.empty(nq, np.uint32)
        cmps = np.empty(nq, np.uint32)
        hops = np.empty(nq, np.uint32)
        lib().orc_search_batch(C.byref(self.c), ptr(queries), queries.strides[0], nq, k, l_search, beam,
                               flavour, threads, ptr(ids), ptr(dists), ptr(counts), ptr(cmps), ptr(hops))
        return ids, dists, counts, cmps, hops


    def search_batch_rerank(self, queries, k, l_search, beam=1, flavour=AVX2, threads=1):
        """PQ (or full-precision) traversal followed by the providers' full-precision Rerank."""
        queries = np.ascontiguousarray(queries)
        nq = queries.shape[0]
        ids = np.empty((nq, k), np.uint32)
        dists = np.empty((nq, k), np.float32)
        counts = np.empty(nq, np.uint32)
        cmps = np.empty(nq, np.uint32)
        hops = np.empty(nq, np.uint32)
        lib().orc_search_batch_rerank(C.byref(self.c), ptr(queries), queries.strides[0], nq, k, l_search, beam,
                                      flavour, threads, ptr(ids), ptr(dists), ptr(counts), ptr(cmps), ptr(hops))
        return ids, dists, counts, cmps, hops


def build_graph(vectors, n_points, n_start, metric, pruned_degree, max_degree, l_build, alpha=1.2, tie_mode=0):
    """tie_mode 1: order exactly tied prune candidates the way the Rust standard library would
    (as far as oracle/graph.cpp restates it); 0: stable."""
    vectors = np.ascontiguousarray(vectors)
    lib().orc_set_pool_tie_mode(tie_mode)
    stride = max_degree + 1
    adj = np.zeros((n_points + n_start, stride), np.uint32)
    lib().orc_build(dtype_code(vectors), metric, vectors.shape[1], n_points, n_start, ptr(vectors),
                    vectors.strides[0], pruned_degree, max_degree, l_build, alpha, ptr(adj), stride)
    lib().orc_set_pool_tie_mode(0)
    return adj


def build_graph_batched(vectors, n_points, n_start, metric, pruned_degree, max_degree, l_build, alpha=1.2, batch_size=0,
                        bootstrap=False):
    """DiskANNIndex::multi_insert over the device build's batch schedule (batch_size 1 == build_graph).
    bootstrap=True: fixed chunks of `batch_size` points and the reference's bootstrap routine under its own condition
    (what the reference's drivers run; the device build does not)."""
    vectors = np.ascontiguousarray(vectors)
    stride = max_degree + 1
    adj = np.zeros((n_points + n_start, stride), np.uint32)
    lib().orc_set_multi_insert_bootstrap(1 if bootstrap else 0)
    lib().orc_build_batched(dtype_code(vectors), metric, vectors.shape[1], n_points, n_start, ptr(vectors),
                            vectors.strides[0], pruned_degree, max_degree, l_build, alpha, batch_size, ptr(adj), stride)
    lib().orc_set_multi_insert_bootstrap(0)
    return adj


def last_bootstrap_counts():
    """(batches of the last build_graph_batched that ran the bootstrap routine, batches for which the reference's condition held)."""
    a, b = C.c_uint64(), C.c_uint64()
    lib().orc_last_bootstrap_counts(C.byref(a), C.byref(b))
    return a.value, b.value


def last_build_counts():
    """(set_neighbors, append_neighbors) of the last build_graph call."""
    a, b = C.c_uint64(), C.c_uint64()
    lib().orc_last_build_counts(C.byref(a), C.byref(b))
    return a.value, b.value


def bruteforce_knn(base, queries, metric, k, threads=None):
    base = np.ascontiguousarray(base)
    queries = np.ascontiguousarray(queries)
    threads = threads or lib().orc_hardware_threads()
    ids = np.empty((queries.shape[0], k), np.uint32)
    dists = np.empty((queries.shape[0], k), np.float32)
    lib().orc_bruteforce_knn(dtype_code(base), metric, base.shape[1], ptr(base), base.shape[0], base.strides[0],
                             ptr(queries), queries.strides[0], queries.shape[0], k, threads, ptr(ids), ptr(dists))
    return ids, dists


def recall(gt, res, counts, k, n):
    gt = np.ascontiguousarray(gt, np.uint32)
    res = np.ascontiguousarray(res, np.uint32)
    c = None if counts is None else np.ascontiguousarray(counts, np.uint32)
    return lib().orc_recall(ptr(gt), gt.shape[1], ptr(res), res.shape[1], ptr(c), gt.shape[0], k, n)


# ---------------------------------------------------------------- MinMax quantizer

def minmax_compress(vectors, nbits, grid_scale=1.0):
    """MinMaxQuantizer::compress for every row: (rows u8 [n, 20 + ceil(dim * nbits / 8)], loss f32 [n], nan flags)."""
    vectors = np.ascontiguousarray(vectors, np.float32)
    n, dim = vectors.shape
    L = lib()
    rb = L.orc_minmax_row_bytes(dim, nbits)
    rows = np.zeros((n, rb), np.uint8)
    loss = np.zeros(n, np.float32)
    nan = np.zeros(n, bool)
    for r in range(n):
        nan[r] = L.orc_minmax_compress(grid_scale, dim, nbits, ptr(vectors[r]), ptr(rows[r]), ptr(loss[r:r + 1])) != 0
    return rows, loss, nan


def minmax_distances(metric, nbits_x, nbits_y, x_rows, y_rows):
    L = lib()
    return np.array([L.orc_minmax_distance(metric, nbits_x, nbits_y, ptr(x_rows[r]), ptr(y_rows[r])) for r in range(x_rows.shape[0])],
                    np.float32)


def minmax_decompress(rows, nbits, dim):
    L = lib()
    out = np.zeros((rows.shape[0], dim), np.float32)
    for r in range(rows.shape[0]):
        L.orc_minmax_decompress(ptr(rows[r]), nbits, ptr(out[r]))
    return out


def minmax_query_distances(metric, nbits, queries, rows):
    """FullQuery x Data distances, all pairs: out[q, r] (the FullQueryMeta of every query is computed first)."""
    L = lib()
    queries = np.ascontiguousarray(queries, np.float32)
    out = np.zeros((queries.shape[0], rows.shape[0]), np.float32)
    for qi in range(queries.shape[0]):
        s, ns = np.zeros(1, np.float32), np.zeros(1, np.float32)
        assert L.orc_minmax_full_query_meta(ptr(queries[qi]), queries.shape[1], ptr(s), ptr(ns)) == 0
        for r in range(rows.shape[0]):
            out[qi, r] = L.orc_minmax_query_distance(metric, nbits, ptr(queries[qi]), float(s[0]), float(ns[0]), ptr(rows[r]))
    return out
