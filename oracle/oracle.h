/*
 * oracle.h — C ABI of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  The oracle is a CPU restatement of the arithmetic and the
 * search / prune loops of the reference (microsoft/DiskANN, Rust workspace v0.56.0) for the
 * batched-distance hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product (diskann_b200/) never does.
 *
 * Parity status: the reference cannot be compiled here (no Rust toolchain), so the oracle is
 * pinned against every portable golden vector the reference's own tests hold for this path
 * (tests/golden/, tests/test_oracle_golden.py): the 2x256 f32 L2 KAT == 429141.2, the
 * Specialize<3> KAT, corner-value broadcasts, the cosine zero-norm rule, the f16 infinity
 * rule, the PQ closed-form table, the update_occlude_factor table, the queue behaviour
 * tests, the SQ doctest and the checked-in grid greedy-search JSON baselines.
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * reference checkout).
 */
#ifndef DISKANN_B200_ORACLE_H
#define DISKANN_B200_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element types (same numbering as include/diskann_b200.h) */
enum { ORC_F32 = 0, ORC_F16 = 1, ORC_I8 = 2, ORC_U8 = 3 };
/* diskann-vector/src/distance/metric.rs:8-20 (#[repr(C)]) */
enum { ORC_COSINE = 0, ORC_INNER_PRODUCT = 1, ORC_L2 = 2, ORC_COSINE_NORMALIZED = 3 };
/* which restatement of the kernel to run */
enum {
    ORC_FLAVOUR_SIMD = 0,   /* lane-exact emulation of the x86-64-v3 (AVX2) schemas, simd.rs */
    ORC_FLAVOUR_SCALAR = 1, /* the scalar definitions in distance/reference.rs              */
    ORC_FLAVOUR_AVX2 = 2    /* real AVX2/FMA/F16C intrinsics in the same order (CPU baseline) */
};

/* ---------------------------------------------------------------- distances */

/* Similarity-score convention of the reference (implementations.rs:217-404):
 *   L2 -> sum (x-y)^2 ; InnerProduct -> -sum xy ; Cosine -> 1 - cos (clamped) ;
 *   CosineNormalized -> 1 - sum xy (floats), == Cosine for integers.
 * Supported (dtype_x, dtype_y): (f32,f32) (f16,f16) (f32,f16) (i8,i8) (u8,u8).
 * Returns NaN and sets *err (if non-null) for unsupported combinations. */
float orc_distance(int flavour, int dtype_x, int dtype_y, int metric, const void* x,
                   const void* y, size_t dim, int* err);

/* 1 query x n rows; rows are `row_stride_bytes` apart. */
void orc_distance_rows(int flavour, int dtype_q, int dtype_rows, int metric, const void* query,
                       const void* rows, size_t row_stride_bytes, size_t n, size_t dim,
                       float* out);

/* f16 <-> f32 (half 2.6 semantics: IEEE binary16, round-to-nearest-even on narrowing) */
float orc_f16_to_f32(uint16_t h);
uint16_t orc_f32_to_f16(float f);

/* ---------------------------------------------------------------- product quantization */

/* diskann-quantization/src/views.rs:226-243: near-equal partition; offsets has n_chunks+1 */
void orc_pq_chunk_offsets(size_t dim, size_t n_chunks, uint64_t* offsets);

/* fixed_chunk_pq_table.rs:152-187: lut[c*n_centers + p]; metric L2 or InnerProduct (-dot) */
void orc_pq_populate_lut(const float* pivots, size_t n_centers, size_t dim,
                         const uint64_t* offsets, size_t n_chunks, int metric,
                         const float* query, float* lut);

/* fixed_chunk_pq_table.rs:82-98 */
float orc_pq_lookup(const uint8_t* code, size_t n_chunks, const float* lut, size_t n_centers);

/* QueryComputer semantics (pq/distance/dynamic.rs:63-87): L2 & CosineNormalized -> TableL2,
 * IP -> TableIP, Cosine -> DirectCosine.  Evaluates n codes for one query. */
void orc_pq_query_distances(const float* pivots, size_t n_centers, size_t dim,
                            const uint64_t* offsets, size_t n_chunks, int metric,
                            const float* query, const uint8_t* codes, size_t n, float* out);

/* fixed_chunk_pq_table.rs:35-59,223-281: full-precision query x code, no LUT */
float orc_pq_direct_distance(const float* pivots, size_t dim, const uint64_t* offsets,
                             size_t n_chunks, int metric, const float* query,
                             const uint8_t* code);

/* fixed_chunk_pq_table.rs:285-361: code x code */
float orc_pq_self_distance(const float* pivots, size_t dim, const uint64_t* offsets,
                           size_t n_chunks, int metric, const uint8_t* left,
                           const uint8_t* right);

/* product/tables/basic.rs:161-194: argmin SquaredL2 per chunk, strict <; returns 0, or
 * 1 + chunk index of the first chunk whose minimum distance is infinite/NaN */
int orc_pq_encode(const float* pivots, size_t n_centers, size_t dim, const uint64_t* offsets,
                  size_t n_chunks, const float* vec, uint8_t* code);

/* ---------------------------------------------------------------- scalar quantization (8-bit) */

/* scalar/quantizer.rs:190-239,407-430.  codes: dim bytes; returns compensation.
 * nbits in {1,2,4,8}; codes are written one per byte (unpacked) — packing is a storage
 * concern, the arithmetic is on the integer codes. */
float orc_sq_compress(const float* shift, float scale, size_t dim, int nbits, const float* vec,
                      uint8_t* codes, int* had_nan);

/* scalar/vectors.rs:206-237 (L2), 310-376 (IP, negated), 380-440 (CosineNormalized) */
float orc_sq_distance(int metric, int nbits, float scale_squared, float shift_square_norm,
                      const uint8_t* x, float comp_x, const uint8_t* y, float comp_y, size_t dim);

/* SQStore::set_vector (providers inmem/scalar.rs:150-175): compress `vec` into one canonical-front row
 * (4 + ceil(dim * nbits / 8) bytes). */
void orc_sq_encode_row(const float* shift, float scale, size_t dim, int nbits, const float* vec, uint8_t* row);

/* PQ codebook training (diskann-providers/src/index/diskann_async.rs:61-89 -> pq_construction.rs:163-243
 * -> diskann-quantization/src/product/train.rs): per chunk k-means++ (algorithms/kmeans/plusplus.rs)
 * and `lloyds_reps` Lloyd iterations (lloyds.rs), arithmetic in the reference's order; the random draws
 * come from SplitMix64(seed + chunk) instead of Rust's StdRng.  pivots: [n_centers][dim], offsets: [n_chunks+1]. */
int orc_pq_train(const float* data, uint64_t n, uint32_t dim, uint32_t n_chunks, uint32_t n_centers,
                 uint32_t lloyds_reps, uint64_t seed, float* pivots, uint64_t* offsets);

/* ---------------------------------------------------------------- graph search / prune */

typedef struct orc_index {
    int dtype;               /* element type of the stored vectors                        */
    int metric;
    uint32_t dim;
    uint64_t n_points;       /* data points: ids [0, n_points)                            */
    uint32_t n_start;        /* frozen start points: ids [n_points, n_points + n_start)   */
    const void* vectors;     /* (n_points + n_start) rows                                 */
    uint64_t row_stride;     /* bytes between rows                                        */
    const uint32_t* adj;     /* (n_points + n_start) rows; row[0] = degree, then ids      */
    uint32_t adj_stride;     /* u32 words between adjacency rows (>= max_degree + 1)      */
    /* optional PQ traversal (QuantAccessor): when pq_codes != NULL search distances are
     * QueryComputer distances over codes instead of full-precision rows */
    const float* pq_pivots;
    const uint64_t* pq_offsets;
    uint32_t pq_chunks;
    uint32_t pq_centers;
    const uint8_t* pq_codes; /* (n_points + n_start) x pq_chunks */
    /* optional scalar-quantized traversal (providers inmem/scalar.rs QuantAccessor): when
     * sq_rows != NULL search distances are Compensated{SquaredL2,IP,CosineNormalized} between the
     * compressed query and the stored rows.  Rows are in the reference's canonical-front layout
     * (meta/vector.rs:478-507): f32 compensation, then the dense N-bit codes (bits/slice.rs:261-323). */
    const uint8_t* sq_rows;  /* (n_points + n_start) x (4 + ceil(dim * sq_nbits / 8)) */
    int sq_nbits;
    const float* sq_shift;   /* [dim] */
    float sq_scale;
    float sq_shift_square_norm;
    float sq_mean_norm;      /* 0: no rescale (quantizer trained without the mean norm) */
} orc_index;

/* diskann/src/graph/index.rs:1933-2000 + neighbor/queue.rs:130-318 + knn_search.rs:170-190.
 * `query` has the index dtype (an f16 query is widened to f32 once, layers/full.rs:421-423).
 * Results exclude start points; returns number of results written (<= k). */
uint32_t orc_search(const orc_index* idx, const void* query, uint32_t k, uint32_t l_search,
                    uint32_t beam_width, int flavour, uint32_t* out_ids, float* out_dists,
                    uint32_t* out_cmps, uint32_t* out_hops);

/* benchmark-core/src/search/api.rs:400-434: contiguous partitions, one thread each. */
void orc_search_batch(const orc_index* idx, const void* queries, uint64_t query_stride,
                      uint32_t nq, uint32_t k, uint32_t l_search, uint32_t beam_width,
                      int flavour, int n_threads, uint32_t* out_ids, float* out_dists,
                      uint32_t* out_counts, uint32_t* out_cmps, uint32_t* out_hops);

/* The same followed by the quantized providers' post-processing Pipeline<FilterStartPoints, Rerank>
 * (diskann-providers/.../inmem/product.rs:391-400, full_precision.rs:356-399): the whole candidate
 * list is re-scored with the full-precision Distance<T, T> and sorted before the first k are taken. */
void orc_search_batch_rerank(const orc_index* idx, const void* queries, uint64_t query_stride,
                             uint32_t nq, uint32_t k, uint32_t l_search, uint32_t beam_width,
                             int flavour, int n_threads, uint32_t* out_ids, float* out_dists,
                             uint32_t* out_counts, uint32_t* out_cmps, uint32_t* out_hops);

/* graph/config/mod.rs:80-103.  kind: 0 TriangleInequality, 1 Occluding */
float orc_update_occlude_factor(int kind, float d_ik, float d_jk, float cur, float alpha);

/* graph/internal/prune.rs:106-259 over a pool already sorted by source distance.
 * pool_ids/pool_dists: candidates; excluded[i] != 0 marks a `None` cache entry.
 * Writes selected pool positions to out_pos, returns count. Distances between candidates
 * are data x data distances (Distance<T,T>). */
uint32_t orc_robust_prune(const orc_index* idx, const uint32_t* pool_ids,
                          const float* pool_dists, const uint8_t* excluded, uint32_t pool_len,
                          uint32_t degree, float alpha, int flavour, uint32_t* out_pos,
                          uint64_t* out_ncmp);

/* Sequential Vamana build by single inserts (index.rs:226-341), for test graphs.
 * vectors: n_points+n_start rows (start rows must be filled by the caller);
 * adj: zero-initialised output, (n_points+n_start) x adj_stride. */
void orc_build(int dtype, int metric, uint32_t dim, uint64_t n_points, uint32_t n_start,
               const void* vectors, uint64_t row_stride, uint32_t pruned_degree,
               uint32_t max_degree, uint32_t l_build, float alpha, uint32_t* adj,
               uint32_t adj_stride);
/* DiskANNIndex::multi_insert (index.rs:815-1050; intra_batch_candidates = None, bootstrap branch not
 * taken) over consecutive id ranges whose sizes follow orc_build_batch_size — the schedule of the
 * device build (batch_size 0: its default).  batch_size 1 == orc_build. */
uint32_t orc_build_batch_size(uint32_t batch_size, uint64_t n_points, uint64_t inserted);
void orc_build_batched(int dtype, int metric, uint32_t dim, uint64_t n_points, uint32_t n_start,
                       const void* vectors, uint64_t row_stride, uint32_t pruned_degree,
                       uint32_t max_degree, uint32_t l_build, float alpha, uint32_t batch_size,
                       uint32_t* adj, uint32_t adj_stride);
/* adjacency writes of the last orc_build: full-list writes and single-edge appends (what the
 * reference's test provider counts as set_neighbors / append_neighbors in the grid_insert baselines) */
void orc_last_build_counts(uint64_t* set_neighbors, uint64_t* append_neighbors);
/* order of exactly tied candidates in a prune pool: 0 (default) = stable sort by distance,
 * 1 = follow the simple parts of Rust's select_nth_unstable_by + sort_unstable_by (graph.cpp) */
void orc_set_pool_tie_mode(int mode);

/* NeighborPriorityQueue exposed for the reference's queue unit tests (queue.rs:607-...) */
typedef struct orc_queue orc_queue;
orc_queue* orc_queue_new(uint32_t capacity);
void orc_queue_free(orc_queue*);
void orc_queue_insert(orc_queue*, uint32_t id, float dist);
int orc_queue_has_notvisited(const orc_queue*);
int orc_queue_closest_notvisited(orc_queue*, uint32_t* id, float* dist);
uint32_t orc_queue_size(const orc_queue*);
void orc_queue_get(const orc_queue*, uint32_t i, uint32_t* id, float* dist, int* visited);

/* ---------------------------------------------------------------- measurement helpers */

/* exact k-NN by brute force with the oracle distance (ties: lower id first) */
void orc_bruteforce_knn(int dtype, int metric, uint32_t dim, const void* base, uint64_t n,
                        uint64_t row_stride, const void* queries, uint64_t query_stride,
                        uint32_t nq, uint32_t k, int n_threads, uint32_t* out_ids,
                        float* out_dists);

/* benchmark-core/src/recall.rs:146-236: k-recall@n = |top-k(gt) ∩ top-n(result)| summed */
double orc_recall(const uint32_t* gt, uint32_t gt_stride, const uint32_t* res,
                  uint32_t res_stride, const uint32_t* res_counts, uint32_t nq, uint32_t k,
                  uint32_t n);

int orc_hardware_threads(void);

/* lloyds(data, centers, max_reps) (diskann-quantization/src/algorithms/kmeans/lloyds.rs:372-460) over whole rows: centers
 * [n_centers][dim] in / out, assignments [n], *loss = residual of the last round.  The loop orc_pq_train runs per chunk;
 * pinned by the reference's closed-form end_to_end_test (lloyds.rs:620-660). */
void orc_lloyds(const float* data, uint64_t n, uint32_t dim, float* centers, uint32_t n_centers, uint32_t reps,
                uint32_t* assignments, float* loss);

/* multi_insert's bootstrap routine (index.rs:589-747, 917-937) inside orc_build_batched: 0 (default) never — batches then
 * follow the device build's growth rule; 1: under the reference's condition, with fixed chunks of `batch_size` like the
 * reference's drivers.  orc_last_bootstrap_counts: batches that ran it / batches for which the condition held. */
void orc_set_multi_insert_bootstrap(int mode);
void orc_last_bootstrap_counts(uint64_t* ran, uint64_t* condition_held);

/* ---------------------------------------------------------------- MinMax quantizer (oracle/minmax.cpp)
 * diskann-quantization/src/minmax/{quantizer.rs, vectors.rs}, Transform::Null.  Rows use the canonical-front layout of
 * Data<NBITS>: MinMaxCompensation {dim u32, b, n, a, norm_squared} (20 bytes), then dense N-bit codes.
 * Pinned by the reference's own closed-form and property tests (quantizer.rs:473-756, vectors.rs:520-700), restated in
 * tests/test_oracle_minmax.py with the reference's tolerances; seeded with Rust's StdRng there, own seeds here. */
size_t orc_minmax_row_bytes(size_t dim, int nbits);
int orc_minmax_compress(float grid_scale, size_t dim, int nbits, const float* v, uint8_t* row, float* loss_out);
int orc_minmax_full_query_meta(const float* v, size_t dim, float* sum_out, float* norm_squared_out);
float orc_minmax_distance(int metric, int nbits_x, int nbits_y, const uint8_t* x_row, const uint8_t* y_row);
void orc_minmax_decompress(const uint8_t* row, int nbits, float* out);
/* FullQueryRef x DataRef<NBITS>: the f32 x N-bit inner product of bits/distances.rs:2295-2725 (x86-64-v3 lane order for
 * 1 / 2 / 4 bits, the scalar loop for 8) + the epilogue of vectors.rs:272-476 */
float orc_minmax_query_distance(int metric, int nbits, const float* query, float q_sum, float q_norm_squared, const uint8_t* row);

#ifdef __cplusplus
}
#endif
#endif
