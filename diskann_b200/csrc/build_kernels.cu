// build_kernels.cu — build-side reuse of the distance path: robust_prune on the device and a
// batched Vamana construction.
//
//   * robust_prune (diskann/src/graph/internal/prune.rs:106-259) runs one warp per pool: the
//     pool is sorted by (source distance, arrival order) in shared memory, the lazy
//     candidate x neighbour distances are data x data Distance<T,T> evaluations in the
//     reference's SIMD order (several neighbours gathered per pass, applied in order with the
//     reference's early break), PruneKind rules from graph/config/mod.rs:80-103.
//   * dab_build inserts points in batches (DiskANNIndex::multi_insert semantics,
//     diskann/src/graph/index.rs:815): the batch is searched against the current graph with
//     the search kernel recording the expanded nodes (VisitedSearchRecord, index.rs:276-282),
//     pruned (index.rs:2349-2380), and the back-edges are grouped by destination (stable radix
//     sort, cub) and applied one at a time per destination exactly like add_edge_and_prune
//     (index.rs:2264-2341) -> robust_prune_list (index.rs:2397-2454).
#include "dab_common.cuh"
#include "distance_device.cuh"

#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cfloat>
#include <vector>

namespace dab {

int run_search(dab_index* idx, const void* d_queries, const uint32_t* d_query_rows, uint32_t nq, uint32_t k,
               uint32_t l_search, uint32_t beam, uint32_t* d_ids, float* d_dists, uint32_t* d_counts, uint32_t* d_cmps,
               uint32_t* d_hops, uint32_t* rec_ids, float* rec_dists, uint32_t* rec_counts, uint32_t rec_cap);

constexpr int kPruneWarps = 4;
constexpr uint32_t kMaxOcclusion = 750;  // graph/config/defaults.rs:13
constexpr int kPairsPerPass = 4;         // neighbour rows gathered per pass

struct PruneSmem {
    uint32_t* ids;      // [P] candidate ids (sorted by distance)
    float* d;           // [P] source distances
    float* occl;        // [P] State::occlude_factor
    uint16_t* last;     // [P] State::last_checked
    uint16_t* nbr;      // [P] State::neighbor
    uint32_t* order;    // [P] sort tie-break (arrival order)
};

__device__ __forceinline__ PruneSmem carve(uint8_t* base, uint32_t P) {
    PruneSmem s;
    s.ids = reinterpret_cast<uint32_t*>(base);
    s.d = reinterpret_cast<float*>(base + 4 * (size_t)P);
    s.occl = reinterpret_cast<float*>(base + 8 * (size_t)P);
    s.order = reinterpret_cast<uint32_t*>(base + 12 * (size_t)P);
    s.last = reinterpret_cast<uint16_t*>(base + 16 * (size_t)P);
    s.nbr = reinterpret_cast<uint16_t*>(base + 18 * (size_t)P);
    return s;
}
__host__ __device__ inline size_t prune_smem_bytes(uint32_t P) { return 20 * (size_t)P; }

// PruneKind::update_occlude_factor, graph/config/mod.rs:80-103 (kind 0 triangle, 1 occluding)
__device__ __forceinline__ float update_occlude(int kind, float d_ik, float d_jk, float cur, float alpha) {
    if (kind == 0) {
        if (d_jk == 0.0f) return FLT_MAX;
        return fmaxf(cur, __fdiv_rn(d_ik, d_jk));
    }
    if (d_jk < __fmul_rn(alpha, d_ik)) return __fadd_rn(alpha, 0.01f);
    return cur;
}

// Bitonic sort of the first n (padded to pow2 P2 <= P) entries by (distance, arrival order):
// SortedNeighbors::new (graph/internal/sorted_neighbors.rs:26-44); ties are unspecified in the
// reference (unstable sort), this kernel breaks them by arrival order (as do the parity tests).
__device__ __forceinline__ void warp_sort_pool(PruneSmem s, uint32_t n, uint32_t P2, int lane, bool preset_order = false) {
    for (uint32_t i = n + lane; i < P2; i += 32) {
        s.d[i] = __int_as_float(0x7F800000);
        s.ids[i] = kNoId;
        s.order[i] = 0xFFFFFFFFu;
    }
    if (!preset_order)
        for (uint32_t i = lane; i < n; i += 32) s.order[i] = i;
    __syncwarp();
    for (uint32_t k = 2; k <= P2; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t t = lane; t < P2; t += 32) {
                const uint32_t o = t ^ j;
                if (o > t) {
                    const float da = s.d[t], db = s.d[o];
                    const uint32_t oa = s.order[t], ob = s.order[o];
                    const bool a_less = da < db || (da == db && oa < ob);
                    const bool asc = (t & k) == 0;
                    if (asc ? !a_less : a_less) {
                        const uint32_t ia = s.ids[t];
                        s.d[t] = db;
                        s.d[o] = da;
                        s.order[t] = ob;
                        s.order[o] = oa;
                        s.ids[t] = s.ids[o];
                        s.ids[o] = ia;
                    }
                }
            }
            __syncwarp();
        }
    }
}

// data x data distance between row a and up to G rows b[g]; returns post-op'ed values on all lanes
template <typename TD, int NA, int KIND, int POST, bool IS_INT, bool SIGNED>
__device__ __forceinline__ void warp_row_distances(const uint8_t* vectors, size_t row_stride, int dim, uint32_t a,
                                                   const uint32_t (&b)[kPairsPerPass], int lane, float (&out)[kPairsPerPass]) {
    if constexpr (IS_INT) {
        const uint8_t* q = vectors + (size_t)a * row_stride;
        const uint8_t* rows[kPairsPerPass];
#pragma unroll
        for (int g = 0; g < kPairsPerPass; ++g) rows[g] = vectors + (size_t)b[g] * row_stride;
        const int qq = KIND == KIND_IP ? 0 : warp_int_self<SIGNED>(q, dim, lane);
        float r[kPairsPerPass];
        warp_int_multi<SIGNED, KIND, kPairsPerPass>(q, rows, dim, lane, qq, r);
#pragma unroll
        for (int g = 0; g < kPairsPerPass; ++g) out[g] = post_op<POST>(r[g]);
    } else {
        constexpr int S = 8 * NA, TEAMS = 32 / S;
        constexpr int U = kPairsPerPass / TEAMS;
        const int team = lane / S, slot = lane % S;
        const TD* q = reinterpret_cast<const TD*>(vectors + (size_t)a * row_stride);
        const TD* rows[U];
#pragma unroll
        for (int u = 0; u < U; ++u) rows[u] = reinterpret_cast<const TD*>(vectors + (size_t)b[u * TEAMS + team] * row_stride);
        float r[U];
        team_float_multi<NA, KIND, U>(q, rows, dim, slot, r);
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int t = 0; t < TEAMS; ++t) out[u * TEAMS + t] = post_op<POST>(__shfl_sync(kFull, r[u], t * S));
        }
    }
}

// robust_prune over the sorted pool in shared memory (prune.rs:106-259); returns `found`,
// selected pool positions in s.nbr[0..found).
template <typename TD, int NA, int KIND, int POST, bool IS_INT, bool SIGNED>
__device__ __forceinline__ uint32_t warp_robust_prune(PruneSmem s, uint32_t n, uint32_t location, uint32_t degree, float alpha,
                                                      int prune_kind, const uint8_t* vectors, size_t row_stride, int dim,
                                                      int lane) {
    for (uint32_t i = lane; i < n; i += 32) {
        s.occl[i] = 0.0f;
        s.last[i] = 0;
        s.nbr[i] = 0;
    }
    __syncwarp();
    uint32_t found = 0;
    float current_alpha = 1.0f;
    const float increment = fminf(alpha, 1.2f);
    while (found < degree) {
        for (uint32_t i = 0; i < n; ++i) {
            if (found >= degree) break;
            float of = s.occl[i];
            uint32_t lc = s.last[i];
            if (of > current_alpha) continue;
            const uint32_t cand = s.ids[i];
            if (cand == location) {  // sorted_cache entry None (index.rs:2608-2614)
                __syncwarp();
                if (lane == 0) s.occl[i] = FLT_MAX;
                __syncwarp();
                continue;
            }
            const float d_ik = s.d[i];
            while (lc != found) {
                // gather the next neighbours that actually need a distance
                uint32_t rows[kPairsPerPass], rpos[kPairsPerPass];
                int cnt = 0;
#pragma unroll
                for (int g = 0; g < kPairsPerPass; ++g) {
                    rpos[g] = lc + g < found ? s.nbr[lc + g] : 0xFFFFFFFFu;
                    const bool need = rpos[g] != 0xFFFFFFFFu && rpos[g] < i;
                    rows[g] = need ? s.ids[rpos[g]] : cand;
                    cnt += rpos[g] != 0xFFFFFFFFu;
                }
                float dist[kPairsPerPass];
                warp_row_distances<TD, NA, KIND, POST, IS_INT, SIGNED>(vectors, row_stride, dim, cand, rows, lane, dist);
                bool stop = false;
#pragma unroll
                for (int g = 0; g < kPairsPerPass; ++g) {
                    if (g < cnt && !stop) {
                        ++lc;
                        if (rpos[g] < i) {
                            of = update_occlude(prune_kind, d_ik, dist[g], of, current_alpha);
                            if (of > current_alpha) stop = true;
                        }
                    }
                }
                if (stop) break;
            }
            __syncwarp();
            if (lane == 0) {
                s.last[i] = (uint16_t)lc;
                if (of > current_alpha) {
                    s.occl[i] = of;
                } else {
                    s.occl[i] = FLT_MAX;
                    s.nbr[found] = (uint16_t)i;
                }
            }
            __syncwarp();
            if (!(of > current_alpha)) ++found;
        }
        if (current_alpha == alpha) break;
        current_alpha = fminf(__fmul_rn(current_alpha, increment), alpha);
    }
    return found;
}

struct PruneParams {
    const uint8_t* vectors;
    size_t row_stride;
    int dim;
    uint32_t P;            // smem slots per warp (pow2)
    // pools
    const uint32_t* pool_ids;   // [n_pools][pool_cap]
    const float* pool_d;        // [n_pools][pool_cap]
    const uint32_t* pool_len;   // [n_pools]
    uint32_t pool_cap;
    const uint32_t* locations;  // [n_pools] id excluded from its own pool
    uint32_t n_pools;
    uint32_t degree;
    float alpha;
    int prune_kind;
    // outputs
    uint32_t* out_ids;          // [n_pools][degree]
    uint32_t* out_counts;       // [n_pools]
    uint32_t* adj;              // optional: also write adj[location] = [count, ids]
    uint32_t adj_stride;
};

template <typename TD, int NA, int KIND, int POST, bool IS_INT, bool SIGNED>
__global__ void __launch_bounds__(kPruneWarps * 32) prune_pools_kernel(const PruneParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    PruneSmem s = carve(smem + (size_t)wib * prune_smem_bytes(p.P), p.P);
    const uint32_t nwarps = gridDim.x * kPruneWarps;
    for (uint32_t w = blockIdx.x * kPruneWarps + wib; w < p.n_pools; w += nwarps) {
        uint32_t n = min(p.pool_len[w], p.pool_cap);
        __syncwarp();
        for (uint32_t i = lane; i < n; i += 32) {
            s.ids[i] = p.pool_ids[(size_t)w * p.pool_cap + i];
            s.d[i] = p.pool_d[(size_t)w * p.pool_cap + i];
        }
        uint32_t P2 = 1;
        while (P2 < n) P2 <<= 1;
        if (P2 < 2) P2 = 2;
        __syncwarp();
        warp_sort_pool(s, n, P2, lane);
        n = min(n, kMaxOcclusion);
        const uint32_t loc = p.locations[w];
        const uint32_t found = warp_robust_prune<TD, NA, KIND, POST, IS_INT, SIGNED>(s, n, loc, p.degree, p.alpha, p.prune_kind,
                                                                                     p.vectors, p.row_stride, p.dim, lane);
        __syncwarp();
        for (uint32_t f = lane; f < found; f += 32) {
            const uint32_t id = s.ids[s.nbr[f]];
            if (p.out_ids) p.out_ids[(size_t)w * p.degree + f] = id;
            if (p.adj) p.adj[(size_t)loc * p.adj_stride + 1 + f] = id;
        }
        if (lane == 0) {
            if (p.out_counts) p.out_counts[w] = found;
            if (p.adj) p.adj[(size_t)loc * p.adj_stride] = found;
        }
    }
}

// ------------------------------------------------------------------ back-edges
struct BackedgeParams {
    const uint8_t* vectors;
    size_t row_stride;
    int dim;
    uint32_t P;
    const uint32_t* keys;   // sorted destinations (kNoId = padding)
    const uint32_t* vals;   // sources
    uint32_t n_pairs;
    uint32_t degree, max_degree;
    float alpha;
    int prune_kind;
    uint32_t* adj;
    uint32_t adj_stride;
    uint32_t* dropped;      // in-edges that did not fit the per-destination list of a batch
};

__global__ void make_pairs_kernel(const uint32_t* __restrict__ batch_ids, const uint32_t* __restrict__ nbr_ids,
                                  const uint32_t* __restrict__ nbr_counts, uint32_t n_batch, uint32_t degree,
                                  uint32_t max_backedges, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t total = n_batch * degree;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t b = i / degree, j = i % degree;
        const bool ok = j < nbr_counts[b] && j < max_backedges;
        keys[i] = ok ? nbr_ids[i] : kNoId;
        vals[i] = batch_ids[b];
    }
}

// One warp per destination segment: add_edge_and_prune (index.rs:2264-2341) with all the incoming
// edges of the batch at once; on overflow robust_prune_list (index.rs:2397-2454).
template <typename TD, int NA, int KIND, int POST, bool IS_INT, bool SIGNED>
__global__ void __launch_bounds__(kPruneWarps * 32) backedge_kernel(const BackedgeParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint8_t* base = smem + (size_t)wib * (prune_smem_bytes(p.P) + 4 * (size_t)p.P);
    PruneSmem s = carve(base, p.P);
    uint32_t* list = reinterpret_cast<uint32_t*>(base + prune_smem_bytes(p.P));
    const uint32_t nwarps = gridDim.x * kPruneWarps;
    const uint32_t nchunks = (p.n_pairs + 31) / 32;
    for (uint32_t c = blockIdx.x * kPruneWarps + wib; c < nchunks; c += nwarps) {
        const uint32_t i = c * 32 + lane;
        const uint32_t key = i < p.n_pairs ? p.keys[i] : kNoId;
        const uint32_t prev = (i > 0 && i < p.n_pairs) ? p.keys[i - 1] : kNoId;
        unsigned heads = __ballot_sync(kFull, key != kNoId && (i == 0 || prev != key));
        while (heads) {
            const int hl = __ffs(heads) - 1;
            heads &= heads - 1;
            const uint32_t start = c * 32 + hl;
            const uint32_t q = __shfl_sync(kFull, key, hl);
            uint32_t* row = p.adj + (size_t)q * p.adj_stride;
            uint32_t deg = min(row[0], p.max_degree);
            __syncwarp();
            for (uint32_t t = lane; t < deg; t += 32) list[t] = row[1 + t];
            __syncwarp();
            // add_edge_and_prune(sorted sources, q) (index.rs:2264-2341, called once per target by
            // multi_insert, index.rs:986-1003): extend_from_slice appends every source that is not yet
            // in the list (sources arrive in ascending id order: the radix sort is stable and the
            // pairs are generated in batch order); if the extended list still fits it is kept,
            // otherwise robust_prune_list (index.rs:2397-2454) runs ONCE over all of it: pool =
            // (id, Distance(q, id)) for id in list, id != q, sorted by distance and cut to the 750
            // closest (SortedNeighbors::new).  A hub can receive thousands of in-edges from one large
            // batch, so the pool is streamed: whenever its P slots are full it is sorted by
            // (distance, arrival index) and cut to 750 — the same 750 a single sort would keep.
            const uint32_t deg0 = deg;
            uint32_t n_new = 0;
            for (uint32_t e = start; e < p.n_pairs && p.keys[e] == q; ++e) {
                const uint32_t src = p.vals[e];
                if (src == q) continue;
                bool present = false;
                for (uint32_t t = lane; t < deg0; t += 32) present |= list[t] == src;
                if (!__any_sync(kFull, present)) ++n_new;
            }
            if (n_new == 0) continue;
            bool changed = true;
            if (deg0 + n_new <= p.max_degree) {
                for (uint32_t e = start; e < p.n_pairs && p.keys[e] == q; ++e) {
                    const uint32_t src = p.vals[e];
                    if (src == q) continue;
                    bool present = false;
                    for (uint32_t t = lane; t < deg0; t += 32) present |= list[t] == src;
                    if (__any_sync(kFull, present)) continue;
                    if (lane == 0) list[deg] = src;
                    ++deg;
                    __syncwarp();
                }
            } else {
                uint32_t n = 0, arrival = 0;
                uint32_t pend[kPairsPerPass];
                int npend = 0;
                auto flush = [&]() {  // distances of the pending ids, appended to the pool in order
                    if (npend == 0) return;
                    uint32_t rows[kPairsPerPass];
#pragma unroll
                    for (int g = 0; g < kPairsPerPass; ++g) rows[g] = pend[g < npend ? g : npend - 1];
                    float dist[kPairsPerPass];
                    warp_row_distances<TD, NA, KIND, POST, IS_INT, SIGNED>(p.vectors, p.row_stride, p.dim, q, rows, lane, dist);
#pragma unroll
                    for (int g = 0; g < kPairsPerPass; ++g) {
                        if (g < npend) {
                            if (n == p.P) {
                                __syncwarp();
                                warp_sort_pool(s, n, p.P, lane, true);
                                n = kMaxOcclusion;
                            }
                            if (lane == 0) {
                                s.ids[n] = rows[g];
                                s.d[n] = dist[g];
                                s.order[n] = arrival;
                            }
                            ++n;
                            ++arrival;
                        }
                    }
                    npend = 0;
                    __syncwarp();
                };
                for (uint32_t t = 0; t < deg0; ++t) {
                    const uint32_t id = list[t];
                    if (id == q) continue;
                    pend[npend++] = id;
                    if (npend == kPairsPerPass) flush();
                }
                for (uint32_t e = start; e < p.n_pairs && p.keys[e] == q; ++e) {
                    const uint32_t src = p.vals[e];
                    if (src == q) continue;
                    bool present = false;
                    for (uint32_t t = lane; t < deg0; t += 32) present |= list[t] == src;
                    if (__any_sync(kFull, present)) continue;
                    pend[npend++] = src;
                    if (npend == kPairsPerPass) flush();
                }
                flush();
                uint32_t P2 = 2;
                while (P2 < n) P2 <<= 1;
                warp_sort_pool(s, n, P2, lane, true);
                n = min(n, kMaxOcclusion);
                const uint32_t found = warp_robust_prune<TD, NA, KIND, POST, IS_INT, SIGNED>(
                    s, n, q, p.degree, p.alpha, p.prune_kind, p.vectors, p.row_stride, p.dim, lane);
                __syncwarp();
                for (uint32_t f = lane; f < found; f += 32) list[f] = s.ids[s.nbr[f]];
                deg = found;
                __syncwarp();
            }
            if (changed) {
                for (uint32_t t = lane; t < deg; t += 32) row[1 + t] = list[t];
                if (lane == 0) row[0] = deg;
            }
            __syncwarp();
        }
    }
}

__global__ void iota_kernel(uint32_t* p, uint32_t first, uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = first + i;
}

// ------------------------------------------------------------------ host dispatch
#define DAB_DATA_DISPATCH(idx, plan, CALL)                                                             \
    do {                                                                                               \
        switch ((idx)->dtype) {                                                                        \
            case DAB_F32:                                                                              \
                if ((plan).kind == KIND_L2) { CALL(float, 4, KIND_L2, POST_ID, false, false); }         \
                else if ((plan).kind == KIND_IP && (plan).post == POST_NEG) { CALL(float, 4, KIND_IP, POST_NEG, false, false); } \
                else if ((plan).kind == KIND_IP) { CALL(float, 4, KIND_IP, POST_ONE_MINUS, false, false); } \
                else { CALL(float, 2, KIND_COS, POST_ONE_MINUS, false, false); }                        \
                break;                                                                                 \
            case DAB_F16: /* f16 x f16: Strategy2x4 for every schema (simd.rs:989, 1752, 2591) */      \
                if ((plan).kind == KIND_L2) { CALL(__half, 2, KIND_L2, POST_ID, false, false); }        \
                else if ((plan).kind == KIND_IP && (plan).post == POST_NEG) { CALL(__half, 2, KIND_IP, POST_NEG, false, false); } \
                else if ((plan).kind == KIND_IP) { CALL(__half, 2, KIND_IP, POST_ONE_MINUS, false, false); } \
                else { CALL(__half, 2, KIND_COS, POST_ONE_MINUS, false, false); }                       \
                break;                                                                                 \
            case DAB_I8:                                                                               \
                if ((plan).kind == KIND_L2) { CALL(uint8_t, 4, KIND_L2, POST_ID, true, true); }         \
                else if ((plan).kind == KIND_IP) { CALL(uint8_t, 4, KIND_IP, POST_NEG, true, true); }   \
                else { CALL(uint8_t, 4, KIND_COS, POST_ONE_MINUS, true, true); }                        \
                break;                                                                                 \
            default:                                                                                   \
                if ((plan).kind == KIND_L2) { CALL(uint8_t, 4, KIND_L2, POST_ID, true, false); }        \
                else if ((plan).kind == KIND_IP) { CALL(uint8_t, 4, KIND_IP, POST_NEG, true, false); }  \
                else { CALL(uint8_t, 4, KIND_COS, POST_ONE_MINUS, true, false); }                       \
                break;                                                                                 \
        }                                                                                              \
    } while (0)

static uint32_t pow2_at_least(uint32_t v) {
    uint32_t p = 2;
    while (p < v) p <<= 1;
    return p;
}

static int launch_prune(const dab_index* idx, PruneParams& p) {
    const bool is_int = idx->dtype == DAB_I8 || idx->dtype == DAB_U8;
    const MetricPlan plan = plan_for(idx->metric, is_int);
    p.vectors = idx->d_vectors;
    p.row_stride = idx->row_stride;
    p.dim = (int)idx->dim;
    p.prune_kind = idx->metric == DAB_INNER_PRODUCT ? 1 : 0;  // PruneKind::from_metric, config/mod.rs:69-76
    p.P = pow2_at_least(std::max<uint32_t>(p.pool_cap, 2));
    const size_t smem = prune_smem_bytes(p.P) * kPruneWarps;
    if (smem > 200 * 1024) return fail(DAB_ERR_INVALID_ARGUMENT, "prune: pool capacity %u too large", p.pool_cap);
    const int grid = (int)std::min<uint64_t>(((uint64_t)p.n_pools + kPruneWarps - 1) / kPruneWarps, (uint64_t)idx->sm_count * 8);
#define CALL(TD, NA, K, P_, II, SG)                                                                        \
    do {                                                                                                   \
        auto kern = prune_pools_kernel<TD, NA, K, P_, II, SG>;                                             \
        DAB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
        kern<<<grid, kPruneWarps * 32, smem, idx->stream>>>(p);                                            \
    } while (0)
    DAB_DATA_DISPATCH(idx, plan, CALL);
#undef CALL
    DAB_LAUNCHED();
    DAB_CUDA(cudaGetLastError());
    return DAB_OK;
}

static int launch_backedges(const dab_index* idx, BackedgeParams& p) {
    const bool is_int = idx->dtype == DAB_I8 || idx->dtype == DAB_U8;
    const MetricPlan plan = plan_for(idx->metric, is_int);
    p.vectors = idx->d_vectors;
    p.row_stride = idx->row_stride;
    p.dim = (int)idx->dim;
    p.prune_kind = idx->metric == DAB_INNER_PRODUCT ? 1 : 0;
    // the list of a destination holds its current neighbours plus every in-edge of the batch
    p.P = std::max<uint32_t>(1024, pow2_at_least(idx->max_degree + 2));
    p.adj = idx->d_adj;
    p.adj_stride = idx->adj_stride;
    p.max_degree = idx->max_degree;
    const size_t smem = (prune_smem_bytes(p.P) + 4 * (size_t)p.P) * kPruneWarps;
    if (smem > 200 * 1024) return fail(DAB_ERR_INVALID_ARGUMENT, "build: max_degree %u too large", idx->max_degree);
    const uint32_t chunks = (p.n_pairs + 31) / 32;
    const int grid = (int)std::min<uint64_t>(((uint64_t)chunks + kPruneWarps - 1) / kPruneWarps, (uint64_t)idx->sm_count * 8);
#define CALL(TD, NA, K, P_, II, SG)                                                                        \
    do {                                                                                                   \
        auto kern = backedge_kernel<TD, NA, K, P_, II, SG>;                                                \
        DAB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
        kern<<<grid, kPruneWarps * 32, smem, idx->stream>>>(p);                                            \
    } while (0)
    DAB_DATA_DISPATCH(idx, plan, CALL);
#undef CALL
    DAB_LAUNCHED();
    DAB_CUDA(cudaGetLastError());
    return DAB_OK;
}

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { cudaFree(p); }
    int alloc(size_t n) {
        cudaFree(p);
        p = nullptr;
        cudaError_t e = cudaMalloc(&p, n ? n : 1);
        if (e != cudaSuccess) return fail(DAB_ERR_OUT_OF_MEMORY, "build: cudaMalloc(%zu) failed: %s", n, cudaGetErrorString(e));
        return DAB_OK;
    }
};

}  // namespace dab

using namespace dab;

extern "C" {

// PruneAccessor::fill + robust_prune over caller-provided pools (unsorted).  Exposed for the
// prune parity tests and for host-driven builds.
int dab_robust_prune(dab_index* idx, const uint32_t* pool_ids, const float* pool_dists, const uint32_t* pool_lens,
                     const uint32_t* locations, uint32_t n_pools, uint32_t pool_cap, uint32_t degree, float alpha,
                     uint32_t* out_ids, uint32_t* out_counts) {
    if (!idx) return fail(DAB_ERR_INVALID_ARGUMENT, "dab_robust_prune: idx is NULL");
    if (!idx->vectors_ready) return fail(DAB_ERR_NOT_READY, "dab_robust_prune: vectors not uploaded");
    if (n_pools == 0) return DAB_OK;
    if (!pool_ids || !pool_dists || !pool_lens || !locations || !out_ids || !out_counts)
        return fail(DAB_ERR_INVALID_ARGUMENT, "dab_robust_prune: NULL argument");
    if (degree == 0 || pool_cap == 0 || pool_cap > 4096) return fail(DAB_ERR_INVALID_ARGUMENT, "dab_robust_prune: bad degree / pool_cap");
    DAB_CUDA(cudaSetDevice(idx->device));
    DevBuf b_ids, b_d, b_len, b_loc, b_out, b_cnt;
    int rc;
    const size_t np = n_pools, pc = pool_cap;
    if ((rc = b_ids.alloc(np * pc * 4)) || (rc = b_d.alloc(np * pc * 4)) || (rc = b_len.alloc(np * 4)) || (rc = b_loc.alloc(np * 4)) ||
        (rc = b_out.alloc(np * degree * 4)) || (rc = b_cnt.alloc(np * 4)))
        return rc;
    DAB_CUDA(cudaMemcpyAsync(b_ids.p, pool_ids, np * pc * 4, cudaMemcpyHostToDevice, idx->stream));
    DAB_CUDA(cudaMemcpyAsync(b_d.p, pool_dists, np * pc * 4, cudaMemcpyHostToDevice, idx->stream));
    DAB_CUDA(cudaMemcpyAsync(b_len.p, pool_lens, np * 4, cudaMemcpyHostToDevice, idx->stream));
    DAB_CUDA(cudaMemcpyAsync(b_loc.p, locations, np * 4, cudaMemcpyHostToDevice, idx->stream));
    DAB_CUDA(cudaMemsetAsync(b_out.p, 0xFF, np * degree * 4, idx->stream));
    PruneParams p;
    memset(&p, 0, sizeof(p));
    p.pool_ids = (const uint32_t*)b_ids.p;
    p.pool_d = (const float*)b_d.p;
    p.pool_len = (const uint32_t*)b_len.p;
    p.pool_cap = pool_cap;
    p.locations = (const uint32_t*)b_loc.p;
    p.n_pools = n_pools;
    p.degree = degree;
    p.alpha = alpha;
    p.out_ids = (uint32_t*)b_out.p;
    p.out_counts = (uint32_t*)b_cnt.p;
    if ((rc = launch_prune(idx, p))) return rc;
    DAB_CUDA(cudaMemcpyAsync(out_ids, b_out.p, np * degree * 4, cudaMemcpyDeviceToHost, idx->stream));
    DAB_CUDA(cudaMemcpyAsync(out_counts, b_cnt.p, np * 4, cudaMemcpyDeviceToHost, idx->stream));
    DAB_CUDA(cudaStreamSynchronize(idx->stream));
    return DAB_OK;
}

int dab_build(dab_index* idx, uint32_t pruned_degree, uint32_t l_build, float alpha, uint32_t batch_size) {
    if (!idx) return fail(DAB_ERR_INVALID_ARGUMENT, "dab_build: idx is NULL");
    if (!idx->vectors_ready) return fail(DAB_ERR_NOT_READY, "dab_build: vectors (including start rows) must be uploaded first");
    if (idx->n_start == 0) return fail(DAB_ERR_INVALID_ARGUMENT, "dab_build: the index needs at least one start point");
    if (pruned_degree == 0 || pruned_degree > idx->max_degree)
        return fail(DAB_ERR_INVALID_ARGUMENT, "dab_build: pruned_degree must be in [1, max_degree]");
    if (l_build == 0) return fail(DAB_ERR_INVALID_ARGUMENT, "dab_build: l_build must be > 0");
    if (!(alpha >= 1.0f)) return fail(DAB_ERR_INVALID_ARGUMENT, "dab_build: alpha must be >= 1");
    DAB_CUDA(cudaSetDevice(idx->device));
    const uint32_t n = (uint32_t)idx->n_points;
    if (batch_size == 0) batch_size = std::max<uint32_t>(1024, std::min<uint32_t>(65536, n / 16));
    // VisitedSearchRecord keeps every expanded node (index.rs:276-282; SortedNeighbors truncates to the 750
    // closest afterwards): the record is sized generously and a search that still outgrows it is reported
    const uint32_t rec_cap = std::min<uint32_t>(2048, 4 * l_build + 64);
    idx->rec_truncated = 0;
    cudaStream_t st = idx->stream;

    DAB_CUDA(cudaMemsetAsync(idx->d_adj, 0, idx->n_total() * (size_t)idx->adj_stride * 4, st));
    idx->graph_ready = true;

    DevBuf b_batch, b_rec_ids, b_rec_d, b_rec_n, b_nbr, b_nbr_n, b_keys, b_vals, b_keys2, b_vals2, b_tmp, b_res_ids, b_res_d, b_dropped;
    int rc;
    const size_t B = batch_size;
    if ((rc = b_batch.alloc(B * 4)) || (rc = b_rec_ids.alloc(B * rec_cap * 4)) || (rc = b_rec_d.alloc(B * rec_cap * 4)) ||
        (rc = b_rec_n.alloc(B * 4)) || (rc = b_nbr.alloc(B * pruned_degree * 4)) || (rc = b_nbr_n.alloc(B * 4)) ||
        (rc = b_keys.alloc(B * pruned_degree * 4)) || (rc = b_vals.alloc(B * pruned_degree * 4)) ||
        (rc = b_keys2.alloc(B * pruned_degree * 4)) || (rc = b_vals2.alloc(B * pruned_degree * 4)) || (rc = b_res_ids.alloc(B * 4)) ||
        (rc = b_res_d.alloc(B * 4)))
        return rc;
    if ((rc = b_dropped.alloc(4))) return rc;
    DAB_CUDA(cudaMemsetAsync(b_dropped.p, 0, 4, st));
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const uint32_t*)b_keys.p, (uint32_t*)b_keys2.p, (const uint32_t*)b_vals.p,
                                    (uint32_t*)b_vals2.p, (int)(B * pruned_degree), 0, 32, st);
    if ((rc = b_tmp.alloc(tmp_bytes))) return rc;

    uint32_t inserted = 0;
    while (inserted < n) {
        // batches grow with the graph so that early points are not all inserted blind
        uint32_t b = std::min<uint32_t>(batch_size, std::max<uint32_t>(1, inserted / 8));
        b = std::min(b, n - inserted);
        iota_kernel<<<(b + 255) / 256, 256, 0, st>>>((uint32_t*)b_batch.p, inserted, b);
        DAB_LAUNCHED();
        // 1. search the batch against the current graph, recording expanded nodes
        if ((rc = run_search(idx, nullptr, (const uint32_t*)b_batch.p, b, 1, l_build, 1, (uint32_t*)b_res_ids.p, (float*)b_res_d.p,
                             nullptr, nullptr, nullptr, (uint32_t*)b_rec_ids.p, (float*)b_rec_d.p, (uint32_t*)b_rec_n.p, rec_cap)))
            return rc;
        // 2. robust_prune each point's visited pool -> out-edges
        PruneParams pp;
        memset(&pp, 0, sizeof(pp));
        pp.pool_ids = (const uint32_t*)b_rec_ids.p;
        pp.pool_d = (const float*)b_rec_d.p;
        pp.pool_len = (const uint32_t*)b_rec_n.p;
        pp.pool_cap = rec_cap;
        pp.locations = (const uint32_t*)b_batch.p;
        pp.n_pools = b;
        pp.degree = pruned_degree;
        pp.alpha = alpha;
        pp.out_ids = (uint32_t*)b_nbr.p;
        pp.out_counts = (uint32_t*)b_nbr_n.p;
        pp.adj = idx->d_adj;
        pp.adj_stride = idx->adj_stride;
        if ((rc = launch_prune(idx, pp))) return rc;
        // 3. back-edges grouped by destination
        const uint32_t n_pairs = b * pruned_degree;
        make_pairs_kernel<<<(n_pairs + 255) / 256, 256, 0, st>>>((const uint32_t*)b_batch.p, (const uint32_t*)b_nbr.p,
                                                                 (const uint32_t*)b_nbr_n.p, b, pruned_degree, pruned_degree,
                                                                 (uint32_t*)b_keys.p, (uint32_t*)b_vals.p);
        DAB_LAUNCHED();
        cudaError_t e = cub::DeviceRadixSort::SortPairs(b_tmp.p, tmp_bytes, (const uint32_t*)b_keys.p, (uint32_t*)b_keys2.p,
                                                        (const uint32_t*)b_vals.p, (uint32_t*)b_vals2.p, (int)n_pairs, 0, 32, st);
        if (e != cudaSuccess) return fail(DAB_ERR_CUDA, "build: radix sort failed: %s", cudaGetErrorString(e));
        DAB_LAUNCHED();
        BackedgeParams bp;
        memset(&bp, 0, sizeof(bp));
        bp.keys = (const uint32_t*)b_keys2.p;
        bp.vals = (const uint32_t*)b_vals2.p;
        bp.n_pairs = n_pairs;
        bp.degree = pruned_degree;
        bp.alpha = alpha;
        bp.dropped = (uint32_t*)b_dropped.p;
        if ((rc = launch_backedges(idx, bp))) return rc;
        inserted += b;
    }
    uint32_t dropped = 0;
    DAB_CUDA(cudaMemcpyAsync(&dropped, b_dropped.p, 4, cudaMemcpyDeviceToHost, st));
    DAB_CUDA(cudaStreamSynchronize(st));
    if (dropped)
        return fail(DAB_ERR_INVALID_ARGUMENT, "dab_build: %u back-edges exceeded the per-destination list of a batch and were dropped "
                    "(the graph is usable but not the reference's; use a smaller batch_size)", dropped);
    if (idx->rec_truncated)
        return fail(DAB_ERR_INVALID_ARGUMENT, "dab_build: %llu insert searches expanded more than %u nodes; their prune pools were cut "
                    "(the graph is usable but not the reference's)", (unsigned long long)idx->rec_truncated, rec_cap);
    return DAB_OK;
}

}  // extern "C"
