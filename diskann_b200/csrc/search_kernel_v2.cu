// search_kernel_v2.cu — latency-restructured batched greedy search for float rows
// (f32 x f32 and f32-widened x f16; L2 / InnerProduct / CosineNormalized schemas, NA = 4).
//
// Same semantics as search_kernel.cu (DiskANNIndex::search_internal, index.rs:1933-2000;
// NeighborPriorityQueue, queue.rs:130-318; expand_beam, provider.rs:436-479) and bit-identical
// results; what changes is how many dependent memory round trips a hop costs:
//
//   * rows of ALL surviving candidates of a hop are fetched with one TMA bulk copy each
//     (cp.async.bulk global -> shared, completion on a per-warp mbarrier): one instruction per
//     row, no registers tied up, every row of the hop in flight at once;
//   * distances are computed from shared memory (lane s <-> SIMD slot s, conflict-free) for 8
//     rows per pass and reduced with a transpose-butterfly in the reference's association
//     (xor 8, 16, [remainder], 4, 2, 1) — 9 shuffles per 8 rows;
//   * the sorted candidate list lives in registers (blocked: lane l owns entries
//     l*QR .. l*QR+QR-1), insert = redux.sync lower bound + one shuffle carry;
//   * all visited-set probes of an adjacency row are issued together; the adjacency row of the
//     next-best unvisited candidate is prefetched into L2 while the current hop runs.
#include "dab_common.cuh"
#include "distance_device.cuh"

#include <algorithm>
#include <cstdlib>

namespace dab {

constexpr int kV2Warps = 4;
constexpr uint32_t kEmptyV2 = 0xFFFFFFFFu;
constexpr uint32_t kFlagV2 = 0x80000000u;
constexpr int kGroup = 8;  // rows reduced together

struct SearchParamsV2 {
    const uint8_t* vectors;
    size_t row_stride;
    const uint32_t* adj;
    uint32_t adj_stride;
    uint64_t n_points;
    uint32_t n_start;
    uint32_t dim;
    uint32_t max_degree;
    const void* queries;
    const uint32_t* query_rows;
    const uint32_t* query_list;
    uint32_t n_work;
    uint32_t k, cap, beam;
    uint32_t* out_ids;
    float* out_dists;
    uint32_t* out_counts;
    uint32_t* out_cmps;
    uint32_t* out_hops;
    uint32_t* tables;
    uint32_t hcap_log2;
    uint32_t* counters;
    uint32_t* overflow_list;
    uint32_t* rec_ids;
    float* rec_dists;
    uint32_t* rec_counts;
    uint32_t rec_cap;
    // per-warp shared memory layout (bytes)
    uint32_t warp_smem, off_q, off_cid, off_cd, off_beam, off_rows, off_bar;
    uint32_t row_bytes;   // bytes copied per row (multiple of 16)
    uint32_t row_slot;    // bytes between staged rows
    uint32_t stage_rows;  // rows staged per round (multiple of kGroup)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t phase) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
            : "=r"(ok)
            : "r"(bar), "r"(phase)
            : "memory");
    }
}
__device__ __forceinline__ void bulk_row(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ uint32_t hash_id_v2(uint32_t id, uint32_t log2cap) { return (id * 0x9E3779B1u) >> (32u - log2cap); }

// ---- register-resident sorted list, blocked layout: entry e lives in lane e / QR, reg e % QR
template <int QR>
struct RegQueue {
    float d[QR];
    uint32_t id[QR];  // bit 31 = visited flag
};

template <int QR>
__device__ __forceinline__ void rq_clear(RegQueue<QR>& q) {
#pragma unroll
    for (int r = 0; r < QR; ++r) {
        q.d[r] = __int_as_float(0x7F800000);
        q.id[r] = kEmptyV2;
    }
}

// NeighborPriorityQueue::insert (queue.rs:130-171); warp-uniform (id, x); `size` uniform
template <int QR>
__device__ __forceinline__ void rq_insert(RegQueue<QR>& q, uint32_t cap, uint32_t& size, uint32_t id, float x, int lane) {
    if (x != x) return;
    if (size == cap) {
        const uint32_t le = cap - 1;
        float last = q.d[0];
#pragma unroll
        for (int r = 1; r < QR; ++r)
            if ((int)(le % QR) == r) last = q.d[r];
        last = __shfl_sync(kFull, last, (int)(le / QR));
        if (last < x) return;
    }
    // lower bound = number of live entries with distance < x
    int c = 0;
#pragma unroll
    for (int r = 0; r < QR; ++r) c += ((uint32_t)(lane * QR + r) < size && q.d[r] < x) ? 1 : 0;
    const uint32_t pos = (uint32_t)__reduce_add_sync(kFull, c);
    const float cd = __shfl_up_sync(kFull, q.d[QR - 1], 1);
    const uint32_t ci = __shfl_up_sync(kFull, q.id[QR - 1], 1);
#pragma unroll
    for (int r = QR - 1; r >= 0; --r) {
        const uint32_t e = (uint32_t)(lane * QR + r);
        if (e > pos) {
            q.d[r] = r > 0 ? q.d[r > 0 ? r - 1 : 0] : cd;
            q.id[r] = r > 0 ? q.id[r > 0 ? r - 1 : 0] : ci;
        } else if (e == pos) {
            q.d[r] = x;
            q.id[r] = id;
        }
    }
    if (size == cap) {
        // the evicted tail moved to index cap: wipe it (falls off the end when cap == 32*QR)
#pragma unroll
        for (int r = 0; r < QR; ++r)
            if ((uint32_t)(lane * QR + r) == cap) {
                q.d[r] = __int_as_float(0x7F800000);
                q.id[r] = kEmptyV2;
            }
    } else {
        ++size;
    }
}

// closest_notvisited (queue.rs:297-313): marks and returns the first unvisited entry below
// `lim`, or kEmptyV2; *dist receives its distance
template <int QR>
__device__ __forceinline__ uint32_t rq_pop(RegQueue<QR>& q, uint32_t lim, float* dist, int lane) {
    int first = QR;
#pragma unroll
    for (int r = QR - 1; r >= 0; --r)
        if ((uint32_t)(lane * QR + r) < lim && !(q.id[r] & kFlagV2)) first = r;
    const unsigned m = __ballot_sync(kFull, first < QR);
    if (!m) return kEmptyV2;
    const int src = __ffs(m) - 1;
    uint32_t id = 0;
    float d = 0.0f;
#pragma unroll
    for (int r = 0; r < QR; ++r)
        if (first == r) {
            id = q.id[r];
            d = q.d[r];
            if (lane == src) q.id[r] = id | kFlagV2;
        }
    *dist = __shfl_sync(kFull, d, src);
    return __shfl_sync(kFull, id, src);
}

// first unvisited entry below lim without marking (for the adjacency prefetch)
template <int QR>
__device__ __forceinline__ uint32_t rq_peek(const RegQueue<QR>& q, uint32_t lim, int lane) {
    int first = QR;
#pragma unroll
    for (int r = QR - 1; r >= 0; --r)
        if ((uint32_t)(lane * QR + r) < lim && !(q.id[r] & kFlagV2)) first = r;
    const unsigned m = __ballot_sync(kFull, first < QR);
    if (!m) return kEmptyV2;
    const int src = __ffs(m) - 1;
    uint32_t id = 0;
#pragma unroll
    for (int r = 0; r < QR; ++r)
        if (first == r) id = q.id[r];
    return __shfl_sync(kFull, id, src);
}

template <int QR>
__device__ __forceinline__ bool rq_has_unvisited(const RegQueue<QR>& q, uint32_t lim, int lane) {
    bool any = false;
#pragma unroll
    for (int r = 0; r < QR; ++r) any |= (uint32_t)(lane * QR + r) < lim && !(q.id[r] & kFlagV2);
    return __any_sync(kFull, any);
}

// transpose-butterfly stage over M live values (see flat_kernels.cu)
template <int M>
__device__ __forceinline__ void bfly8(float (&v)[kGroup], int lane, int bit) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < M / 2; ++i) {
        const float keep = up ? v[M / 2 + i] : v[i];
        const float send = up ? v[i] : v[M / 2 + i];
        v[i] = __fadd_rn(keep, __shfl_xor_sync(kFull, send, bit));
    }
}

// distances of 8 staged rows (shared memory) against the query (shared memory, f32); returns on
// every lane the value of row u = ((lane>>3)&1)<<2 | ((lane>>4)&1)<<1 | ((lane>>2)&1)
template <typename TD, int KIND>
__device__ __forceinline__ float group_distance(const float* __restrict__ q, const uint8_t* __restrict__ rows,
                                                uint32_t row_slot, int dim, int lane) {
    float v[kGroup];
#pragma unroll
    for (int g = 0; g < kGroup; ++g) v[g] = 0.0f;
    const int full8 = dim & ~7, rem = dim & 7;
    for (int e = lane; e < full8; e += 32) {
        const float x = q[e];
#pragma unroll
        for (int g = 0; g < kGroup; ++g) {
            const float y = to_f32(reinterpret_cast<const TD*>(rows + (size_t)g * row_slot)[e]);
            if (KIND == KIND_L2) {
                const float c = __fsub_rn(x, y);
                v[g] = __fmaf_rn(c, c, v[g]);
            } else {
                v[g] = __fmaf_rn(x, y, v[g]);
            }
        }
    }
    bfly8<8>(v, lane, 8);
    bfly8<4>(v, lane, 16);
    if (rem) {
        // two live values: index i | b4 << 1 | b3 << 2 (b3 = lane bit 3, b4 = lane bit 4)
        const int hi = (((lane >> 4) & 1) << 1) | (((lane >> 3) & 1) << 2);
        const int l = lane & 7;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int g = i | hi;
            const float x = l < rem ? q[full8 + l] : 0.0f;
            const float y = l < rem ? to_f32(reinterpret_cast<const TD*>(rows + (size_t)g * row_slot)[full8 + l]) : 0.0f;
            if (KIND == KIND_L2) {
                const float c = __fsub_rn(x, y);
                v[i] = __fmaf_rn(c, c, v[i]);
            } else {
                v[i] = __fmaf_rn(x, y, v[i]);
            }
        }
    }
    bfly8<2>(v, lane, 4);
    float a = v[0];
    a = __fadd_rn(a, __shfl_xor_sync(kFull, a, 2));
    a = __fadd_rn(a, __shfl_xor_sync(kFull, a, 1));
    return a;
}

template <typename TD, int KIND, int POST, int QR>
__global__ void __launch_bounds__(kV2Warps * 32) search_kernel_v2(const SearchParamsV2 p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint8_t* base = smem + (size_t)wib * p.warp_smem;
    float* qf = reinterpret_cast<float*>(base + p.off_q);
    uint32_t* cid = reinterpret_cast<uint32_t*>(base + p.off_cid);
    float* cd = reinterpret_cast<float*>(base + p.off_cd);
    uint32_t* beam_ids = reinterpret_cast<uint32_t*>(base + p.off_beam);
    uint8_t* rows = base + p.off_rows;
    const uint32_t bar = smem_u32(base + p.off_bar);
    const uint32_t rows_a = smem_u32(rows);

    const uint32_t warp_slot = blockIdx.x * kV2Warps + wib;
    uint32_t* table = p.tables + ((size_t)warp_slot << p.hcap_log2);
    const uint32_t hcap = 1u << p.hcap_log2, hmask = hcap - 1;
    const uint32_t hlimit = hcap - (hcap >> 2);
    const uint64_t n_total = p.n_points + p.n_start;
    const int dim = (int)p.dim;
    uint32_t phase = 0;
    if (lane == 0) mbar_init(bar);
    __syncwarp();

    for (;;) {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(p.counters, 1u);
        w = __shfl_sync(kFull, w, 0);
        if (w >= p.n_work) break;
        const uint32_t qidx = p.query_list ? p.query_list[w] : w;

        __syncwarp();
        {
            const TD* s = p.query_rows ? reinterpret_cast<const TD*>(p.vectors + (size_t)p.query_rows[qidx] * p.row_stride)
                                       : reinterpret_cast<const TD*>(p.queries) + (size_t)qidx * dim;
            for (int e = lane; e < dim; e += 32) qf[e] = to_f32(s[e]);
            uint4 e4 = make_uint4(kEmptyV2, kEmptyV2, kEmptyV2, kEmptyV2);
            uint4* t4 = reinterpret_cast<uint4*>(table);
            for (uint32_t i = lane; i < (hcap >> 2); i += 32) t4[i] = e4;
        }
        __syncwarp();

        RegQueue<QR> best;
        rq_clear(best);
        uint32_t size = 0, cmps = 0, hops = 0, nvisited = 0, nrec = 0;
        bool overflow = false;

        // stages `n` candidate rows (ids in cid[c0..)) and computes their distances into cd[]
        auto distances = [&](uint32_t c0, uint32_t n) {
            if (lane == 0) mbar_expect(bar, n * p.row_bytes);
            __syncwarp();
            if ((uint32_t)lane < n) bulk_row(rows_a + lane * p.row_slot, p.vectors + (size_t)cid[c0 + lane] * p.row_stride, p.row_bytes, bar);
            for (uint32_t j = 32 + lane; j < n; j += 32)
                bulk_row(rows_a + j * p.row_slot, p.vectors + (size_t)cid[c0 + j] * p.row_stride, p.row_bytes, bar);
            mbar_wait(bar, phase);
            phase ^= 1;
            for (uint32_t g0 = 0; g0 < n; g0 += kGroup) {
                const float r = group_distance<TD, KIND>(qf, rows + (size_t)g0 * p.row_slot, p.row_slot, dim, lane);
                const uint32_t u = (((lane >> 3) & 1) << 2) | (((lane >> 4) & 1) << 1) | ((lane >> 2) & 1);
                if ((lane & 3) == 0 && g0 + u < n) cd[c0 + g0 + u] = post_op<POST>(r);
            }
            __syncwarp();
        };

        // ---- start points
        for (uint32_t s0 = 0; s0 < p.n_start; s0 += p.stage_rows) {
            const uint32_t n = min(p.stage_rows, p.n_start - s0);
            for (uint32_t j = lane; j < n; j += 32) {
                const uint32_t id = (uint32_t)p.n_points + s0 + j;
                cid[j] = id;
                uint32_t h = hash_id_v2(id, p.hcap_log2);
                for (;;) {
                    uint32_t old = atomicCAS(table + h, kEmptyV2, id);
                    if (old == kEmptyV2 || old == id) break;
                    h = (h + 1) & hmask;
                }
            }
            __syncwarp();
            distances(0, n);
            for (uint32_t j = 0; j < n; ++j) rq_insert(best, p.cap, size, cid[j], cd[j], lane);
            nvisited += n;
            cmps += n;
            __syncwarp();
        }

        // ---- greedy loop
        while (rq_has_unvisited(best, min(p.cap, size), lane)) {
            uint32_t nb = 0;
            while (nb < p.beam) {
                float nd;
                const uint32_t id = rq_pop(best, min(p.cap, size), &nd, lane);
                if (id == kEmptyV2) break;
                if (lane == 0) {
                    beam_ids[nb] = id;
                    if (p.rec_ids && nrec < p.rec_cap) {
                        p.rec_ids[(size_t)qidx * p.rec_cap + nrec] = id;
                        p.rec_dists[(size_t)qidx * p.rec_cap + nrec] = nd;
                    }
                }
                ++nrec;
                ++nb;
            }
            __syncwarp();
            {
                // speculative: the next hop most likely expands the now-first unvisited entry
                const uint32_t nxt = rq_peek(best, min(p.cap, size), lane);
                if (nxt != kEmptyV2 && lane < 3) prefetch_l2(p.adj + (size_t)nxt * p.adj_stride + lane * 32);
            }

            uint32_t ncand = 0;
            for (uint32_t b = 0; b < nb; ++b) {
                const uint32_t node = beam_ids[b];
                const uint32_t* row = p.adj + (size_t)node * p.adj_stride;
                uint32_t wd[3];
                wd[0] = __ldg(row + lane);
                wd[1] = 32 + lane < p.adj_stride ? __ldg(row + 32 + lane) : kEmptyV2;
                wd[2] = 64 + lane < p.adj_stride ? __ldg(row + 64 + lane) : kEmptyV2;
                const uint32_t deg = min(__shfl_sync(kFull, wd[0], 0), p.max_degree);
                // first probes of all three chunks in flight together
                bool valid[3];
                uint32_t h[3], seen[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const uint32_t j = c * 32 + lane;
                    valid[c] = j >= 1 && j <= deg;
                    h[c] = hash_id_v2(wd[c], p.hcap_log2);
                    seen[c] = valid[c] ? __ldcg(table + h[c]) : wd[c];
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if ((uint32_t)c * 32 > deg) break;
                    bool inserted = false;
                    if (valid[c]) {
                        uint32_t old = seen[c], hh = h[c];
                        for (;;) {
                            if (old == kEmptyV2) old = atomicCAS(table + hh, kEmptyV2, wd[c]);
                            if (old == kEmptyV2) {
                                inserted = true;
                                break;
                            }
                            if (old == wd[c]) break;
                            hh = (hh + 1) & hmask;
                            old = __ldcg(table + hh);
                        }
                    }
                    const bool isnew = inserted && wd[c] < n_total;
                    const unsigned mi = __ballot_sync(kFull, inserted);
                    const unsigned mn = __ballot_sync(kFull, isnew);
                    if (isnew) cid[ncand + __popc(mn & ((1u << lane) - 1u))] = wd[c];
                    ncand += __popc(mn);
                    nvisited += __popc(mi);
                }
                // adjacency rows longer than 95 neighbours (max_degree > 95): remaining chunks
                for (uint32_t c0 = 96; c0 < deg + 1; c0 += 32) {
                    const uint32_t j = c0 + lane;
                    const uint32_t word = j < p.adj_stride ? __ldg(row + j) : kEmptyV2;
                    const bool v = j <= deg;
                    bool inserted = false;
                    if (v) {
                        uint32_t hh = hash_id_v2(word, p.hcap_log2);
                        for (;;) {
                            uint32_t old = __ldcg(table + hh);
                            if (old == kEmptyV2) old = atomicCAS(table + hh, kEmptyV2, word);
                            if (old == kEmptyV2) {
                                inserted = true;
                                break;
                            }
                            if (old == word) break;
                            hh = (hh + 1) & hmask;
                        }
                    }
                    const bool isnew = inserted && word < n_total;
                    const unsigned mi = __ballot_sync(kFull, inserted);
                    const unsigned mn = __ballot_sync(kFull, isnew);
                    if (isnew) cid[ncand + __popc(mn & ((1u << lane) - 1u))] = word;
                    ncand += __popc(mn);
                    nvisited += __popc(mi);
                }
                if (nvisited + p.max_degree + 32 > hlimit) overflow = true;
            }
            if (overflow) break;
            __syncwarp();

            for (uint32_t c0 = 0; c0 < ncand; c0 += p.stage_rows) distances(c0, min(p.stage_rows, ncand - c0));

            // best.insert in adjacency order; candidates that cannot enter a full list are skipped
            for (uint32_t c0 = 0; c0 < ncand; c0 += 32) {
                const uint32_t j = c0 + lane;
                const float dj = j < ncand ? cd[j] : __int_as_float(0x7FC00000);
                const uint32_t ij = j < ncand ? cid[j] : 0;
                float worst = __int_as_float(0x7F800000);
                if (size == p.cap) {
                    const uint32_t le = p.cap - 1;
                    float t = best.d[0];
#pragma unroll
                    for (int r = 1; r < QR; ++r)
                        if ((int)(le % QR) == r) t = best.d[r];
                    worst = __shfl_sync(kFull, t, (int)(le / QR));
                }
                unsigned m = __ballot_sync(kFull, j < ncand && !(worst < dj));
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    rq_insert(best, p.cap, size, __shfl_sync(kFull, ij, src), __shfl_sync(kFull, dj, src), lane);
                }
            }
            cmps += ncand;
            hops += nb;
        }

        if (overflow) {
            if (lane == 0) {
                uint32_t o = atomicAdd(p.counters + 1, 1u);
                p.overflow_list[o] = qidx;
            }
            continue;
        }

        // ---- post-process: drop start points, first k
        {
            const uint32_t n = min(p.cap, size);
            uint32_t count = 0;
            // entries are blocked per lane (index order = lane-major): exclusive prefix of the
            // kept flags over (lane, r) gives each kept entry its output position
            int keptc = 0;
            bool keep[QR];
#pragma unroll
            for (int r = 0; r < QR; ++r) {
                const uint32_t e = (uint32_t)(lane * QR + r);
                keep[r] = e < n && (best.id[r] & ~kFlagV2) < p.n_points;
                keptc += keep[r] ? 1 : 0;
            }
            int incl = keptc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += t;
            }
            uint32_t pos = (uint32_t)(incl - keptc);
            count = (uint32_t)__shfl_sync(kFull, incl, 31);
#pragma unroll
            for (int r = 0; r < QR; ++r) {
                if (keep[r]) {
                    if (pos < p.k) {
                        p.out_ids[(size_t)qidx * p.k + pos] = best.id[r] & ~kFlagV2;
                        p.out_dists[(size_t)qidx * p.k + pos] = best.d[r];
                    }
                    ++pos;
                }
            }
            count = min(count, p.k);
            for (uint32_t i = count + lane; i < p.k; i += 32) {
                p.out_ids[(size_t)qidx * p.k + i] = kEmptyV2;
                p.out_dists[(size_t)qidx * p.k + i] = __int_as_float(0x7F800000);
            }
            if (lane == 0) {
                atomicMax(p.counters + 2, nvisited);
                if (p.out_counts) p.out_counts[qidx] = count;
                if (p.out_cmps) p.out_cmps[qidx] = cmps;
                if (p.out_hops) p.out_hops[qidx] = hops;
                if (p.rec_counts) p.rec_counts[qidx] = min(nrec, p.rec_cap);
            }
        }
    }
}

// ------------------------------------------------------------------ host side
struct V2Launch {
    void (*kern)(const SearchParamsV2);
    size_t smem_block;
    int grid;
};

// Returns 1 when this configuration is not covered by v2 (caller falls back to v1), 0 on
// success with `out` filled, or a negative DAB error code.
int v2_prepare(const dab_index* idx, uint32_t l_search, uint32_t beam, SearchParamsV2& p, V2Launch& out) {
    if (getenv("DAB_DISABLE_V2")) return 1;
    if (idx->dtype != DAB_F32 && idx->dtype != DAB_F16) return 1;
    const MetricPlan plan = plan_for(idx->metric, false);
    if (plan.kind == KIND_COS) return 1;
    const uint32_t cap = l_search + idx->n_start;
    if (cap > 256 || idx->max_degree > 1000) return 1;
    const uint32_t row_bytes = (uint32_t)round_up((size_t)idx->dim * elem_size(idx->dtype), 16);
    if (row_bytes > idx->row_stride) return 1;
    const uint32_t row_slot = row_bytes + 16;  // +16 B: rows start on different banks
    size_t off = 0;
    p.off_q = (uint32_t)off;
    off += round_up((size_t)idx->dim * 4, 16);
    const size_t ncand_max = (size_t)beam * idx->max_degree;
    p.off_cid = (uint32_t)off;
    off += round_up(std::max<size_t>(ncand_max, idx->n_start) * 4, 16);
    p.off_cd = (uint32_t)off;
    off += round_up(std::max<size_t>(ncand_max, idx->n_start) * 4, 16);
    p.off_beam = (uint32_t)off;
    off += round_up((size_t)beam * 4, 16);
    p.off_bar = (uint32_t)off;
    off += 16;
    off = round_up(off, 128);
    p.off_rows = (uint32_t)off;
    const size_t fixed = off;
    // rows staged per round: as many as fit ~6 KB per warp, a multiple of the reduce group
    uint32_t stage = (uint32_t)std::max<size_t>(kGroup, (6144 / row_slot) / kGroup * kGroup);
    stage = std::min<uint32_t>(stage, 32);
    p.stage_rows = stage;
    p.row_bytes = row_bytes;
    p.row_slot = row_slot;
    p.warp_smem = (uint32_t)round_up(fixed + (size_t)stage * row_slot, 128);
    out.smem_block = (size_t)p.warp_smem * kV2Warps;
    if (out.smem_block > 200 * 1024) return 1;

#define PICK2(TD, K, P, Q) out.kern = search_kernel_v2<TD, K, P, Q>
#define PICK_Q(TD, K, P)                 \
    do {                                 \
        if (cap <= 128) PICK2(TD, K, P, 4); \
        else PICK2(TD, K, P, 8);         \
    } while (0)
#define PICK_T(TD)                                                       \
    do {                                                                 \
        if (plan.kind == KIND_L2) PICK_Q(TD, KIND_L2, POST_ID);           \
        else if (plan.post == POST_NEG) PICK_Q(TD, KIND_IP, POST_NEG);    \
        else PICK_Q(TD, KIND_IP, POST_ONE_MINUS);                         \
    } while (0)
    if (idx->dtype == DAB_F32) PICK_T(float);
    else PICK_T(__half);
#undef PICK_T
#undef PICK_Q
#undef PICK2
    if (cudaFuncSetAttribute(out.kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)out.smem_block) != cudaSuccess) {
        cudaGetLastError();
        return 1;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, out.kern, kV2Warps * 32, out.smem_block) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        return 1;
    }
    out.grid = per_sm * idx->sm_count;
    return 0;
}

}  // namespace dab
