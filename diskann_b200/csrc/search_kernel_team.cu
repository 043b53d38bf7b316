// search_kernel_team.cu — batched greedy search with 8-lane teams: four queries per warp.
//
// Same semantics and bit-identical results as search_kernel.cu / search_kernel_v2.cu
// (DiskANNIndex::search_internal, index.rs:1933-2000; NeighborPriorityQueue, queue.rs:130-318;
// expand_beam, provider.rs:436-479), restructured for the regime the profile showed: a hop has
// only ~10-40 new candidates, so a full warp per query spends most of its instructions on
// bookkeeping that keeps 3/4 of the lanes idle.  Here a team of 8 lanes owns a query and the 4
// teams of a warp run in lock step, so every bookkeeping instruction, every memory round trip
// and every scheduler slot serves 4 queries, and a 10K-query batch is resident in one wave.
//
// Distance arithmetic: lane t of a team owns SIMD lane t of all four accumulators of the
// reference's Strategy4x1/4x2 schemas (simd.rs:321-422): element e goes to accumulator (e/8)%4,
// lane e%8, sequential FMA chain -> the accumulators are combined in-lane ((s0+s1)+(s2+s3)), the
// masked remainder is added on the combined value, and sum_tree is xor 4, 2, 1 inside the team.
//
// Scope: f32 / f16 rows (queries widened to f32), L2 / InnerProduct / CosineNormalized,
// beam width 1, L + #start <= 128, max_degree <= 95.  Everything else runs v2 / the generic kernel.
#include "dab_common.cuh"
#include "distance_device.cuh"

#include <algorithm>
#include <cstdlib>

namespace dab {

constexpr int kTeamWarps = 2;       // warps per CTA
constexpr int kTeamsPerWarp = 4;
constexpr uint32_t kE = 0xFFFFFFFFu;
constexpr uint32_t kVis = 0x80000000u;
#ifndef DAB_TEAM_ROWS
#define DAB_TEAM_ROWS 4
#endif
#ifndef DAB_TEAM_MIN_CTAS
#define DAB_TEAM_MIN_CTAS 10
#endif
constexpr int kRowsU = DAB_TEAM_ROWS;  // rows in flight per team

struct SearchParamsTeam {
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
    uint32_t k, cap;
    uint32_t* out_ids;
    float* out_dists;
    uint32_t* out_counts;
    uint32_t* out_cmps;
    uint32_t* out_hops;
    uint32_t* tables;
    uint32_t n_buckets;
    uint32_t* counters;
    uint32_t* overflow_list;
    uint32_t* rec_ids;
    float* rec_dists;
    uint32_t* rec_counts;
    uint32_t rec_cap;
    // per-team shared memory layout (bytes)
    uint32_t team_smem, off_q, off_qd, off_qi, off_cid, off_cd;
};

__device__ __forceinline__ uint32_t team_bucket_of(uint32_t id, uint32_t n_buckets) { return __umulhi(id * 0x9E3779B1u, n_buckets); }

// exact visited-set insert, one 32-byte bucket per probe (see search_kernel_v2.cu)
__device__ __forceinline__ bool team_bucket_insert(uint32_t* table, uint32_t n_buckets, uint32_t b, uint4 lo4, uint4 hi4, uint32_t id) {
    for (;;) {
        const uint32_t s[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
        bool found = false;
        int empty = -1;
#pragma unroll
        for (int k = 7; k >= 0; --k) {
            found |= s[k] == id;
            if (s[k] == kE) empty = k;
        }
        if (found) return false;
        uint32_t* bp = table + (size_t)b * 8;
        if (empty >= 0) {
            const uint32_t old = atomicCAS(bp + empty, kE, id);
            if (old == kE) return true;
            if (old == id) return false;
        } else {
            b = b + 1 == n_buckets ? 0 : b + 1;
            bp = table + (size_t)b * 8;
        }
        lo4 = __ldcg(reinterpret_cast<const uint4*>(bp));
        hi4 = __ldcg(reinterpret_cast<const uint4*>(bp) + 1);
    }
}

// team-local ballot: bit i = predicate of team lane i
__device__ __forceinline__ uint32_t team_ballot(bool pred, int g) { return (__ballot_sync(kFull, pred) >> (8 * g)) & 0xFFu; }

// 4 consecutive elements of a row / of the f32 query
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4(const __half* p) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    const __half2 lo = *reinterpret_cast<const __half2*>(&v.x), hi = *reinterpret_cast<const __half2*>(&v.y);
    return make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
}

template <int KIND>
__device__ __forceinline__ void acc4(float (&a)[4], const float4 x, const float4 y) {
    if (KIND == KIND_L2) {
        const float c0 = __fsub_rn(x.x, y.x), c1 = __fsub_rn(x.y, y.y), c2 = __fsub_rn(x.z, y.z), c3 = __fsub_rn(x.w, y.w);
        a[0] = __fmaf_rn(c0, c0, a[0]);
        a[1] = __fmaf_rn(c1, c1, a[1]);
        a[2] = __fmaf_rn(c2, c2, a[2]);
        a[3] = __fmaf_rn(c3, c3, a[3]);
    } else {
        a[0] = __fmaf_rn(x.x, y.x, a[0]);
        a[1] = __fmaf_rn(x.y, y.y, a[1]);
        a[2] = __fmaf_rn(x.z, y.z, a[2]);
        a[3] = __fmaf_rn(x.w, y.w, a[3]);
    }
}

// U rows x one query per team.  Lane t = (accumulator a = t >> 1, half h = t & 1) owns SIMD
// lanes 4h..4h+3 of accumulator a: per 32-element block it loads the 16 bytes at element
// 8a + 4h, so the team reads 128 contiguous bytes per row per block and every (accumulator,
// lane) slot is still one sequential FMA chain in increasing element order.
template <typename TD, int KIND, int U>
__device__ __forceinline__ void team_rows(const float* __restrict__ q, const TD* const (&rows)[U], int dim, int t, float (&out)[U]) {
    const int a_idx = t >> 1, h = t & 1;
    float a[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) a[u][0] = a[u][1] = a[u][2] = a[u][3] = 0.0f;
    const int full8 = dim & ~7, rem = dim & 7;
    const int nvec = full8 >> 3;
    const int blocks = nvec >> 2, ep = nvec & 3;  // full 32-element blocks, leftover full vectors
    const int lane_off = 8 * a_idx + 4 * h;
    for (int k = 0; k < blocks; ++k) {
        const int e = 32 * k + lane_off;
        const float4 x = *reinterpret_cast<const float4*>(q + e);
        float4 y[U];
#pragma unroll
        for (int u = 0; u < U; ++u) y[u] = ld4(rows[u] + e);
#pragma unroll
        for (int u = 0; u < U; ++u) acc4<KIND>(a[u], x, y[u]);
    }
    if (a_idx < ep) {  // the j-th leftover full vector goes to accumulator j (simd.rs:352-362)
        const int e = 32 * blocks + lane_off;
        const float4 x = *reinterpret_cast<const float4*>(q + e);
#pragma unroll
        for (int u = 0; u < U; ++u) acc4<KIND>(a[u], x, ld4(rows[u] + e));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        float c[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            // (s0 + s1) + (s2 + s3): accumulators live in lanes differing in bits 1 and 2
            float v = __fadd_rn(a[u][i], __shfl_xor_sync(kFull, a[u][i], 2));
            c[i] = __fadd_rn(v, __shfl_xor_sync(kFull, v, 4));
        }
        if (rem) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int l = 4 * h + i;
                const float x = l < rem ? q[full8 + l] : 0.0f;
                const float y = l < rem ? to_f32(rows[u][full8 + l]) : 0.0f;
                if (KIND == KIND_L2) {
                    const float d = __fsub_rn(x, y);
                    c[i] = __fmaf_rn(d, d, c[i]);
                } else {
                    c[i] = __fmaf_rn(x, y, c[i]);
                }
            }
        }
        // sum_tree: x[l] + x[l+4] pairs sit in lanes differing in bit 0
#pragma unroll
        for (int i = 0; i < 4; ++i) c[i] = __fadd_rn(c[i], __shfl_xor_sync(kFull, c[i], 1));
        out[u] = __fadd_rn(__fadd_rn(c[0], c[2]), __fadd_rn(c[1], c[3]));
    }
}

template <typename TD, int KIND, int POST>
__global__ void __launch_bounds__(kTeamWarps * 32, DAB_TEAM_MIN_CTAS) search_kernel_team(const SearchParamsTeam p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int g = lane >> 3, t = lane & 7;
    uint8_t* base = smem + (size_t)(wib * kTeamsPerWarp + g) * p.team_smem;
    float* qf = reinterpret_cast<float*>(base + p.off_q);
    float* qd = reinterpret_cast<float*>(base + p.off_qd);
    uint32_t* qi = reinterpret_cast<uint32_t*>(base + p.off_qi);
    uint32_t* cid = reinterpret_cast<uint32_t*>(base + p.off_cid);
    float* cd = reinterpret_cast<float*>(base + p.off_cd);

    const uint32_t team_slot = (blockIdx.x * kTeamWarps + wib) * kTeamsPerWarp + g;
    const uint32_t nbk = p.n_buckets;
    uint32_t* table = p.tables + (size_t)team_slot * nbk * 8;
    const uint32_t hlimit = nbk * 7;
    const uint64_t n_total = p.n_points + p.n_start;
    const int dim = (int)p.dim;
    const uint32_t cap = p.cap;

    // distances of candidates cid[0..n) -> cd[]; n is per team, loop bound is warp-uniform
    auto distances = [&](uint32_t n) {
        const uint32_t nmax = __reduce_max_sync(kFull, n);
        for (uint32_t c0 = 0; c0 < nmax; c0 += kRowsU) {
            const TD* rows[kRowsU];
#pragma unroll
            for (int u = 0; u < kRowsU; ++u) {
                const uint32_t c = c0 + u;
                const uint32_t id = c < n ? cid[c] : (uint32_t)p.n_points;  // any valid row
                rows[u] = reinterpret_cast<const TD*>(p.vectors + (size_t)id * p.row_stride);
            }
            float r[kRowsU];
            team_rows<TD, KIND, kRowsU>(qf, rows, dim, t, r);
#pragma unroll
            for (int u = 0; u < kRowsU; ++u)
                if (t == 0 && c0 + u < n) cd[c0 + u] = post_op<POST>(r[u]);
        }
        __syncwarp();
    };

    // batched rank-merge of candidates cid/cd[0..n) into the sorted list (see
    // search_kernel_v2.cu merge_round for the equivalence argument), 8 candidates per round
    auto merge = [&](uint32_t n, uint32_t& size, uint32_t& cursor_lo) {
        const uint32_t nmax = __reduce_max_sync(kFull, n);
        for (uint32_t c0 = 0; c0 < nmax; c0 += 8) {
            const uint32_t j = c0 + t;
            const float dj = j < n ? cd[j] : __int_as_float(0x7FC00000);
            const uint32_t idj = j < n ? cid[j] : 0;
            const float worst = size == cap ? qd[cap - 1] : __int_as_float(0x7F800000);
            const bool valid = j < n && dj == dj && !(worst < dj);
            const uint32_t vm = team_ballot(valid, g);
            if (!__any_sync(kFull, valid)) continue;
            // lower bound among old entries
            uint32_t lo = 0, hi = valid ? size : 0;
            while (__any_sync(kFull, lo < hi)) {
                const uint32_t mid = (lo + hi) >> 1;
                if (lo < hi) {
                    if (qd[mid] < dj) lo = mid + 1;
                    else hi = mid;
                }
            }
            // rank among this round's valid candidates + smallest valid distance of the team
            uint32_t rn = 0;
            float dmin = __int_as_float(0x7F800000);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float di = __shfl_sync(kFull, dj, (g << 3) | i);
                if ((vm >> i) & 1u) {
                    rn += (di < dj || (di == dj && i > t)) ? 1u : 0u;
                    dmin = fminf(dmin, di);
                }
            }
            const uint32_t pos = lo + rn;
            const bool keep_new = valid && pos < cap;
            const uint32_t nvalid = __popc(vm);
            // first old entry that moves = lower bound of dmin with `<=` shifted... entries with
            // d_old >= dmin move; find it as the minimum `lo` over the valid candidates
            uint32_t first_move = valid ? lo : 0xFFFFFFFFu;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) first_move = min(first_move, __shfl_xor_sync(kFull, first_move, o));
            // shift old entries, highest first so a write never lands on an unread entry
            const uint32_t sz = size;
            const int s_hi = vm ? (int)((sz + 7) >> 3) - 1 : -1;
            const int s_lo = vm ? (int)(first_move >> 3) : 0;
            const int s_hi_w = __reduce_max_sync(kFull, s_hi);
            for (int s = s_hi_w; s >= 0; --s) {
                const bool mine = s <= s_hi && s >= s_lo;
                const uint32_t e = (uint32_t)s * 8 + t;
                float od = 0.0f;
                uint32_t oi = 0, sh = 0;
                const bool live = mine && e < sz;
                if (live) {
                    od = qd[e];
                    oi = qi[e];
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float di = __shfl_sync(kFull, dj, (g << 3) | i);
                    if ((vm >> i) & 1u) sh += di <= od ? 1u : 0u;
                }
                __syncwarp();
                if (live && sh != 0 && e + sh < cap) {
                    qd[e + sh] = od;
                    qi[e + sh] = oi;
                }
                __syncwarp();
                if (!__any_sync(kFull, s_lo <= s_hi && s > s_lo)) break;  // some team still has lower steps to move
            }
            if (keep_new) {
                qd[pos] = dj;
                qi[pos] = idj;
            }
            size = min(cap, size + nvalid);
            uint32_t mp = keep_new ? pos : 0xFFFFFFFFu;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) mp = min(mp, __shfl_xor_sync(kFull, mp, o));
            cursor_lo = min(cursor_lo, mp);
            __syncwarp();
        }
    };

    for (;;) {
        uint32_t w0 = 0;
        if (lane == 0) w0 = atomicAdd(p.counters, (uint32_t)kTeamsPerWarp);
        w0 = __shfl_sync(kFull, w0, 0);
        if (w0 >= p.n_work) break;
        const bool active = w0 + g < p.n_work;
        const uint32_t qidx = active ? (p.query_list ? p.query_list[w0 + g] : w0 + g) : 0;

        __syncwarp();
        if (active) {
            const TD* s = p.query_rows ? reinterpret_cast<const TD*>(p.vectors + (size_t)p.query_rows[qidx] * p.row_stride)
                                       : reinterpret_cast<const TD*>(p.queries) + (size_t)qidx * dim;
            for (int e = t; e < dim; e += 8) qf[e] = to_f32(s[e]);
            const uint4 e4 = make_uint4(kE, kE, kE, kE);
            uint4* t4 = reinterpret_cast<uint4*>(table);
            for (uint32_t i = t; i < nbk * 2; i += 8) t4[i] = e4;
        }
        __syncwarp();

        uint32_t size = 0, cursor_lo = 0, cmps = 0, hops = 0, nvisited = 0, nrec = 0;
        bool overflow = false;
        bool done = !active;

        // ---- start points
        {
            uint32_t n = active ? min(p.n_start, 8u) : 0;
            if ((uint32_t)t < n) {
                const uint32_t id = (uint32_t)p.n_points + t;
                cid[t] = id;
                const uint32_t b = team_bucket_of(id, nbk);
                const uint4* bp = reinterpret_cast<const uint4*>(table + (size_t)b * 8);
                team_bucket_insert(table, nbk, b, __ldcg(bp), __ldcg(bp + 1), id);
            }
            __syncwarp();
            distances(n);
            merge(n, size, cursor_lo);
            nvisited += n;
            cmps += n;
        }

        // ---- greedy loop, all four teams in lock step
        while (__any_sync(kFull, !done)) {
            // closest_notvisited (queue.rs:297-313)
            const uint32_t lim = min(cap, size);
            uint32_t idx = 0xFFFFFFFFu;
            {
                const int steps = __reduce_max_sync(kFull, done ? 0 : (int)((lim + 7) >> 3));
                const int s0 = __reduce_min_sync(kFull, done ? 0x7FFFFFFF : (int)(cursor_lo >> 3));
                for (int s = s0; s < steps; ++s) {
                    const uint32_t e = (uint32_t)s * 8 + t;
                    const bool u = !done && idx == 0xFFFFFFFFu && e >= cursor_lo && e < lim && !(qi[e] & kVis);
                    const uint32_t m = team_ballot(u, g);
                    if (m && idx == 0xFFFFFFFFu) idx = (uint32_t)s * 8 + __ffs(m) - 1;
                    if (!__any_sync(kFull, !done && idx == 0xFFFFFFFFu && (uint32_t)(s + 1) * 8 < lim)) break;
                }
            }
            if (!done && idx == 0xFFFFFFFFu) done = true;
            if (!__any_sync(kFull, !done)) break;
            uint32_t node = 0;
            if (!done) {
                node = qi[idx];
#ifdef DAB_TEAM_DEBUG
                if (node >= n_total && t == 0)
                    printf("BAD node=%u qidx=%u team=%d idx=%u size=%u lim=%u cursor_lo=%u hops=%u cmps=%u nwork=%u q0=%u q1=%u\n", node, qidx, g, idx,
                           size, lim, cursor_lo, hops, cmps, p.n_work, qi[0], qi[1]);
#endif
                if (t == 0) {
                    qi[idx] = node | kVis;
                    if (p.rec_ids && nrec < p.rec_cap) {
                        p.rec_ids[(size_t)qidx * p.rec_cap + nrec] = node;
                        p.rec_dists[(size_t)qidx * p.rec_cap + nrec] = qd[idx];
                    }
                }
                cursor_lo = idx + 1;
                ++nrec;
            }
            __syncwarp();

            // adjacency row: words [0, 96) as three 16-byte loads per lane; word 0 = degree
            uint32_t wd[3][4];
            {
                const uint4* row4 = reinterpret_cast<const uint4*>(p.adj + (size_t)node * p.adj_stride);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    uint4 v = make_uint4(kE, kE, kE, kE);
                    if (!done && (uint32_t)(k * 32 + t * 4) < p.adj_stride) v = __ldg(row4 + k * 8 + t);
                    wd[k][0] = v.x;
                    wd[k][1] = v.y;
                    wd[k][2] = v.z;
                    wd[k][3] = v.w;
                }
            }
            const uint32_t deg_word = __shfl_sync(kFull, wd[0][0], g << 3);  // all lanes take part
            const uint32_t deg = done ? 0 : min(deg_word, p.max_degree);

            // visited filter, adjacency order = (k, t, c); 4 probes in flight per lane per round
            uint32_t ncand = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                bool valid[4];
                uint32_t bk[4];
                uint4 lo4[4], hi4[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t w = (uint32_t)(k * 32 + t * 4 + c);
                    valid[c] = !done && w >= 1 && w <= deg;
                    bk[c] = team_bucket_of(wd[k][c], nbk);
                    if (valid[c]) {
                        const uint4* bp = reinterpret_cast<const uint4*>(table + (size_t)bk[c] * 8);
                        lo4[c] = __ldcg(bp);
                        hi4[c] = __ldcg(bp + 1);
                    }
                }
                uint32_t newbits = 0, insbits = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    bool inserted = false;
                    if (valid[c]) inserted = team_bucket_insert(table, nbk, bk[c], lo4[c], hi4[c], wd[k][c]);
                    if (inserted) insbits |= 1u << c;
                    if (inserted && wd[k][c] < n_total) newbits |= 1u << c;
                }
                // exclusive prefix of the per-lane new counts inside the team
                const uint32_t mine = __popc(newbits);
                uint32_t incl = mine;
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(kFull, incl, o, 8);
                    if (t >= o) incl += v;
                }
                uint32_t pos = ncand + incl - mine;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if ((newbits >> c) & 1u) cid[pos++] = wd[k][c];
                ncand += __shfl_sync(kFull, incl, (g << 3) | 7);
                uint32_t ins = __popc(insbits);
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) ins += __shfl_xor_sync(kFull, ins, o);
                nvisited += ins;
            }
            if (!done && nvisited + p.max_degree > hlimit) {
                overflow = true;
                done = true;
                ncand = 0;
            }
            __syncwarp();

            distances(ncand);
            merge(ncand, size, cursor_lo);
            if (!done) {
                cmps += ncand;
                hops += 1;
            }
        }

        if (active && overflow) {
            if (t == 0) {
                const uint32_t o = atomicAdd(p.counters + 1, 1u);
                p.overflow_list[o] = qidx;
            }
        }
        // ---- post-process: drop start points, first k (provider.rs:907-950)
        {
            const bool emit = active && !overflow;
            const uint32_t n = emit ? min(cap, size) : 0;
            uint32_t count = 0;
            const int steps = __reduce_max_sync(kFull, (int)((n + 7) >> 3));
            for (int s = 0; s < steps; ++s) {
                const uint32_t e = (uint32_t)s * 8 + t;
                const uint32_t id = e < n ? (qi[e] & ~kVis) : kE;
                const bool keep = e < n && id < p.n_points;
                const uint32_t m = team_ballot(keep, g);
                const uint32_t pos = count + __popc(m & ((1u << t) - 1u));
                if (keep && pos < p.k) {
                    p.out_ids[(size_t)qidx * p.k + pos] = id;
                    p.out_dists[(size_t)qidx * p.k + pos] = qd[e];
                }
                count += __popc(m);
            }
            if (emit) {
                count = min(count, p.k);
                for (uint32_t i = count + t; i < p.k; i += 8) {
                    p.out_ids[(size_t)qidx * p.k + i] = kE;
                    p.out_dists[(size_t)qidx * p.k + i] = __int_as_float(0x7F800000);
                }
                if (t == 0) {
                    atomicMax(p.counters + 2, nvisited);
                    if (p.out_counts) p.out_counts[qidx] = count;
                    if (p.out_cmps) p.out_cmps[qidx] = cmps;
                    if (p.out_hops) p.out_hops[qidx] = hops;
                    if (p.rec_counts) p.rec_counts[qidx] = min(nrec, p.rec_cap);
                }
            }
        }
    }
}

// ------------------------------------------------------------------ host side
struct TeamLaunch {
    void (*kern)(const SearchParamsTeam);
    size_t smem_block;
    int grid;  // CTAs
};

// 0: covered (out/p filled); 1: not covered -> caller uses v2 / generic
int team_prepare(const dab_index* idx, uint32_t l_search, uint32_t beam, SearchParamsTeam& p, TeamLaunch& out) {
    if (getenv("DAB_DISABLE_TEAM")) return 1;
    if (idx->dtype != DAB_F32 && idx->dtype != DAB_F16) return 1;
    const MetricPlan plan = plan_for(idx->metric, false);
    if (plan.kind == KIND_COS) return 1;
    const uint32_t cap = l_search + idx->n_start;
    if (beam != 1 || cap > 128 || idx->max_degree > 95 || idx->n_start > 8 || idx->adj_stride > 96) return 1;
    size_t off = 0;
    p.off_q = (uint32_t)off;
    off += round_up((size_t)idx->dim * 4, 16);
    const size_t cap_pad = round_up(cap, 8) + 8;
    p.off_qd = (uint32_t)off;
    off += cap_pad * 4;
    p.off_qi = (uint32_t)off;
    off += cap_pad * 4;
    const size_t maxc = std::max<size_t>(96, idx->n_start);
    p.off_cid = (uint32_t)off;
    off += maxc * 4;
    p.off_cd = (uint32_t)off;
    off += maxc * 4;
    p.team_smem = (uint32_t)round_up(off, 16);
    out.smem_block = (size_t)p.team_smem * kTeamsPerWarp * kTeamWarps;
    if (out.smem_block > 200 * 1024) return 1;
#define PICKT(TD)                                                                          \
    do {                                                                                   \
        if (plan.kind == KIND_L2) out.kern = search_kernel_team<TD, KIND_L2, POST_ID>;      \
        else if (plan.post == POST_NEG) out.kern = search_kernel_team<TD, KIND_IP, POST_NEG>; \
        else out.kern = search_kernel_team<TD, KIND_IP, POST_ONE_MINUS>;                    \
    } while (0)
    if (idx->dtype == DAB_F32) PICKT(float);
    else PICKT(__half);
#undef PICKT
    if (cudaFuncSetAttribute(out.kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)out.smem_block) != cudaSuccess) {
        cudaGetLastError();
        return 1;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, out.kern, kTeamWarps * 32, out.smem_block) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        return 1;
    }
    if (const char* e = getenv("DAB_TEAM_CTAS_PER_SM")) {
        const int v = atoi(e);
        if (v >= 1 && v < per_sm) per_sm = v;
    }
    out.grid = per_sm * idx->sm_count;
    return 0;
}

}  // namespace dab
