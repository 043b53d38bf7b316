// oracle/kmeans.cpp — CPU restatement of the PQ codebook training the reference benchmark runs
// before a quantized build: per chunk k-means++ seeding followed by Lloyd iterations.
// TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// Follows, with the arithmetic in the reference's order:
//   * diskann-providers/src/index/diskann_async.rs:61-89 (train_pq: 256 centres, 5 Lloyd reps, no centring),
//     model/pq/pq_construction.rs:163-243, diskann-quantization/src/product/train.rs (per-chunk thunk);
//   * algorithms/kmeans/common.rs (square_norm: four 8-lane accumulators for full 32-blocks combined
//     (s0+s1)+(s2+s3), then 8-blocks, a zero-filled remainder, sum_tree);
//   * algorithms/kmeans/plusplus.rs:238-320, 381-498 (update_distances: d = (norm_i + norm_c) + (-2 * dot),
//     dot = FMA chain over the dimensions; running minimum with `<`; the block sums enter an f64 rolling
//     sum; selection: first i with rolling >= threshold, d_i > 0, not yet picked);
//   * algorithms/kmeans/lloyds.rs:27-330, 345-366, 372-426 (assignment by n_c - s - s + n_i with first minimum in
//     centre order, centroid = f64 sum in data order / max(count, 1)).
// The reference draws from Rust's StdRng (ChaCha12), which is not restated: the random choices
// come from SplitMix64 streams (one per chunk, seeded `seed + chunk`), documented in DESIGN.md.
// "parity unpinned" for the random draws; the deterministic arithmetic above is what the GPU
// kernels are compared with.

#include "oracle.h"

#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace {

struct SplitMix64 {
    uint64_t s;
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    uint64_t below(uint64_t n) { return (uint64_t)(((unsigned __int128)next() * n) >> 64); }
    double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

float tree8(const float v[8]) { return ((v[0] + v[4]) + (v[2] + v[6])) + ((v[1] + v[5]) + (v[3] + v[7])); }

// common.rs square_norm over a strided chunk view (len = chunk dims)
float square_norm(const float* x, size_t len) {
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    size_t i = 0;
    if (i + 32 <= len) {
        float a[4][8];
        std::memset(a, 0, sizeof(a));
        while (i + 32 <= len) {
            for (int k = 0; k < 4; ++k)
                for (int l = 0; l < 8; ++l) a[k][l] = std::fmaf(x[i + 8 * k + l], x[i + 8 * k + l], a[k][l]);
            i += 32;
        }
        for (int l = 0; l < 8; ++l) s[l] = (a[0][l] + a[1][l]) + (a[2][l] + a[3][l]);
    }
    while (i + 8 <= len) {
        for (int l = 0; l < 8; ++l) s[l] = std::fmaf(x[i + l], x[i + l], s[l]);
        i += 8;
    }
    const size_t rem = len - i;
    if (rem) {
        for (size_t l = 0; l < 8; ++l) {
            const float v = l < rem ? x[i + l] : 0.0f;
            s[l] = std::fmaf(v, v, s[l]);
        }
    }
    return tree8(s);
}

float dot_chain(const float* row, const float* c, size_t len) {  // s = fma(c[d], x[d], s) over d
    float s = 0.0f;
    for (size_t d = 0; d < len; ++d) s = std::fmaf(row[d], c[d], s);
    return s;
}

}  // namespace

extern "C" {

// lloyds_inner (lloyds.rs:372-426) over the chunk [lo, lo + len) of the rows of `data`: `reps` rounds of assignment
// (distances_in_place: d = ((n_c - s) - s) + n_i with s the FMA chain over the dimensions, first minimum in centre order)
// and update_centroids (:345-366: f64 sums in data order / max(count, 1)); centre norms refreshed between rounds.
// *residual (optional) = the sum of the winning distances of the last round.
static void run_lloyd(const float* data, uint64_t n, uint32_t dim, size_t lo, size_t len, float* pivots, uint32_t n_centers,
                      uint32_t lloyds_reps, const std::vector<float>& norms, std::vector<float>& cn, std::vector<uint32_t>& assign,
                      float* residual) {
    auto row = [&](uint64_t i) { return data + i * dim + lo; };
    auto center = [&](uint32_t p) { return pivots + (size_t)p * dim + lo; };
    for (uint32_t p = 0; p < n_centers; ++p) cn[p] = square_norm(center(p), len);
    std::vector<double> sums((size_t)n_centers * len);
    std::vector<uint32_t> counts(n_centers);
    for (uint32_t rep = 0; rep < lloyds_reps; ++rep) {
        float res = 0.0f;
        for (uint64_t i = 0; i < n; ++i) {
            float best = std::numeric_limits<float>::infinity();
            uint32_t arg = 0xFFFFFFFFu;
            for (uint32_t p = 0; p < n_centers; ++p) {
                const float sdot = dot_chain(center(p), row(i), len);  // c.mul_add(d, s): same product, same chain
                const float d = ((cn[p] - sdot) - sdot) + norms[i];
                if (d < best) {
                    best = d;
                    arg = p;
                }
            }
            assign[i] = arg;
            res += best;  // (the reference sums its lanes in SIMD order; only used where every term is exact)
        }
        if (residual) *residual = res;
        std::fill(sums.begin(), sums.end(), 0.0);
        std::fill(counts.begin(), counts.end(), 0u);
        for (uint64_t i = 0; i < n; ++i) {
            const uint32_t a = assign[i];
            if (a == 0xFFFFFFFFu) continue;  // all-NaN row: the reference would index out of bounds
            ++counts[a];
            for (size_t d = 0; d < len; ++d) sums[(size_t)a * len + d] += (double)row(i)[d];
        }
        for (uint32_t p = 0; p < n_centers; ++p) {
            const double c = (double)std::max<uint32_t>(counts[p], 1);
            for (size_t d = 0; d < len; ++d) center(p)[d] = (float)(sums[(size_t)p * len + d] / c);
        }
        if (rep != lloyds_reps - 1)
            for (uint32_t p = 0; p < n_centers; ++p) cn[p] = square_norm(center(p), len);
    }
}

// data: [n][dim] f32 row-major.  pivots out: [n_centers][dim] (full_pivot_data layout: centre p's
// chunk c occupies columns offsets[c]..offsets[c+1]).  Returns 0, or -1 when a chunk could not be
// seeded with n_centers distinct points (KMeansPlusPlusError; the reference tolerates the
// numerically recoverable kinds and continues with the zero-filled rows, train.rs).
int orc_pq_train(const float* data, uint64_t n, uint32_t dim, uint32_t n_chunks, uint32_t n_centers,
                 uint32_t lloyds_reps, uint64_t seed, float* pivots, uint64_t* offsets) {
    orc_pq_chunk_offsets(dim, n_chunks, offsets);
    int status = 0;
    std::vector<float> norms(n), mins(n), cn(n_centers);
    std::vector<uint32_t> assign(n);
    std::vector<uint8_t> picked(n);
    for (uint32_t ch = 0; ch < n_chunks; ++ch) {
        const size_t lo = offsets[ch], len = offsets[ch + 1] - offsets[ch];
        auto row = [&](uint64_t i) { return data + i * dim + lo; };
        auto center = [&](uint32_t p) { return pivots + (size_t)p * dim + lo; };
        for (uint64_t i = 0; i < n; ++i) norms[i] = square_norm(row(i), len);
        for (uint32_t p = 0; p < n_centers; ++p) std::memset(center(p), 0, len * sizeof(float));
        SplitMix64 rng{seed + ch};
        // ---- k-means++ (plusplus.rs:381-498)
        std::fill(mins.begin(), mins.end(), std::numeric_limits<float>::infinity());
        std::fill(picked.begin(), picked.end(), 0);
        uint64_t first = n ? rng.below(n) : 0;
        if (n == 0) return -1;
        std::memcpy(center(0), row(first), len * sizeof(float));
        picked[first] = 1;
        float prev_norm = norms[first];
        uint32_t selected = 1;
        const uint32_t want = (uint32_t)std::min<uint64_t>(n_centers, n);
        for (uint32_t cur = 1; cur < want; ++cur) {
            const float* last = center(cur - 1);
            // update_distances: blocks of 16 rows; per block the 16 minima enter the f64 rolling sum
            // as sum_j (d0[j] + d1[j]) (lanes beyond the data contribute 0)
            double s = 0.0;
            for (uint64_t b = 0; b < n; b += 16) {
                double blk = 0.0;
                float cur_d[16];
                for (int l = 0; l < 16; ++l) {
                    const uint64_t i = b + l;
                    if (i < n) {
                        const float inter = dot_chain(row(i), last, len) * -2.0f;
                        const float d = (norms[i] + prev_norm) + inter;
                        if (d < mins[i]) mins[i] = d;
                        cur_d[l] = mins[i];
                    } else {
                        cur_d[l] = 0.0f;
                    }
                }
                for (int j = 0; j < 8; ++j) blk += (double)cur_d[j] + (double)cur_d[8 + j];
                s = s + blk;
            }
            if (!(s > 0.0)) break;  // Uniform::new(0, s) is empty: skip (and fail below)
            if (!std::isfinite(s)) return -1;
            const double threshold = rng.unit() * s;
            double rolling = 0.0;
            bool got = false;
            for (uint64_t i = 0; i < n; ++i) {
                rolling += (double)mins[i];
                if (rolling >= threshold && mins[i] > 0.0f && !picked[i]) {
                    std::memcpy(center(cur), row(i), len * sizeof(float));
                    picked[i] = 1;
                    prev_norm = norms[i];
                    selected = cur + 1;
                    got = true;
                    break;
                }
            }
            if (!got) break;
        }
        if (selected != n_centers) status = -1;
        // ---- Lloyd (lloyds.rs:372-426)
        run_lloyd(data, n, dim, lo, len, pivots, n_centers, lloyds_reps, norms, cn, assign, nullptr);
    }
    return status;
}

// lloyds(data, centers, max_reps) (lloyds.rs:428-460) over whole rows: centers [n_centers][dim] in / out, assignments [n],
// *loss = the residual of the last round.
void orc_lloyds(const float* data, uint64_t n, uint32_t dim, float* centers, uint32_t n_centers, uint32_t reps,
                uint32_t* assignments, float* loss) {
    std::vector<float> norms(n), cn(n_centers);
    std::vector<uint32_t> assign(n);
    for (uint64_t i = 0; i < n; ++i) norms[i] = square_norm(data + i * dim, dim);
    float res = 0.0f;
    run_lloyd(data, n, dim, 0, dim, centers, n_centers, reps, norms, cn, assign, &res);
    for (uint64_t i = 0; i < n; ++i) assignments[i] = assign[i];
    if (loss) *loss = res;
}

}  // extern "C"
