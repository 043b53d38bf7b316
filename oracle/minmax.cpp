// minmax.cpp — CPU restatement of the MinMax quantizer (test infrastructure only, see oracle.h).
//
// Follows diskann-quantization/src/minmax/{quantizer.rs, vectors.rs}: per-vector N-bit quantization with the
// compensation coefficients stored in front of the codes, and the distances between two compressed vectors.
// Transform::Null only (the Hadamard / random-rotation transforms of algorithms/transforms are outside the path).
// Everything here is scalar and sequential in the reference, so the order below IS the reference's order.
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "oracle.h"

namespace {

// f32::min / f32::max: the other operand when one is NaN
inline float rmin(float a, float b) { return std::fmin(a, b); }
inline float rmax(float a, float b) { return std::fmax(a, b); }

// MinMaxQuantizer::get_range (quantizer.rs:117-151)
void get_range(const float* v, size_t dim, int nbits, float grid_scale, float* lo, float* hi) {
    float mn, mx;
    if (nbits == 1) {
        // values below / not below the mean, each side's average clamped towards the mean
        float sum = -0.0f;  // <f32 as Sum>::sum folds from -0.0
        for (size_t i = 0; i < dim; ++i) sum = sum + v[i];
        const float mean = sum / (float)dim;
        float a = 0.0f, ac = 0.0f, b = 0.0f, bc = 0.0f;
        for (size_t i = 0; i < dim; ++i) {
            const float m = v[i] < mean ? 1.0f : 0.0f;
            a += m * v[i];
            ac += m;
            b += (1.0f - m) * v[i];
            bc += 1.0f - m;
        }
        mn = rmin(a / ac, mean);
        mx = rmax(b / bc, mean);
    } else {
        mn = mx = std::numeric_limits<float>::quiet_NaN();
        for (size_t i = 0; i < dim; ++i) {
            mn = rmin(mn, v[i]);
            mx = rmax(mx, v[i]);
        }
    }
    const float width = (mx - mn) / 2.0f;
    const float mid = mn + width;
    *lo = mid - width * grid_scale;
    *hi = mid + width * grid_scale;
}

inline uint32_t code_at(const uint8_t* codes, size_t i, int nbits) {
    const size_t bit = i * (size_t)nbits;
    return (uint32_t)(codes[bit / 8] >> (bit % 8)) & ((1u << nbits) - 1u);
}

}  // namespace

extern "C" {

size_t orc_minmax_row_bytes(size_t dim, int nbits) { return 20 + (dim * (size_t)nbits + 7) / 8; }

// MinMaxQuantizer::compress (quantizer.rs:153-228) into the canonical-front layout of Data<NBITS>
// (meta/vector.rs:377-392): MinMaxCompensation {dim: u32, b, n, a, norm_squared} (vectors.rs:43-52, 20 bytes), then
// the dense codes, value i at bit i * nbits (bits/slice.rs:261-305).  Returns 1 when the input holds a NaN
// (InputContainsNaN — the row is written all the same, like the reference's set_meta before the error); *loss = the
// squared reconstruction error (L2Loss).
int orc_minmax_compress(float grid_scale, size_t dim, int nbits, const float* v, uint8_t* row, float* loss_out) {
    float lo, hi;
    get_range(v, dim, nbits, grid_scale, &lo, &hi);
    const float domain_max = (float)((1u << nbits) - 1u);
    const float inverse_scale = rmax(hi - lo, 1e-8f) / domain_max;
    float norm_squared = 0.0f, code_sum = 0.0f, loss = 0.0f;
    bool nan = false;
    const size_t bytes = (dim * (size_t)nbits + 7) / 8;
    uint8_t* codes = row + 20;
    memset(codes, 0, bytes);
    for (size_t i = 0; i < dim; ++i) {
        nan |= v[i] != v[i];
        float t = (v[i] - lo) / inverse_scale;
        // f32::clamp keeps NaN; round = half away from zero
        float code = t != t ? t : (t < 0.0f ? 0.0f : (t > domain_max ? domain_max : t));
        code = std::round(code);
        const float vr = (code * inverse_scale) + lo;
        norm_squared += vr * vr;
        code_sum += code;
        const float e = vr - v[i];
        loss += e * e;
        const uint32_t c = code != code ? 0u : (uint32_t)code;  // `as u8`: NaN -> 0
        const size_t bit = i * (size_t)nbits;
        codes[bit / 8] |= (uint8_t)(c << (bit % 8));
    }
    const uint32_t d32 = (uint32_t)dim;
    const float n = inverse_scale * code_sum;
    memcpy(row, &d32, 4);
    memcpy(row + 4, &lo, 4);
    memcpy(row + 8, &n, 4);
    memcpy(row + 12, &inverse_scale, 4);
    memcpy(row + 16, &norm_squared, 4);
    if (loss_out) *loss_out = loss;
    return nan ? 1 : 0;
}

// CompressInto<&[T], FullQueryMut> (quantizer.rs:369-417): the query stays f32; meta = {sum, norm_squared}
// (FullQueryMeta, vectors.rs:180-186), both sequential sums.  Returns 1 on NaN input.
int orc_minmax_full_query_meta(const float* v, size_t dim, float* sum_out, float* norm_squared_out) {
    for (size_t i = 0; i < dim; ++i)
        if (v[i] != v[i]) return 1;
    float ns = -0.0f, s = -0.0f;
    for (size_t i = 0; i < dim; ++i) ns = ns + v[i] * v[i];
    for (size_t i = 0; i < dim; ++i) s = s + v[i];
    *sum_out = s;
    *norm_squared_out = ns;
    return 0;
}

// Distances between two compressed vectors (vectors.rs:206-228 `kernel`, :231-262 MinMaxIP, :308-345 MinMaxL2Squared,
// :395-415 MinMaxCosine, :437-455 MinMaxCosineNormalized), similarity-score convention (IP negated).  x has nbits_x
// bits per value, y nbits_y (the reference instantiates N x N and 8 x N); the integer inner product is exact.
float orc_minmax_distance(int metric, int nbits_x, int nbits_y, const uint8_t* x_row, const uint8_t* y_row) {
    uint32_t dx, dy;
    float xb, xn, xa, xnorm, yb, yn, ya, ynorm;
    memcpy(&dx, x_row, 4);
    memcpy(&xb, x_row + 4, 4);
    memcpy(&xn, x_row + 8, 4);
    memcpy(&xa, x_row + 12, 4);
    memcpy(&xnorm, x_row + 16, 4);
    memcpy(&dy, y_row, 4);
    memcpy(&yb, y_row + 4, 4);
    memcpy(&yn, y_row + 8, 4);
    memcpy(&ya, y_row + 12, 4);
    memcpy(&ynorm, y_row + 16, 4);
    if (dx != dy) return std::numeric_limits<float>::quiet_NaN();  // UnequalLengths
    uint32_t ip = 0;
    for (size_t i = 0; i < dx; ++i) ip += code_at(x_row + 20, i, nbits_x) * code_at(y_row + 20, i, nbits_y);
    const float term0 = xa * ya * (float)ip;
    const float term1_x = xn * yb;
    const float term1_y = yn * xb;
    const float term2 = xb * yb * (float)dx;
    const float v = term0 + term1_x + term1_y + term2;
    switch (metric) {
        case ORC_INNER_PRODUCT: return -v;
        case ORC_L2: return -2.0f * v + xnorm + ynorm;
        case ORC_COSINE: return 1.0f - v / (std::sqrt(xnorm) * std::sqrt(ynorm));
        default: return 1.0f - v;  // CosineNormalized
    }
}

// DataRef::decompress_into (vectors.rs:120-140): x_i = code_i * a + b
void orc_minmax_decompress(const uint8_t* row, int nbits, float* out) {
    uint32_t d;
    float b, a;
    memcpy(&d, row, 4);
    memcpy(&b, row + 4, 4);
    memcpy(&a, row + 12, 4);
    for (size_t i = 0; i < d; ++i) out[i] = (float)code_at(row + 20, i, nbits) * a + b;
}

}  // extern "C"

// ---------------------------------------------------------------- full-precision query x compressed vector
// InnerProduct::evaluate(&[f32], BitSlice<NBITS>) (bits/distances.rs): the x86-64-v3 kernels for 1 / 2 / 4 bits
// (:2295-2436, :2438-2595, :2603-2665: eight f32 lanes, FMA, one or two accumulators, zero-filled remainder loads,
// sum_tree) and the scalar loop the 8-bit instantiation retargets to (:2668-2725).  Each SIMD lane is emulated as
// its own sequential chain.
namespace {

inline float tree8(const float (&s)[8]) {  // diskann-wide/src/traits.rs:583-595
    return ((s[0] + s[4]) + (s[2] + s[6])) + ((s[1] + s[5]) + (s[3] + s[7]));
}

// the little-endian value of the first `nbytes` (<= 4) bytes at p (load_one .. load_four, distances.rs:180-234)
inline uint32_t load_bytes(const uint8_t* p, size_t nbytes) {
    uint32_t v = 0;
    for (size_t i = 0; i < nbytes; ++i) v |= (uint32_t)p[i] << (8 * i);
    return v;
}

float full_ip(const float* x, const uint8_t* codes, size_t len, int nbits) {
    if (nbits == 8) {
        float s = 0.0f;
        for (size_t i = 0; i < len; ++i) s += x[i] * (float)codes[i];
        return s;
    }
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const size_t tail = len % 8 == 0 ? 8 : len % 8;
    if (nbits == 4) {
        const size_t blocks = len / 8;
        for (size_t b = 0; b < blocks; ++b) {
            const uint32_t w = load_bytes(codes + 4 * b, 4);
            for (int l = 0; l < 8; ++l) s[l] = std::fmaf(x[8 * b + l], (float)((w >> (4 * l)) & 15u), s[l]);
        }
        const size_t rem = len % 8;
        if (rem) {
            const uint32_t w = load_bytes(codes + 4 * blocks, (rem + 1) / 2);
            for (int l = 0; l < 8; ++l) s[l] = std::fmaf((size_t)l < rem ? x[8 * blocks + l] : 0.0f, (float)((w >> (4 * l)) & 15u), s[l]);
        }
        return tree8(s);
    }
    if (nbits == 2) {
        const size_t blocks = len / 16;
        if (blocks) {
            float s0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (size_t b = 0; b < blocks; ++b) {
                const uint32_t w = load_bytes(codes + 4 * b, 4);
                for (int l = 0; l < 8; ++l) {
                    s0[l] = std::fmaf(x[16 * b + l], (float)((w >> (2 * l)) & 3u), s0[l]);
                    s1[l] = std::fmaf(x[16 * b + 8 + l], (float)((w >> (16 + 2 * l)) & 3u), s1[l]);
                }
            }
            for (int l = 0; l < 8; ++l) s[l] = s0[l] + s1[l];
        }
        const size_t rem = len % 16;
        if (rem) {
            const uint32_t w = load_bytes(codes + 4 * blocks, (rem + 3) / 4);
            const float* px = x + 16 * blocks;
            if (rem <= 8) {
                for (int l = 0; l < 8; ++l) s[l] = std::fmaf((size_t)l < tail ? px[l] : 0.0f, (float)((w >> (2 * l)) & 3u), s[l]);
            } else {
                for (int l = 0; l < 8; ++l) s[l] = std::fmaf(px[l], (float)((w >> (2 * l)) & 3u), s[l]);
                for (int l = 0; l < 8; ++l) s[l] = std::fmaf((size_t)l < tail ? px[8 + l] : 0.0f, (float)((w >> (16 + 2 * l)) & 3u), s[l]);
            }
        }
        return tree8(s);
    }
    // one bit
    const size_t blocks = len / 32;
    if (blocks) {
        float s0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (size_t b = 0; b < blocks; ++b) {
            const uint32_t w = load_bytes(codes + 4 * b, 4);
            for (int l = 0; l < 8; ++l) {
                s0[l] = std::fmaf(x[32 * b + l], (float)((w >> l) & 1u), s0[l]);
                s1[l] = std::fmaf(x[32 * b + 8 + l], (float)((w >> (8 + l)) & 1u), s1[l]);
                s0[l] = std::fmaf(x[32 * b + 16 + l], (float)((w >> (16 + l)) & 1u), s0[l]);
                s1[l] = std::fmaf(x[32 * b + 24 + l], (float)((w >> (24 + l)) & 1u), s1[l]);
            }
        }
        for (int l = 0; l < 8; ++l) s[l] = s0[l] + s1[l];
    }
    const size_t rem = len % 32;
    if (rem) {
        const size_t groups = (rem + 7) / 8;
        const uint32_t w = load_bytes(codes + 4 * blocks, groups);
        const float* px = x + 32 * blocks;
        for (size_t j = 0; j < groups; ++j)
            for (int l = 0; l < 8; ++l) {
                const bool in = j + 1 < groups || (size_t)l < tail;
                s[l] = std::fmaf(in ? px[8 * j + l] : 0.0f, (float)((w >> (8 * j + l)) & 1u), s[l]);
            }
    }
    return tree8(s);
}

}  // namespace

extern "C" {

// MinMax{IP, L2Squared, Cosine, CosineNormalized}::evaluate(FullQueryRef, DataRef<NBITS>) (vectors.rs:272-305, 347-392,
// 417-436, 457-476), similarity-score convention.  q_sum / q_norm_squared = FullQueryMeta of the query.
float orc_minmax_query_distance(int metric, int nbits, const float* query, float q_sum, float q_norm_squared, const uint8_t* row) {
    uint32_t d;
    float yb, ya, ynorm;
    memcpy(&d, row, 4);
    memcpy(&yb, row + 4, 4);
    memcpy(&ya, row + 12, 4);
    memcpy(&ynorm, row + 16, 4);
    const float raw = full_ip(query, row + 20, d, nbits);
    const float ip = raw * ya + q_sum * yb;
    switch (metric) {
        case ORC_INNER_PRODUCT: return -ip;
        case ORC_L2: return q_norm_squared + ynorm - 2.0f * ip;
        case ORC_COSINE: return 1.0f - ip / (std::sqrt(q_norm_squared) * std::sqrt(ynorm));
        default: return 1.0f - ip;
    }
}

}  // extern "C"
