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
