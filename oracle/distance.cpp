// oracle/distance.cpp — CPU restatement of the reference distance kernels.
// TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// Three flavours of every kernel:
//   * SIMD   : lane-exact emulation of the x86-64-v3 (AVX2) schemas of
//              diskann-vector/src/distance/simd.rs — same accumulator count, same element ->
//              (accumulator, lane) assignment, same combine and sum_tree order, FMA where the
//              reference uses FMA.  The reference's default target is x86-64-v3
//              (.cargo/config.toml:7-8) and the inmem query path runs the compile-time ARCH
//              (diskann-inmem/src/layers/full.rs:333), so V3 is the order pinned here.
//   * SCALAR : the naive folds of diskann-vector/src/distance/reference.rs.
//   * AVX2   : the same V3 order with real intrinsics (used as the timed CPU baseline);
//              tests assert AVX2 == SIMD bit for bit.
//
// Compile with -ffp-contract=off: every fused multiply-add below is explicit.

#include "oracle.h"

#include <cmath>
#include <cstring>
#include <limits>

#if defined(__AVX2__) && defined(__FMA__) && defined(__F16C__)
#include <immintrin.h>
#define ORC_HAVE_AVX2 1
#else
#define ORC_HAVE_AVX2 0
#endif

namespace {

// ------------------------------------------------------------------ f16
// half 2.6 / diskann-wide cast_f16_to_f32 (diskann-wide/src/reference.rs): IEEE binary16.
inline float f16_to_f32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu;
    uint32_t man = h & 0x3FFu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else {  // subnormal: normalise
            int e = -1;
            do {
                man <<= 1;
                ++e;
            } while ((man & 0x400u) == 0);
            man &= 0x3FFu;
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7F800000u | (man << 13);
    } else {
        bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
    }
    float f;
    std::memcpy(&f, &bits, 4);
    return f;
}

inline uint16_t f32_to_f16(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t man = x & 0x007FFFFFu;
    int32_t exp = (int32_t)((x >> 23) & 0xFFu);
    if (exp == 255) return (uint16_t)(sign | 0x7C00u | (man ? (0x200u | (man >> 13)) : 0u));
    int32_t e = exp - 127 + 15;
    if (e >= 31) return (uint16_t)(sign | 0x7C00u);
    if (e <= 0) {
        if (e < -10) return (uint16_t)sign;
        man |= 0x00800000u;
        uint32_t shift = (uint32_t)(14 - e);
        uint32_t half_man = man >> shift;
        uint32_t rem = man & ((1u << shift) - 1u);
        uint32_t halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (half_man & 1u))) ++half_man;
        return (uint16_t)(sign | half_man);
    }
    uint32_t half_man = man >> 13;
    uint32_t rem = man & 0x1FFFu;
    uint16_t out = (uint16_t)(sign | ((uint32_t)e << 10) | half_man);
    if (rem > 0x1000u || (rem == 0x1000u && (half_man & 1u))) ++out;  // carries into exponent
    return out;
}

// ------------------------------------------------------------------ element loaders
struct LoadF32 {
    const float* p;
    float operator()(size_t i) const { return p[i]; }
};
struct LoadF16 {
    const uint16_t* p;
    float operator()(size_t i) const { return f16_to_f32(p[i]); }
};

// ------------------------------------------------------------------ SIMD-order emulation
// simd.rs:686-747 (simd_op) with W = 8 lanes and the MainLoop strategies of simd.rs:245-483.
// For Strategy4x1 / 4x2 / 2x4 the k-th full 8-wide vector always lands in accumulator
// k % NA (NA = 4, 4, 2): see the load(block, offset) arithmetic at simd.rs:170-186 and the
// epilogue assignments at :352-362, :409-419, :468-480.  After the accumulators are combined
// ((s0+s1)+(s2+s3) or s0+s1) the masked remainder (len % 8, load_simd_first zero-fills) is
// accumulated on top (simd.rs:733-744) and the result reduced with sum_tree
// (diskann-wide/src/traits.rs:583-595): ((x0+x4)+(x2+x6)) + ((x1+x5)+(x3+x7)).
struct Vec8 {
    float v[8];
};
inline Vec8 vzero() {
    Vec8 r;
    for (int l = 0; l < 8; ++l) r.v[l] = 0.0f;
    return r;
}
inline Vec8 vadd(const Vec8& a, const Vec8& b) {
    Vec8 r;
    for (int l = 0; l < 8; ++l) r.v[l] = a.v[l] + b.v[l];
    return r;
}
inline float sum_tree(const Vec8& x) {
    float a0 = x.v[0] + x.v[4], a1 = x.v[1] + x.v[5], a2 = x.v[2] + x.v[6], a3 = x.v[3] + x.v[7];
    float b0 = a0 + a2, b1 = a1 + a3;
    return b0 + b1;
}

// One accumulate step of a schema on one lane.
struct OpL2 {  // simd.rs:833-836: c = x - y; c.mul_add(c, acc)
    static float step(float x, float y, float acc) {
        float c = x - y;
        return std::fmaf(c, c, acc);
    }
};
struct OpIP {  // simd.rs:1601-1608: x.mul_add(y, acc)
    static float step(float x, float y, float acc) { return std::fmaf(x, y, acc); }
};

// Runs main loop + epilogues + remainder and returns the combined 8-lane accumulator
// (what a Resumable schema receives in combine_with, simd.rs:637-671).
template <int NA, class Op, class LX, class LY>
Vec8 simd_accumulate(const LX& x, const LY& y, size_t len) {
    Vec8 s[NA];
    for (int a = 0; a < NA; ++a) s[a] = vzero();
    const size_t full = len / 8;
    for (size_t k = 0; k < full; ++k) {
        Vec8& acc = s[k % NA];
        for (int l = 0; l < 8; ++l) acc.v[l] = Op::step(x(8 * k + l), y(8 * k + l), acc.v[l]);
    }
    Vec8 c = (NA == 4) ? vadd(vadd(s[0], s[1]), vadd(s[2 % NA], s[3 % NA])) : vadd(s[0], s[1 % NA]);
    const size_t rem = len % 8;
    if (rem != 0) {
        for (int l = 0; l < 8; ++l) {
            float xv = (size_t)l < rem ? x(8 * full + l) : 0.0f;
            float yv = (size_t)l < rem ? y(8 * full + l) : 0.0f;
            c.v[l] = Op::step(xv, yv, c.v[l]);
        }
    }
    return c;
}

// FullCosineAccumulator, simd.rs:2281-2383; all float cosine schemas on V3 are Strategy2x4.
struct Cos3 {
    Vec8 nx, ny, xy;
};
template <class LX, class LY>
Cos3 cosine_accumulate(const LX& x, const LY& y, size_t len) {
    Cos3 s[2];
    for (int a = 0; a < 2; ++a) s[a].nx = s[a].ny = s[a].xy = vzero();
    const size_t full = len / 8;
    for (size_t k = 0; k < full; ++k) {
        Cos3& acc = s[k % 2];
        for (int l = 0; l < 8; ++l) {
            float xv = x(8 * k + l), yv = y(8 * k + l);
            acc.nx.v[l] = std::fmaf(xv, xv, acc.nx.v[l]);
            acc.ny.v[l] = std::fmaf(yv, yv, acc.ny.v[l]);
            acc.xy.v[l] = std::fmaf(xv, yv, acc.xy.v[l]);
        }
    }
    Cos3 c;
    c.nx = vadd(s[0].nx, s[1].nx);
    c.ny = vadd(s[0].ny, s[1].ny);
    c.xy = vadd(s[0].xy, s[1].xy);
    const size_t rem = len % 8;
    if (rem != 0) {
        for (int l = 0; l < 8; ++l) {
            float xv = (size_t)l < rem ? x(8 * full + l) : 0.0f;
            float yv = (size_t)l < rem ? y(8 * full + l) : 0.0f;
            c.nx.v[l] = std::fmaf(xv, xv, c.nx.v[l]);
            c.ny.v[l] = std::fmaf(yv, yv, c.ny.v[l]);
            c.xy.v[l] = std::fmaf(xv, yv, c.xy.v[l]);
        }
    }
    return c;
}

// FullCosineAccumulator::sum, simd.rs:2328-2364
inline float cosine_finish(float normx, float normy, float prod) {
    float denominator = std::sqrt(normx) * std::sqrt(normy);
    if (normx < std::numeric_limits<float>::min() || normy < std::numeric_limits<float>::min())
        return 0.0f;
    float v = prod / denominator;
    // (-1.0f32).max(1.0f32.min(v)): Rust min/max return the non-NaN operand
    float m = std::fmin(1.0f, v);
    return std::fmax(-1.0f, m);
}

enum Kind { K_L2, K_IP, K_COS };

// Mathematical value (pre post-op) in V3 order for float pairs.
template <class LX, class LY>
float simd_float(Kind kind, int na_l2ip, const LX& x, const LY& y, size_t len) {
    if (kind == K_COS) {
        Cos3 c = cosine_accumulate(x, y, len);
        return cosine_finish(sum_tree(c.nx), sum_tree(c.ny), sum_tree(c.xy));
    }
    Vec8 c;
    if (kind == K_L2)
        c = na_l2ip == 4 ? simd_accumulate<4, OpL2>(x, y, len) : simd_accumulate<2, OpL2>(x, y, len);
    else
        c = na_l2ip == 4 ? simd_accumulate<4, OpIP>(x, y, len) : simd_accumulate<2, OpIP>(x, y, len);
    return sum_tree(c);
}

// ------------------------------------------------------------------ scalar definitions
// reference.rs:67-115 (L2), :250-305 (IP), :340-437 (cosine)
template <class LX, class LY>
float scalar_float(Kind kind, const LX& x, const LY& y, size_t len) {
    if (kind == K_L2) {
        float acc = 0.0f;
        for (size_t i = 0; i < len; ++i) {
            float d = x(i) - y(i);
            acc = std::fmaf(d, d, acc);
        }
        return acc;
    }
    if (kind == K_IP) {
        float acc = 0.0f;
        for (size_t i = 0; i < len; ++i) acc = std::fmaf(x(i), y(i), acc);
        return acc;
    }
    float nx = 0.0f, ny = 0.0f, xy = 0.0f;
    for (size_t i = 0; i < len; ++i) {
        float a = x(i), b = y(i);
        nx = std::fmaf(a, a, nx);
        ny = std::fmaf(b, b, ny);
        xy = std::fmaf(a, b, xy);
    }
    if (nx < std::numeric_limits<float>::min() || ny < std::numeric_limits<float>::min()) return 0.0f;
    float v = xy / (std::sqrt(nx) * std::sqrt(ny));
    return std::fmax(-1.0f, std::fmin(1.0f, v));
}

// ------------------------------------------------------------------ integers (exact)
// simd.rs:1157-1225 / 1335-1400 (L2), 1913-2146 (IP), 2750-3035 (cosine); reference.rs:67-90.
// i32 accumulation is exact, so lane order is irrelevant; wrapping arithmetic like the SIMD.
template <class T>
float int_kernel(Kind kind, const T* x, const T* y, size_t len) {
    if (kind == K_L2) {
        uint32_t acc = 0;
        for (size_t i = 0; i < len; ++i) {
            int32_t d = (int32_t)x[i] - (int32_t)y[i];
            acc += (uint32_t)(d * d);
        }
        return (float)(int32_t)acc;
    }
    if (kind == K_IP) {
        uint32_t acc = 0;
        for (size_t i = 0; i < len; ++i) acc += (uint32_t)((int32_t)x[i] * (int32_t)y[i]);
        return (float)(int32_t)acc;
    }
    uint32_t nx = 0, ny = 0, xy = 0;
    for (size_t i = 0; i < len; ++i) {
        int32_t a = x[i], b = y[i];
        nx += (uint32_t)(a * a);
        ny += (uint32_t)(b * b);
        xy += (uint32_t)(a * b);
    }
    return cosine_finish((float)(int32_t)nx, (float)(int32_t)ny, (float)(int32_t)xy);
}

// ------------------------------------------------------------------ AVX2 (same order)
#if ORC_HAVE_AVX2
inline __m256 ld_f32(const float* p) { return _mm256_loadu_ps(p); }
inline __m256 ld_f16(const uint16_t* p) {
    return _mm256_cvtph_ps(_mm_loadu_si128((const __m128i*)p));
}
inline __m256 ld_first_f32(const float* p, size_t n) {
    alignas(32) float t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (size_t i = 0; i < n; ++i) t[i] = p[i];
    return _mm256_load_ps(t);
}
inline __m256 ld_first_f16(const uint16_t* p, size_t n) {
    alignas(16) uint16_t t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (size_t i = 0; i < n; ++i) t[i] = p[i];
    return _mm256_cvtph_ps(_mm_load_si128((const __m128i*)t));
}
inline float hsum_tree(__m256 x) {  // diskann-wide/src/arch/x86_64/v3/f32x8_.rs:187-212
    __m128 hi = _mm256_extractf128_ps(x, 1);
    __m128 lo = _mm256_castps256_ps128(x);
    __m128 q = _mm_add_ps(lo, hi);
    __m128 d = _mm_add_ps(q, _mm_movehl_ps(q, q));
    __m128 s = _mm_add_ss(d, _mm_shuffle_ps(d, d, 0x1));
    return _mm_cvtss_f32(s);
}
struct PF32 {
    const float* p;
    __m256 full(size_t k) const { return ld_f32(p + 8 * k); }
    __m256 first(size_t k, size_t n) const { return ld_first_f32(p + 8 * k, n); }
};
struct PF16 {
    const uint16_t* p;
    __m256 full(size_t k) const { return ld_f16(p + 8 * k); }
    __m256 first(size_t k, size_t n) const { return ld_first_f16(p + 8 * k, n); }
};

template <int NA, bool IS_L2, class PX, class PY>
float avx2_l2ip(const PX& x, const PY& y, size_t len) {
    __m256 s[4] = {_mm256_setzero_ps(), _mm256_setzero_ps(), _mm256_setzero_ps(),
                   _mm256_setzero_ps()};
    const size_t full = len / 8;
    size_t k = 0;
    if (NA == 4) {
        for (; k + 4 <= full; k += 4) {
            for (int a = 0; a < 4; ++a) {
                __m256 xv = x.full(k + a), yv = y.full(k + a);
                if (IS_L2) {
                    __m256 c = _mm256_sub_ps(xv, yv);
                    s[a] = _mm256_fmadd_ps(c, c, s[a]);
                } else {
                    s[a] = _mm256_fmadd_ps(xv, yv, s[a]);
                }
            }
        }
    }
    for (; k < full; ++k) {
        int a = (int)(k % NA);
        __m256 xv = x.full(k), yv = y.full(k);
        if (IS_L2) {
            __m256 c = _mm256_sub_ps(xv, yv);
            s[a] = _mm256_fmadd_ps(c, c, s[a]);
        } else {
            s[a] = _mm256_fmadd_ps(xv, yv, s[a]);
        }
    }
    __m256 c = NA == 4 ? _mm256_add_ps(_mm256_add_ps(s[0], s[1]), _mm256_add_ps(s[2], s[3]))
                       : _mm256_add_ps(s[0], s[1]);
    const size_t rem = len % 8;
    if (rem) {
        __m256 xv = x.first(full, rem), yv = y.first(full, rem);
        if (IS_L2) {
            __m256 d = _mm256_sub_ps(xv, yv);
            c = _mm256_fmadd_ps(d, d, c);
        } else {
            c = _mm256_fmadd_ps(xv, yv, c);
        }
    }
    return hsum_tree(c);
}

template <class PX, class PY>
float avx2_cos(const PX& x, const PY& y, size_t len) {
    __m256 nx[2] = {_mm256_setzero_ps(), _mm256_setzero_ps()};
    __m256 ny[2] = {_mm256_setzero_ps(), _mm256_setzero_ps()};
    __m256 xy[2] = {_mm256_setzero_ps(), _mm256_setzero_ps()};
    const size_t full = len / 8;
    for (size_t k = 0; k < full; ++k) {
        int a = (int)(k & 1);
        __m256 xv = x.full(k), yv = y.full(k);
        nx[a] = _mm256_fmadd_ps(xv, xv, nx[a]);
        ny[a] = _mm256_fmadd_ps(yv, yv, ny[a]);
        xy[a] = _mm256_fmadd_ps(xv, yv, xy[a]);
    }
    __m256 cnx = _mm256_add_ps(nx[0], nx[1]);
    __m256 cny = _mm256_add_ps(ny[0], ny[1]);
    __m256 cxy = _mm256_add_ps(xy[0], xy[1]);
    const size_t rem = len % 8;
    if (rem) {
        __m256 xv = x.first(full, rem), yv = y.first(full, rem);
        cnx = _mm256_fmadd_ps(xv, xv, cnx);
        cny = _mm256_fmadd_ps(yv, yv, cny);
        cxy = _mm256_fmadd_ps(xv, yv, cxy);
    }
    return cosine_finish(hsum_tree(cnx), hsum_tree(cny), hsum_tree(cxy));
}

// i8 / u8 via widen to i16 + _mm256_madd_epi16 (diskann-wide v3/i32x8_.rs:186-195)
template <bool SIGNED>
inline __m256i widen16(const void* p) {
    __m128i b = _mm_loadu_si128((const __m128i*)p);
    return SIGNED ? _mm256_cvtepi8_epi16(b) : _mm256_cvtepu8_epi16(b);
}
inline int32_t hsum_i32(__m256i v) {
    __m128i s = _mm_add_epi32(_mm256_castsi256_si128(v), _mm256_extracti128_si256(v, 1));
    s = _mm_add_epi32(s, _mm_shuffle_epi32(s, 0x4E));
    s = _mm_add_epi32(s, _mm_shuffle_epi32(s, 0xB1));
    return _mm_cvtsi128_si32(s);
}
template <bool SIGNED, class T>
float avx2_int(Kind kind, const T* x, const T* y, size_t len) {
    __m256i a0 = _mm256_setzero_si256(), a1 = a0, a2 = a0;
    size_t i = 0;
    for (; i + 16 <= len; i += 16) {
        __m256i xv = widen16<SIGNED>(x + i), yv = widen16<SIGNED>(y + i);
        if (kind == K_L2) {
            __m256i c = _mm256_sub_epi16(xv, yv);
            a0 = _mm256_add_epi32(a0, _mm256_madd_epi16(c, c));
        } else if (kind == K_IP) {
            a0 = _mm256_add_epi32(a0, _mm256_madd_epi16(xv, yv));
        } else {
            a0 = _mm256_add_epi32(a0, _mm256_madd_epi16(xv, xv));
            a1 = _mm256_add_epi32(a1, _mm256_madd_epi16(yv, yv));
            a2 = _mm256_add_epi32(a2, _mm256_madd_epi16(xv, yv));
        }
    }
    uint32_t r0 = (uint32_t)hsum_i32(a0), r1 = (uint32_t)hsum_i32(a1), r2 = (uint32_t)hsum_i32(a2);
    for (; i < len; ++i) {
        int32_t a = x[i], b = y[i];
        if (kind == K_L2) {
            int32_t d = a - b;
            r0 += (uint32_t)(d * d);
        } else if (kind == K_IP) {
            r0 += (uint32_t)(a * b);
        } else {
            r0 += (uint32_t)(a * a);
            r1 += (uint32_t)(b * b);
            r2 += (uint32_t)(a * b);
        }
    }
    if (kind == K_COS) return cosine_finish((float)(int32_t)r0, (float)(int32_t)r1, (float)(int32_t)r2);
    return (float)(int32_t)r0;
}
#endif  // ORC_HAVE_AVX2

// ------------------------------------------------------------------ dispatch
// Strategy per (types, schema) on V3 — simd.rs impl headers:
//   f32xf32 L2 :817 4x1, IP :1588 4x1 ; f16xf16 L2 :989 2x4, IP :1752 2x4 ;
//   f32xf16 L2 :1121 4x2, IP :1878 4x2 ; all float cosine 2x4 (:2430, :2591, :2716).
inline int na_for(int dx, int dy) { return (dx == ORC_F16 && dy == ORC_F16) ? 2 : 4; }

float mathematical(int flavour, int dx, int dy, Kind kind, const void* x, const void* y,
                   size_t len, int* err) {
    if (err) *err = 0;
    if (dx == ORC_I8 && dy == ORC_I8) {
#if ORC_HAVE_AVX2
        if (flavour == ORC_FLAVOUR_AVX2) return avx2_int<true>(kind, (const int8_t*)x, (const int8_t*)y, len);
#endif
        return int_kernel(kind, (const int8_t*)x, (const int8_t*)y, len);
    }
    if (dx == ORC_U8 && dy == ORC_U8) {
#if ORC_HAVE_AVX2
        if (flavour == ORC_FLAVOUR_AVX2) return avx2_int<false>(kind, (const uint8_t*)x, (const uint8_t*)y, len);
#endif
        return int_kernel(kind, (const uint8_t*)x, (const uint8_t*)y, len);
    }
    const bool ff = dx == ORC_F32 && dy == ORC_F32;
    const bool hh = dx == ORC_F16 && dy == ORC_F16;
    const bool fh = dx == ORC_F32 && dy == ORC_F16;
    if (!(ff || hh || fh)) {
        if (err) *err = 1;
        return std::numeric_limits<float>::quiet_NaN();
    }
    if (flavour == ORC_FLAVOUR_SCALAR) {
        if (ff) return scalar_float(kind, LoadF32{(const float*)x}, LoadF32{(const float*)y}, len);
        if (hh) return scalar_float(kind, LoadF16{(const uint16_t*)x}, LoadF16{(const uint16_t*)y}, len);
        return scalar_float(kind, LoadF32{(const float*)x}, LoadF16{(const uint16_t*)y}, len);
    }
#if ORC_HAVE_AVX2
    if (flavour == ORC_FLAVOUR_AVX2) {
        if (ff) {
            PF32 a{(const float*)x}, b{(const float*)y};
            if (kind == K_L2) return avx2_l2ip<4, true>(a, b, len);
            if (kind == K_IP) return avx2_l2ip<4, false>(a, b, len);
            return avx2_cos(a, b, len);
        }
        if (hh) {
            PF16 a{(const uint16_t*)x}, b{(const uint16_t*)y};
            if (kind == K_L2) return avx2_l2ip<2, true>(a, b, len);
            if (kind == K_IP) return avx2_l2ip<2, false>(a, b, len);
            return avx2_cos(a, b, len);
        }
        PF32 a{(const float*)x};
        PF16 b{(const uint16_t*)y};
        if (kind == K_L2) return avx2_l2ip<4, true>(a, b, len);
        if (kind == K_IP) return avx2_l2ip<4, false>(a, b, len);
        return avx2_cos(a, b, len);
    }
#endif
    const int na = na_for(dx, dy);
    if (ff) return simd_float(kind, na, LoadF32{(const float*)x}, LoadF32{(const float*)y}, len);
    if (hh) return simd_float(kind, na, LoadF16{(const uint16_t*)x}, LoadF16{(const uint16_t*)y}, len);
    return simd_float(kind, na, LoadF32{(const float*)x}, LoadF16{(const uint16_t*)y}, len);
}

}  // namespace

// Shared with pq.cpp: combined 8-lane accumulators for resumable f32xf32 schemas.
namespace orc_detail {
void l2_partial(const float* x, const float* y, size_t len, float acc[8]) {
    Vec8 c = simd_accumulate<4, OpL2>(LoadF32{x}, LoadF32{y}, len);
    for (int l = 0; l < 8; ++l) acc[l] = acc[l] + c.v[l];
}
void ip_partial(const float* x, const float* y, size_t len, float acc[8]) {
    Vec8 c = simd_accumulate<4, OpIP>(LoadF32{x}, LoadF32{y}, len);
    for (int l = 0; l < 8; ++l) acc[l] = acc[l] + c.v[l];
}
void cos_partial(const float* x, const float* y, size_t len, float nx[8], float ny[8], float xy[8]) {
    Cos3 c = cosine_accumulate(LoadF32{x}, LoadF32{y}, len);
    for (int l = 0; l < 8; ++l) {
        nx[l] = nx[l] + c.nx.v[l];
        ny[l] = ny[l] + c.ny.v[l];
        xy[l] = xy[l] + c.xy.v[l];
    }
}
float tree8(const float v[8]) {
    Vec8 t;
    for (int l = 0; l < 8; ++l) t.v[l] = v[l];
    return sum_tree(t);
}
float cos_finish(float nx, float ny, float xy) { return cosine_finish(nx, ny, xy); }
}  // namespace orc_detail

extern "C" {

float orc_f16_to_f32(uint16_t h) { return f16_to_f32(h); }
uint16_t orc_f32_to_f16(float f) { return f32_to_f16(f); }

// Post-ops: implementations.rs:217-404; integer CosineNormalized == Cosine
// (distance_provider.rs:275-297).
float orc_distance(int flavour, int dx, int dy, int metric, const void* x, const void* y,
                   size_t dim, int* err) {
    const bool is_int = dx == ORC_I8 || dx == ORC_U8;
    switch (metric) {
        case ORC_L2:
            return mathematical(flavour, dx, dy, K_L2, x, y, dim, err);
        case ORC_INNER_PRODUCT:
            return -mathematical(flavour, dx, dy, K_IP, x, y, dim, err);
        case ORC_COSINE:
            return 1.0f - mathematical(flavour, dx, dy, K_COS, x, y, dim, err);
        case ORC_COSINE_NORMALIZED:
            if (is_int) return 1.0f - mathematical(flavour, dx, dy, K_COS, x, y, dim, err);
            return 1.0f - mathematical(flavour, dx, dy, K_IP, x, y, dim, err);
        default:
            if (err) *err = 2;
            return std::numeric_limits<float>::quiet_NaN();
    }
}

void orc_distance_rows(int flavour, int dq, int dr, int metric, const void* query,
                       const void* rows, size_t row_stride_bytes, size_t n, size_t dim,
                       float* out) {
    const char* p = (const char*)rows;
    for (size_t i = 0; i < n; ++i)
        out[i] = orc_distance(flavour, dq, dr, metric, query, p + i * row_stride_bytes, dim, nullptr);
}

}  // extern "C"
