// oracle/pq.cpp — CPU restatement of the product-quantization and scalar-quantization
// distance paths.  TEST INFRASTRUCTURE ONLY (see oracle.h).

#include "oracle.h"

#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace orc_detail {
void l2_partial(const float* x, const float* y, size_t len, float acc[8]);
void ip_partial(const float* x, const float* y, size_t len, float acc[8]);
void cos_partial(const float* x, const float* y, size_t len, float nx[8], float ny[8], float xy[8]);
float tree8(const float v[8]);
float cos_finish(float nx, float ny, float xy);
}  // namespace orc_detail

extern "C" {

// diskann-quantization/src/views.rs:226-243: the first dim % n_chunks chunks get one extra.
void orc_pq_chunk_offsets(size_t dim, size_t n_chunks, uint64_t* offsets) {
    size_t base = dim / n_chunks, extra = dim % n_chunks, pos = 0;
    offsets[0] = 0;
    for (size_t c = 0; c < n_chunks; ++c) {
        pos += base + (c < extra ? 1 : 0);
        offsets[c + 1] = pos;
    }
}

// fixed_chunk_pq_table.rs:152-187.  T::evaluate is SquaredL2 (populate_chunk_distances,
// :194-203) or InnerProduct (:209-218) whose f32 return applies the negating post-op
// (implementations.rs:309-314), i.e. IP tables hold -dot per chunk.
void orc_pq_populate_lut(const float* pivots, size_t n_centers, size_t dim,
                         const uint64_t* offsets, size_t n_chunks, int metric,
                         const float* query, float* lut) {
    for (size_t p = 0; p < n_centers; ++p) {
        const float* row = pivots + p * dim;
        for (size_t c = 0; c < n_chunks; ++c) {
            size_t start = offsets[c], stop = offsets[c + 1];
            lut[c * n_centers + p] =
                orc_distance(ORC_FLAVOUR_SIMD, ORC_F32, ORC_F32,
                             metric == ORC_INNER_PRODUCT ? ORC_INNER_PRODUCT : ORC_L2,
                             query + start, row + start, stop - start, nullptr);
        }
    }
}

// fixed_chunk_pq_table.rs:82-98: sequential f32 adds in chunk order starting from 0.0
float orc_pq_lookup(const uint8_t* code, size_t n_chunks, const float* lut, size_t n_centers) {
    float accum = 0.0f;
    for (size_t c = 0; c < n_chunks; ++c) accum += lut[c * n_centers + code[c]];
    return accum;
}

// direct_distance_impl, fixed_chunk_pq_table.rs:35-59: one Resumable accumulator, each chunk's
// simd_op result (the combined 8-lane accumulator) is added lane-wise, sum_tree at the end.
static float direct(const float* pivots, size_t dim, const uint64_t* offsets, size_t n_chunks,
                    int kind /*0 l2, 1 ip, 2 cos*/, const float* left_full, const uint8_t* left_code,
                    const uint8_t* right_code) {
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, b[8] = {0, 0, 0, 0, 0, 0, 0, 0},
          c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (size_t ch = 0; ch < n_chunks; ++ch) {
        size_t start = offsets[ch], stop = offsets[ch + 1];
        const float* l = left_full ? left_full + start : pivots + dim * left_code[ch] + start;
        const float* r = pivots + dim * right_code[ch] + start;
        if (kind == 0)
            orc_detail::l2_partial(l, r, stop - start, a);
        else if (kind == 1)
            orc_detail::ip_partial(l, r, stop - start, a);
        else
            orc_detail::cos_partial(l, r, stop - start, a, b, c);
    }
    if (kind == 2) return orc_detail::cos_finish(orc_detail::tree8(a), orc_detail::tree8(b), orc_detail::tree8(c));
    return orc_detail::tree8(a);
}

// fixed_chunk_pq_table.rs:223-281 (l2_distance, cosine_distance, cosine_normalized_distance
// == cosine_distance, inner_product == -raw); VTable mapping pq/distance/dynamic.rs:117-124.
float orc_pq_direct_distance(const float* pivots, size_t dim, const uint64_t* offsets,
                             size_t n_chunks, int metric, const float* query,
                             const uint8_t* code) {
    switch (metric) {
        case ORC_L2:
            return direct(pivots, dim, offsets, n_chunks, 0, query, nullptr, code);
        case ORC_INNER_PRODUCT:
            return -direct(pivots, dim, offsets, n_chunks, 1, query, nullptr, code);
        default:
            return 1.0f - direct(pivots, dim, offsets, n_chunks, 2, query, nullptr, code);
    }
}

// fixed_chunk_pq_table.rs:285-361; VTable mapping dynamic.rs:126-131 (CosineNormalized ->
// qq_cosine_distance).
float orc_pq_self_distance(const float* pivots, size_t dim, const uint64_t* offsets,
                           size_t n_chunks, int metric, const uint8_t* left,
                           const uint8_t* right) {
    switch (metric) {
        case ORC_L2:
            return direct(pivots, dim, offsets, n_chunks, 0, nullptr, left, right);
        case ORC_INNER_PRODUCT:
            return -direct(pivots, dim, offsets, n_chunks, 1, nullptr, left, right);
        default:
            return 1.0f - direct(pivots, dim, offsets, n_chunks, 2, nullptr, left, right);
    }
}

// QueryComputer::new, pq/distance/dynamic.rs:63-87
void orc_pq_query_distances(const float* pivots, size_t n_centers, size_t dim,
                            const uint64_t* offsets, size_t n_chunks, int metric,
                            const float* query, const uint8_t* codes, size_t n, float* out) {
    if (metric == ORC_COSINE) {  // DirectCosine, pq/distance/cosine.rs:16-70
        for (size_t i = 0; i < n; ++i)
            out[i] = orc_pq_direct_distance(pivots, dim, offsets, n_chunks, ORC_COSINE, query,
                                            codes + i * n_chunks);
        return;
    }
    std::vector<float> lut(n_chunks * n_centers);
    orc_pq_populate_lut(pivots, n_centers, dim, offsets, n_chunks,
                        metric == ORC_INNER_PRODUCT ? ORC_INNER_PRODUCT : ORC_L2, query, lut.data());
    for (size_t i = 0; i < n; ++i)
        out[i] = orc_pq_lookup(codes + i * n_chunks, n_chunks, lut.data(), n_centers);
}

// BasicTable::compress_into, diskann-quantization/src/product/tables/basic.rs:161-194
int orc_pq_encode(const float* pivots, size_t n_centers, size_t dim, const uint64_t* offsets,
                  size_t n_chunks, const float* vec, uint8_t* code) {
    for (size_t c = 0; c < n_chunks; ++c) {
        size_t start = offsets[c], stop = offsets[c + 1];
        float min_distance = std::numeric_limits<float>::infinity();
        size_t min_index = (size_t)-1;
        for (size_t p = 0; p < n_centers; ++p) {
            float d = orc_distance(ORC_FLAVOUR_SIMD, ORC_F32, ORC_F32, ORC_L2, vec + start,
                                   pivots + p * dim + start, stop - start, nullptr);
            if (d < min_distance) {
                min_distance = d;
                min_index = p;
            }
        }
        if (std::isinf(min_distance)) return 1 + (int)c;
        code[c] = (uint8_t)min_index;
    }
    return 0;
}

// ------------------------------------------------------------------ scalar quantization
// scalar/quantizer.rs:190-239 (compress) and :407-430 (compensation).
// Rust f32::round rounds half away from zero == roundf; clamp then round; NaN -> code 0.
float orc_sq_compress(const float* shift, float scale, size_t dim, int nbits, const float* vec,
                      uint8_t* codes, int* had_nan) {
    const float max = (float)((1u << nbits) - 1u);       // scalar/mod.rs:129-131
    const float inverse_scale = max / scale;             // bit_scale / self.scale
    const float inverse_bit_scale = 1.0f / max;          // scalar/mod.rs:133-135
    float dot = 0.0f;
    int nan = 0;
    for (size_t i = 0; i < dim; ++i) {
        float f = vec[i];
        nan |= std::isnan(f) ? 1 : 0;
        float t = (f - shift[i]) * inverse_scale;
        // f32::clamp: NaN stays NaN
        float code = std::isnan(t) ? t : (t < 0.0f ? 0.0f : (t > max ? max : t));
        code = std::round(code);
        dot = std::fmaf(code, shift[i], dot);
        codes[i] = std::isnan(code) ? 0 : (uint8_t)code;
    }
    if (had_nan) *had_nan = nan;
    return scale * inverse_bit_scale * dot;
}

// SQStore::set_vector -> compress_into(MutCompensatedVectorRef::from_canonical_front_mut):
// compensation first (meta/vector.rs:495-507), then Dense-packed codes, value i at bit i * nbits
// (bits/slice.rs:261-305).
void orc_sq_encode_row(const float* shift, float scale, size_t dim, int nbits, const float* vec, uint8_t* row) {
    std::vector<uint8_t> codes(dim);
    const float comp = orc_sq_compress(shift, scale, dim, nbits, vec, codes.data(), nullptr);
    memcpy(row, &comp, 4);
    const size_t bytes = (dim * (size_t)nbits + 7) / 8;
    memset(row + 4, 0, bytes);
    for (size_t i = 0; i < dim; ++i) {
        const size_t bit = i * (size_t)nbits;
        row[4 + bit / 8] |= (uint8_t)(codes[i] << (bit % 8));
    }
}

// scalar/vectors.rs:206-237 (CompensatedSquaredL2), :310-376 (CompensatedIP, Result<f32>
// negates), :380-440 (CompensatedCosineNormalized: 1 - l2/2 mathematical; the similarity
// wrapper is 1 - that).  Integer cores: bits/distances.rs:397, 979 (exact u32).
float orc_sq_distance(int metric, int nbits, float scale_squared, float shift_square_norm,
                      const uint8_t* x, float comp_x, const uint8_t* y, float comp_y, size_t dim) {
    const float ibs = 1.0f / (float)((1u << nbits) - 1u);
    const float bit_scale = ibs * ibs;
    uint32_t l2 = 0, ip = 0;
    for (size_t i = 0; i < dim; ++i) {
        int32_t a = x[i], b = y[i];
        l2 += (uint32_t)((a - b) * (a - b));
        ip += (uint32_t)(a * b);
    }
    switch (metric) {
        case ORC_L2:
            return bit_scale * scale_squared * (float)l2;
        case ORC_INNER_PRODUCT: {
            float r = std::fmaf(bit_scale * scale_squared, (float)ip, shift_square_norm) + (comp_y + comp_x);
            return -r;
        }
        default: {
            float l = bit_scale * scale_squared * (float)l2;
            float mathematical = 1.0f - l / 2.0f;
            return 1.0f - mathematical;
        }
    }
}

}  // extern "C"
