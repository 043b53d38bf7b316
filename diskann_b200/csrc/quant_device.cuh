// quant_device.cuh — thread-serial emulation of the reference's simd_op for short f32 chunks
// (PQ LUT entries, encode), shared by quant_kernels.cu and search_kernel_pq.cu.
#pragma once

#include "distance_device.cuh"

namespace dab {

// ------------------------------------------------------------------ thread-serial simd_op
// One thread emulates simd_op (simd.rs:686-747) for f32 x f32 with NA accumulators of 8
// lanes and returns the *combined* 8-lane accumulator (what Resumable::combine_with receives,
// simd.rs:637-671).  Used where work items are tiny (PQ chunks of 2..16 dims).
template <int NA, int KIND /*L2 or IP*/>
__device__ __forceinline__ void thread_simd_combined(const float* __restrict__ x, const float* __restrict__ y,
                                                     int len, float (&c)[8]) {
    float s[NA][8];
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int l = 0; l < 8; ++l) s[a][l] = 0.0f;
    const int full = len >> 3;
    const int groups = full / NA;
    for (int g = 0; g < groups; ++g) {
#pragma unroll
        for (int a = 0; a < NA; ++a) {
            const int base = (g * NA + a) * 8;
#pragma unroll
            for (int l = 0; l < 8; ++l) {
                float xv = x[base + l], yv = y[base + l];
                if (KIND == KIND_L2) {
                    float d = __fsub_rn(xv, yv);
                    s[a][l] = __fmaf_rn(d, d, s[a][l]);
                } else {
                    s[a][l] = __fmaf_rn(xv, yv, s[a][l]);
                }
            }
        }
    }
    const int ep = full - groups * NA;
#pragma unroll
    for (int a = 0; a < NA - 1; ++a) {
        if (a < ep) {
            const int base = (groups * NA + a) * 8;
#pragma unroll
            for (int l = 0; l < 8; ++l) {
                float xv = x[base + l], yv = y[base + l];
                if (KIND == KIND_L2) {
                    float d = __fsub_rn(xv, yv);
                    s[a][l] = __fmaf_rn(d, d, s[a][l]);
                } else {
                    s[a][l] = __fmaf_rn(xv, yv, s[a][l]);
                }
            }
        }
    }
#pragma unroll
    for (int l = 0; l < 8; ++l) {
        if (NA == 4)
            c[l] = __fadd_rn(__fadd_rn(s[0][l], s[1][l]), __fadd_rn(s[2 % NA][l], s[3 % NA][l]));
        else
            c[l] = __fadd_rn(s[0][l], s[1 % NA][l]);
    }
    const int rem = len & 7;
    if (rem) {
        const int base = full * 8;
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            float xv = l < rem ? x[base + l] : 0.0f, yv = l < rem ? y[base + l] : 0.0f;
            if (KIND == KIND_L2) {
                float d = __fsub_rn(xv, yv);
                c[l] = __fmaf_rn(d, d, c[l]);
            } else {
                c[l] = __fmaf_rn(xv, yv, c[l]);
            }
        }
    }
}

__device__ __forceinline__ float thread_tree8(const float (&c)[8]) {
    float a0 = __fadd_rn(c[0], c[4]), a1 = __fadd_rn(c[1], c[5]), a2 = __fadd_rn(c[2], c[6]), a3 = __fadd_rn(c[3], c[7]);
    return __fadd_rn(__fadd_rn(a0, a2), __fadd_rn(a1, a3));
}

template <int KIND>
__device__ __forceinline__ float thread_simd_l2ip(const float* x, const float* y, int len) {
    float c[8];
    thread_simd_combined<4, KIND>(x, y, len, c);
    return thread_tree8(c);
}

}  // namespace dab
