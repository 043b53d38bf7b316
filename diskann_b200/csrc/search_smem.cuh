// search_smem.cuh — device helpers shared by the search kernels that keep the visited set in
// shared memory (search_kernel_v3.cu).
#pragma once

#include "distance_device.cuh"
#include "search_common.cuh"

namespace dab {
namespace {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- shared-memory visited set: 16 tags of 16 bits per 32-byte bucket -----------------------
// Entries fill a bucket from slot 0 upwards and are never removed, so an id that is absent
// from its home bucket while that bucket has a free slot is new; an id displaced to the d-th
// following bucket (d <= 2) carries d in its top two bits.  0xFFFF marks an empty slot.
__device__ __forceinline__ void load_bucket_smem(const uint32_t* bp, uint32_t (&s)[8]) {
    const uint4 lo = reinterpret_cast<const uint4*>(bp)[0];
    const uint4 hi = reinterpret_cast<const uint4*>(bp)[1];
    s[0] = lo.x, s[1] = lo.y, s[2] = lo.z, s[3] = lo.w, s[4] = hi.x, s[5] = hi.y, s[6] = hi.z, s[7] = hi.w;
}

// true when the id was newly inserted (HashSet::insert); `ovf` is raised when the home bucket
// and the two after it are full
__device__ __forceinline__ bool smem16_insert(uint32_t* table, uint32_t n_buckets, uint32_t b, uint32_t tag, bool& ovf) {
    uint32_t d = 0;
    for (;;) {
        uint32_t* bp = table + (size_t)b * 8;
        uint32_t s[8];
        load_bucket_smem(bp, s);
        const uint32_t want = (d << 14) | tag, want2 = want * 0x10001u;
        // "some 16-bit half of x is zero" <=> ((x - 0x00010001) & ~x & 0x80008000) != 0
        uint32_t hit = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t x = s[k] ^ want2;
            hit |= (x - 0x00010001u) & ~x & 0x80008000u;
        }
        if (hit) return false;
        // first free slot: slots fill in order, so it is the number of occupied halves
        int ew = -1;
        uint32_t old = 0;
#pragma unroll
        for (int k = 7; k >= 0; --k) {
            if ((s[k] >> 16) == 0xFFFFu) {
                ew = k;
                old = s[k];
            }
        }
        if (ew >= 0) {
            const uint32_t neu = (old & 0xFFFFu) == 0xFFFFu ? (0xFFFF0000u | want) : ((old & 0xFFFFu) | (want << 16));
            if (atomicCAS(bp + ew, old, neu) == old) return true;
            continue;  // another lane of this warp changed the word: look at the bucket again
        }
        if (++d > 2) {
            ovf = true;
            return false;
        }
        b = b + 1 == n_buckets ? 0 : b + 1;
    }
}

// ---- linear-probing variant: one 16-bit entry per slot ---------------------------------------
// id -> h = (id * odd) mod 2^K (a bijection), home slot = h mod n_slots, tag = h div n_slots.
// An entry stores (displacement << tag_bits) | tag, so an entry found d slots after its home
// is unambiguous; entries are never removed, hence everything between an id's home and its
// slot stays occupied and a lookup can stop at the first empty slot.  0xFFFF marks empty.
// One probe step is one 32-bit shared-memory load and two compares (against ~90 instructions for
// a 16-entry bucket scan); at the load the search runs at (<= 87 %, typically 30 %) a probe
// takes 1.2 - 2 steps.
struct Lp16Map {
    uint32_t kmask;     // 2^K - 1
    uint32_t n_slots;
    uint32_t magic;     // ceil(2^(K+s) / n_slots), s = ceil(log2 n_slots): exact h / n_slots for h < 2^K
    uint32_t shift;     // K + s
    uint32_t tag_bits;  // bits of the largest tag
    uint32_t dmax;      // largest displacement an entry can record
};

// true when the id was newly inserted (HashSet::insert); `ovf` is raised when the id would
// need a displacement beyond dmax
__device__ __forceinline__ bool lp16_insert(uint32_t* table, const Lp16Map& m, uint32_t id, bool& ovf) {
    const uint32_t h = (id * 0x9E3779B1u) & m.kmask;
    const uint32_t tag = (uint32_t)(((uint64_t)h * m.magic) >> m.shift);
    uint32_t s = h - tag * m.n_slots;
    uint32_t want = tag;
    const uint32_t step = 1u << m.tag_bits;
    for (uint32_t d = 0;;) {
        uint32_t* wp = table + (s >> 1);
        const uint32_t w = *wp;
        const uint32_t cur = (s & 1u) ? (w >> 16) : (w & 0xFFFFu);
        if (cur == want) return false;
        if (cur == 0xFFFFu) {
            const uint32_t neu = (s & 1u) ? ((w & 0xFFFFu) | (want << 16)) : ((w & 0xFFFF0000u) | want);
            if (atomicCAS(wp, w, neu) == w) return true;
            continue;  // another lane of this warp changed the word: look at the slot again
        }
        if (++d > m.dmax) {
            ovf = true;
            return false;
        }
        want += step;
        s = s + 1 == m.n_slots ? 0 : s + 1;
    }
}

// Packed f32x2 arithmetic (FADD2 / FFMA2): each half is an IEEE round-to-nearest operation.
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
template <int KIND>
__device__ __forceinline__ uint64_t step2(uint64_t acc, uint64_t x2, uint64_t y2) {
    if (KIND == KIND_L2) {
        uint64_t c2;
        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(c2) : "l"(x2), "l"(y2));
        asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(acc) : "l"(c2), "l"(acc));
    } else {
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(acc) : "l"(x2), "l"(y2), "l"(acc));
    }
    return acc;
}

__device__ __forceinline__ uint4 ldg16(const uint8_t* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// All candidate rows of a hop are requested from HBM at once with one bulk L2 prefetch per row
// (no registers, no shared memory); the register passes below then overlap with the fills and
// find all but the first rows in L2.
__device__ __forceinline__ void prefetch_rows(const uint8_t* __restrict__ vectors, size_t row_stride, const uint32_t* __restrict__ cid,
                                              uint32_t n, uint32_t row_bytes16, int lane) {
    for (uint32_t j = lane; j < n; j += 32) {
        const uint8_t* src = vectors + (size_t)cid[j] * row_stride;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(row_bytes16) : "memory");
    }
}


}  // namespace
}  // namespace dab
