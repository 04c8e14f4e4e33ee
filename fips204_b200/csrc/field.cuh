// field.cuh -- Z_q arithmetic for q = 8380417 on int32 (host + device).
// Follows src/helpers.rs of the reference: signed Montgomery reduction with R = 2^32 and
// QINV = 58728449 (helpers.rs:156-165), the shift-based partial reductions (helpers.rs:61-76) and
// centred representatives (helpers.rs:84-91).  mont_reduce here is bit-identical to the
// reference's: (a - t*Q) is divisible by 2^32, so its high word equals hi(a) - hi(t*Q).
#pragma once
#include "params.h"

namespace mldsa {

#ifdef __CUDA_ARCH__
#define MLDSA_MULHI(a, b) __mulhi((a), (b))
#else
#define MLDSA_MULHI(a, b) ((int32_t)(((int64_t)(a) * (int64_t)(b)) >> 32))
#endif

// a * 2^-32 mod q, result in (-q, q) for |a| < 2^31 * q          (helpers.rs:156-165)
__host__ __device__ __forceinline__ int32_t mont_reduce(int64_t a)
{
    const int32_t t = (int32_t)((uint32_t)a * (uint32_t)QINV);
    return (int32_t)(a >> 32) - MLDSA_MULHI(t, Q);
}
// mont_reduce(a * b) with bq = b * QINV mod 2^32 precomputed: IMAD + 2 x IMAD.HI + IADD
__host__ __device__ __forceinline__ int32_t mont_mul_pre(int32_t a, int32_t b, int32_t bq)
{
    const int32_t t = (int32_t)((uint32_t)a * (uint32_t)bq);
    return MLDSA_MULHI(a, b) - MLDSA_MULHI(t, Q);
}
// mont_reduce(a * b), nothing precomputed
__host__ __device__ __forceinline__ int32_t mont_mul(int32_t a, int32_t b)
{
    const int32_t lo = (int32_t)((uint32_t)a * (uint32_t)b);
    const int32_t t = (int32_t)((uint32_t)lo * (uint32_t)QINV);
    return MLDSA_MULHI(a, b) - MLDSA_MULHI(t, Q);
}
// helpers.rs:61-67: into (-q, q) for |a| < 2143289344
__host__ __device__ __forceinline__ int32_t reduce32(int32_t a)
{
#if defined(__CUDA_ARCH__) && defined(RED_ALU)
    // t q = (t << 23) - (t << 13) + t on the ALU pipe (the FMA-heavy pipe is what bounds the arithmetic kernels)
    const uint32_t t = (uint32_t)((a + (1 << 22)) >> 23);
    uint32_t r;
    asm("{\n\t.reg .u32 s13, s23, u;\n\tshl.b32 s13, %2, 13;\n\tshl.b32 s23, %2, 23;\n\tsub.u32 u, s13, %2;\n\tsub.u32 u, u, s23;\n\tadd.u32 %0, %1, u;\n\t}"
        : "=r"(r) : "r"((uint32_t)a), "r"(t));
    return (int32_t)r;
#else
    return a - ((a + (1 << 22)) >> 23) * Q;
#endif
}
// add q if negative
__host__ __device__ __forceinline__ int32_t caddq(int32_t a) { return a + ((a >> 31) & Q); }
// helpers.rs:70-76: canonical [0, q)
__host__ __device__ __forceinline__ int32_t freeze(int32_t a) { return caddq(reduce32(a)); }
// helpers.rs:84-91 on a canonical input t in [0, q): centred representative in (-q/2, q/2]
__host__ __device__ __forceinline__ int32_t center(int32_t t) { return t - ((((Q / 2) - t) >> 31) & Q); }


// high_low.rs:66-96 (decompose) for a canonical r in [0, q); returns r1, writes r0
template <int32_t GAMMA2>
__host__ __device__ __forceinline__ int32_t decompose(int32_t r, int32_t& r0)
{
    int32_t r1 = (r + 127) >> 7;
    if (GAMMA2 == (Q - 1) / 88) {
        r1 = (r1 * 11275 + (1 << 23)) >> 24;
        r1 ^= ((43 - r1) >> 31) & r1;
    } else {
        r1 = (r1 * 1025 + (1 << 21)) >> 22;
        r1 &= 15;
    }
    r0 = r - r1 * 2 * GAMMA2;
    r0 -= (((Q - 1) / 2 - r0) >> 31) & Q;
    return r1;
}
// high_low.rs:155-192 (use_hint) for canonical r
template <int32_t GAMMA2>
__host__ __device__ __forceinline__ int32_t use_hint(int32_t h, int32_t r)
{
    int32_t r0;
    const int32_t r1 = decompose<GAMMA2>(r, r0);
    if (!h) return r1;
    if (GAMMA2 == (Q - 1) / 88) {
        if (r0 > 0) return r1 == 43 ? 0 : r1 + 1;
        return r1 == 0 ? 43 : r1 - 1;
    }
    return r0 > 0 ? (r1 + 1) & 15 : (r1 - 1) & 15;
}

}  // namespace mldsa
