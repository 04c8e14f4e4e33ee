// ntt.cuh -- one warp per 256-coefficient polynomial, 8 coefficients per lane, staged through a
// per-warp shared-memory tile so that every butterfly is register-local.
//
// Replaces src/ntt.rs:14-76 (ntt) and :85-161 (inv_ntt) of the reference for the FUSED kernels, which need
// results mod q only (Shoup butterflies below).  The standalone, representative-exact forward transform (the
// reference's signed Montgomery butterfly ct_bfly, same twiddle table, no reductions between stages) and the
// standalone inverse run with 16 coefficients per lane: ntt16.cuh.
//
// Index bits 7..0 of a coefficient are split over three register layouts (t = lane, i = 0..7):
//   A: idx = 32*i + t              bits 7,6,5 local  -> stages len = 128, 64, 32
//   M: idx = 32*(t>>2) + 4*i + t&3 bits 4,3,2 local  -> stages len = 16, 8, 4
//   B: idx = 8*t + i               bits 2,1,0 local  -> stages len = 2, 1 (and all (un)packing)
// Layout changes go through the tile in natural order with rows padded 32 -> 36 words, which makes
// all three access patterns bank-conflict free (B uses 128-bit accesses).
#pragma once
#include "field.cuh"

namespace mldsa {

// per-lane packed twiddles of one transform direction (generated into zetas.inc): B stage .ba / .y0 / .y1 by lane,
// M stage .m[vector][lane >> 2] (vector-major: a warp's load of one vector is a single 128-byte line)
struct TwPack { int4 ba[32]; int4 y0[32]; int4 y1[32]; int4 m[4][8]; };
// per-lane packed twiddles of the 16-coefficients-per-lane transforms (ntt16.cuh): vector j of lane t (t = lane & 15)
struct Tw16 { int4 v[8][16]; };

#include "zetas.inc"

__device__ __forceinline__ int32_t mulhi_noadd(int32_t a, int32_t b);

constexpr int NTT_ROW = 36;
constexpr int NTT_TILE_WORDS = 8 * NTT_ROW;  // 288 int32 = 1152 B per warp

// t = mont_reduce(z * b) = hi(b z) - hi(lo(b zq) q) (ntt.rs:45) is never formed: both high products feed the two
// three-input adds directly, a' = a + h1 - h2 (ntt.rs:51), b' = a - h1 + h2 (ntt.rs:48).  Same integers (two's
// complement), but with a single use of each high product ptxas folds the subtraction into an IMAD.HI with a 64-bit
// addend and pays three extra instructions per butterfly to build that register pair.
__device__ __forceinline__ void ct_bfly(int32_t& a, int32_t& b, int32_t z, int32_t zq)
{
#ifdef __CUDA_ARCH__
    const uint32_t lo = (uint32_t)b * (uint32_t)zq;
    const uint32_t h1 = (uint32_t)mulhi_noadd(b, z), h2 = (uint32_t)mulhi_noadd((int32_t)lo, Q);
    const uint32_t ua = (uint32_t)a;
    a = (int32_t)(ua + h1 - h2);
    b = (int32_t)(ua - h1 + h2);
#else
    const int32_t t = mont_mul_pre(b, z, zq);
    b = a - t;
    a = a + t;
#endif
}
// Per-lane twiddles of the B stage (len = 2, 1 forward / 1, 2 inverse) and per-row twiddles of the M stage
// (len = 16, 8, 4 / 4, 8, 16) from the packed tables of zetas.inc: 3 + 4 contiguous 128-bit loads per transform.
struct TwB { int2 za, zb, y0, y1, y2, y3; };
struct TwM { int2 z8, za, zb, y0, y1, y2, y3; };
__device__ __forceinline__ TwB load_twB(const TwPack& T, int lane)
{
    const int4* p = T.ba + lane;  // .y0[lane] and .y1[lane] are 32 and 64 vectors further on
    const int4 a = __ldg(p), b = __ldg(p + 32), c = __ldg(p + 64);
    return TwB{make_int2(a.x, a.y), make_int2(a.z, a.w), make_int2(b.x, b.y), make_int2(b.z, b.w), make_int2(c.x, c.y),
               make_int2(c.z, c.w)};
}
__device__ __forceinline__ TwM load_twM(const TwPack& T, int lane)
{
    const int4* p = &T.m[0][lane >> 2];
    const int4 a = __ldg(p), b = __ldg(p + 8), c = __ldg(p + 16), d = __ldg(p + 24);
    return TwM{make_int2(a.x, a.y), make_int2(a.z, a.w), make_int2(b.x, b.y), make_int2(b.z, b.w), make_int2(c.x, c.y),
               make_int2(c.z, c.w), make_int2(d.x, d.y)};
}

// ---- tile <-> register moves (caller provides __syncwarp() ordering unless noted) ----
__device__ __forceinline__ void tile_load_A(const int32_t* tile, int lane, int32_t (&v)[8])
{
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = tile[NTT_ROW * i + lane];
}
__device__ __forceinline__ void tile_store_A(int32_t* tile, int lane, const int32_t (&v)[8])
{
#pragma unroll
    for (int i = 0; i < 8; i++) tile[NTT_ROW * i + lane] = v[i];
}
__device__ __forceinline__ void tile_load_M(const int32_t* tile, int lane, int32_t (&v)[8])
{
    const int base = NTT_ROW * (lane >> 2) + (lane & 3);
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = tile[base + 4 * i];
}
__device__ __forceinline__ void tile_store_M(int32_t* tile, int lane, const int32_t (&v)[8])
{
    const int base = NTT_ROW * (lane >> 2) + (lane & 3);
#pragma unroll
    for (int i = 0; i < 8; i++) tile[base + 4 * i] = v[i];
}
__device__ __forceinline__ void tile_load_B(const int32_t* tile, int lane, int32_t (&v)[8])
{
    const int4* p = reinterpret_cast<const int4*>(tile + NTT_ROW * (lane >> 2) + 8 * (lane & 3));
    const int4 x = p[0], y = p[1];
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
}
__device__ __forceinline__ void tile_store_B(int32_t* tile, int lane, const int32_t (&v)[8])
{
    int4* p = reinterpret_cast<int4*>(tile + NTT_ROW * (lane >> 2) + 8 * (lane & 3));
    p[0] = make_int4(v[0], v[1], v[2], v[3]);
    p[1] = make_int4(v[4], v[5], v[6], v[7]);
}

// A polynomial row parked in shared memory by its owning warp (y-hat / z-hat between the NTT and the
// matrix-vector product).  Lane t's coefficients 8t..8t+3 go to words [4t, 4t+4) and 8t+4..8t+7 to
// words [128 + 4t, 128 + 4t + 4): every 128-bit access of a quarter-warp is one contiguous 128 B line,
// so both halves are bank-conflict free (a plain 8t layout would be 2-way conflicted).
__device__ __forceinline__ void row_store(int32_t* row, int lane, const int32_t (&v)[8])
{
    int4* p = reinterpret_cast<int4*>(row);
    p[lane] = make_int4(v[0], v[1], v[2], v[3]);
    p[32 + lane] = make_int4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void row_load(const int32_t* row, int lane, int32_t (&v)[8])
{
    const int4* p = reinterpret_cast<const int4*>(row);
    const int4 x = p[lane], y = p[32 + lane];
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
}

// =================================================================================================
// Shoup-multiplication transforms for the FUSED kernels (keygen / sign / verify), which need results
// mod q only.  Measured on B200 (tools/pipe_probe.cu): IMAD 2.1 cycles per warp instruction, IMAD.HI
// without an addend 4.0 (and the ALU pipe keeps issuing beside it), IMAD.HI WITH a register addend 5.1
// and it stalls ALU issue -- which is what nvcc makes of the second high product of a Montgomery
// multiplication.  A product with a precomputed quotient z' = round(z 2^32 / q) needs ONE high product:
//     t = w z - mulhi(w, z') q          (IMAD.HI + 2 x IMAD = 8.2 cycles instead of 11.2 + stalls)
// With |z| <= q/2 and any int32 w the result is in (-q/4, 5q/4): mulhi(w, z') = floor(w z / q + e),
// |e| < 1/4.  Ranges: a forward transform adds < 1.25 q per level (< 11 q after 8 levels from |in| < q);
// the inverse doubles its sum path per level, so it reduces once after five levels and its output is
// canonical.  The inverse does NOT scale by 1/256: its callers fold that constant into one operand of the
// preceding pointwise product (A-hat, s-hat, t-hat are prescaled when they are built).
// =================================================================================================
__device__ __forceinline__ int32_t mulhi_noadd(int32_t a, int32_t b)
{
    int32_t r;
    asm("mul.hi.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));  // inline PTX: keeps the addend-free IMAD.HI form
    return r;
}
__device__ __forceinline__ int32_t shoup_mul(int32_t w, int32_t z, int32_t zs)
{
    const int32_t qh = mulhi_noadd(w, zs);
    return (int32_t)((uint32_t)w * (uint32_t)z - (uint32_t)qh * (uint32_t)Q);
}
// z' for a run-time multiplicand z in (-q/2, q/2] (c-hat, a fresh operand every signing iteration):
// 2^32 / q = 513 - g with g = 0.49957...; z' = 513 z - floor(z g) is good to +-1, which costs |w| 2^-32 q of
// extra range.  The intermediate 513 z may wrap; the result fits (|z'| <= 2^31 - 255).
__device__ __forceinline__ int32_t shoup_companion(int32_t z)
{
    constexpr int64_t F32 = (((int64_t)1 << 32) % Q) * ((int64_t)1 << 32) / Q;  // frac(2^32 / q) * 2^32
    constexpr int32_t G = (int32_t)(((int64_t)1 << 32) - F32);                  // (1 - frac) * 2^32 < 2^31
    return (int32_t)((uint32_t)z * 513u - (uint32_t)mulhi_noadd(z, G));
}
// a' = a + t is formed by the two IMADs of the Shoup product themselves (the addend rides along), b' = 2a - a' is one
// three-input IADD3: one ALU-pipe instruction per butterfly instead of two, the FMA-pipe count unchanged.
__device__ __forceinline__ void ct_bfly_s(int32_t& a, int32_t& b, int32_t z, int32_t zs)
{
    const int32_t qh = mulhi_noadd(b, zs);
    const int32_t s = (int32_t)((uint32_t)b * (uint32_t)z + (uint32_t)a - (uint32_t)qh * (uint32_t)Q);
    b = (int32_t)((uint32_t)a + (uint32_t)a - (uint32_t)s);
    a = s;
}
// z0: a register that holds 0 at run time but is opaque to ptxas (a kernel argument).  As a third addend it pins the
// butterfly's sum and difference to IADD3 on the ALU pipe.  Left to itself ptxas turns part of them into IMAD.IADD
// (an add on the FMA-heavy pipe) wherever its per-function pipe balance says so -- e.g. 51 of the 64 adds of the
// inverse transform inside sign_z, a kernel that is bound by exactly that pipe (ncu: fmaheavy 75 %, ALU 47 %).
// The literal 0 (the default) folds away and leaves the choice to ptxas, which is right for the kernels that run
// beside Keccak warps (ALU-bound).
__device__ __forceinline__ void gs_bfly_s(int32_t& a, int32_t& b, int32_t z, int32_t zs, int32_t z0 = 0)
{
    const int32_t t = a;
    a = t + b + z0;
    b = shoup_mul(t - b + z0, z, zs);
}

// Forward NTT, layout A -> layout B, output = NTT(v) mod q as representatives in (-|in| - 10 q, |in| + 10 q).
__device__ __forceinline__ void sntt_fwd_A2B(int32_t (&v)[8], int32_t* tile, int lane)
{
    ct_bfly_s(v[0], v[4], ZSF1, ZSF1S); ct_bfly_s(v[1], v[5], ZSF1, ZSF1S);
    ct_bfly_s(v[2], v[6], ZSF1, ZSF1S); ct_bfly_s(v[3], v[7], ZSF1, ZSF1S);
    ct_bfly_s(v[0], v[2], ZSF2, ZSF2S); ct_bfly_s(v[1], v[3], ZSF2, ZSF2S);
    ct_bfly_s(v[4], v[6], ZSF3, ZSF3S); ct_bfly_s(v[5], v[7], ZSF3, ZSF3S);
    ct_bfly_s(v[0], v[1], ZSF4, ZSF4S); ct_bfly_s(v[2], v[3], ZSF5, ZSF5S);
    ct_bfly_s(v[4], v[5], ZSF6, ZSF6S); ct_bfly_s(v[6], v[7], ZSF7, ZSF7S);
    __syncwarp();
    tile_store_A(tile, lane, v);
    __syncwarp();
    tile_load_M(tile, lane, v);
    {
        const TwM w = load_twM(ZSP_FWD, lane);
        ct_bfly_s(v[0], v[4], w.z8.x, w.z8.y); ct_bfly_s(v[1], v[5], w.z8.x, w.z8.y);
        ct_bfly_s(v[2], v[6], w.z8.x, w.z8.y); ct_bfly_s(v[3], v[7], w.z8.x, w.z8.y);
        ct_bfly_s(v[0], v[2], w.za.x, w.za.y); ct_bfly_s(v[1], v[3], w.za.x, w.za.y);
        ct_bfly_s(v[4], v[6], w.zb.x, w.zb.y); ct_bfly_s(v[5], v[7], w.zb.x, w.zb.y);
        ct_bfly_s(v[0], v[1], w.y0.x, w.y0.y); ct_bfly_s(v[2], v[3], w.y1.x, w.y1.y);
        ct_bfly_s(v[4], v[5], w.y2.x, w.y2.y); ct_bfly_s(v[6], v[7], w.y3.x, w.y3.y);
    }
    tile_store_M(tile, lane, v);
    __syncwarp();
    tile_load_B(tile, lane, v);
    {
        const TwB w = load_twB(ZSP_FWD, lane);
        ct_bfly_s(v[0], v[2], w.za.x, w.za.y); ct_bfly_s(v[1], v[3], w.za.x, w.za.y);
        ct_bfly_s(v[4], v[6], w.zb.x, w.zb.y); ct_bfly_s(v[5], v[7], w.zb.x, w.zb.y);
        ct_bfly_s(v[0], v[1], w.y0.x, w.y0.y); ct_bfly_s(v[2], v[3], w.y1.x, w.y1.y);
        ct_bfly_s(v[4], v[5], w.y2.x, w.y2.y); ct_bfly_s(v[6], v[7], w.y3.x, w.y3.y);
    }
    __syncwarp();
}

// Inverse NTT WITHOUT the 1/256 scaling.  In: layout B, |v| < 2 q.  Out: layout A, canonical [0, q),
// = 256 * INTT(v) mod q (callers fold the 1/256 into an operand; the standalone, scaled inverse is in ntt16.cuh).
__device__ __forceinline__ void sntt_inv_B2A(int32_t (&v)[8], int32_t* tile, int lane, int32_t z0 = 0)
{
    {
        const TwB w = load_twB(ZSP_INV, lane);
        gs_bfly_s(v[0], v[1], w.y0.x, w.y0.y, z0); gs_bfly_s(v[2], v[3], w.y1.x, w.y1.y, z0);
        gs_bfly_s(v[4], v[5], w.y2.x, w.y2.y, z0); gs_bfly_s(v[6], v[7], w.y3.x, w.y3.y, z0);
        gs_bfly_s(v[0], v[2], w.za.x, w.za.y, z0); gs_bfly_s(v[1], v[3], w.za.x, w.za.y, z0);
        gs_bfly_s(v[4], v[6], w.zb.x, w.zb.y, z0); gs_bfly_s(v[5], v[7], w.zb.x, w.zb.y, z0);
    }
    __syncwarp();
    tile_store_B(tile, lane, v);
    __syncwarp();
    tile_load_M(tile, lane, v);
    {
        const TwM w = load_twM(ZSP_INV, lane);
        gs_bfly_s(v[0], v[1], w.y0.x, w.y0.y, z0); gs_bfly_s(v[2], v[3], w.y1.x, w.y1.y, z0);
        gs_bfly_s(v[4], v[5], w.y2.x, w.y2.y, z0); gs_bfly_s(v[6], v[7], w.y3.x, w.y3.y, z0);
        gs_bfly_s(v[0], v[2], w.za.x, w.za.y, z0); gs_bfly_s(v[1], v[3], w.za.x, w.za.y, z0);
        gs_bfly_s(v[4], v[6], w.zb.x, w.zb.y, z0); gs_bfly_s(v[5], v[7], w.zb.x, w.zb.y, z0);
        gs_bfly_s(v[0], v[4], w.z8.x, w.z8.y, z0); gs_bfly_s(v[1], v[5], w.z8.x, w.z8.y, z0);
        gs_bfly_s(v[2], v[6], w.z8.x, w.z8.y, z0); gs_bfly_s(v[3], v[7], w.z8.x, w.z8.y, z0);
    }
    // five doubling levels done: |v| < 64 q.  Back under q so that the last three stay inside int32.
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = reduce32(v[i]);
    tile_store_M(tile, lane, v);
    __syncwarp();
    tile_load_A(tile, lane, v);
    gs_bfly_s(v[0], v[1], ZSI7, ZSI7S, z0); gs_bfly_s(v[2], v[3], ZSI6, ZSI6S, z0);
    gs_bfly_s(v[4], v[5], ZSI5, ZSI5S, z0); gs_bfly_s(v[6], v[7], ZSI4, ZSI4S, z0);
    gs_bfly_s(v[0], v[2], ZSI3, ZSI3S, z0); gs_bfly_s(v[1], v[3], ZSI3, ZSI3S, z0);
    gs_bfly_s(v[4], v[6], ZSI2, ZSI2S, z0); gs_bfly_s(v[5], v[7], ZSI2, ZSI2S, z0);
    gs_bfly_s(v[0], v[4], ZSI1, ZSI1S, z0); gs_bfly_s(v[1], v[5], ZSI1, ZSI1S, z0);
    gs_bfly_s(v[2], v[6], ZSI1, ZSI1S, z0); gs_bfly_s(v[3], v[7], ZSI1, ZSI1S, z0);
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = freeze(v[i]);
    __syncwarp();
}
__device__ __forceinline__ void sntt_fwd_B2B(int32_t (&v)[8], int32_t* tile, int lane)
{
    __syncwarp();
    tile_store_B(tile, lane, v);
    __syncwarp();
    tile_load_A(tile, lane, v);
    sntt_fwd_A2B(v, tile, lane);
}
__device__ __forceinline__ void sntt_inv_B2B(int32_t (&v)[8], int32_t* tile, int lane, int32_t z0 = 0)
{
    sntt_inv_B2A(v, tile, lane, z0);
    tile_store_A(tile, lane, v);
    __syncwarp();
    tile_load_B(tile, lane, v);
    __syncwarp();
}

}  // namespace mldsa
