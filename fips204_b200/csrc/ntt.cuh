// ntt.cuh -- one warp per 256-coefficient polynomial, 8 coefficients per lane, staged through a
// per-warp shared-memory tile so that every butterfly is register-local.
//
// Replaces src/ntt.rs:14-76 (ntt) and :85-161 (inv_ntt) of the reference.  The forward transform
// is representative-exact with the reference (same signed Montgomery butterfly, same twiddle table,
// no reductions between stages); the inverse returns canonical [0, q) like the reference.
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

#include "zetas.inc"

constexpr int NTT_ROW = 36;
constexpr int NTT_TILE_WORDS = 8 * NTT_ROW;  // 288 int32 = 1152 B per warp

__device__ __forceinline__ void ct_bfly(int32_t& a, int32_t& b, int32_t z, int32_t zq)
{
    const int32_t t = mont_mul_pre(b, z, zq);  // ntt.rs:45
    b = a - t;                                 // ntt.rs:48
    a = a + t;                                 // ntt.rs:51
}
__device__ __forceinline__ void gs_bfly(int32_t& a, int32_t& b, int32_t z, int32_t zq)
{
    const int32_t t = a;                       // ntt.rs:124
    a = t + b;                                 // ntt.rs:127
    b = mont_mul_pre(t - b, z, zq);            // ntt.rs:130-133
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

// Forward NTT.  In: v in layout A (coefficient domain).  Out: v in layout B (NTT domain, the
// reference's in-place array order).  Uses the warp's tile; ends with the tile free.
__device__ __forceinline__ void ntt_fwd_A2B(int32_t (&v)[8], int32_t* tile, int lane)
{
    // len = 128, 64, 32: twiddles ZETA[1], ZETA[2..3], ZETA[4..7] are compile-time constants
    ct_bfly(v[0], v[4], ZF1, ZF1Q); ct_bfly(v[1], v[5], ZF1, ZF1Q);
    ct_bfly(v[2], v[6], ZF1, ZF1Q); ct_bfly(v[3], v[7], ZF1, ZF1Q);
    ct_bfly(v[0], v[2], ZF2, ZF2Q); ct_bfly(v[1], v[3], ZF2, ZF2Q);
    ct_bfly(v[4], v[6], ZF3, ZF3Q); ct_bfly(v[5], v[7], ZF3, ZF3Q);
    ct_bfly(v[0], v[1], ZF4, ZF4Q); ct_bfly(v[2], v[3], ZF5, ZF5Q);
    ct_bfly(v[4], v[5], ZF6, ZF6Q); ct_bfly(v[6], v[7], ZF7, ZF7Q);
    __syncwarp();
    tile_store_A(tile, lane, v);
    __syncwarp();
    tile_load_M(tile, lane, v);
    {   // len = 16, 8, 4
        const int r = lane >> 2;
        const int2 z8 = __ldg(&ZETA_FWD[8 + r]);
        ct_bfly(v[0], v[4], z8.x, z8.y); ct_bfly(v[1], v[5], z8.x, z8.y);
        ct_bfly(v[2], v[6], z8.x, z8.y); ct_bfly(v[3], v[7], z8.x, z8.y);
        const int2 za = __ldg(&ZETA_FWD[16 + 2 * r]), zb = __ldg(&ZETA_FWD[17 + 2 * r]);
        ct_bfly(v[0], v[2], za.x, za.y); ct_bfly(v[1], v[3], za.x, za.y);
        ct_bfly(v[4], v[6], zb.x, zb.y); ct_bfly(v[5], v[7], zb.x, zb.y);
        const int2 y0 = __ldg(&ZETA_FWD[32 + 4 * r]), y1 = __ldg(&ZETA_FWD[33 + 4 * r]);
        const int2 y2 = __ldg(&ZETA_FWD[34 + 4 * r]), y3 = __ldg(&ZETA_FWD[35 + 4 * r]);
        ct_bfly(v[0], v[1], y0.x, y0.y); ct_bfly(v[2], v[3], y1.x, y1.y);
        ct_bfly(v[4], v[5], y2.x, y2.y); ct_bfly(v[6], v[7], y3.x, y3.y);
    }
    tile_store_M(tile, lane, v);  // same addresses this lane just read
    __syncwarp();
    tile_load_B(tile, lane, v);
    {   // len = 2, 1
        const int2 za = __ldg(&ZETA_FWD[64 + 2 * lane]), zb = __ldg(&ZETA_FWD[65 + 2 * lane]);
        ct_bfly(v[0], v[2], za.x, za.y); ct_bfly(v[1], v[3], za.x, za.y);
        ct_bfly(v[4], v[6], zb.x, zb.y); ct_bfly(v[5], v[7], zb.x, zb.y);
        const int2 y0 = __ldg(&ZETA_FWD[128 + 4 * lane]), y1 = __ldg(&ZETA_FWD[129 + 4 * lane]);
        const int2 y2 = __ldg(&ZETA_FWD[130 + 4 * lane]), y3 = __ldg(&ZETA_FWD[131 + 4 * lane]);
        ct_bfly(v[0], v[1], y0.x, y0.y); ct_bfly(v[2], v[3], y1.x, y1.y);
        ct_bfly(v[4], v[5], y2.x, y2.y); ct_bfly(v[6], v[7], y3.x, y3.y);
    }
    __syncwarp();
}

// Inverse NTT.  In: v in layout B (NTT domain), |v| < q.  Out: v in layout A, canonical [0, q),
// scaled by mont(F * .) -- F = F_MONT for a plain transform (ntt.rs:152-154), F_MONT2 when the input
// carries a stray 2^-32 from an un-compensated Montgomery product.
template <int32_t F>
__device__ __forceinline__ void ntt_inv_B2A(int32_t (&v)[8], int32_t* tile, int lane)
{
    {   // len = 1, 2   (twiddle index counts down: m = 2*nb - 1 - block, ntt.rs:112-118)
        const int2 y0 = __ldg(&ZETA_INV[255 - 4 * lane]), y1 = __ldg(&ZETA_INV[254 - 4 * lane]);
        const int2 y2 = __ldg(&ZETA_INV[253 - 4 * lane]), y3 = __ldg(&ZETA_INV[252 - 4 * lane]);
        gs_bfly(v[0], v[1], y0.x, y0.y); gs_bfly(v[2], v[3], y1.x, y1.y);
        gs_bfly(v[4], v[5], y2.x, y2.y); gs_bfly(v[6], v[7], y3.x, y3.y);
        const int2 za = __ldg(&ZETA_INV[127 - 2 * lane]), zb = __ldg(&ZETA_INV[126 - 2 * lane]);
        gs_bfly(v[0], v[2], za.x, za.y); gs_bfly(v[1], v[3], za.x, za.y);
        gs_bfly(v[4], v[6], zb.x, zb.y); gs_bfly(v[5], v[7], zb.x, zb.y);
    }
    __syncwarp();
    tile_store_B(tile, lane, v);
    __syncwarp();
    tile_load_M(tile, lane, v);
    {   // len = 4, 8, 16
        const int r = lane >> 2;
        const int2 y0 = __ldg(&ZETA_INV[63 - 4 * r]), y1 = __ldg(&ZETA_INV[62 - 4 * r]);
        const int2 y2 = __ldg(&ZETA_INV[61 - 4 * r]), y3 = __ldg(&ZETA_INV[60 - 4 * r]);
        gs_bfly(v[0], v[1], y0.x, y0.y); gs_bfly(v[2], v[3], y1.x, y1.y);
        gs_bfly(v[4], v[5], y2.x, y2.y); gs_bfly(v[6], v[7], y3.x, y3.y);
        const int2 za = __ldg(&ZETA_INV[31 - 2 * r]), zb = __ldg(&ZETA_INV[30 - 2 * r]);
        gs_bfly(v[0], v[2], za.x, za.y); gs_bfly(v[1], v[3], za.x, za.y);
        gs_bfly(v[4], v[6], zb.x, zb.y); gs_bfly(v[5], v[7], zb.x, zb.y);
        const int2 z8 = __ldg(&ZETA_INV[15 - r]);
        gs_bfly(v[0], v[4], z8.x, z8.y); gs_bfly(v[1], v[5], z8.x, z8.y);
        gs_bfly(v[2], v[6], z8.x, z8.y); gs_bfly(v[3], v[7], z8.x, z8.y);
    }
    tile_store_M(tile, lane, v);
    __syncwarp();
    tile_load_A(tile, lane, v);
    // len = 32, 64, 128
    gs_bfly(v[0], v[1], ZI7, ZI7Q); gs_bfly(v[2], v[3], ZI6, ZI6Q);
    gs_bfly(v[4], v[5], ZI5, ZI5Q); gs_bfly(v[6], v[7], ZI4, ZI4Q);
    gs_bfly(v[0], v[2], ZI3, ZI3Q); gs_bfly(v[1], v[3], ZI3, ZI3Q);
    gs_bfly(v[4], v[6], ZI2, ZI2Q); gs_bfly(v[5], v[7], ZI2, ZI2Q);
    gs_bfly(v[0], v[4], ZI1, ZI1Q); gs_bfly(v[1], v[5], ZI1, ZI1Q);
    gs_bfly(v[2], v[6], ZI1, ZI1Q); gs_bfly(v[3], v[7], ZI1, ZI1Q);
    constexpr int32_t FQ = (int32_t)((uint32_t)F * (uint32_t)QINV);
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = freeze(mont_mul_pre(v[i], F, FQ));
    __syncwarp();
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

// Convenience wrappers used by the fused kernels: coefficient-domain data lives in layout B
// (lane-local bit packing), NTT-domain data too.
__device__ __forceinline__ void ntt_fwd_B2B(int32_t (&v)[8], int32_t* tile, int lane)
{
    __syncwarp();
    tile_store_B(tile, lane, v);
    __syncwarp();
    tile_load_A(tile, lane, v);
    ntt_fwd_A2B(v, tile, lane);
}
template <int32_t F>
__device__ __forceinline__ void ntt_inv_B2B(int32_t (&v)[8], int32_t* tile, int lane)
{
    ntt_inv_B2A<F>(v, tile, lane);
    tile_store_A(tile, lane, v);
    __syncwarp();
    tile_load_B(tile, lane, v);
    __syncwarp();
}

}  // namespace mldsa
