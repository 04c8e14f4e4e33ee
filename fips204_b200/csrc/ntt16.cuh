// ntt16.cuh -- the Shoup transforms of ntt.cuh with 16 coefficients per lane: a HALF-warp per polynomial, two
// polynomials per warp (lanes 0-15 and 16-31 run the same instruction stream on their own polynomial).
//
// Why: the warp-per-item kernels are held by the L1 data pipe.  With 8 coefficients per lane the 8 index
// bits need three register layouts (3 + 3 + 2 local bits) and a transform between two byte-packed forms moves its 1 KiB
// through shared memory three times (48 wavefronts); with 16 per lane two layouts suffice:
//   Y: idx = 16 t + i   bits 3..0 local -> levels len = 8, 4, 2, 1, all (un)packing (16 contiguous coefficients)
//   X: idx = 16 i + t   bits 7..4 local -> levels len = 128, 64, 32, 16, whose twiddles do not depend on the lane at all
// i.e. two exchanges (32 wavefronts per polynomial) and 20 instead of 36 shared-memory instructions per polynomial; the
// per-lane twiddles of the Y levels (1 + 2 + 4 + 8 entries from index 16 + t, 32 + 2t, 64 + 4t, 128 + 8t of the plain
// tables; mirrored for the inverse) come from per-lane packed copies, eight 128-bit loads serving both polynomials of
// the warp; and every level offers 8 independent butterflies per lane instead of 4.
// Index maps, twiddle indices and the packed tables are checked against a direct transform by a Python model of this
// file (tests/test_ntt16_tables.py, no GPU needed).
//
// Same arithmetic as ntt.cuh (src/ntt.rs:14-161 mod q): ct_bfly_s / gs_bfly_s, results mod q as lazily reduced int32.
#pragma once
#include "ntt.cuh"

namespace mldsa {

// Per half-warp tile: lane t's 16 words start at 20 t (16 + 4 words of padding: the 128-bit Y accesses of a quarter-warp
// then cover all 32 banks, and the 32-bit X accesses -- word 20 i + t -- are 16 consecutive banks).  The second
// half-warp's tile starts 336 words further (= 16 mod 32), so both halves of a warp instruction hit disjoint banks.
constexpr int T16_BLK = 20;
constexpr int T16_POLY = 16 * T16_BLK + 16;
constexpr int T16_WORDS = 2 * T16_POLY;  // per warp: 672 words = 2 688 B

__device__ __forceinline__ void t16_store_Y(int32_t* tile, int t, const int32_t (&v)[16])
{
    int4* p = reinterpret_cast<int4*>(tile + T16_BLK * t);
#pragma unroll
    for (int c = 0; c < 4; c++) p[c] = make_int4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}
__device__ __forceinline__ void t16_load_Y(const int32_t* tile, int t, int32_t (&v)[16])
{
    const int4* p = reinterpret_cast<const int4*>(tile + T16_BLK * t);
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const int4 x = p[c];
        v[4 * c] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w;
    }
}
__device__ __forceinline__ void t16_store_X(int32_t* tile, int t, const int32_t (&v)[16])
{
#pragma unroll
    for (int i = 0; i < 16; i++) tile[T16_BLK * i + t] = v[i];
}
__device__ __forceinline__ void t16_load_X(const int32_t* tile, int t, int32_t (&v)[16])
{
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = tile[T16_BLK * i + t];
}

// one level on the local index bit `STRIDE` with ONE twiddle per block of 2 * STRIDE registers: tw[b] = (z, z')
#define T16_CT(a, b, Z, ZS) ct_bfly_s(v[a], v[b], Z, ZS)
#define T16_GS(a, b, Z, ZS) gs_bfly_s(v[a], v[b], Z, ZS, z0)

// Forward NTT, layout Y -> layout Y.  |out| < |in| + 10 q (no reductions, as in ntt.cuh).  `tile`: this half-warp's tile.
__device__ __forceinline__ void sntt16_fwd(int32_t (&v)[16], int32_t* tile, int t)
{
    __syncwarp();
    t16_store_Y(tile, t, v);
    __syncwarp();
    t16_load_X(tile, t, v);
    // len = 128, 64, 32, 16: twiddle ZETA[2^(7-s) + (i >> (s - 3))], compile-time constants
#pragma unroll
    for (int i = 0; i < 8; i++) T16_CT(i, i + 8, ZSF1, ZSF1S);
#pragma unroll
    for (int i = 0; i < 4; i++) { T16_CT(i, i + 4, ZSF2, ZSF2S); T16_CT(i + 8, i + 12, ZSF3, ZSF3S); }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        T16_CT(i, i + 2, ZSF4, ZSF4S); T16_CT(i + 4, i + 6, ZSF5, ZSF5S);
        T16_CT(i + 8, i + 10, ZSF6, ZSF6S); T16_CT(i + 12, i + 14, ZSF7, ZSF7S);
    }
    T16_CT(0, 1, ZSF8, ZSF8S); T16_CT(2, 3, ZSF9, ZSF9S); T16_CT(4, 5, ZSF10, ZSF10S); T16_CT(6, 7, ZSF11, ZSF11S);
    T16_CT(8, 9, ZSF12, ZSF12S); T16_CT(10, 11, ZSF13, ZSF13S); T16_CT(12, 13, ZSF14, ZSF14S); T16_CT(14, 15, ZSF15, ZSF15S);
    t16_store_X(tile, t, v);  // the addresses this lane just read
    __syncwarp();
    t16_load_Y(tile, t, v);
    // len = 8: ZETA[16 + t]; len = 4: ZETA[32 + 2t + (i >> 3)]; len = 2: ZETA[64 + 4t + (i >> 2)]; len = 1: ZETA[128 + 8t + (i >> 1)]
    // -- the lane's 15 (z, z') pairs in this order, as eight contiguous-across-lanes 128-bit vectors (ZS16_FWD, zetas.inc)
    const int4* q = &ZS16_FWD.v[0][t];
    {
        const int4 w0 = __ldg(q), w1 = __ldg(q + 16), w2 = __ldg(q + 32), w3 = __ldg(q + 48);
#pragma unroll
        for (int i = 0; i < 8; i++) T16_CT(i, i + 8, w0.x, w0.y);
#pragma unroll
        for (int i = 0; i < 4; i++) { T16_CT(i, i + 4, w0.z, w0.w); T16_CT(i + 8, i + 12, w1.x, w1.y); }
#pragma unroll
        for (int i = 0; i < 2; i++) {
            T16_CT(i, i + 2, w1.z, w1.w); T16_CT(i + 4, i + 6, w2.x, w2.y);
            T16_CT(i + 8, i + 10, w2.z, w2.w); T16_CT(i + 12, i + 14, w3.x, w3.y);
        }
        T16_CT(0, 1, w3.z, w3.w);
    }
    {
        const int4 w4 = __ldg(q + 64), w5 = __ldg(q + 80), w6 = __ldg(q + 96), w7 = __ldg(q + 112);
        T16_CT(2, 3, w4.x, w4.y); T16_CT(4, 5, w4.z, w4.w); T16_CT(6, 7, w5.x, w5.y); T16_CT(8, 9, w5.z, w5.w);
        T16_CT(10, 11, w6.x, w6.y); T16_CT(12, 13, w6.z, w6.w); T16_CT(14, 15, w7.x, w7.y);
    }
}

// Inverse NTT WITHOUT the 1/256 (callers fold it into an operand, ntt.cuh).  In: layout Y, |v| < 2 q.  Out: layout Y,
// canonical [0, q).  Ranges: a sum output adds the bounds of its inputs, a difference output leaves its product in
// (-q/4, 5q/4); after the four Y levels the bound of register i is 32 q, 10 q, 5 q, 5 q, 2.5 q x 4, 1.25 q x 8, and in
// layout X every register of lane t has the bound of Y-register t, which the four X levels multiply by at most 16: only
// Y-register 0 (32 q -> 512 q) has to be reduced on the way, everything else stays below 160 q < 2^31 / q = 256 q.
__device__ __forceinline__ void sntt16_inv(int32_t (&v)[16], int32_t* tile, int t, int32_t z0 = 0)
{
    // len = 1: -ZETA[255 - 8t - (i >> 1)]; len = 2: -ZETA[127 - 4t - (i >> 2)]; len = 4: -ZETA[63 - 2t - (i >> 3)]; len = 8:
    // -ZETA[31 - t] -- the lane's 15 pairs in this order (ZS16_INV)
    const int4* q = &ZS16_INV.v[0][t];
    {
        const int4 w0 = __ldg(q), w1 = __ldg(q + 16), w2 = __ldg(q + 32), w3 = __ldg(q + 48);
        T16_GS(0, 1, w0.x, w0.y); T16_GS(2, 3, w0.z, w0.w); T16_GS(4, 5, w1.x, w1.y); T16_GS(6, 7, w1.z, w1.w);
        T16_GS(8, 9, w2.x, w2.y); T16_GS(10, 11, w2.z, w2.w); T16_GS(12, 13, w3.x, w3.y); T16_GS(14, 15, w3.z, w3.w);
    }
    {
        const int4 w4 = __ldg(q + 64), w5 = __ldg(q + 80), w6 = __ldg(q + 96), w7 = __ldg(q + 112);
#pragma unroll
        for (int i = 0; i < 2; i++) {
            T16_GS(i, i + 2, w4.x, w4.y); T16_GS(i + 4, i + 6, w4.z, w4.w);
            T16_GS(i + 8, i + 10, w5.x, w5.y); T16_GS(i + 12, i + 14, w5.z, w5.w);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) { T16_GS(i, i + 4, w6.x, w6.y); T16_GS(i + 8, i + 12, w6.z, w6.w); }
#pragma unroll
        for (int i = 0; i < 8; i++) T16_GS(i, i + 8, w7.x, w7.y);
    }
    v[0] = reduce32(v[0]);
    __syncwarp();
    t16_store_Y(tile, t, v);
    __syncwarp();
    t16_load_X(tile, t, v);
    // len = 16, 32, 64, 128: -ZETA[15 - (i >> 1)], [7 - (i >> 2)], [3 - (i >> 3)], [1]
    T16_GS(0, 1, ZSI15, ZSI15S); T16_GS(2, 3, ZSI14, ZSI14S); T16_GS(4, 5, ZSI13, ZSI13S); T16_GS(6, 7, ZSI12, ZSI12S);
    T16_GS(8, 9, ZSI11, ZSI11S); T16_GS(10, 11, ZSI10, ZSI10S); T16_GS(12, 13, ZSI9, ZSI9S); T16_GS(14, 15, ZSI8, ZSI8S);
#pragma unroll
    for (int i = 0; i < 2; i++) {
        T16_GS(i, i + 2, ZSI7, ZSI7S); T16_GS(i + 4, i + 6, ZSI6, ZSI6S);
        T16_GS(i + 8, i + 10, ZSI5, ZSI5S); T16_GS(i + 12, i + 14, ZSI4, ZSI4S);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) { T16_GS(i, i + 4, ZSI3, ZSI3S); T16_GS(i + 8, i + 12, ZSI2, ZSI2S); }
#pragma unroll
    for (int i = 0; i < 8; i++) T16_GS(i, i + 8, ZSI1, ZSI1S);
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = freeze(v[i]);
    t16_store_X(tile, t, v);
    __syncwarp();
    t16_load_Y(tile, t, v);
}

// ---- the standalone transforms of the microbench entry points (mldsa_ntt_batch / mldsa_intt_batch): ONE exchange each,
// because their global accesses can use either layout (X: 64 contiguous bytes per half-warp and instruction; Y: 64
// contiguous bytes per lane) ----

// Forward NTT, representative-exact with src/ntt.rs:14-76: the reference's signed Montgomery butterfly in the reference's
// order per coefficient, so every output is the very int32 the reference computes.  In: layout X.  Out: layout Y.
#define T16_CTM(a, b, Z, ZQ) ct_bfly(v[a], v[b], Z, ZQ)
__device__ __forceinline__ void ntt16_fwd_mont_X2Y(int32_t (&v)[16], int32_t* tile, int t)
{
#pragma unroll
    for (int i = 0; i < 8; i++) T16_CTM(i, i + 8, ZF1, ZF1Q);
#pragma unroll
    for (int i = 0; i < 4; i++) { T16_CTM(i, i + 4, ZF2, ZF2Q); T16_CTM(i + 8, i + 12, ZF3, ZF3Q); }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        T16_CTM(i, i + 2, ZF4, ZF4Q); T16_CTM(i + 4, i + 6, ZF5, ZF5Q);
        T16_CTM(i + 8, i + 10, ZF6, ZF6Q); T16_CTM(i + 12, i + 14, ZF7, ZF7Q);
    }
    T16_CTM(0, 1, ZF8, ZF8Q); T16_CTM(2, 3, ZF9, ZF9Q); T16_CTM(4, 5, ZF10, ZF10Q); T16_CTM(6, 7, ZF11, ZF11Q);
    T16_CTM(8, 9, ZF12, ZF12Q); T16_CTM(10, 11, ZF13, ZF13Q); T16_CTM(12, 13, ZF14, ZF14Q); T16_CTM(14, 15, ZF15, ZF15Q);
    __syncwarp();  // (the tile's previous readers: the other lanes' layout-Y loads of the last polynomial)
    t16_store_X(tile, t, v);
    __syncwarp();
    t16_load_Y(tile, t, v);
    const int4* q = &ZM16_FWD.v[0][t];
    {
        const int4 w0 = __ldg(q), w1 = __ldg(q + 16), w2 = __ldg(q + 32), w3 = __ldg(q + 48);
#pragma unroll
        for (int i = 0; i < 8; i++) T16_CTM(i, i + 8, w0.x, w0.y);
#pragma unroll
        for (int i = 0; i < 4; i++) { T16_CTM(i, i + 4, w0.z, w0.w); T16_CTM(i + 8, i + 12, w1.x, w1.y); }
#pragma unroll
        for (int i = 0; i < 2; i++) {
            T16_CTM(i, i + 2, w1.z, w1.w); T16_CTM(i + 4, i + 6, w2.x, w2.y);
            T16_CTM(i + 8, i + 10, w2.z, w2.w); T16_CTM(i + 12, i + 14, w3.x, w3.y);
        }
        T16_CTM(0, 1, w3.z, w3.w);
    }
    {
        const int4 w4 = __ldg(q + 64), w5 = __ldg(q + 80), w6 = __ldg(q + 96), w7 = __ldg(q + 112);
        T16_CTM(2, 3, w4.x, w4.y); T16_CTM(4, 5, w4.z, w4.w); T16_CTM(6, 7, w5.x, w5.y); T16_CTM(8, 9, w5.z, w5.w);
        T16_CTM(10, 11, w6.x, w6.y); T16_CTM(12, 13, w6.z, w6.w); T16_CTM(14, 15, w7.x, w7.y);
    }
}
#undef T16_CTM

// Inverse NTT WITH the 1/256, canonical output like the reference's (src/ntt.rs:85-161 returns full_reduce32 values, so
// only the residue is defined and the Shoup butterflies apply).  In: layout Y, canonical or |v| < 2 q.  Out: layout X.
// The 1/256 rides on the last level: difference outputs through the twiddle zeta / 256, sum outputs one product each.
__device__ __forceinline__ void sntt16_inv_scaled_Y2X(int32_t (&v)[16], int32_t* tile, int t)
{
    constexpr int32_t z0 = 0;
    const int4* q = &ZS16_INV.v[0][t];
    {
        const int4 w0 = __ldg(q), w1 = __ldg(q + 16), w2 = __ldg(q + 32), w3 = __ldg(q + 48);
        T16_GS(0, 1, w0.x, w0.y); T16_GS(2, 3, w0.z, w0.w); T16_GS(4, 5, w1.x, w1.y); T16_GS(6, 7, w1.z, w1.w);
        T16_GS(8, 9, w2.x, w2.y); T16_GS(10, 11, w2.z, w2.w); T16_GS(12, 13, w3.x, w3.y); T16_GS(14, 15, w3.z, w3.w);
    }
    {
        const int4 w4 = __ldg(q + 64), w5 = __ldg(q + 80), w6 = __ldg(q + 96), w7 = __ldg(q + 112);
#pragma unroll
        for (int i = 0; i < 2; i++) {
            T16_GS(i, i + 2, w4.x, w4.y); T16_GS(i + 4, i + 6, w4.z, w4.w);
            T16_GS(i + 8, i + 10, w5.x, w5.y); T16_GS(i + 12, i + 14, w5.z, w5.w);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) { T16_GS(i, i + 4, w6.x, w6.y); T16_GS(i + 8, i + 12, w6.z, w6.w); }
#pragma unroll
        for (int i = 0; i < 8; i++) T16_GS(i, i + 8, w7.x, w7.y);
    }
    v[0] = reduce32(v[0]);
    __syncwarp();  // (the tile's previous readers: the other lanes' layout-X loads of the last polynomial)
    t16_store_Y(tile, t, v);
    __syncwarp();
    t16_load_X(tile, t, v);
    T16_GS(0, 1, ZSI15, ZSI15S); T16_GS(2, 3, ZSI14, ZSI14S); T16_GS(4, 5, ZSI13, ZSI13S); T16_GS(6, 7, ZSI12, ZSI12S);
    T16_GS(8, 9, ZSI11, ZSI11S); T16_GS(10, 11, ZSI10, ZSI10S); T16_GS(12, 13, ZSI9, ZSI9S); T16_GS(14, 15, ZSI8, ZSI8S);
#pragma unroll
    for (int i = 0; i < 2; i++) {
        T16_GS(i, i + 2, ZSI7, ZSI7S); T16_GS(i + 4, i + 6, ZSI6, ZSI6S);
        T16_GS(i + 8, i + 10, ZSI5, ZSI5S); T16_GS(i + 12, i + 14, ZSI4, ZSI4S);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) { T16_GS(i, i + 4, ZSI3, ZSI3S); T16_GS(i + 8, i + 12, ZSI2, ZSI2S); }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int32_t a = v[i];
        v[i] = freeze(shoup_mul(a + v[i + 8], SC_INV256, SC_INV256S));
        v[i + 8] = freeze(shoup_mul(a - v[i + 8], ZSI1_256, ZSI1_256S));
    }
}

#undef T16_CT
#undef T16_GS

}  // namespace mldsa
