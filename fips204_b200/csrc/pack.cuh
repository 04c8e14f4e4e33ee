// pack.cuh -- lane-local bit (un)packing in register layout B (lane t owns coefficients 8t..8t+7)
// plus warp-cooperative byte movers for byte-aligned regions of keys and signatures.
//
// Replaces src/conversion.rs:143-182 (bit_pack), :227-262 (bit_unpack), :277-328 (hint_bit_pack),
// :340-414 (hint_bit_unpack) and the layouts of src/encodings.rs.  All packing is LSB-first,
// little-endian, c bits per coefficient, so 8 coefficients are exactly c bytes: a lane's byte
// window is [c*t, c*t + c) of the 32*c-byte polynomial and no cross-lane traffic is needed.
#pragma once
#include "field.cuh"

namespace mldsa {

// Read NB bytes at an arbitrarily aligned address into little-endian words.  Never touches an
// aligned word that does not contain at least one requested byte.
template <int NB>
__device__ __forceinline__ void load_span(const uint8_t* p, uint32_t (&w)[(NB + 3) / 4])
{
    constexpr int NW = (NB + 3) / 4;
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const unsigned off = (unsigned)(a & 3);
    const uint32_t* q = reinterpret_cast<const uint32_t*>(a - off);
    uint32_t x[NW + 1];
#pragma unroll
    for (int k = 0; k < NW; k++) x[k] = __ldg(q + k);
    x[NW] = (off + NB > 4 * NW) ? __ldg(q + NW) : 0u;
#pragma unroll
    for (int k = 0; k < NW; k++) w[k] = __funnelshift_r(x[k], x[k + 1], off * 8);
}

// field i (compile-time) of width BITS from a little-endian word string
template <int BITS>
__device__ __forceinline__ uint32_t get_field(const uint32_t* w, int i)
{
    const int bit = BITS * i, k = bit >> 5, s = bit & 31;
    const uint32_t v = (s + BITS <= 32) ? (w[k] >> s) : __funnelshift_r(w[k], w[k + 1], s);
    return v & ((1u << BITS) - 1);
}
// 8 fields of BITS bits -> BITS bytes = ceil(BITS/4) words (upper bytes of the last word are zero)
template <int BITS>
__device__ __forceinline__ void pack_fields(const uint32_t (&f)[8], uint32_t (&w)[(BITS + 3) / 4])
{
#pragma unroll
    for (int k = 0; k < (BITS + 3) / 4; k++) w[k] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int bit = BITS * i, k = bit >> 5, s = bit & 31;
        w[k] |= f[i] << s;
        if (s + BITS > 32) w[k + 1] |= f[i] >> (32 - s);
    }
}

// Store a lane's NB packed bytes (little-endian words w) at dst.  Picks the widest store the
// alignment of dst and NB allows; the choice is resolved per call site, not per byte.
template <int NB>
__device__ __forceinline__ void store_span(uint8_t* dst, const uint32_t (&w)[(NB + 3) / 4])
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(dst);
    if (NB % 4 == 0 && (a & 3) == 0) {
#pragma unroll
        for (int k = 0; k < NB / 4; k++) reinterpret_cast<uint32_t*>(dst)[k] = w[k];
    } else if (NB % 2 == 0 && (a & 1) == 0) {
#pragma unroll
        for (int k = 0; k < NB / 2; k++)
            reinterpret_cast<uint16_t*>(dst)[k] = (uint16_t)(w[k >> 1] >> (16 * (k & 1)));
    } else {
#pragma unroll
        for (int k = 0; k < NB; k++) dst[k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
    }
}

// A lane's 8 coefficients (32 bytes, 32-byte aligned: one sector) in ONE 256-bit access (sm_100: LDG / STG .256).
// As 2 x 128-bit each store instruction of a warp filled half of every sector it touched, i.e. two write requests per
// sector towards L2 (ncu, sign_key: 33 GB leave L1 for 18 GB of rows).  Used for the streamed per-lane rows; rows that
// are meant to stay hot in L1 (a shared key's A-hat / s-hat) keep the 128-bit __ldg path.  -DNO_V8: the old form.
struct Row8 { int32_t v[8]; };
__device__ __forceinline__ Row8 ld_row8(const int32_t* p)  // coherent (data this kernel also writes)
{
    Row8 r;
#ifndef NO_V8
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
#else
    const int4 x = reinterpret_cast<const int4*>(p)[0], y = reinterpret_cast<const int4*>(p)[1];
    r.v[0] = x.x; r.v[1] = x.y; r.v[2] = x.z; r.v[3] = x.w; r.v[4] = y.x; r.v[5] = y.y; r.v[6] = y.z; r.v[7] = y.w;
#endif
    return r;
}
__device__ __forceinline__ Row8 ldg_row8(const int32_t* p)  // read-only for the lifetime of the kernel
{
    Row8 r;
#ifndef NO_V8
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
#else
    const int4 x = __ldg(reinterpret_cast<const int4*>(p)), y = __ldg(reinterpret_cast<const int4*>(p) + 1);
    r.v[0] = x.x; r.v[1] = x.y; r.v[2] = x.z; r.v[3] = x.w; r.v[4] = y.x; r.v[5] = y.y; r.v[6] = y.z; r.v[7] = y.w;
#endif
    return r;
}
__device__ __forceinline__ void st_row8(int32_t* p, const int32_t (&v)[8])
{
#ifndef NO_V8
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]),
                 "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
#else
    reinterpret_cast<int4*>(p)[0] = make_int4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<int4*>(p)[1] = make_int4(v[4], v[5], v[6], v[7]);
#endif
}

// 16 contiguous coefficients of a lane (ntt16.cuh, layout Y: 64 bytes = two sectors) as two 256-bit accesses
__device__ __forceinline__ void st_row16(int32_t* p, const int32_t (&v)[16])
{
    const int32_t a[8] = {v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]};
    const int32_t b[8] = {v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]};
    st_row8(p, a);
    st_row8(p + 8, b);
}
// Warp-cooperative L2 prefetch of `bytes` starting at p (128-byte lines): the streamed per-lane operands
// (rows of A-hat, s-hat, w) are pulled from HBM into L2 while the warp is busy with an NTT, so the
// demand loads that follow see L2 latency instead of DRAM latency.
__device__ __forceinline__ void warp_prefetch_l2(const void* p, int bytes, int lane)
{
    const char* c = reinterpret_cast<const char*>(p);
    for (int off = lane * 128; off < bytes; off += 32 * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(c + off));
}

// Warp-cooperative zero fill of a 16-byte aligned region whose size is a multiple of 16 bytes
__device__ __forceinline__ void warp_zero16(void* p, int bytes, int lane)
{
    int4* q = reinterpret_cast<int4*>(p);
    for (int i = lane; i < bytes / 16; i += 32) q[i] = make_int4(0, 0, 0, 0);
}

// Warp-cooperative byte copies (any alignment, n small: hint regions, c~, seeds)
__device__ __forceinline__ void warp_copy_bytes(uint8_t* dst, const uint8_t* src, int n, int lane)
{
    for (int i = lane; i < n; i += 32) dst[i] = src[i];
}

// Warp-cooperative copy of n bytes from a word-aligned shared staging buffer to an arbitrarily
// aligned global address: aligned interior goes out as 32-bit stores (funnel-shifted), the <= 3
// head/tail bytes as byte stores.  The staging buffer must be readable one word past n.
__device__ __forceinline__ void warp_copy_out(uint8_t* dst, const uint32_t* stage, int n, int lane)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(dst);
    const int nh = (int)((4 - (a & 3)) & 3);          // head bytes until dst is word aligned
    const uint8_t* sb = reinterpret_cast<const uint8_t*>(stage);
    if (lane < nh && lane < n) dst[lane] = sb[lane];
    if (n <= nh) return;
    const int nw = (n - nh) >> 2;
    uint32_t* dw = reinterpret_cast<uint32_t*>(dst + nh);
    for (int j = lane; j < nw; j += 32) {
        // source bytes nh + 4j .. nh + 4j + 3
        dw[j] = __funnelshift_r(stage[j], stage[j + 1], nh * 8);
    }
    const int done = nh + 4 * nw;
    if (lane < n - done) dst[done + lane] = sb[done + lane];
}

}  // namespace mldsa
