// k_fused.cu -- ExpandA fused with the matrix-vector product: A-hat is never written to memory.
//
// Keygen (t = A s1 + s2, src/ml_dsa.rs:78-92), Signer::get_public_key (src/ml_dsa.rs:502-563) and verification under
// a per-lane public key (w' = A z - c t1 2^d, src/ml_dsa.rs:360, :406-416) use every polynomial of A-hat exactly ONCE.
// Materialising it costs K L KiB of HBM writes and the same again in reads per lane (30 / 56 KiB for ML-DSA-65 / 87:
// 12 x the algorithmic bytes of a keygen), and the consumer kernel then waits on those reads.  Here the thread that
// squeezes a SHAKE128 stream (thread-per-stream Keccak, keccak.cuh) multiplies every accepted coefficient with the
// matching coefficient of s1-hat / z-hat from shared memory and adds the product into the row accumulator t-hat[k] in
// shared memory (one shared-memory atomic per coefficient: the FMA and LSU pipes this ALU-bound kernel leaves idle).
// The warp-per-polynomial arithmetic around it (forward NTTs before, inverse NTTs / rounding / packing after) runs in
// the same CTA on the same shared rows; CTAs of one SM drift out of phase, so the arithmetic of one hides under the
// Keccak of the others.
//
// A CTA pass handles FG = 4 keys: 4 K L streams = 64 / 120 / 224 threads (ML-DSA-44 / 65 / 87).  Thread t owns the
// stream of A-hat[k][l] of key g with (g, k) = t mod 4K and l = t / 4K, so the 32 threads of a warp add into (almost)
// distinct accumulator rows; rows are 257 words apart, so equal coefficient indices of different rows fall into
// different banks (streams reject 0.1 % of their candidates: the indices of a warp stay within one or two of each other).
#include <atomic>

#include "engine.h"
#include "keccak.cuh"
#include "ntt.cuh"
#include "pack.cuh"
#include "sampling.cuh"

namespace mldsa {

constexpr int FG = 4;      // keys per CTA pass
constexpr int FROW = 257;  // words per polynomial row in shared memory (skewed by one bank per row)

template <class P>
struct FusedCfg {
    static constexpr int STREAMS = FG * P::K * P::L;
    static constexpr int THREADS = (STREAMS + 31) / 32 * 32;
    static constexpr int WARPS = THREADS / 32;
    static constexpr int SROWS = FG * P::L, AROWS = FG * P::K;
    static constexpr int MINB = P::PS == 44 ? 6 : (P::PS == 65 ? 4 : 3);
    static_assert(((SROWS + AROWS) * FROW) % 4 == 0, "the NTT tiles behind the rows must stay 16-byte aligned");
};

// One SHAKE128 stream per thread: RejNTTPoly(rho || s || r) (src/hashing.rs:111-146, :225-239; CoeffFromThreeBytes,
// src/conversion.rs:40-61).  Coefficient j of the stream, a_j in [0, q), is consumed on the spot:
//     arow[j] += a_j * srow[j] * 2^-32 mod q          (srow carries the 2^32 and the 1/256 of the un-scaled inverse NTT)
// A 6th squeeze block (p ~ 1e-39) is handled by the same loop.  All 32 threads of a warp must call this.
__device__ __forceinline__ void expand_a_mac(bool valid, const uint8_t* __restrict__ rho, uint32_t s_col, uint32_t r_row,
                                             const int32_t* __restrict__ srow, int32_t* __restrict__ arow)
{
    uint64_t a[25];
#pragma unroll
    for (int i = 0; i < 25; i++) a[i] = 0;
    if (valid) {
        uint32_t w[8];
        load_span<32>(rho, w);
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
        a[4] = (uint64_t)(s_col | (r_row << 8) | (0x1Fu << 16));  // rho || s || r (hashing.rs:236), then the SHAKE pad
        a[20] = 0x8000000000000000ULL;                           // byte 167 of the 168-byte rate
    }
    // Running shared-memory byte addresses instead of an index: sp walks the s-hat row, the accumulator row sits at the
    // thread-constant distance `delta` from it.  Once a stream has its 256 coefficients, sp parks on the row's pad word
    // (index 256, never read), so the per-candidate test is `c < q` alone; address arithmetic is issued as IMAD (FMA
    // pipe, idle here) -- the ALU pipe, which bounds this kernel, sees the candidate extraction and two ops per candidate.
    uint32_t sp = (uint32_t)__cvta_generic_to_shared(srow);
    const uint32_t end = sp + 256 * 4, delta = (uint32_t)__cvta_generic_to_shared(arow) - sp;
    if (!valid) sp = end;
    while (__any_sync(0xffffffffu, sp < end)) {
        keccak_f1600(a);
        // Branch-free candidate: the product is formed and accumulated for EVERY candidate -- a rejected one (c >= q,
        // 0.1 % of them) adds zero to the slot of the next coefficient and does not advance sp -- so ptxas can overlap
        // the shared-memory and IMAD latencies of neighbouring candidates instead of fencing each one with a
        // reconvergence point (a predicated accumulation is turned back into a branch).
#define FA_TAKE(c)                                                                                  \
        asm volatile("{\n\t"                                                                        \
                     ".reg .pred p;\n\t"                                                            \
                     ".reg .s32 sv, lo, tq, h1, h2;\n\t"                                            \
                     ".reg .u32 ap, nx;\n\t"                                                        \
                     "setp.lt.u32 p, %1, 8380417;\n\t"                                              \
                     "ld.shared.s32 sv, [%0];\n\t"                                                  \
                     "mad.lo.u32 ap, %0, 1, %2;\n\t"                                                \
                     "mul.lo.s32 lo, %1, sv;\n\t"                                                   \
                     "mul.lo.s32 tq, lo, 58728449;\n\t"                                             \
                     "mul.hi.s32 h1, %1, sv;\n\t"                                                   \
                     "mul.hi.s32 h2, tq, 8380417;\n\t"                                              \
                     "sub.s32 h1, h1, h2;\n\t"                                                      \
                     "selp.s32 h1, h1, 0, p;\n\t"                                                   \
                     "red.shared.add.s32 [ap], h1;\n\t"                                             \
                     "add.u32 nx, %0, 4;\n\t"                                                       \
                     "min.u32 nx, nx, %3;\n\t"                                                      \
                     "@p mov.u32 %0, nx;\n\t"                                                       \
                     "}"                                                                            \
                     : "+r"(sp)                                                                     \
                     : "r"(c), "r"(delta), "r"(end)                                                 \
                     : "memory");
#pragma unroll
        for (int g = 0; g < 14; g++) {  // 14 groups of 4 candidates = 3 words each
            const uint64_t l0 = a[(3 * g) >> 1], l1 = a[((3 * g) >> 1) + 1];
            uint32_t w0, w1, w2;
            if ((g & 1) == 0) { w0 = (uint32_t)l0; w1 = (uint32_t)(l0 >> 32); w2 = (uint32_t)l1; }
            else              { w0 = (uint32_t)(l0 >> 32); w1 = (uint32_t)l1; w2 = (uint32_t)(l1 >> 32); }
            const uint32_t c0 = w0 & 0x7FFFFF;
            const uint32_t c1 = __funnelshift_r(w0, w1, 24) & 0x7FFFFF;
            const uint32_t c2 = __funnelshift_r(w1, w2, 16) & 0x7FFFFF;
            const uint32_t c3 = (w2 >> 8) & 0x7FFFFF;
            FA_TAKE(c0) FA_TAKE(c1) FA_TAKE(c2) FA_TAKE(c3)
        }
#undef FA_TAKE
    }
}

// layout-B registers (lane t owns coefficients 8t .. 8t+7) <-> a skewed shared row
__device__ __forceinline__ void frow_store(int32_t* row, int lane, const int32_t (&v)[8])
{
#pragma unroll
    for (int i = 0; i < 8; i++) row[8 * lane + i] = v[i];
}
__device__ __forceinline__ void frow_load(const int32_t* row, int lane, int32_t (&v)[8])
{
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = row[8 * lane + i];
}

// The dynamic shared-memory limit is a per-device function attribute.
template <class Kernel>
static cudaError_t ensure_smem(Kernel kernel, size_t bytes, std::atomic<uint64_t>& done)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 64 || !((done.load(std::memory_order_acquire) >> dev) & 1)) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        if (dev < 64) done.fetch_or(uint64_t(1) << dev, std::memory_order_release);
    }
    return cudaSuccess;
}

// ------------------------------------------------------------------------------------------------
// Keygen arithmetic with A-hat on the fly.  FROM_SK = false: s1, s2 are the int8 polynomials RejBoundedPoly wrote
// (k_keygen.cu), outputs are pk and the sk fields rho, K, s1, s2, t0 (tr follows in its own kernel).  FROM_SK = true:
// s1, s2 are unpacked from packed secret keys and only pk is written (Signer::get_public_key).
// ------------------------------------------------------------------------------------------------
template <class P, bool FROM_SK>
__global__ void __launch_bounds__(FusedCfg<P>::THREADS, FusedCfg<P>::MINB)
kg_fused_kernel(size_t n, const uint8_t* __restrict__ seeds, const int8_t* __restrict__ sall, const uint8_t* __restrict__ skin,
                uint8_t* __restrict__ pk, uint8_t* __restrict__ sk)
{
    using LO = Layout<P>;
    using CF = FusedCfg<P>;
    constexpr int K = P::K, L = P::L;
    extern __shared__ __align__(16) int32_t fsm[];
    int32_t* const srows = fsm;
    int32_t* const arows = srows + CF::SROWS * FROW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t* const tile = arows + CF::AROWS * FROW + warp * NTT_TILE_WORDS;

    for (size_t g0 = (size_t)blockIdx.x * FG; g0 < n; g0 += (size_t)gridDim.x * FG) {
        // ---- phase 0: s1-hat rows (scaled for the Montgomery accumulation), cleared accumulators, seed fields ----
        for (int i = threadIdx.x; i < CF::AROWS * FROW; i += CF::THREADS) arows[i] = 0;
        for (int p = warp; p < FG * L; p += CF::WARPS) {
            const int g = p / L, l = p % L;
            const size_t key = g0 + g;
            if (key >= n) continue;
            int32_t v[8];
            if (FROM_SK) {
                uint32_t w[(P::ETABITS + 3) / 4];
                load_span<P::ETABITS>(skin + key * (size_t)P::SK_LEN + LO::SK_S1 + l * 32 * P::ETABITS + P::ETABITS * lane, w);
#pragma unroll
                for (int i = 0; i < 8; i++) v[i] = P::ETA - (int32_t)get_field<P::ETABITS>(w, i);
            } else {
                const uint2 raw = *reinterpret_cast<const uint2*>(sall + (key * (size_t)(K + L) + l) * 256 + 8 * lane);
                uint32_t f[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int32_t c = (int32_t)(int8_t)(((i < 4 ? raw.x : raw.y) >> (8 * (i & 3))) & 0xFF);
                    v[i] = c;
                    f[i] = (uint32_t)(P::ETA - c);  // BitPack(s1, eta, eta): field = eta - s   (conversion.rs:164)
                }
                uint32_t w[(P::ETABITS + 3) / 4];
                pack_fields<P::ETABITS>(f, w);
                store_span<P::ETABITS>(sk + key * (size_t)P::SK_LEN + LO::SK_S1 + l * 32 * P::ETABITS + P::ETABITS * lane, w);
            }
            sntt_fwd_B2B(v, tile, lane);
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = shoup_mul(v[i], SC_MONT_INV256, SC_MONT_INV256S);
            frow_store(srows + p * FROW, lane, v);
        }
        for (int g = warp; g < FG; g += CF::WARPS) {
            const size_t key = g0 + g;
            if (key >= n) continue;
            if (FROM_SK) {
                pk[key * (size_t)P::PK_LEN + lane] = skin[key * (size_t)P::SK_LEN + lane];  // rho
            } else {  // rho -> pk[0..32], sk[0..32]; K -> sk[32..64]          (encodings.rs:27, :112-113)
                const uint8_t* sd = seeds + 128 * key;
                pk[key * (size_t)P::PK_LEN + lane] = sd[lane];
                sk[key * (size_t)P::SK_LEN + lane] = sd[lane];
                sk[key * (size_t)P::SK_LEN + 32 + lane] = sd[96 + lane];
            }
        }
        __syncthreads();
        // ---- phase 1: every stream of the FG keys' A-hat, multiplied and accumulated as it is squeezed ----
        {
            const int t = threadIdx.x;
            const int l = t / CF::AROWS, ra = t % CF::AROWS, g = ra / K, k = ra % K;
            const size_t key = g0 + g;
            const bool valid = t < CF::STREAMS && key < n;
            const uint8_t* rho = FROM_SK ? skin + (valid ? key : 0) * (size_t)P::SK_LEN : seeds + 128 * (valid ? key : 0);
            expand_a_mac(valid, rho, (uint32_t)l, (uint32_t)k, srows + (valid ? g * L + l : 0) * FROW, arows + ra * FROW);
        }
        __syncthreads();
        // ---- phase 2: t = INTT(t-hat) + s2, Power2Round, pkEncode / skEncode                ml_dsa.rs:88-92 ----
        for (int p = warp; p < FG * K; p += CF::WARPS) {
            const int g = p / K, k = p % K;
            const size_t key = g0 + g;
            if (key >= n) continue;
            int32_t t[8];
            frow_load(arows + p * FROW, lane, t);
#pragma unroll
            for (int i = 0; i < 8; i++) t[i] = reduce32(t[i]);
            sntt_inv_B2B(t, tile, lane);
            int32_t s2[8];
            if (FROM_SK) {
                uint32_t w2[(P::ETABITS + 3) / 4];
                load_span<P::ETABITS>(skin + key * (size_t)P::SK_LEN + LO::SK_S2 + k * 32 * P::ETABITS + P::ETABITS * lane, w2);
#pragma unroll
                for (int i = 0; i < 8; i++) s2[i] = P::ETA - (int32_t)get_field<P::ETABITS>(w2, i);
            } else {
                const uint2 raw = *reinterpret_cast<const uint2*>(sall + (key * (size_t)(K + L) + L + k) * 256 + 8 * lane);
#pragma unroll
                for (int i = 0; i < 8; i++) s2[i] = (int32_t)(int8_t)(((i < 4 ? raw.x : raw.y) >> (8 * (i & 3))) & 0xFF);
            }
            uint32_t fs[8], f1[8], f0[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                fs[i] = (uint32_t)(P::ETA - s2[i]);
                const int32_t tt = freeze(t[i] + s2[i]);                // ml_dsa.rs:88-91
                const int32_t r1 = (tt + (1 << (D - 1)) - 1) >> D;      // high_low.rs:26-27
                const int32_t r0 = tt - (r1 << D);                      // high_low.rs:30-31
                f1[i] = (uint32_t)r1;
                f0[i] = (uint32_t)((1 << (D - 1)) - r0);                // BitPack(t0, 2^12 - 1, 2^12)
            }
            uint32_t w1[3];
            pack_fields<10>(f1, w1);
            store_span<10>(pk + key * (size_t)P::PK_LEN + 32 + 320 * k + 10 * lane, w1);
            if (!FROM_SK) {
                uint8_t* sko = sk + key * (size_t)P::SK_LEN;
                uint32_t ws[(P::ETABITS + 3) / 4], w0[4];
                pack_fields<P::ETABITS>(fs, ws);
                store_span<P::ETABITS>(sko + LO::SK_S2 + k * 32 * P::ETABITS + P::ETABITS * lane, ws);
                pack_fields<13>(f0, w0);
                store_span<13>(sko + LO::SK_T0 + 416 * k + 13 * lane, w0);
            }
        }
        // s1-hat must not outlive the pass in shared memory any longer than needed (and the next pass overwrites it)
        __syncthreads();
    }
    // scrub what is left of the last pass (s1-hat rows, t-hat accumulators): nothing secret-derived stays in shared memory
    for (int i = threadIdx.x; i < (CF::SROWS + CF::AROWS) * FROW; i += CF::THREADS) fsm[i] = 0;
}

template <class P>
static size_t kg_fused_smem()
{
    using CF = FusedCfg<P>;
    return ((size_t)(CF::SROWS + CF::AROWS) * FROW + (size_t)CF::WARPS * NTT_TILE_WORDS) * 4;
}

template <class P, bool FROM_SK>
static cudaError_t launch_kg_fused_t(cudaStream_t st, size_t n, const uint8_t* seeds, const int8_t* s, const uint8_t* skin,
                                     uint8_t* pk, uint8_t* sk)
{
    using CF = FusedCfg<P>;
    static std::atomic<uint64_t> done{0};
    const size_t smem = kg_fused_smem<P>();
    cudaError_t e = ensure_smem(kg_fused_kernel<P, FROM_SK>, smem, done);
    if (e != cudaSuccess) return e;
    const unsigned grid = capped_grid((n + FG - 1) / FG, tuning().kcap > 0 ? tuning().kcap : 4 * CF::MINB);
    {
        KLaunch kl(FROM_SK ? KID_SK2PK : KID_KG_ARITH, st);
        kg_fused_kernel<P, FROM_SK><<<grid, CF::THREADS, smem, st>>>(n, seeds, s, skin, pk, sk);
    }
    return cudaGetLastError();
}

cudaError_t launch_kg_fused(int ps, cudaStream_t st, size_t n, const uint8_t* seeds, const int8_t* s, uint8_t* pk, uint8_t* sk)
{
    if (n == 0) return cudaSuccess;
    switch (ps) {
    case 44: return launch_kg_fused_t<Params<44>, false>(st, n, seeds, s, nullptr, pk, sk);
    case 65: return launch_kg_fused_t<Params<65>, false>(st, n, seeds, s, nullptr, pk, sk);
    case 87: return launch_kg_fused_t<Params<87>, false>(st, n, seeds, s, nullptr, pk, sk);
    }
    return cudaErrorInvalidValue;
}
cudaError_t launch_sk_to_pk_fused(int ps, cudaStream_t st, size_t n, const uint8_t* skin, uint8_t* pk)
{
    if (n == 0) return cudaSuccess;
    switch (ps) {
    case 44: return launch_kg_fused_t<Params<44>, true>(st, n, nullptr, nullptr, skin, pk, nullptr);
    case 65: return launch_kg_fused_t<Params<65>, true>(st, n, nullptr, nullptr, skin, pk, nullptr);
    case 87: return launch_kg_fused_t<Params<87>, true>(st, n, nullptr, nullptr, skin, pk, nullptr);
    }
    return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------------------------------
// Verification under a per-lane packed public key with A-hat on the fly (src/ml_dsa.rs:351-437 between the mu /
// SampleInBall hashes of verify_pre and the commitment-hash comparison of ctilde):
//   phase 0   HintBitUnpack + its malformed exits, z unpack + norm check + NTT, c-hat, accumulators = -c-hat o t1-hat
//   phase 1   accumulators += A-hat o z-hat, stream by stream
//   phase 2   w' = INTT(accumulator), UseHint, w1Encode into the mu || w1' row
// ------------------------------------------------------------------------------------------------
template <class P>
struct VerifyFusedExtra {
    alignas(16) int32_t chat[FG][256];                  // c-hat per lane, layout-B order
    uint32_t hbits[FG][P::K * 8];                       // hint bitmaps, 256 bits per polynomial
    uint32_t stage[FG][(P::OMEGA + P::K + 3) / 4 + 1];  // hint bytes
    uint32_t bad[FG];
};

template <class P>
__global__ void __launch_bounds__(FusedCfg<P>::THREADS, FusedCfg<P>::MINB)
verify_fused_kernel(size_t n, const uint8_t* __restrict__ pk, size_t pk_stride, const uint8_t* __restrict__ sig,
                    const uint8_t* __restrict__ cmask, uint8_t* __restrict__ mw1, uint8_t* __restrict__ pre_ok)
{
    using LO = Layout<P>;
    using CF = FusedCfg<P>;
    constexpr int K = P::K, L = P::L;
    extern __shared__ __align__(16) int32_t fsm[];
    int32_t* const srows = fsm;
    int32_t* const arows = srows + CF::SROWS * FROW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t* const tile = arows + CF::AROWS * FROW + warp * NTT_TILE_WORDS;
    VerifyFusedExtra<P>& ex = *reinterpret_cast<VerifyFusedExtra<P>*>(arows + CF::AROWS * FROW + CF::WARPS * NTT_TILE_WORDS);

    for (size_t g0 = (size_t)blockIdx.x * FG; g0 < n; g0 += (size_t)gridDim.x * FG) {
        // ---- phase 0a: per lane: hints (conversion.rs:340-414) and c-hat ----
        for (int g = warp; g < FG; g += CF::WARPS) {
            const size_t ln = g0 + g;
            if (ln >= n) continue;
            const uint8_t* sg = sig + ln * (size_t)P::SIG_LEN;
            uint8_t* hb = reinterpret_cast<uint8_t*>(ex.stage[g]);
            warp_copy_bytes(hb, sg + LO::SIG_H, P::OMEGA + K, lane);
            if (lane < K * 8) ex.hbits[g][lane] = 0;
            if (K * 8 > 32 && lane + 32 < K * 8) ex.hbits[g][lane + 32] = 0;
            __syncwarp();
            bool bad = false;
            if (lane < K) {
                const int cur = hb[P::OMEGA + lane], prev = lane ? hb[P::OMEGA + lane - 1] : 0;
                bad = cur < prev || cur > P::OMEGA;                                   // conversion.rs:357
            }
            bad = __any_sync(0xffffffffu, bad);
            if (!bad) {
                const int total = hb[P::OMEGA + K - 1];
                for (int p = lane; p < P::OMEGA; p += 32) {
                    if (p >= total) {
                        bad |= hb[p] != 0;                                            // conversion.rs:398
                    } else {
                        int poly = 0, first = 0;
#pragma unroll
                        for (int k = 0; k < K; k++) {
                            const int c = hb[P::OMEGA + k];
                            if (p >= c) { poly = k + 1; first = c; }
                        }
                        if (p > first) bad |= hb[p - 1] >= hb[p];                     // conversion.rs:373
                        atomicOr(&ex.hbits[g][poly * 8 + (hb[p] >> 5)], 1u << (hb[p] & 31));
                    }
                }
                bad = __any_sync(0xffffffffu, bad);
            }
            if (lane == 0) ex.bad[g] = bad ? 1u : 0u;
            int32_t ch[8];
            load_c_layoutA(cmask + ln * (size_t)CMASK_STRIDE, lane, ch);
            sntt_fwd_A2B(ch, tile, lane);
#pragma unroll
            for (int i = 0; i < 8; i++) ch[i] = reduce32(ch[i]);
            int4* cp = reinterpret_cast<int4*>(ex.chat[g] + 8 * lane);
            cp[0] = make_int4(ch[0], ch[1], ch[2], ch[3]);
            cp[1] = make_int4(ch[4], ch[5], ch[6], ch[7]);
        }
        __syncthreads();
        // ---- phase 0b: z = BitUnpack(gamma1 - 1, gamma1), ||z||inf check, z-hat rows ----
        for (int p = warp; p < FG * L; p += CF::WARPS) {
            const int g = p / L, l = p % L;
            const size_t ln = g0 + g;
            if (ln >= n) continue;
            uint32_t w[(P::ZBITS + 3) / 4];
            load_span<P::ZBITS>(sig + ln * (size_t)P::SIG_LEN + LO::SIG_Z + l * 32 * P::ZBITS + P::ZBITS * lane, w);
            int32_t zv[8];
            bool zbad = false;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int32_t z = P::GAMMA1 - (int32_t)get_field<P::ZBITS>(w, i);  // conversion.rs:250
                zbad |= (z < 0 ? -z : z) >= P::GAMMA1 - P::BETA;                   // ml_dsa.rs:434
                zv[i] = z;
            }
            if (__any_sync(0xffffffffu, zbad) && lane == 0) atomicOr(&ex.bad[g], 1u);
            sntt_fwd_B2B(zv, tile, lane);
#pragma unroll
            for (int i = 0; i < 8; i++) zv[i] = shoup_mul(zv[i], SC_MONT_INV256, SC_MONT_INV256S);
            frow_store(srows + p * FROW, lane, zv);
        }
        // ---- phase 0c: accumulator rows start at -c-hat o NTT(t1 2^d) (pk_decode + ml_dsa.rs:490-495, :406-416) ----
        for (int p = warp; p < FG * K; p += CF::WARPS) {
            const int g = p / K, k = p % K;
            const size_t ln = g0 + g;
            int32_t t[8];
#pragma unroll
            for (int i = 0; i < 8; i++) t[i] = 0;
            if (ln < n) {
                uint32_t w[3];
                load_span<10>(pk + ln * pk_stride + 32 + 320 * k + 10 * lane, w);
                int32_t t1v[8];
#pragma unroll
                for (int i = 0; i < 8; i++) t1v[i] = (int32_t)(get_field<10>(w, i) << D);
                sntt_fwd_B2B(t1v, tile, lane);
                const int4* cp = reinterpret_cast<const int4*>(ex.chat[g] + 8 * lane);
                const int4 x = cp[0], y = cp[1];
                const int32_t ch[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int32_t ts = shoup_mul(t1v[i], SC_MONT_INV256, SC_MONT_INV256S);
                    t[i] = mont_reduce(-(int64_t)ch[i] * ts);
                }
            }
            frow_store(arows + p * FROW, lane, t);
        }
        __syncthreads();
        // ---- phase 1 ----
        {
            const int t = threadIdx.x;
            const int l = t / CF::AROWS, ra = t % CF::AROWS, g = ra / K, k = ra % K;
            const size_t ln = g0 + g;
            const bool valid = t < CF::STREAMS && ln < n;
            expand_a_mac(valid, pk + (valid ? ln : 0) * pk_stride, (uint32_t)l, (uint32_t)k,
                         srows + (valid ? g * L + l : 0) * FROW, arows + ra * FROW);
        }
        __syncthreads();
        // ---- phase 2: w' = INTT(.), UseHint, w1Encode (ml_dsa.rs:416-425, encodings.rs:338-360) ----
        for (int p = warp; p < FG * K; p += CF::WARPS) {
            const int g = p / K, k = p % K;
            const size_t ln = g0 + g;
            if (ln >= n) continue;
            if (ex.bad[g]) {  // verify returns false (ml_dsa.rs:368-376, :434)
                if (k == 0 && lane == 0) pre_ok[ln] = 0;
                continue;
            }
            int32_t t[8];
            frow_load(arows + p * FROW, lane, t);
#pragma unroll
            for (int i = 0; i < 8; i++) t[i] = reduce32(t[i]);
            sntt_inv_B2B(t, tile, lane);
            const uint32_t hbyte = (ex.hbits[g][k * 8 + (lane >> 2)] >> (8 * (lane & 3))) & 0xFF;
            uint32_t f[8];
#pragma unroll
            for (int i = 0; i < 8; i++) f[i] = (uint32_t)use_hint<P::GAMMA2>((hbyte >> i) & 1, t[i]);
            uint32_t w[(P::W1BITS + 3) / 4];
            pack_fields<P::W1BITS>(f, w);
            store_span<P::W1BITS>(mw1 + ln * (size_t)(MU_BYTES + P::W1_LEN) + MU_BYTES + k * 32 * P::W1BITS + P::W1BITS * lane, w);
        }
        __syncthreads();
    }
}

template <class P>
static cudaError_t launch_verify_fused_t(cudaStream_t st, size_t n, const uint8_t* pk, size_t pk_stride, const uint8_t* sig,
                                         const uint8_t* cmask, uint8_t* mw1, uint8_t* pre_ok)
{
    using CF = FusedCfg<P>;
    static std::atomic<uint64_t> done{0};
    const size_t smem = kg_fused_smem<P>() + sizeof(VerifyFusedExtra<P>);
    cudaError_t e = ensure_smem(verify_fused_kernel<P>, smem, done);
    if (e != cudaSuccess) return e;
    const unsigned grid = capped_grid((n + FG - 1) / FG, tuning().kcap > 0 ? tuning().kcap : 4 * CF::MINB);
    {
        KLaunch kl(KID_VERIFY_ARITH, st);
        verify_fused_kernel<P><<<grid, CF::THREADS, smem, st>>>(n, pk, pk_stride, sig, cmask, mw1, pre_ok);
    }
    return cudaGetLastError();
}

cudaError_t launch_verify_fused(int ps, cudaStream_t st, size_t n, const uint8_t* pk, size_t pk_stride, const uint8_t* sig,
                                const uint8_t* cmask, uint8_t* mw1, uint8_t* pre_ok)
{
    if (n == 0) return cudaSuccess;
    switch (ps) {
    case 44: return launch_verify_fused_t<Params<44>>(st, n, pk, pk_stride, sig, cmask, mw1, pre_ok);
    case 65: return launch_verify_fused_t<Params<65>>(st, n, pk, pk_stride, sig, cmask, mw1, pre_ok);
    case 87: return launch_verify_fused_t<Params<87>>(st, n, pk, pk_stride, sig, cmask, mw1, pre_ok);
    }
    return cudaErrorInvalidValue;
}

}  // namespace mldsa
