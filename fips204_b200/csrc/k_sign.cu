// k_sign.cu -- batched ML-DSA.Sign_internal with the rejection loop run as lock-step rounds over a
// compacted list of still-active lanes.
// Replaces src/ml_dsa.rs:153-337 (sign_internal) with expand_private (ml_dsa.rs:445-469) and
// sk_decode / sig_encode (src/encodings.rs:168-224, :238-276) folded in.
//
// A lane (one (sk, message, rnd)) is owned by ONE WARP in the arithmetic kernels, so a rejection is a
// warp-uniform branch: nothing diverges inside a warp.  After every round the rejected lanes are
// re-queued into a dense list (compact_kernel) and only those are launched again; lanes finish in
// any order, and a lane's bytes depend on its own inputs only, so scheduling never affects results.
//
//   once per chunk
//     sign_prep     thread/lane  mu = H(tr || M'), rho'' = H(K || rnd || mu)        ml_dsa.rs:185-201
//     sign_key      warp/key     s1-hat, s2-hat, t0-hat from the packed sk           ml_dsa.rs:447-453
//     expand_a      thread/poly  A-hat                                               ml_dsa.rs:181
//   per round, over the active list; every active lane tries D consecutive candidates (kappa, kappa + L, ...)
//   in one round ("depth", speculative: candidate j only counts if candidates 0..j-1 were rejected, exactly the
//   reference's sequential order).  The D candidates of a lane run on neighbouring warps and share the lane's
//   A-hat through L2, which halves (D = 2) the HBM traffic of the matrix-vector product and the round count.
//     mask          thread/poly  y bytes = H(rho'' || LE16(kappa + r), 32c)          hashing.rs:281-313
//     sign_w        warp/lane    w = INTT(A-hat o NTT(y)); w1 = HighBits(w); w1Encode ml_dsa.rs:215-232
//     ctilde        thread/lane  c~ = H(mu || w1), c = SampleInBall(c~) as bit masks   ml_dsa.rs:233-237
//     sign_z        warp/lane    c-hat, r0 / z / ct0 / hint checks, sigEncode          ml_dsa.rs:240-336
//     compact       thread/lane  rebuild the active list
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "engine.h"
#include "keccak.cuh"
#include "ntt.cuh"
#include "ntt16.cuh"
#include "pack.cuh"
#include "sampling.cuh"
#include "ctilde.cuh"
#include "tma.cuh"

namespace mldsa {

template <class P>
__global__ void __launch_bounds__(128)
sign_prep_kernel(size_t n, const uint8_t* __restrict__ sk, size_t sk_stride, MsgDesc md,
                 const uint8_t* __restrict__ rnd, uint8_t* __restrict__ mu_out, uint8_t* __restrict__ rhopp,
                 uint32_t* __restrict__ kappa, uint8_t* __restrict__ done, uint32_t* __restrict__ act0,
                 uint32_t* __restrict__ ctl, int* __restrict__ ctx_err, const uint32_t* __restrict__ kidx)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == 0) ctl[0] = (uint32_t)n;  // live count of the first active list
    const uint8_t* key = sk + (kidx ? (size_t)kidx[i] : i) * sk_stride;  // rho || K || tr || ... (only K and tr are read)
    Sponge<SHAKE256_RATE> sp;
    if (!absorb_mu(sp, key + 64, md, i)) *ctx_err = 1;  // lib.rs:274 -> whole call fails with -2
    uint64_t mu[8];
    sp.out_words(mu, 8);
    {  // sign_w copies mu in front of every candidate's w1 (the input row of the c~ hash)
        uint64_t* mo = reinterpret_cast<uint64_t*>(mu_out + i * (size_t)MU_BYTES);
#pragma unroll
        for (int k = 0; k < 8; k++) mo[k] = mu[k];
    }
    sp.init();                                            // rho'' = H(K || rnd || mu, 64)  ml_dsa.rs:199-201
    sp.absorb(key + 32, 32);
    sp.absorb(rnd + 32 * i, 32);
    sp.absorb(reinterpret_cast<const uint8_t*>(mu), 64);
    sp.finish();
    sp.out_words(reinterpret_cast<uint64_t*>(rhopp + 64 * i), 8);
    kappa[i] = 0;
    done[i] = 0;
    act0[i] = (uint32_t)i;
}

constexpr int SG_WARPS = 4;
#ifndef SZ_MINB
#define SZ_MINB 5   // sign_z: 96 registers hold the early loads of the w row / y without spills (6: 80 registers, 30.0 ms; 5: 29.5 ms)
#endif

// s2-hat, s1-hat, t0-hat for every key: rows [0, K), [K, K + L), [K + L, 2K + L) of the key's block of sh
// (SignWs::sh: the order sign_z walks them).  Warp per key, TWO polynomials per warp (ntt16.cuh: a half-warp per
// polynomial, 16 coefficients per lane): 32 instead of 48 shared-memory wavefronts per transform and one set of twiddle
// loads for both -- with 8 coefficients per lane this kernel sat at 93 % of the L1 data pipe's wavefront peak (6.36 ms per
// 2^20 ML-DSA-65 keys; now 6.17, ML-DSA-87 9.09 -> 8.20, ML-DSA-44 4.73 -> 4.00).  Half-warp h takes polynomial pp + h of
// the pair; an odd polynomial count leaves the second half of the last pair transforming a duplicate that is not stored.
// The rows are stored prescaled by 1/256 (the inverse transforms of c*s1, c*s2, c*t0 in sign_z do not scale): 1/256 =
// -32736 mod q is small, so the factor goes onto the INPUT of the (linear) transform as part of the unpacking
// multiply-add -- |input| < 2^12 * 32736 = 16 q, well inside the transform's int32 range -- instead of modular products
// after it; sign_z's products take any int32 operand.
#ifndef SK_MINB
#define SK_MINB 6
#endif
template <class P>
__global__ void __launch_bounds__(SG_WARPS * 32, SK_MINB)
sign_key_kernel(size_t nkeys, const uint8_t* __restrict__ sk, size_t sk_stride, int32_t* __restrict__ sh)
{
    using LO = Layout<P>;
    constexpr int K = P::K, L = P::L, NP = L + 2 * K;
    __shared__ __align__(16) int32_t tiles[SG_WARPS][T16_WORDS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, h = lane >> 4, t = lane & 15;
    int32_t* tile = tiles[warp] + h * T16_POLY;
    const size_t stride = (size_t)gridDim.x * SG_WARPS;
    for (size_t kn = (size_t)blockIdx.x * SG_WARPS + warp; kn < nkeys; kn += stride) {
        const uint8_t* key = sk + kn * sk_stride;
        int32_t* const rows = sh + kn * (size_t)NP * 256;
        // packed bytes of this lane's 16 coefficients of polynomial p: 2 * eta-bits bytes (s1 / s2) or 26 bytes (t0)
        auto fetch = [&](int p, uint32_t (&w)[7]) {
            if (p < L + K) {
                uint32_t e[(2 * P::ETABITS + 3) / 4];
                load_span<2 * P::ETABITS>(key + LO::SK_S1 + p * 32 * P::ETABITS + 2 * P::ETABITS * t, e);
#pragma unroll
                for (int i = 0; i < (2 * P::ETABITS + 3) / 4; i++) w[i] = e[i];
            } else {
                load_span<26>(key + LO::SK_T0 + (p - L - K) * 416 + 26 * t, w);
            }
        };
        auto poly_of = [&](int pp) { return pp + h < NP ? pp + h : NP - 1; };
        uint32_t cur[7] = {0, 0, 0, 0, 0, 0, 0}, nxt[7] = {0, 0, 0, 0, 0, 0, 0};
        fetch(poly_of(0), cur);
#pragma unroll 1
        for (int pp = 0; pp < NP; pp += 2) {  // in the packed key's order: s1, s2, t0
            const int p = poly_of(pp);
            if (pp + 2 < NP) fetch(poly_of(pp + 2), nxt);  // in flight during this pair's transform
            constexpr int32_t C256 = SC_INV256;  // centred representative of 1/256
            static_assert(C256 > -(1 << 15) && C256 < (1 << 15), "1/256 mod q as a small centred constant");
            int32_t v[16];
            int row;
            if (p < L + K) {  // BitUnpack(eta, eta): s = eta - field            (encodings.rs:185-198)
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] = (P::ETA - (int32_t)get_field<P::ETABITS>(cur, i)) * C256;
                row = p < L ? K + p : p - L;
            } else {          // BitUnpack(2^12 - 1, 2^12): t0 = 2^12 - field     (encodings.rs:203-210)
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] = ((1 << (D - 1)) - (int32_t)get_field<13>(cur, i)) * C256;
                row = p;
            }
#pragma unroll
            for (int i = 0; i < 7; i++) cur[i] = nxt[i];
            sntt16_fwd(v, tile, t);
            if (pp + h < NP) st_row16(rows + row * 256 + 16 * t, v);
        }
    }
}

// ---- per-stage jobs.  Every stage is a grid-stride loop over WORK UNITS (what one 128-thread CTA does in one
// step: 128 Keccak streams, or one item per warp) counted from the live lane count in device memory.
// Candidate scratch (yraw, w, mw1, ctilde, cmask) is addressed by SLOT.  pos = 0: slot = lane * depth + candidate.
// pos = 1 (tail rounds with a deeper speculation than the workspace was carved for): slot = position of the lane
// in the active list * depth + candidate -- the few live lanes share the low end of the candidate buffers.
struct MaskJob {
    const uint32_t* nact; const uint32_t* act; const uint8_t* rhopp; const uint32_t* kappa; uint8_t* yraw; int depth; int pos;
    size_t cap;  // pos = 1: candidate slots of the tail region; lanes whose slots would lie beyond it sit the round out
};
struct SignWJob {
    const uint32_t* nact; const uint32_t* act; const int32_t* A; const uint8_t* yraw; int32_t* w; uint8_t* mw1; int depth;
    const uint8_t* mu; int pos;
    const uint32_t* kidx;  // resident key set: key of lane ln = kidx[ln] (nullptr: ln)
    size_t cap;            // as MaskJob::cap
};
struct SignZJob {
    const uint32_t* nact; const uint32_t* act; const int32_t* sh;  // SignWs::sh
    const uint8_t* yraw; int32_t* w; const uint8_t* ctilde; const uint8_t* cmask; uint32_t* kappa; uint8_t* done;
    uint8_t* rhopp; uint8_t* sig; uint32_t* iters; int depth;
    int pos; int depth0; size_t tail_slots;  // pos = 1: the lane's stale slots of the carved depth0 are zeroed on accept
                                             // too, unless they lie in the shared tail region [0, tail_slots)
    const uint32_t* kidx;  // resident key set: key of lane ln = kidx[ln] (nullptr: ln)
    int wipe_keys;         // 1: the key's s-hat rows belong to this call and are zeroed when the lane is accepted
    int zero;              // 0, but only the host knows: pins the butterfly adds of the inverse transform to the ALU pipe (ntt.cuh)
};

// ExpandMask: one thread per (active lane, r) SHAKE256 stream; the squeezed bytes ARE the packed y.
template <class P>
__device__ __forceinline__ size_t mask_units(const MaskJob& j) { return ((size_t)*j.nact * P::L * j.depth + 127) / 128; }
template <class P>
__device__ __forceinline__ void mask_unit(const MaskJob& j, size_t unit)
{
    constexpr int YW = 32 * P::ZBITS / 8;  // 72 or 80 output words
    const int L = P::L * j.depth;          // polynomial r of a lane belongs to candidate r / P::L: nonce kappa + r
    const size_t nstreams = (size_t)*j.nact * L;
    const size_t q = unit * 128 + threadIdx.x;
    if (q >= nstreams) return;
    const size_t li = q / L;
    if (j.pos && (li + 1) * (size_t)j.depth > j.cap) return;  // (only if the host's estimate of the live count was low)
    const size_t ln = j.act[li];
    const uint32_t r = (uint32_t)(q % L);
    const uint64_t* rp = reinterpret_cast<const uint64_t*>(j.rhopp + 64 * ln);
    uint64_t a[25];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = rp[i];
    const uint32_t nonce = (j.kappa[ln] + r) & 0xFFFF;  // u16 counter (ml_dsa.rs:204, hashing.rs:296)
    a[8] = (uint64_t)(nonce | (0x1Fu << 16));
#pragma unroll
    for (int i = 9; i < 25; i++) a[i] = 0;
    a[16] = 0x8000000000000000ULL;
    uint64_t* out = reinterpret_cast<uint64_t*>(j.yraw + ((j.pos ? li : ln) * L + r) * (size_t)(YW * 8));
#pragma unroll 1
    for (int b = 0; b < (YW + 16) / 17; b++) {
        keccak_f1600(a);
#pragma unroll
        for (int k = 0; k < 17; k++)
            if (b * 17 + k < YW) out[b * 17 + k] = a[k];
    }
}
template <class P>
__global__ void __launch_bounds__(128, 4) mask_kernel(MaskJob j)
{
    const size_t nu = mask_units<P>(j);
    for (size_t u = blockIdx.x; u < nu; u += gridDim.x) mask_unit<P>(j, u);
}

// y (layout B regs) from the packed mask bytes: BitUnpack(gamma1 - 1, gamma1)
template <class P>
__device__ __forceinline__ void load_y(const uint8_t* ypoly, int lane, int32_t (&y)[8])
{
    uint32_t w[(P::ZBITS + 3) / 4];
    load_span<P::ZBITS>(ypoly + P::ZBITS * lane, w);
#pragma unroll
    for (int i = 0; i < 8; i++) y[i] = P::GAMMA1 - (int32_t)get_field<P::ZBITS>(w, i);
}

// ---- sign_w with A-hat rows staged through shared memory by the TMA unit --------------------------------
// One bulk copy (cp.async.bulk, completion on an mbarrier) brings row k + 1 of A-hat (L KiB, contiguous) into
// shared memory while the warp runs the inverse NTT of row k, so the matrix-vector product never waits on
// HBM / L2 (ncu: 40 % of the un-staged kernel's stall samples were long-scoreboard waits on these loads).
// With an even speculation depth the two candidates of a lane sit on neighbouring warps and SHARE the staged
// row (G = 2): the row is fetched once for both, and whichever warp finishes its products last issues the next copy.
// sign_w occupancy, per parameter set.  The kernel is bound by latency and by the shared-memory data pipe together
// (ncu: L1 data-pipe wavefronts 91 %, FMA-heavy pipe 67 %, issue 61 %), so what pays is more resident warps: YREG rows
// of y-hat stay in registers instead of shared memory during the matrix-vector product (1 KiB of shared memory per
// warp and 10 wavefronts per row of A-hat each), and MINB CTAs per SM bound the register allocation to match what
// then fits.  Shared memory per warp: (L - YREG) KiB of y-hat + L / 2 KiB of the staged A-hat row + the NTT tile.
// Measured, 2^20 lanes (YREG / CTAs per SM: ms): ML-DSA-65 0 / 6 35.64, 1 / 6 35.41, 2 / 5 35.57, 5 / 4 36.96, 1 / 7 34.93;
// ML-DSA-87 0 / 4 44.58, 1 / 5 41.79, 2 / 5 41.16; ML-DSA-44 0 / 6 21.61, 0 / 7 21.58, 1 / 8 22.02 (spills), 1 / 7 21.27.
#ifndef SW_YREG_44
#define SW_YREG_44 1
#define SW_MINB_44 7
#endif
#ifndef SW_YREG_65
#define SW_YREG_65 1
#define SW_MINB_65 7
#endif
#ifndef SW_YREG_87
#define SW_YREG_87 2
#define SW_MINB_87 5
#endif
template <int YREG_, int MINB_>
struct SwTuneT { static constexpr int YREG = YREG_, MINB = MINB_; };
template <class P> struct SwTune;
template <> struct SwTune<Params<44>> : SwTuneT<SW_YREG_44, SW_MINB_44> {};
template <> struct SwTune<Params<65>> : SwTuneT<SW_YREG_65, SW_MINB_65> {};
template <> struct SwTune<Params<87>> : SwTuneT<SW_YREG_87, SW_MINB_87> {};
template <class P>
struct SwYreg { static constexpr int N = SwTune<P>::YREG < P::L ? SwTune<P>::YREG : P::L; };
template <class P, int G>
struct alignas(128) SignWStage {
    alignas(128) int32_t arow[SG_WARPS / G][P::L * 256];  // one staged row of A-hat per group of G warps
    alignas(16) int32_t yh[SG_WARPS][P::L - SwYreg<P>::N > 0 ? P::L - SwYreg<P>::N : 1][256];
    alignas(16) int32_t tile[SG_WARPS][NTT_TILE_WORDS];
    alignas(8) uint64_t full[SG_WARPS / G];
    uint32_t cnt[SG_WARPS / G];
};

template <class P, bool SHARED, int G>
__global__ void __launch_bounds__(SG_WARPS * 32, SwTune<P>::MINB) sign_w_staged_kernel(SignWJob j)
{
    constexpr int K = P::K, L = P::L;
    constexpr uint32_t ROW_BYTES = L * 1024;
    extern __shared__ __align__(128) unsigned char sm_raw[];
    SignWStage<P, G>& sm = *reinterpret_cast<SignWStage<P, G>*>(sm_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = warp / G;
    if (threadIdx.x < SG_WARPS / G) {
        mbar_init(&sm.full[threadIdx.x], 1);
        sm.cnt[threadIdx.x] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t phase = 0;
    int32_t* tile = sm.tile[warp];
    // The staged row is in natural order, so lane t's coefficients 8t..8t+7 are the int4 slots 2t and 2t+1 -- a
    // 2-way bank conflict for 128-bit loads.  Lanes with bit 2 set read their two slots in the opposite order
    // (slot numbers of a quarter-warp are then distinct mod 8), keep their y-hat rows swapped the same way, and
    // swap the two halves of the accumulator back before the inverse transform.
    const int swz = (lane >> 2) & 1;
    const int depth = j.depth;
    const size_t nitems = (size_t)*j.nact * depth;
    const size_t nu = (nitems + SG_WARPS - 1) / SG_WARPS;
    for (size_t u = blockIdx.x; u < nu; u += gridDim.x) {
        const size_t ai = u * SG_WARPS + warp;
        if (ai >= nitems) continue;  // G = 2 needs an even depth, so both warps of a group skip together
        if (j.pos && (ai / depth + 1) * (size_t)depth > j.cap) continue;  // whole lanes (all warps of a group) sit out
        const size_t ln = j.act[ai / depth];
        const size_t slot = j.pos ? ai : ln * depth + ai % depth;
        const int32_t* A = j.A + (SHARED ? 0 : (j.kidx ? (size_t)j.kidx[ln] : ln) * (size_t)(K * L) * 256);
        const uint8_t* yp = j.yraw + slot * (size_t)(L * 32 * P::ZBITS);
        int32_t* wo = j.w + slot * (size_t)K * 256;
        uint8_t* w1o = j.mw1 + slot * (size_t)(MU_BYTES + P::W1_LEN) + MU_BYTES;
        if (lane < 8)  // mu || w1 is the input row of the c~ hash
            reinterpret_cast<uint64_t*>(w1o - MU_BYTES)[lane] = reinterpret_cast<const uint64_t*>(j.mu + ln * (size_t)MU_BYTES)[lane];
        // Every warp of the group reports "done reading the staged row"; the last one to report starts the next copy.
        // Ordering: a warp's reads of the row precede its (release) arrival; the last arriver's (acquire) arrival
        // precedes the copy it issues; the proxy fence orders those generic-proxy accesses before the async-proxy
        // write of the TMA unit.  (compute-sanitizer racecheck does not model completion through an mbarrier and
        // still reports the copy against the reads of the PREVIOUS row as a hazard.)
        auto release_and_fetch = [&](int row) {
            __syncwarp();
            if (lane == 0) {
                bool last = true;
                if (G > 1) last = atom_add_acq_rel_cta(&sm.cnt[grp], 1u) == (uint32_t)(G - 1);
                if (last) {
                    if (G > 1) atomicExch(&sm.cnt[grp], 0u);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx(&sm.full[grp], ROW_BYTES);
                    bulk_g2s(sm.arow[grp], A + (size_t)row * L * 256, ROW_BYTES, &sm.full[grp]);
                }
            }
        };
        release_and_fetch(0);
#ifdef SW_YPF
        if (lane < (L - 1) * 32 * P::ZBITS / 128)  // the later y polynomials stream into L2 behind the first one
            asm volatile("prefetch.global.L2 [%0];" ::"l"(yp + 32 * P::ZBITS + 128 * lane));
#endif
        uint32_t yw[2][(P::ZBITS + 3) / 4];  // packed y of the next polynomial is in flight during an NTT
        load_span<P::ZBITS>(yp + P::ZBITS * lane, yw[0]);
        constexpr int YR = SwYreg<P>::N;
        int32_t yreg[YR > 0 ? YR : 1][8];
#pragma unroll
        for (int l = 0; l < L; l++) {
            if (l + 1 < L) load_span<P::ZBITS>(yp + (l + 1) * 32 * P::ZBITS + P::ZBITS * lane, yw[(l + 1) & 1]);
            int32_t y[8];
#pragma unroll
            for (int i = 0; i < 8; i++) y[i] = P::GAMMA1 - (int32_t)get_field<P::ZBITS>(yw[l & 1], i);
            sntt_fwd_B2B(y, tile, lane);
            if (swz) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int32_t x = y[i];
                    y[i] = y[i + 4];
                    y[i + 4] = x;
                }
            }
            if (l < YR) {
#pragma unroll
                for (int i = 0; i < 8; i++) yreg[l][i] = y[i];
            } else {
                row_store(sm.yh[warp][l - YR], lane, y);  // only this lane reads these words back
            }
        }
#pragma unroll 1
        for (int k = 0; k < K; k++) {
            mbar_wait(&sm.full[grp], phase);
            phase ^= 1;
            int64_t acc[8];
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = 0;
#pragma unroll
            for (int l = 0; l < L; l++) {
                const int4* ap = reinterpret_cast<const int4*>(sm.arow[grp] + l * 256 + 8 * lane);
                const int4 x = ap[swz], y = ap[1 - swz];
                const int32_t av[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
                int32_t yv[8];
                if (l < YR) {
#pragma unroll
                    for (int i = 0; i < 8; i++) yv[i] = yreg[l][i];
                } else {
                    row_load(sm.yh[warp][l - YR], lane, yv);
                }
#pragma unroll
                for (int i = 0; i < 8; i++) acc[i] += (int64_t)av[i] * yv[i];
            }
            if (k + 1 < K) release_and_fetch(k + 1);
            int32_t t[8];
#pragma unroll
            for (int i = 0; i < 8; i++) t[i] = mont_reduce(acc[i]);
            if (swz) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int32_t x = t[i];
                    t[i] = t[i + 4];
                    t[i + 4] = x;
                }
            }
            sntt_inv_B2B(t, tile, lane);  // w[k], canonical (A-hat is prescaled by 2^32 / 256)
            st_row8(wo + 256 * k + 8 * lane, t);
            uint32_t f[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                int32_t r0;
                f[i] = (uint32_t)decompose<P::GAMMA2>(t[i], r0);  // HighBits (ml_dsa.rs:225-226)
            }
            uint32_t w[(P::W1BITS + 3) / 4];
            pack_fields<P::W1BITS>(f, w);
            store_span<P::W1BITS>(w1o + k * 32 * P::W1BITS + P::W1BITS * lane, w);
        }
    }
}

template <class P, bool SHARED, int G>
static cudaError_t launch_sign_w_staged(cudaStream_t st, const SignWJob& j, size_t items, int cap_per_sm)
{
    // The dynamic shared-memory limit is a PER-DEVICE function attribute: one call per instantiation and device
    // (callers on different devices run on different host threads).
    static std::atomic<uint64_t> configured{0};
    constexpr size_t SMEM = sizeof(SignWStage<P, G>);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 64 || !((configured.load(std::memory_order_acquire) >> dev) & 1)) {
        e = cudaFuncSetAttribute(sign_w_staged_kernel<P, SHARED, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        if (e != cudaSuccess) return e;
        if (dev < 64) configured.fetch_or(uint64_t(1) << dev, std::memory_order_release);
    }
    KLaunch kl(KID_SIGN_W, st);
    sign_w_staged_kernel<P, SHARED, G><<<capped_grid((items + SG_WARPS - 1) / SG_WARPS, cap_per_sm), SG_WARPS * 32, SMEM, st>>>(j);
    return cudaGetLastError();
}

template <class P>
struct alignas(16) SignSmem {
    int32_t tile[NTT_TILE_WORDS];
    uint32_t stage[32 * P::ZBITS / 4 + 1];  // packed z polynomial (576/640 B) or the hint bytes
};

// One pass of the second half of a rejection-loop iteration for one lane.  The three vector products
// c*s2, c*s1, c*t0 share ONE instance of the inverse-NTT code (a single loop over 2K+L polynomials); the
// checks run most-selective-first (r0, then z, then ct0/hints) so a rejected iteration stops early.  The
// order of the checks does not affect the outcome: the reference rejects if ANY fails (ml_dsa.rs:277-319).
template <class P>
__device__ __forceinline__ size_t sign_z_units(const SignZJob& j) { return ((size_t)*j.nact + SG_WARPS - 1) / SG_WARPS; }
template <class P, bool SHARED>
__device__ __forceinline__ void sign_z_unit(const SignZJob& j, size_t unit, SignSmem<P>* sm_all)
{
    using LO = Layout<P>;
    constexpr int K = P::K, L = P::L;
    const uint32_t* __restrict__ act = j.act;
    const int32_t* __restrict__ sh_all = j.sh;
    const uint8_t* __restrict__ yraw = j.yraw;
    int32_t* __restrict__ wall = j.w;
    const uint8_t* __restrict__ ctilde = j.ctilde;
    const uint8_t* __restrict__ cmask = j.cmask;
    uint32_t* __restrict__ kappa = j.kappa;
    uint8_t* __restrict__ done = j.done;
    uint8_t* __restrict__ rhopp = j.rhopp;
    uint8_t* __restrict__ sig = j.sig;
    uint32_t* __restrict__ iters = j.iters;
    const int depth = j.depth;
    const size_t nact = *j.nact;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t z0 = j.zero;  // 0, opaque to ptxas: keeps the inverse transform's adds on the ALU pipe (ntt.cuh: gs_bfly_s)
    SignSmem<P>& sm = sm_all[warp];
    int32_t* tile = sm.tile;
    uint8_t* hb = reinterpret_cast<uint8_t*>(sm.stage);
    const size_t ai = unit * SG_WARPS + warp;
    if (ai < nact && !(j.pos && (ai + 1) * (size_t)depth > j.tail_slots)) {  // (a lane without slots sits the round out)
        const size_t ln = act[ai];
        const size_t kn = SHARED ? 0 : (j.kidx ? (size_t)j.kidx[ln] : ln);
        const size_t sbase = (j.pos ? ai : ln) * depth;  // first candidate slot of this lane
        uint8_t* sg = sig + ln * (size_t)P::SIG_LEN;
        bool rej = true;
        int cand = 0;
        // the lane's candidates in the reference's order: the first one that passes every check is the signature
#pragma unroll 1
        for (; cand < depth && rej; cand++) {
        const size_t slot = sbase + cand;
        const uint8_t* yp = yraw + slot * (size_t)(L * 32 * P::ZBITS);
        int32_t* wl = wall + slot * (size_t)K * 256;

        // the first w row and the first y polynomial stream into L2 while c-hat is computed
        if (lane < 13) {
            const char* nxt = lane < 8 ? reinterpret_cast<const char*>(wl) + 128 * lane
                                       : reinterpret_cast<const char*>(yp) + 128 * (lane - 8);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt));
        }
        rej = false;
        int hcount = 0;
        // operand row of step p: s2-hat[p], then s1-hat[p - K], then t0-hat[p - K - L] -- consecutive rows of the key's block
        // (a shared key's rows stay hot in L1: 128-bit __ldg; per-lane keys stream: one 256-bit load)
        const int32_t* rp = sh_all + kn * (size_t)(2 * K + L) * 256 + 8 * lane;
        auto load_row = [&](const int32_t* r) {
            if (!SHARED) return ldg_row8(r);
            Row8 o;
            const int4 x = __ldg(reinterpret_cast<const int4*>(r)), y = __ldg(reinterpret_cast<const int4*>(r) + 1);
            o.v[0] = x.x; o.v[1] = x.y; o.v[2] = x.z; o.v[3] = x.w; o.v[4] = y.x; o.v[5] = y.y; o.v[6] = y.z; o.v[7] = y.w;
            return o;
        };
        Row8 nx = load_row(rp);  // in flight during the forward NTT of c
        // c from the challenge masks, c-hat = NTT(c) (ml_dsa.rs:237-240) centred, with its Shoup companion
        int32_t ch[8], chs[8];
        load_c_layoutA(cmask + slot * (size_t)CMASK_STRIDE, lane, ch);
        sntt_fwd_A2B(ch, tile, lane);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            ch[i] = center(freeze(ch[i]));
            chs[i] = shoup_companion(ch[i]);
        }
#pragma unroll 1
        for (int p = 0; p < 2 * K + L && !rej; p++) {
            const int phase = p < K ? 0 : (p < K + L ? 1 : 2);
            const int idx = phase == 0 ? p : (phase == 1 ? p - K : p - K - L);
            const Row8 sv = nx;
            if (p + 1 < 2 * K + L) {  // the next step's operand row loads while this step's inverse NTT runs
                rp += 256;
                nx = load_row(rp);
            }
            // the w rows of a candidate are contiguous: pull the next one into L2
            if (phase == 0 && lane < 8 && idx + 1 < K)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(wl + 256 * (idx + 1)) + 128 * lane));
            if (phase == 1 && lane < 5 && idx + 1 < L)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(yp + (idx + 1) * 32 * P::ZBITS + 128 * lane));
            // What the checks after the inverse transform consume -- the w row (phases 0 and 2) or the packed y
            // polynomial (phase 1) -- is requested BEFORE the transform: issued after it, these loads were the largest
            // single stall of the kernel (ncu source page: 9 % of the not-issued samples on their first uses).
            Row8 wrow{};
            uint32_t yw[(P::ZBITS + 3) / 4];
            int32_t* const wp = wl + 256 * idx + 8 * lane;
            if (phase == 1) {
                load_span<P::ZBITS>(yp + idx * 32 * P::ZBITS + P::ZBITS * lane, yw);
            } else {
                wrow = ld_row8(wp);
            }
            int32_t t[8];
#pragma unroll
            for (int i = 0; i < 8; i++) t[i] = shoup_mul(sv.v[i], ch[i], chs[i]);  // s-hat rows are prescaled by 1/256
            sntt_inv_B2B(t, tile, lane, z0);  // c*s2[idx] / c*s1[idx] / c*t0[idx], canonical

            bool bad = false;
            if (phase == 0) {
                // r0 = LowBits(w - c*s2), ||r0||inf < gamma2 - beta              ml_dsa.rs:253-260, 268-272
                int32_t v[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    v[i] = caddq(wrow.v[i] - t[i]);  // w - cs2, canonical
                    int32_t r0;
                    decompose<P::GAMMA2>(v[i], r0);
                    bad |= (r0 < 0 ? -r0 : r0) >= P::GAMMA2 - P::BETA;
                }
                rej = __any_sync(0xffffffffu, bad);
                if (!rej) st_row8(wp, v);  // w is only ever needed as w - cs2 from here on
            } else if (phase == 1) {
                // z = y + c*s1, ||z||inf < gamma1 - beta                          ml_dsa.rs:243-250, 263-265
                uint32_t f[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int32_t y = P::GAMMA1 - (int32_t)get_field<P::ZBITS>(yw, i);  // BitUnpack(gamma1 - 1, gamma1)
                    const int32_t z = center(freeze(y + t[i]));                    // z mod+- q (ml_dsa.rs:334-335)
                    bad |= (z < 0 ? -z : z) >= P::GAMMA1 - P::BETA;
                    f[i] = (uint32_t)(P::GAMMA1 - z);                              // BitPack(z, gamma1 - 1, gamma1)
                }
                rej = __any_sync(0xffffffffu, bad);
                if (!rej) {  // sigEncode z[idx] into the output; a later rejection just overwrites it
                    uint32_t w[(P::ZBITS + 3) / 4];
                    pack_fields<P::ZBITS>(f, w);
                    __syncwarp();
                    store_span<P::ZBITS>(reinterpret_cast<uint8_t*>(sm.stage) + P::ZBITS * lane, w);
                    __syncwarp();
                    warp_copy_out(sg + LO::SIG_Z + idx * 32 * P::ZBITS, sm.stage, 32 * P::ZBITS, lane);
                    if (idx == L - 1) {  // entering the hint phase: clear the hint staging area
                        __syncwarp();
                        for (int i = lane; i < P::OMEGA + K; i += 32) hb[i] = 0;
                        __syncwarp();
                    }
                }
            } else {
                // ||c*t0||inf < gamma2; h = MakeHint(-ct0, w - cs2 + ct0)          ml_dsa.rs:288-319
                const int32_t (&v)[8] = wrow.v;
                uint32_t hbits = 0;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int32_t c0 = center(t[i]);
                    bad |= (c0 < 0 ? -c0 : c0) >= P::GAMMA2;                       // ml_dsa.rs:314
                    int32_t r0;
                    int32_t u = v[i] + t[i];                                       // w - cs2 + ct0 in [0, 2q)
                    u -= (u >= Q) ? Q : 0;
                    const int32_t h1 = decompose<P::GAMMA2>(u, r0);
                    const int32_t h0 = decompose<P::GAMMA2>(v[i], r0);             // HighBits(r + z) = HighBits(w - cs2)
                    hbits |= (uint32_t)(h1 != h0) << i;                            // high_low.rs:134-144
                }
                // HintBitPack (conversion.rs:277-328): ordered compaction by warp prefix sum
                const int cnt = __popc(hbits);
                int incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int up = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += up;
                }
                int pos = hcount + incl - cnt;
                const int total = __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
                for (int i = 0; i < 8; i++)
                    if ((hbits >> i) & 1) {
                        if (pos < P::OMEGA) hb[pos] = (uint8_t)(8 * lane + i);
                        pos++;
                    }
                hcount += total;
                rej = __any_sync(0xffffffffu, bad) || hcount > P::OMEGA;           // ml_dsa.rs:312-319
                if (!rej && lane == 0) hb[P::OMEGA + idx] = (uint8_t)hcount;
            }
        }
        __syncwarp();
        }  // candidates
        if (rej) {
            if (lane == 0) kappa[ln] += L * depth;                                 // ml_dsa.rs:281, 317
            return;
        }
        cand--;  // the accepted candidate
        // accept: c~ and hints complete the signature                            encodings.rs:254, 273
        __syncwarp();
        warp_copy_out(sg + LO::SIG_H, sm.stage, P::OMEGA + K, lane);
        warp_copy_bytes(sg, ctilde + (sbase + cand) * 64, P::CTILDE, lane);
        if (lane == 0) {
            done[ln] = 1;
            if (iters) iters[ln] = kappa[ln] / L + cand + 1;
        }
        // The lane is finished: zero everything it derived from the secret key and the mask (y, w - c*s2, rho'',
        // and its own s1-hat / s2-hat / t0-hat) right here, where the stores hide behind other warps' arithmetic,
        // instead of a 27 GB cudaMemset after the last round (src/types.rs:19 ZeroizeOnDrop analogue).
        warp_zero16(wall + sbase * (size_t)K * 256, depth * K * 1024, lane);
        warp_zero16(const_cast<uint8_t*>(yraw) + sbase * (size_t)(L * 32 * P::ZBITS), depth * L * 32 * P::ZBITS, lane);
        if (j.pos && ln * (size_t)j.depth0 >= j.tail_slots) {
            // what the lane's last rejected candidates left in its own slots (the tail region itself, which other
            // live lanes are using right now, is cleared by the host after the last round)
            warp_zero16(wall + ln * (size_t)j.depth0 * K * 256, j.depth0 * K * 1024, lane);
            warp_zero16(const_cast<uint8_t*>(yraw) + ln * (size_t)j.depth0 * (L * 32 * P::ZBITS), j.depth0 * L * 32 * P::ZBITS, lane);
        }
        warp_zero16(rhopp + ln * 64, 64, lane);
        if (!SHARED && j.wipe_keys) warp_zero16(const_cast<int32_t*>(sh_all) + kn * (size_t)(2 * K + L) * 256, (2 * K + L) * 1024, lane);
        __syncwarp();
    }
}

template <class P, bool SHARED>
__global__ void __launch_bounds__(SG_WARPS * 32, SZ_MINB) sign_z_kernel(SignZJob j)
{
    __shared__ __align__(16) SignSmem<P> sm_all[SG_WARPS];
    const size_t nu = sign_z_units<P>(j);
    for (size_t u = blockIdx.x; u < nu; u += gridDim.x) {
        sign_z_unit<P, SHARED>(j, u, sm_all);
        __syncwarp();
    }
}

// Device -> host mailbox without the copy engine: the pinned mailbox is mapped into the device address space and a
// one-warp kernel stores the control words there.  (A 4-byte cudaMemcpyAsync D2H queues behind any bulk D2H copy
// that is in flight -- e.g. the previous chunk's signatures -- and stalled the sign stream for its whole duration.)
__global__ void publish_kernel(const uint32_t* __restrict__ src, volatile uint32_t* __restrict__ dst_mapped, int nwords)
{
    if ((int)threadIdx.x < nwords) dst_mapped[threadIdx.x] = src[threadIdx.x];
    __threadfence_system();
}

__global__ void __launch_bounds__(256)
compact_kernel(const uint32_t* __restrict__ nact_dev, const uint32_t* __restrict__ act_in,
               const uint8_t* __restrict__ done, uint32_t* __restrict__ act_out, uint32_t* __restrict__ nout)
{
    // grid-stride over the LIVE count (whole warps per trip): the grid is sized from the host's estimate of that
    // count, and a lane must never be dropped from the list if the estimate was low
    const size_t nact = *nact_dev;
    const size_t span = (size_t)gridDim.x * blockDim.x;
    for (size_t base = (size_t)blockIdx.x * blockDim.x; base < nact; base += span) {
        const size_t i = base + threadIdx.x;
        const bool keep = i < nact && !done[act_in[i]];
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (m == 0) continue;
        const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
        uint32_t pos = 0;
        if (lane == leader) pos = atomicAdd(nout, (uint32_t)__popc(m));
        pos = __shfl_sync(0xffffffffu, pos, leader);
        if (keep) act_out[pos + __popc(m & ((1u << lane) - 1))] = act_in[i];
    }
}

constexpr int CTL_WORDS = 16;       // SignWs::nact: [buffer] active counts of the two ping-pong lists, [8] ctx error flag
constexpr int CTL_CTX_ERR = 8;

// ---- the rejection loop of one chunk, in three host phases ------------------------------------------------
//   sign_begin   set-up kernels + the BULK rounds (while the expected number of live lanes is above `tail_lanes`)
//   sign_tail    the remaining rounds that almost surely suffice + publication of the control words
//   sign_finish  the only host wait: read the live count, run straggler rounds until it is zero
// No phase before sign_finish waits for the device, and the tail may be enqueued on another stream than the bulk
// (the caller orders the two with an event): the host entry point runs chunk i's latency-bound tail (a few
// thousand lanes, ~100 dependent small launches) on a high-priority stream UNDER chunk i+1's bulk rounds.
// Every kernel is a grid-stride loop over the live count it reads from device memory, so grids are sized from
// the EXPECTED number of survivors: a round rejects a lane with probability (1 - p)^D, p >= 0.19 for all three
// parameter sets (FIPS 204 table 1: 4.25 / 5.1 / 3.85 expected iterations).
template <class P>
static cudaError_t sign_setup_t(cudaStream_t st, SignRun& r, const uint8_t* sk, size_t sk_stride, const MsgDesc& md,
                                const uint8_t* rnd)
{
    constexpr int K = P::K, L = P::L;
    const SignWs& ws = r.ws;
    const size_t n = r.n, nkeys = r.shared ? 1 : n;
    const Tuning tn = tuning();
    int* ctx_err = reinterpret_cast<int*>(ws.nact + CTL_CTX_ERR);
    cudaError_t e = cudaMemsetAsync(ws.nact, 0, CTL_WORDS * 4, st);
    if (e != cudaSuccess) return e;
    {
        KLaunch kl(KID_SIGN_PREP, st);
        sign_prep_kernel<P><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, sk, sk_stride, md, rnd, ws.mu, ws.rhopp,
                                                                        ws.kappa, ws.done, ws.act[0], ws.nact,
                                                                        ctx_err, ws.kidx);
    }
    if (ws.keys_resident) return cudaGetLastError();  // A-hat and the s-hat rows come from a loaded key set
    {
        KLaunch kl(KID_SIGN_KEY, st);
        sign_key_kernel<P><<<capped_grid((nkeys + SG_WARPS - 1) / SG_WARPS, arith_cap()), SG_WARPS * 32, 0, st>>>(
            nkeys, sk, sk_stride, ws.sh);
    }
    return launch_expand_a(st, K, L, sk, sk_stride, nkeys, ws.A);
}

// Speculation depth of a round over (at most) sz live lanes: the carved depth D while the candidates fill the
// device; once sz * d candidates fit in the tail region (the low tail_slots slots of the candidate buffers), the
// deepest d in {16, 8, 4} allowed -- the device is under-filled there and a round costs its fixed latency, so
// extra speculative candidates are free and the number of rounds (dependent launches) drops with log(0.81^d).
static int round_depth(const SignRun& r, size_t sz)
{
    const int D = r.ws.depth;
    if (r.tail_slots)
        for (int d = 16; d > D; d >>= 1)
            if (d <= r.tail_depth && sz * (size_t)d <= r.tail_slots) return d;
    return D;
}

// one round over (at most) sz live lanes
template <class P>
static cudaError_t sign_round_t(cudaStream_t st, SignRun& r, size_t sz)
{
    constexpr int L = P::L;
    const SignWs& ws = r.ws;
    const Tuning tn = tuning();
    const int D = round_depth(r, sz);
    const int pos = D != ws.depth ? 1 : 0;
    if (pos) r.used_tail = true;
    const bool shared = r.shared;
    const uint32_t* nact = ws.nact + r.cur;
    const uint32_t* act = ws.act[r.cur];
    auto warp_grid = [&](size_t items) { return capped_grid((items + SG_WARPS - 1) / SG_WARPS, arith_cap()); };
    {
        KLaunch kl(KID_MASK, st);
        mask_kernel<P><<<capped_grid((sz * L * D + 127) / 128, tn.kcap), 128, 0, st>>>(
            MaskJob{nact, act, ws.rhopp, ws.kappa, ws.yraw, D, pos, r.tail_slots});
    }
    cudaError_t ce = cudaSuccess;
    const SignWJob wj{nact, act, ws.A, ws.yraw, ws.w, ws.mw1, D, ws.mu, pos, ws.kidx, r.tail_slots};
    {   // an even depth pairs the two candidates of a lane on one staged A-hat row (G = 2); an odd one stages per warp
        const int cap = arith_cap();
        if (D % 2 == 0)
            ce = shared ? launch_sign_w_staged<P, true, 2>(st, wj, sz * D, cap) : launch_sign_w_staged<P, false, 2>(st, wj, sz * D, cap);
        else
            ce = shared ? launch_sign_w_staged<P, true, 1>(st, wj, sz * D, cap) : launch_sign_w_staged<P, false, 1>(st, wj, sz * D, cap);
        if (ce != cudaSuccess) return ce;
    }
    // pos: rows are the positions themselves (act = nullptr)
    ce = launch_ctilde(P::PS, st, sz * D, nact, pos ? nullptr : act, D, ws.mw1, nullptr, nullptr, nullptr, ws.ctilde, ws.cmask,
                       pos ? r.tail_slots : ~size_t(0));
    if (ce != cudaSuccess) return ce;
    {
        const SignZJob zj{nact, act, ws.sh, ws.yraw, ws.w, ws.ctilde, ws.cmask, ws.kappa, ws.done, ws.rhopp,
                          r.sig, r.iters, D, pos, ws.depth, r.tail_slots, ws.kidx, ws.keys_resident ? 0 : 1, 0};
        KLaunch kl(KID_SIGN_Z, st);
        if (shared) sign_z_kernel<P, true><<<warp_grid(sz), SG_WARPS * 32, 0, st>>>(zj);
        else sign_z_kernel<P, false><<<warp_grid(sz), SG_WARPS * 32, 0, st>>>(zj);
    }
    // re-queue the rejected lanes into the other list
    uint32_t* nout = ws.nact + (r.cur ^ 1);
    ce = cudaMemsetAsync(nout, 0, 4, st);
    if (ce != cudaSuccess) return ce;
    {
        KLaunch kl(KID_COMPACT, st);
        compact_kernel<<<(unsigned)((sz + 255) / 256) + 1, 256, 0, st>>>(nact, act, ws.done, ws.act[r.cur ^ 1], nout);
    }
    r.cur ^= 1;
    r.bound = sz;
    r.rounds++;
    return cudaGetLastError();
}
static cudaError_t sign_round(cudaStream_t st, SignRun& r, size_t sz)
{
    switch (r.ps) {
    case 44: return sign_round_t<Params<44>>(st, r, sz);
    case 65: return sign_round_t<Params<65>>(st, r, sz);
    case 87: return sign_round_t<Params<87>>(st, r, sz);
    }
    return cudaErrorInvalidValue;
}
static cudaError_t sign_rounds_until(cudaStream_t st, SignRun& r, double stop_at)
{
    const int max_rounds = tuning().max_rounds;
    // test hook: MLDSA_EST_SCALE < 1 makes the host UNDER-estimate the live count, which exercises what must then still
    // hold -- grid-stride kernels and compaction over the true count, lanes without a tail slot sitting a round out
    static const double est_scale = getenv("MLDSA_EST_SCALE") ? atof(getenv("MLDSA_EST_SCALE")) : 1.0;
    while (r.expect > stop_at) {
        if (max_rounds > 0 && r.rounds >= max_rounds) break;  // test hook: fixed number of rounds
        size_t sz = (size_t)(r.expect * 1.25 * est_scale) + (est_scale < 1.0 ? 1 : 256);
        if (sz > r.n) sz = r.n;
        const double keep = pow(1.0 - 0.19, (double)round_depth(r, sz));
        cudaError_t e = sign_round(st, r, sz);
        if (e != cudaSuccess) return e;
        r.expect *= keep;
    }
    return cudaSuccess;
}

// expand_private for a key set (src/ml_dsa.rs:445-469): A-hat and s1-hat / s2-hat / t0-hat of nkeys packed secret keys
template <class P>
static cudaError_t sign_expand_keys_t(cudaStream_t st, size_t nkeys, const uint8_t* sk, int32_t* A, int32_t* sh, bool with_A)
{
    {
        KLaunch kl(KID_SIGN_KEY, st);
        sign_key_kernel<P><<<capped_grid((nkeys + SG_WARPS - 1) / SG_WARPS, arith_cap()), SG_WARPS * 32, 0, st>>>(
            nkeys, sk, P::SK_LEN, sh);
    }
    if (!with_A) return cudaGetLastError();  // the caller already holds this key's A-hat (e.g. from keygen)
    return launch_expand_a(st, P::K, P::L, sk, P::SK_LEN, nkeys, A);
}
cudaError_t sign_expand_keys(int ps, cudaStream_t st, size_t nkeys, const uint8_t* sk, int32_t* A, int32_t* sh, bool with_A)
{
    if (nkeys == 0) return cudaSuccess;
    switch (ps) {
    case 44: return sign_expand_keys_t<Params<44>>(st, nkeys, sk, A, sh, with_A);
    case 65: return sign_expand_keys_t<Params<65>>(st, nkeys, sk, A, sh, with_A);
    case 87: return sign_expand_keys_t<Params<87>>(st, nkeys, sk, A, sh, with_A);
    }
    return cudaErrorInvalidValue;
}

cudaError_t sign_begin(int ps, cudaStream_t st, size_t n, const uint8_t* sk, size_t sk_stride, const MsgDesc& md,
                       const uint8_t* rnd, uint8_t* sig, uint32_t* iters, const SignWs& ws, size_t tail_lanes,
                       SignRun* run)
{
    SignRun& r = *run;
    r = SignRun{};
    r.ps = ps;
    r.n = n;
    r.shared = sk_stride == 0;
    r.sig = sig;
    r.iters = iters;
    r.ws = ws;
    r.expect = (double)n;
    r.tail_stream = st;
    {
        // tail region: the low slots of the candidate buffers, shared by the few live lanes of the last rounds
        const Tuning tn = tuning();
        size_t slots = n * (size_t)ws.depth;
        if (slots > (size_t)tn.tail_slots) slots = (size_t)tn.tail_slots;
        slots -= slots % (size_t)(16 * ws.depth);  // whole lanes of the carved depth, whole lanes of every tail depth
        r.tail_slots = tn.tail_depth > ws.depth ? slots : 0;
        r.tail_depth = tn.tail_depth;
        const int K = ps == 44 ? 4 : (ps == 65 ? 6 : 8), L = ps == 44 ? 4 : (ps == 65 ? 5 : 7), zbits = ps == 44 ? 18 : 20;
        r.slot_w_bytes = (size_t)K * 1024;
        r.slot_y_bytes = (size_t)L * 32 * zbits;
    }
    if (n == 0) {
        r.finished = true;
        return cudaSuccess;
    }
    cudaError_t e;
    switch (ps) {
    case 44: e = sign_setup_t<Params<44>>(st, r, sk, sk_stride, md, rnd); break;
    case 65: e = sign_setup_t<Params<65>>(st, r, sk, sk_stride, md, rnd); break;
    case 87: e = sign_setup_t<Params<87>>(st, r, sk, sk_stride, md, rnd); break;
    default: return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    return sign_rounds_until(st, r, tail_lanes > 0 ? (double)tail_lanes : 0.01);
}

cudaError_t sign_tail(cudaStream_t st, SignRun* run)
{
    SignRun& r = *run;
    r.tail_stream = st;
    if (r.finished) return cudaSuccess;
    cudaError_t e = sign_rounds_until(st, r, 0.01);  // P(a lane is still active after these rounds) < 1 %
    if (e != cudaSuccess) return e;
    publish_kernel<<<1, 32, 0, st>>>(r.ws.nact, r.ws.h_nact, CTL_WORDS);
    return cudaGetLastError();
}

cudaError_t sign_finish(SignRun* run)
{
    SignRun& r = *run;
    if (r.finished) return cudaSuccess;
    cudaStream_t st = r.tail_stream;
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return e;
    if (r.ws.h_nact[CTL_CTX_ERR] != 0) {  // ctx too long: the caller maps this to -2 (lib.rs:274)
        r.ctx_err = true;
        r.finished = true;
        return cudaErrorInvalidValue;
    }
    r.bound = r.ws.h_nact[r.cur];
    if (tuning().max_rounds > 0 && r.bound > 0) {  // test hook: the caller wipes the unfinished lanes' state
        r.incomplete = true;
        r.finished = true;
        return cudaErrorNotReady;
    }
    while (r.bound > 0) {  // stragglers beyond the estimate (rare): a few rounds at a time
        const size_t sz = r.bound;
        for (int k = 0; k < 4; k++) {
            e = sign_round(st, r, sz);
            if (e != cudaSuccess) return e;
        }
        publish_kernel<<<1, 32, 0, st>>>(r.ws.nact, r.ws.h_nact, CTL_WORDS);
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return e;
        r.bound = r.ws.h_nact[r.cur];
    }
    if (r.used_tail) {  // rejected candidates of the tail rounds (y, w - c*s2) may be left in the shared tail region
        e = cudaMemsetAsync(r.ws.w, 0, r.tail_slots * r.slot_w_bytes, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(r.ws.yraw, 0, r.tail_slots * r.slot_y_bytes, st);
        if (e != cudaSuccess) return e;
    }
    r.finished = true;
    return cudaSuccess;
}

}  // namespace mldsa
