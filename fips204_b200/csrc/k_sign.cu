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
#include "engine.h"
#include "keccak.cuh"
#include "ntt.cuh"
#include "pack.cuh"
#include "sampling.cuh"

namespace mldsa {

template <class P>
__global__ void __launch_bounds__(128)
sign_prep_kernel(size_t n, const uint8_t* __restrict__ sk, size_t sk_stride, MsgDesc md,
                 const uint8_t* __restrict__ rnd, uint8_t* __restrict__ mw1, uint8_t* __restrict__ rhopp,
                 uint32_t* __restrict__ kappa, uint8_t* __restrict__ done, uint32_t* __restrict__ act0,
                 uint32_t* __restrict__ nact0, int* __restrict__ ctx_err, int depth)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == 0) *nact0 = (uint32_t)n;
    const uint8_t* key = sk + i * sk_stride;
    Sponge<SHAKE256_RATE> sp;
    if (!absorb_mu(sp, key + 64, md, i)) *ctx_err = 1;  // lib.rs:274 -> whole call fails with -2
    uint64_t mu[8];
    sp.out_words(mu, 8);
    for (int j = 0; j < depth; j++) {  // every candidate row starts with mu (the input row of the c~ hash)
        uint64_t* mo = reinterpret_cast<uint64_t*>(mw1 + (i * depth + j) * (size_t)(MU_BYTES + P::W1_LEN));
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

// s1-hat, s2-hat, t0-hat (plain NTT domain) for every key.  warp per key.
template <class P>
__global__ void __launch_bounds__(SG_WARPS * 32)
sign_key_kernel(size_t nkeys, const uint8_t* __restrict__ sk, size_t sk_stride, int32_t* __restrict__ s1h,
                int32_t* __restrict__ s2h, int32_t* __restrict__ t0h)
{
    using LO = Layout<P>;
    constexpr int K = P::K, L = P::L;
    __shared__ __align__(16) int32_t tiles[SG_WARPS][NTT_TILE_WORDS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t* tile = tiles[warp];
    const size_t stride = (size_t)gridDim.x * SG_WARPS;
    for (size_t kn = (size_t)blockIdx.x * SG_WARPS + warp; kn < nkeys; kn += stride) {
        const uint8_t* key = sk + kn * sk_stride;
#pragma unroll 1
        for (int p = 0; p < L + 2 * K; p++) {
            int32_t v[8];
            int32_t* dst;
            if (p < L + K) {  // BitUnpack(eta, eta): s = eta - field            (encodings.rs:185-198)
                uint32_t w[(P::ETABITS + 3) / 4];
                load_span<P::ETABITS>(key + LO::SK_S1 + p * 32 * P::ETABITS + P::ETABITS * lane, w);
#pragma unroll
                for (int i = 0; i < 8; i++) v[i] = P::ETA - (int32_t)get_field<P::ETABITS>(w, i);
                dst = p < L ? s1h + (kn * L + p) * 256 : s2h + (kn * K + (p - L)) * 256;
            } else {          // BitUnpack(2^12 - 1, 2^12): t0 = 2^12 - field     (encodings.rs:203-210)
                uint32_t w[4];
                load_span<13>(key + LO::SK_T0 + (p - L - K) * 416 + 13 * lane, w);
#pragma unroll
                for (int i = 0; i < 8; i++) v[i] = (1 << (D - 1)) - (int32_t)get_field<13>(w, i);
                dst = t0h + (kn * K + (p - L - K)) * 256;
            }
            sntt_fwd_B2B(v, tile, lane);
            // prescaled by 1/256: the inverse transforms of c*s1, c*s2, c*t0 in sign_z do not scale
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = shoup_mul(v[i], SC_INV256, SC_INV256S);
            int4* o = reinterpret_cast<int4*>(dst + 8 * lane);
            o[0] = make_int4(v[0], v[1], v[2], v[3]);
            o[1] = make_int4(v[4], v[5], v[6], v[7]);
        }
    }
}

// ExpandMask: one thread per (active lane, r) SHAKE256 stream; the squeezed bytes ARE the packed y.
template <class P>
__global__ void __launch_bounds__(128, 4)
mask_kernel(const uint32_t* __restrict__ nact_dev, const uint32_t* __restrict__ act,
            const uint8_t* __restrict__ rhopp, const uint32_t* __restrict__ kappa, uint8_t* __restrict__ yraw, int depth)
{
    constexpr int YW = 32 * P::ZBITS / 8;  // 72 or 80 output words
    const int L = P::L * depth;            // polynomial r of a lane belongs to candidate r / P::L: nonce kappa + r
    const size_t nstreams = (size_t)*nact_dev * L;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < nstreams; q += (size_t)gridDim.x * blockDim.x) {
    const size_t ln = act[q / L];
    const uint32_t r = (uint32_t)(q % L);
    const uint64_t* rp = reinterpret_cast<const uint64_t*>(rhopp + 64 * ln);
    uint64_t a[25];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = rp[i];
    const uint32_t nonce = (kappa[ln] + r) & 0xFFFF;  // u16 counter (ml_dsa.rs:204, hashing.rs:296)
    a[8] = (uint64_t)(nonce | (0x1Fu << 16));
#pragma unroll
    for (int i = 9; i < 25; i++) a[i] = 0;
    a[16] = 0x8000000000000000ULL;
    uint64_t* out = reinterpret_cast<uint64_t*>(yraw + (ln * L + r) * (size_t)(YW * 8));
#pragma unroll 1
    for (int b = 0; b < (YW + 16) / 17; b++) {
        keccak_f1600(a);
#pragma unroll
        for (int k = 0; k < 17; k++)
            if (b * 17 + k < YW) out[b * 17 + k] = a[k];
    }
    }
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

template <class P>
struct alignas(16) SignWSmem {
    int32_t tile[NTT_TILE_WORDS];
    alignas(16) int32_t yh[P::L][256];  // y-hat, NTT domain, natural order (lane t owns words 8t..8t+7 of each row)
};

template <class P, bool SHARED>
__global__ void __launch_bounds__(SG_WARPS * 32, 5)
sign_w_kernel(const uint32_t* __restrict__ nact_dev, const uint32_t* __restrict__ act,
              const int32_t* __restrict__ Aall, const uint8_t* __restrict__ yraw, int32_t* __restrict__ wall,
              uint8_t* __restrict__ mw1, int depth)
{
    constexpr int K = P::K, L = P::L;
    const size_t nact = (size_t)*nact_dev * depth;  // work items: (active lane, candidate)
    __shared__ __align__(16) SignWSmem<P> sm_all[SG_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SignWSmem<P>& sm = sm_all[warp];
    int32_t* tile = sm.tile;
    const size_t stride = (size_t)gridDim.x * SG_WARPS;
    for (size_t ai = (size_t)blockIdx.x * SG_WARPS + warp; ai < nact; ai += stride) {
        const size_t ln = act[ai / depth];
        const size_t slot = ln * depth + ai % depth;  // neighbouring warps hold the candidates of one lane
        const int32_t* A = Aall + (SHARED ? 0 : ln * (size_t)(K * L) * 256);
        const uint8_t* yp = yraw + slot * (size_t)(L * 32 * P::ZBITS);
        int32_t* wo = wall + slot * (size_t)K * 256;
        uint8_t* w1o = mw1 + slot * (size_t)(MU_BYTES + P::W1_LEN) + MU_BYTES;
        __syncwarp();
        warp_prefetch_l2(A, 2 * L * 1024, lane);  // rows 0 and 1 of A-hat stream in during the forward NTTs
        uint32_t yw[2][(P::ZBITS + 3) / 4];       // packed y of the next polynomial is in flight during an NTT
        load_span<P::ZBITS>(yp + P::ZBITS * lane, yw[0]);
#pragma unroll
        for (int l = 0; l < L; l++) {
            if (l + 1 < L) load_span<P::ZBITS>(yp + (l + 1) * 32 * P::ZBITS + P::ZBITS * lane, yw[(l + 1) & 1]);
            int32_t y[8];
#pragma unroll
            for (int i = 0; i < 8; i++) y[i] = P::GAMMA1 - (int32_t)get_field<P::ZBITS>(yw[l & 1], i);
            sntt_fwd_B2B(y, tile, lane);
            row_store(sm.yh[l], lane, y);  // only this lane reads these words back
        }
#pragma unroll 1
        for (int k = 0; k < K; k++) {
            if (k + 2 < K) warp_prefetch_l2(A + (size_t)(k + 2) * L * 256, L * 1024, lane);
            int64_t acc[8];
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = 0;
#pragma unroll
            for (int l = 0; l < L; l++) {
                const int4* ap = reinterpret_cast<const int4*>(A + (size_t)(k * L + l) * 256 + 8 * lane);
                const int4 x = __ldg(ap), y = __ldg(ap + 1);
                const int32_t av[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
                int32_t yv[8];
                row_load(sm.yh[l], lane, yv);
#pragma unroll
                for (int i = 0; i < 8; i++) acc[i] += (int64_t)av[i] * yv[i];
            }
            int32_t t[8];
#pragma unroll
            for (int i = 0; i < 8; i++) t[i] = mont_reduce(acc[i]);
            sntt_inv_B2B(t, tile, lane);  // w[k], canonical (A-hat is prescaled by 2^32 / 256)
            int4* o = reinterpret_cast<int4*>(wo + 256 * k + 8 * lane);
            o[0] = make_int4(t[0], t[1], t[2], t[3]);
            o[1] = make_int4(t[4], t[5], t[6], t[7]);
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

template <class P>
struct alignas(16) SignSmem {
    int32_t tile[NTT_TILE_WORDS];
    uint32_t stage[32 * P::ZBITS / 4 + 1];  // packed z polynomial (576/640 B) or the hint bytes
};

// One pass of the second half of a rejection-loop iteration for one lane.  The three vector products
// c*s2, c*s1, c*t0 share ONE instance of the inverse-NTT code (a single loop over 2K+L polynomials); the
// checks run most-selective-first (r0, then z, then ct0/hints) so a rejected iteration stops early.  The
// order of the checks does not affect the outcome: the reference rejects if ANY fails (ml_dsa.rs:277-319).
template <class P, bool SHARED>
__global__ void __launch_bounds__(SG_WARPS * 32, 5)
sign_z_kernel(const uint32_t* __restrict__ nact_dev, const uint32_t* __restrict__ act,
              const int32_t* __restrict__ s1h_all,
              const int32_t* __restrict__ s2h_all, const int32_t* __restrict__ t0h_all,
              const uint8_t* __restrict__ yraw, int32_t* __restrict__ wall, const uint8_t* __restrict__ ctilde,
              const uint8_t* __restrict__ cmask, uint32_t* __restrict__ kappa, uint8_t* __restrict__ done,
              uint8_t* __restrict__ rhopp, uint8_t* __restrict__ sig, uint32_t* __restrict__ iters, int depth)
{
    using LO = Layout<P>;
    constexpr int K = P::K, L = P::L;
    const size_t nact = *nact_dev;
    __shared__ __align__(16) SignSmem<P> sm_all[SG_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SignSmem<P>& sm = sm_all[warp];
    int32_t* tile = sm.tile;
    uint8_t* hb = reinterpret_cast<uint8_t*>(sm.stage);
    const size_t stride = (size_t)gridDim.x * SG_WARPS;
    for (size_t ai = (size_t)blockIdx.x * SG_WARPS + warp; ai < nact; ai += stride) {
        const size_t ln = act[ai];
        const size_t kn = SHARED ? 0 : ln;
        uint8_t* sg = sig + ln * (size_t)P::SIG_LEN;
        bool rej = true;
        int cand = 0;
        // the lane's candidates in the reference's order: the first one that passes every check is the signature
#pragma unroll 1
        for (; cand < depth && rej; cand++) {
        const size_t slot = ln * depth + cand;
        const uint8_t* yp = yraw + slot * (size_t)(L * 32 * P::ZBITS);
        int32_t* wl = wall + slot * (size_t)K * 256;

        // first operands of the r0 phase stream in while c-hat is computed
        if (lane < 16) {
            const char* nxt = lane < 8 ? reinterpret_cast<const char*>(s2h_all + kn * (size_t)K * 256) + 128 * lane
                                       : reinterpret_cast<const char*>(wl) + 128 * (lane - 8);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt));
        }
        // c from the challenge masks, c-hat = NTT(c) (ml_dsa.rs:237-240) centred, with its Shoup companion
        int32_t ch[8], chs[8];
        load_c_layoutA(cmask + slot * (size_t)CMASK_STRIDE, lane, ch);
        sntt_fwd_A2B(ch, tile, lane);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            ch[i] = center(freeze(ch[i]));
            chs[i] = shoup_companion(ch[i]);
        }

        rej = false;
        int hcount = 0;
#pragma unroll 1
        for (int p = 0; p < 2 * K + L && !rej; p++) {
            const int phase = p < K ? 0 : (p < K + L ? 1 : 2);
            const int idx = phase == 0 ? p : (phase == 1 ? p - K : p - K - L);
            const int32_t* src = phase == 0 ? s2h_all + (kn * K + idx) * 256
                                            : (phase == 1 ? s1h_all + (kn * L + idx) * 256 : t0h_all + (kn * K + idx) * 256);
            const int4* sp = reinterpret_cast<const int4*>(src + 8 * lane);
            const int4 x = __ldg(sp), y4 = __ldg(sp + 1);
            const int32_t sv[8] = {x.x, x.y, x.z, x.w, y4.x, y4.y, y4.z, y4.w};
            // the s2-hat rows and w rows of one key / lane are contiguous: pull the next ones into L2
            if (phase == 0 && lane < 16) {
                const char* nxt = lane < 8 ? reinterpret_cast<const char*>(src + 256) + 128 * lane
                                           : reinterpret_cast<const char*>(wl + 256 * (idx + 1)) + 128 * (lane - 8);
                if (idx + 1 < K) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt));
            }
            int32_t t[8];
#pragma unroll
            for (int i = 0; i < 8; i++) t[i] = shoup_mul(sv[i], ch[i], chs[i]);  // s-hat rows are prescaled by 1/256
            sntt_inv_B2B(t, tile, lane);  // c*s2[idx] / c*s1[idx] / c*t0[idx], canonical

            bool bad = false;
            if (phase == 0) {
                // r0 = LowBits(w - c*s2), ||r0||inf < gamma2 - beta              ml_dsa.rs:253-260, 268-272
                int4* wp = reinterpret_cast<int4*>(wl + 256 * idx + 8 * lane);
                const int4 wx = wp[0], wy = wp[1];
                const int32_t wv[8] = {wx.x, wx.y, wx.z, wx.w, wy.x, wy.y, wy.z, wy.w};
                int32_t v[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    v[i] = caddq(wv[i] - t[i]);  // w - cs2, canonical
                    int32_t r0;
                    decompose<P::GAMMA2>(v[i], r0);
                    bad |= (r0 < 0 ? -r0 : r0) >= P::GAMMA2 - P::BETA;
                }
                rej = __any_sync(0xffffffffu, bad);
                if (!rej) {  // w is only ever needed as w - cs2 from here on
                    wp[0] = make_int4(v[0], v[1], v[2], v[3]);
                    wp[1] = make_int4(v[4], v[5], v[6], v[7]);
                }
            } else if (phase == 1) {
                // z = y + c*s1, ||z||inf < gamma1 - beta                          ml_dsa.rs:243-250, 263-265
                int32_t y[8];
                load_y<P>(yp + idx * 32 * P::ZBITS, lane, y);
                uint32_t f[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int32_t z = center(freeze(y[i] + t[i]));                 // z mod+- q (ml_dsa.rs:334-335)
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
                const int4* wp = reinterpret_cast<const int4*>(wl + 256 * idx + 8 * lane);
                const int4 wx = wp[0], wy = wp[1];
                const int32_t v[8] = {wx.x, wx.y, wx.z, wx.w, wy.x, wy.y, wy.z, wy.w};
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
            continue;
        }
        cand--;  // the accepted candidate
        // accept: c~ and hints complete the signature                            encodings.rs:254, 273
        __syncwarp();
        warp_copy_out(sg + LO::SIG_H, sm.stage, P::OMEGA + K, lane);
        warp_copy_bytes(sg, ctilde + (ln * depth + cand) * 64, P::CTILDE, lane);
        if (lane == 0) {
            done[ln] = 1;
            if (iters) iters[ln] = kappa[ln] / L + cand + 1;
        }
        // The lane is finished: zero everything it derived from the secret key and the mask (y, w - c*s2, rho'',
        // and its own s1-hat / s2-hat / t0-hat) right here, where the stores hide behind other warps' arithmetic,
        // instead of a 27 GB cudaMemset after the last round (src/types.rs:19 ZeroizeOnDrop analogue).
        warp_zero16(wall + ln * (size_t)depth * K * 256, depth * K * 1024, lane);
        warp_zero16(const_cast<uint8_t*>(yraw) + ln * (size_t)depth * (L * 32 * P::ZBITS), depth * L * 32 * P::ZBITS, lane);
        warp_zero16(rhopp + ln * 64, 64, lane);
        if (!SHARED) {
            warp_zero16(const_cast<int32_t*>(s1h_all) + ln * (size_t)L * 256, L * 1024, lane);
            warp_zero16(const_cast<int32_t*>(s2h_all) + ln * (size_t)K * 256, K * 1024, lane);
            warp_zero16(const_cast<int32_t*>(t0h_all) + ln * (size_t)K * 256, K * 1024, lane);
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256)
compact_kernel(const uint32_t* __restrict__ nact_dev, const uint32_t* __restrict__ act_in,
               const uint8_t* __restrict__ done, uint32_t* __restrict__ act_out, uint32_t* __restrict__ nout)
{
    const size_t nact = *nact_dev;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool keep = i < nact && !done[act_in[i]];
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (m == 0) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(nout, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (keep) act_out[base + __popc(m & ((1u << lane) - 1))] = act_in[i];
}

template <class P>
static cudaError_t sign_run_t(cudaStream_t st, size_t n, const uint8_t* sk, size_t sk_stride, const MsgDesc& md,
                              const uint8_t* rnd, uint8_t* sig, uint32_t* iters, const SignWs& ws, int* rounds_out)
{
    constexpr int K = P::K, L = P::L;
    const bool shared = sk_stride == 0;
    const size_t nkeys = shared ? 1 : n;
    const Tuning tn = tuning();
    const int D = ws.depth;
    auto warp_grid = [&](size_t items) {
        return capped_grid((items + SG_WARPS - 1) / SG_WARPS, tn.acap > 0 ? tn.acap : 16);
    };
    cudaError_t e;
    int* ctx_err = reinterpret_cast<int*>(ws.nact + 2);
    e = cudaMemsetAsync(ws.nact, 0, 16, st);
    if (e != cudaSuccess) return e;
    {
        KLaunch kl(KID_SIGN_PREP, st);
        sign_prep_kernel<P><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, sk, sk_stride, md, rnd, ws.mw1, ws.rhopp,
                                                                        ws.kappa, ws.done, ws.act[0], ws.nact, ctx_err, D);
    }
    {
        KLaunch kl(KID_SIGN_KEY, st);
        sign_key_kernel<P><<<warp_grid(nkeys), SG_WARPS * 32, 0, st>>>(nkeys, sk, sk_stride, ws.s1h, ws.s2h, ws.t0h);
    }
    e = launch_expand_a(st, K, L, sk, sk_stride, nkeys, ws.A);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(ws.h_nact, ctx_err, 4, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return e;
    if (ws.h_nact[0] != 0) {  // ctx too long: the caller maps this marker to -2 (lib.rs:274)
        ws.h_nact[1] = CTX_ERR_MARK;
        return cudaErrorInvalidValue;
    }

    // Round loop.  The kernels read the live active-lane count from device memory (nact[cur]); the host
    // only needs an upper bound to size the grids, so once the active set is small it re-reads the count
    // every few rounds instead of stalling the stream after each one.
    size_t bound = n;
    int cur = 0, rounds = 0, since_sync = 0;
    while (bound > 0) {
        {
            KLaunch kl(KID_MASK, st);
            mask_kernel<P><<<capped_grid((bound * L * D + 127) / 128, tn.kcap), 128, 0, st>>>(
                ws.nact + cur, ws.act[cur], ws.rhopp, ws.kappa, ws.yraw, D);
        }
        {
            KLaunch kl(KID_SIGN_W, st);
            if (shared)
                sign_w_kernel<P, true><<<warp_grid(bound * D), SG_WARPS * 32, 0, st>>>(ws.nact + cur, ws.act[cur], ws.A,
                                                                                     ws.yraw, ws.w, ws.mw1, D);
            else
                sign_w_kernel<P, false><<<warp_grid(bound * D), SG_WARPS * 32, 0, st>>>(ws.nact + cur, ws.act[cur], ws.A,
                                                                                      ws.yraw, ws.w, ws.mw1, D);
        }
        e = launch_ctilde(P::PS, st, bound * D, ws.nact + cur, ws.act[cur], D, ws.mw1, nullptr, nullptr, nullptr,
                          ws.ctilde, ws.cmask);
        if (e != cudaSuccess) return e;
        {
            KLaunch kl(KID_SIGN_Z, st);
            if (shared)
                sign_z_kernel<P, true><<<warp_grid(bound), SG_WARPS * 32, 0, st>>>(
                    ws.nact + cur, ws.act[cur], ws.s1h, ws.s2h, ws.t0h, ws.yraw, ws.w, ws.ctilde, ws.cmask, ws.kappa,
                    ws.done, ws.rhopp, sig, iters, D);
            else
                sign_z_kernel<P, false><<<warp_grid(bound), SG_WARPS * 32, 0, st>>>(
                    ws.nact + cur, ws.act[cur], ws.s1h, ws.s2h, ws.t0h, ws.yraw, ws.w, ws.ctilde, ws.cmask, ws.kappa,
                    ws.done, ws.rhopp, sig, iters, D);
        }
        e = cudaMemsetAsync(ws.nact + (cur ^ 1), 0, 4, st);
        if (e != cudaSuccess) return e;
        {
            KLaunch kl(KID_COMPACT, st);
            compact_kernel<<<(unsigned)((bound + 255) / 256), 256, 0, st>>>(ws.nact + cur, ws.act[cur], ws.done,
                                                                           ws.act[cur ^ 1], ws.nact + (cur ^ 1));
        }
        cur ^= 1;
        rounds++;
        const int every = bound > 65536 ? 1 : (bound > 4096 ? 3 : 6);
        if (++since_sync >= every) {
            e = cudaMemcpyAsync(ws.h_nact, ws.nact + cur, 4, cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) return e;
            e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) return e;
            bound = ws.h_nact[0];
            since_sync = 0;
        }
    }
    if (rounds_out) *rounds_out = rounds;
    return cudaGetLastError();
}

cudaError_t sign_run(int ps, cudaStream_t st, size_t n, const uint8_t* sk, size_t sk_stride, const MsgDesc& md,
                     const uint8_t* rnd, uint8_t* sig, uint32_t* iters, const SignWs& ws, int* rounds_out)
{
    if (rounds_out) *rounds_out = 0;
    if (n == 0) return cudaSuccess;
    switch (ps) {
    case 44: return sign_run_t<Params<44>>(st, n, sk, sk_stride, md, rnd, sig, iters, ws, rounds_out);
    case 65: return sign_run_t<Params<65>>(st, n, sk, sk_stride, md, rnd, sig, iters, ws, rounds_out);
    case 87: return sign_run_t<Params<87>>(st, n, sk, sk_stride, md, rnd, sig, iters, ws, rounds_out);
    }
    return cudaErrorInvalidValue;
}

}  // namespace mldsa
