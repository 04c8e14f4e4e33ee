// k_verify.cu -- batched ML-DSA.Verify_internal over packed bytes.
// Replaces src/ml_dsa.rs:351-437 (verify_internal) with expand_public (ml_dsa.rs:477-498) and
// sig_decode / pk_decode (src/encodings.rs:292-328, :55-80) folded in.
//
// Pipeline over a chunk of lanes (one lane = one (pk, message, signature)):
//   [key prep]    one shared key: expand_a, tr = H(pk) and NTT(t1 * 2^d), once; resident key sets bring theirs;
//                 per-lane packed keys: none -- verify_fused (k_fused.cu) squeezes A-hat inside the product
//   verify_pre    thread/lane  tr (per-lane keys), mu = H(tr || M'), c = SampleInBall(c~) as bit masks
//   verify_arith  warp/lane    hint decode + checks, z unpack + norm, NTTs,
//                              w' = INTT(A-hat o z-hat - c-hat o t1-hat), UseHint, w1Encode
//   ctilde_check  thread/lane  c~' = H(mu || w1'), compare with c~, AND with the pre-checks
#include "engine.h"
#include "keccak.cuh"
#include "ntt.cuh"
#include "pack.cuh"
#include "sampling.cuh"

namespace mldsa {

template <class P>
__global__ void __launch_bounds__(128)
verify_pre_kernel(size_t n, const uint8_t* __restrict__ pk, size_t pk_stride, const uint8_t* __restrict__ tr_shared,
                  MsgDesc md, const uint8_t* __restrict__ sig, uint8_t* __restrict__ mw1, uint8_t* __restrict__ cmask,
                  uint8_t* __restrict__ pre_ok, size_t tr_stride, const uint32_t* __restrict__ kidx)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Sponge<SHAKE256_RATE> sp;
    uint64_t tr[8];
    if (tr_shared) {  // cached tr: one shared key (tr_stride 0) or a resident key set (64-byte rows, key kidx[i] or i)
        const uint64_t* row = reinterpret_cast<const uint64_t*>(tr_shared + (kidx ? (size_t)kidx[i] : i) * tr_stride);
#pragma unroll
        for (int k = 0; k < 8; k++) tr[k] = row[k];
    } else {
        sp.init();
        sp.absorb(pk + i * pk_stride, P::PK_LEN);  // ml_dsa.rs:486-488
        sp.finish();
        sp.out_words(tr, 8);
    }
    uint8_t* mu = mw1 + i * (size_t)(MU_BYTES + P::W1_LEN);
    const bool ctx_ok = absorb_mu(sp, reinterpret_cast<const uint8_t*>(tr), md, i);  // ml_dsa.rs:386-397
    sp.out_words(reinterpret_cast<uint64_t*>(mu), 8);
    pre_ok[i] = ctx_ok ? 1 : 0;  // ctx > 255 -> false (lib.rs:368-370)
    // c = SampleInBall(c~) (ml_dsa.rs:400, hashing.rs:43-100), stored as nonzero / negative masks
    sp.init();
    const uint8_t* ct = sig + i * (size_t)P::SIG_LEN;
    for (int k = 0; k < P::CTILDE; k++) sp.absorb_byte(ct[k]);
    sp.finish();
    SibWalk walk;
    walk.init(P::TAU);
    while (!walk.consume(sp.s, P::TAU)) sp.next_block();
    walk.store(cmask + i * (size_t)CMASK_STRIDE);
}

// NTT(t1 * 2^d) for one polynomial of a public key, from the 10-bit packed bytes (layout B regs)
__device__ __forceinline__ void load_t1_hat(const uint8_t* pkpoly, int32_t (&v)[8], int32_t* tile, int lane)
{
    uint32_t w[3];
    load_span<10>(pkpoly + 10 * lane, w);
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = (int32_t)(get_field<10>(w, i) << D);  // pk_decode + ml_dsa.rs:490-495
    sntt_fwd_B2B(v, tile, lane);
    // carries the same 2^32 / 256 as the prescaled A-hat: both terms of A z - c t1 then share one mont_reduce and an
    // un-scaled inverse transform
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = shoup_mul(v[i], SC_MONT_INV256, SC_MONT_INV256S);
}

template <class P>
__global__ void __launch_bounds__(128)
t1hat_kernel(const uint8_t* __restrict__ pk, int32_t* __restrict__ t1h)
{
    __shared__ __align__(16) int32_t tiles[4][NTT_TILE_WORDS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = warp; k < P::K; k += 4) {
        int32_t v[8];
        load_t1_hat(pk + 32 + 320 * k, v, tiles[warp], lane);
        st_row8(t1h + 256 * k + 8 * lane, v);
    }
}

// NTT(t1 * 2^d) for every polynomial of nkeys public keys (expand_public, src/ml_dsa.rs:490-495): warp per polynomial
template <class P>
__global__ void __launch_bounds__(128)
t1hat_batch_kernel(size_t nkeys, const uint8_t* __restrict__ pk, int32_t* __restrict__ t1h)
{
    __shared__ __align__(16) int32_t tiles[4][NTT_TILE_WORDS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t npoly = nkeys * P::K;
    for (size_t q = (size_t)blockIdx.x * 4 + warp; q < npoly; q += (size_t)gridDim.x * 4) {
        const size_t key = q / P::K;
        const int k = (int)(q % P::K);
        int32_t v[8];
        load_t1_hat(pk + key * (size_t)P::PK_LEN + 32 + 320 * k, v, tiles[warp], lane);
        st_row8(t1h + q * 256 + 8 * lane, v);
    }
}

constexpr int VF_WARPS = 4;

template <class P>
struct alignas(16) VerifySmem {
    int32_t tile[NTT_TILE_WORDS];
    uint32_t hbits[P::K * 8];                       // hint bitmap, 256 bits per polynomial
    uint32_t stage[(P::OMEGA + P::K + 3) / 4 + 1];  // hint bytes
    alignas(16) int32_t zh[P::L][256];              // z-hat, NTT domain, natural order
};

// MODE 1: one shared key (A-hat, t1-hat cached once); MODE 2: resident key set (A-hat, t1-hat cached per key, lane
// ln uses key kidx[ln] or ln) -- what the reference's PublicKey struct caches (src/types.rs:35-41).  Per-lane packed
// keys go through verify_fused_kernel (k_fused.cu), which never materialises A-hat.
template <class P, int MODE>
__global__ void __launch_bounds__(VF_WARPS * 32)
verify_arith_kernel(size_t n, const uint8_t* __restrict__ sig,
                    const int32_t* __restrict__ Aall, const int32_t* __restrict__ t1h_shared,
                    const uint8_t* __restrict__ cmask, uint8_t* __restrict__ mw1, uint8_t* __restrict__ pre_ok,
                    const uint32_t* __restrict__ kidx)
{
    using LO = Layout<P>;
    constexpr int K = P::K, L = P::L;
    constexpr bool SHARED = MODE == 1;
    __shared__ __align__(16) VerifySmem<P> sm_all[VF_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    VerifySmem<P>& sm = sm_all[warp];
    int32_t* tile = sm.tile;
    const size_t stride = (size_t)gridDim.x * VF_WARPS;
    for (size_t ln = (size_t)blockIdx.x * VF_WARPS + warp; ln < n; ln += stride) {
        const uint8_t* sg = sig + ln * (size_t)P::SIG_LEN;
        const size_t key = MODE == 2 ? (kidx ? (size_t)kidx[ln] : ln) : ln;
        const int32_t* A = Aall + (SHARED ? 0 : key * (size_t)(K * L) * 256);
        const int32_t* t1h = t1h_shared + (MODE == 2 ? key * (size_t)K * 256 : 0);
        uint8_t* w1o = mw1 + ln * (size_t)(MU_BYTES + P::W1_LEN) + MU_BYTES;
        __syncwarp();
        // the signature this warp handles next streams from HBM into L2 while this one is verified
        if (ln + stride < n) warp_prefetch_l2(sig + (ln + stride) * (size_t)P::SIG_LEN, P::SIG_LEN, lane);

        // ---- HintBitUnpack with its three malformed-input exits (conversion.rs:357, :373, :398) ----
        uint8_t* hb = reinterpret_cast<uint8_t*>(sm.stage);
        warp_copy_bytes(hb, sg + LO::SIG_H, P::OMEGA + K, lane);
        if (lane < K * 8) sm.hbits[lane] = 0;
        if (K * 8 > 32 && lane + 32 < K * 8) sm.hbits[lane + 32] = 0;
        __syncwarp();
        bool bad = false;
        if (lane < K) {
            const int cur = hb[P::OMEGA + lane], prev = lane ? hb[P::OMEGA + lane - 1] : 0;
            bad = cur < prev || cur > P::OMEGA;
        }
        bad = __any_sync(0xffffffffu, bad);
        if (!bad) {
            const int total = hb[P::OMEGA + K - 1];
            for (int p = lane; p < P::OMEGA; p += 32) {
                if (p >= total) {
                    bad |= hb[p] != 0;
                } else {
                    int poly = 0, first = 0;
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        const int c = hb[P::OMEGA + k];
                        if (p >= c) { poly = k + 1; first = c; }
                    }
                    if (p > first) bad |= hb[p - 1] >= hb[p];
                    atomicOr(&sm.hbits[poly * 8 + (hb[p] >> 5)], 1u << (hb[p] & 31));
                }
            }
            bad = __any_sync(0xffffffffu, bad);
        }
        if (bad) {  // verify returns false (ml_dsa.rs:368-376); warp-uniform exit
            if (lane == 0) pre_ok[ln] = 0;
            continue;
        }

        // ---- z = BitUnpack(gamma1 - 1, gamma1), ||z||inf check, z-hat (kept in shared memory) ----
        // the packed bytes of polynomial l + 1 are in flight during the transform of polynomial l
        bool zbad = false;
        uint32_t zw[2][(P::ZBITS + 3) / 4];
        load_span<P::ZBITS>(sg + LO::SIG_Z + P::ZBITS * lane, zw[0]);
#pragma unroll
        for (int l = 0; l < L; l++) {
            if (l + 1 < L) load_span<P::ZBITS>(sg + LO::SIG_Z + (l + 1) * 32 * P::ZBITS + P::ZBITS * lane, zw[(l + 1) & 1]);
            int32_t zv[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int32_t z = P::GAMMA1 - (int32_t)get_field<P::ZBITS>(zw[l & 1], i);  // conversion.rs:250
                zbad |= (z < 0 ? -z : z) >= P::GAMMA1 - P::BETA;                           // ml_dsa.rs:434
                zv[i] = z;
            }
            sntt_fwd_B2B(zv, tile, lane);
            row_store(sm.zh[l], lane, zv);
        }
        zbad = __any_sync(0xffffffffu, zbad);
        if (zbad) {
            if (lane == 0) pre_ok[ln] = 0;
            continue;
        }

        // ---- c from the challenge masks, c-hat = NTT(c) ----
        int32_t ch[8];
        load_c_layoutA(cmask + ln * (size_t)CMASK_STRIDE, lane, ch);
        sntt_fwd_A2B(ch, tile, lane);
#pragma unroll
        for (int i = 0; i < 8; i++) ch[i] = reduce32(ch[i]);

        // ---- per row: w' = INTT(sum_l A[k][l] o z-hat[l] - c-hat o t1-hat[k]); UseHint; w1Encode ----
        if (!SHARED) warp_prefetch_l2(A, 2 * L * 1024, lane);
#pragma unroll 1
        for (int k = 0; k < K; k++) {
            if (!SHARED && k + 2 < K) warp_prefetch_l2(A + (size_t)(k + 2) * L * 256, L * 1024, lane);
            // one shared key: its 36 KiB of A-hat / t1-hat stay hot in L1 (128-bit __ldg); a resident key set streams
            // each lane's rows once (256-bit loads, pack.cuh)
            auto load_row = [&](const int32_t* r) {
                if (!SHARED) return ldg_row8(r);
                Row8 o;
                const int4 x = __ldg(reinterpret_cast<const int4*>(r)), y = __ldg(reinterpret_cast<const int4*>(r) + 1);
                o.v[0] = x.x; o.v[1] = x.y; o.v[2] = x.z; o.v[3] = x.w; o.v[4] = y.x; o.v[5] = y.y; o.v[6] = y.z; o.v[7] = y.w;
                return o;
            };
            const Row8 t1v = load_row(t1h + 256 * k + 8 * lane);
            int64_t acc[8];
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = -(int64_t)ch[i] * t1v.v[i];
#pragma unroll
            for (int l = 0; l < L; l++) {
                const Row8 av = load_row(A + (size_t)(k * L + l) * 256 + 8 * lane);
                int32_t zz[8];
                row_load(sm.zh[l], lane, zz);
#pragma unroll
                for (int i = 0; i < 8; i++) acc[i] += (int64_t)av.v[i] * zz[i];
            }
            int32_t t[8];
#pragma unroll
            for (int i = 0; i < 8; i++) t[i] = mont_reduce(acc[i]);
            sntt_inv_B2B(t, tile, lane);
            const uint32_t hbyte = (sm.hbits[k * 8 + (lane >> 2)] >> (8 * (lane & 3))) & 0xFF;
            uint32_t f[8];
#pragma unroll
            for (int i = 0; i < 8; i++) f[i] = (uint32_t)use_hint<P::GAMMA2>((hbyte >> i) & 1, t[i]);
            uint32_t w[(P::W1BITS + 3) / 4];
            pack_fields<P::W1BITS>(f, w);                                           // encodings.rs:338-360
            store_span<P::W1BITS>(w1o + k * 32 * P::W1BITS + P::W1BITS * lane, w);
        }
    }
}

template <class P>
static cudaError_t verify_run_t(cudaStream_t st, size_t n, const uint8_t* pk, size_t pk_stride, const MsgDesc& md,
                                const uint8_t* sig, uint8_t* ok, const VerifyWs& ws)
{
    constexpr int K = P::K, L = P::L;
    const bool resident = ws.keys_resident, shared = !resident && pk_stride == 0;
    cudaError_t e;
    if (resident) {
        // A-hat, t1-hat and tr of every key were built by verify_expand_keys
    } else if (shared) {
        e = launch_expand_a(st, K, L, pk, 0, 1, ws.A);
        if (e != cudaSuccess) return e;
        e = launch_shake256_rows(st, pk, 0, P::PK_LEN, 1, ws.tr, 64, 64);
        if (e != cudaSuccess) return e;
        {
            KLaunch kl(KID_T1HAT, st);
            t1hat_kernel<P><<<1, 128, 0, st>>>(pk, ws.t1h);
        }
    }
    {
        KLaunch kl(KID_VERIFY_PRE, st);
        verify_pre_kernel<P><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(
            n, pk, pk_stride, (shared || resident) ? ws.tr : nullptr, md, sig, ws.mw1, ws.cmask, ws.pre_ok,
            resident ? (size_t)64 : 0, resident ? ws.kidx : nullptr);
    }
    // One shared key: few long-lived CTAs (measured at 2^20 lanes, ML-DSA-65: 16 per SM -> 10.9 ms, 192 per SM ->
    // 24.4 ms: the key's 36 KiB of A-hat / t1-hat stay hot in each SM's L1); per-lane keys: many short CTAs (no partly
    // empty last wave)
    // per-lane packed keys: A-hat is consumed as it is squeezed, never written (k_fused.cu)
    if (!resident && !shared) {
        e = launch_verify_fused(P::PS, st, n, pk, pk_stride, sig, ws.cmask, ws.mw1, ws.pre_ok);
        if (e != cudaSuccess) return e;
    } else {
        const unsigned grid = capped_grid((n + VF_WARPS - 1) / VF_WARPS, shared ? 16 : arith_cap());
        KLaunch kl(KID_VERIFY_ARITH, st);
        if (resident)
            verify_arith_kernel<P, 2><<<grid, VF_WARPS * 32, 0, st>>>(n, sig, ws.A, ws.t1h, ws.cmask, ws.mw1, ws.pre_ok, ws.kidx);
        else
            verify_arith_kernel<P, 1><<<grid, VF_WARPS * 32, 0, st>>>(n, sig, ws.A, ws.t1h, ws.cmask, ws.mw1, ws.pre_ok, nullptr);
    }
    e = launch_ctilde(P::PS, st, n, nullptr, nullptr, 1, ws.mw1, /*verify*/ sig, ws.pre_ok, ok, nullptr, nullptr);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

// expand_public (src/ml_dsa.rs:477-498) for nkeys packed public keys: A-hat, NTT(t1 * 2^d), tr = H(pk)
template <class P>
static cudaError_t verify_expand_keys_t(cudaStream_t st, size_t nkeys, const uint8_t* pk, int32_t* A, int32_t* t1h, uint8_t* tr,
                                        bool with_A, bool with_tr)
{
    cudaError_t e = cudaSuccess;
    if (with_A) e = launch_expand_a(st, P::K, P::L, pk, P::PK_LEN, nkeys, A);
    if (e != cudaSuccess) return e;
    if (with_tr) e = launch_shake256_rows(st, pk, P::PK_LEN, P::PK_LEN, nkeys, tr, 64, 64);
    if (e != cudaSuccess) return e;
    {
        KLaunch kl(KID_T1HAT, st);
        t1hat_batch_kernel<P><<<capped_grid((nkeys * P::K + 3) / 4, arith_cap()), 128, 0, st>>>(nkeys, pk, t1h);
    }
    return cudaGetLastError();
}
cudaError_t verify_expand_keys(int ps, cudaStream_t st, size_t nkeys, const uint8_t* pk, int32_t* A, int32_t* t1h, uint8_t* tr,
                               bool with_A, bool with_tr)
{
    if (nkeys == 0) return cudaSuccess;
    switch (ps) {
    case 44: return verify_expand_keys_t<Params<44>>(st, nkeys, pk, A, t1h, tr, with_A, with_tr);
    case 65: return verify_expand_keys_t<Params<65>>(st, nkeys, pk, A, t1h, tr, with_A, with_tr);
    case 87: return verify_expand_keys_t<Params<87>>(st, nkeys, pk, A, t1h, tr, with_A, with_tr);
    }
    return cudaErrorInvalidValue;
}

cudaError_t verify_run(int ps, cudaStream_t st, size_t n, const uint8_t* pk, size_t pk_stride, const MsgDesc& md,
                       const uint8_t* sig, uint8_t* ok, const VerifyWs& ws)
{
    if (n == 0) return cudaSuccess;
    switch (ps) {
    case 44: return verify_run_t<Params<44>>(st, n, pk, pk_stride, md, sig, ok, ws);
    case 65: return verify_run_t<Params<65>>(st, n, pk, pk_stride, md, sig, ok, ws);
    case 87: return verify_run_t<Params<87>>(st, n, pk, pk_stride, md, sig, ok, ws);
    }
    return cudaErrorInvalidValue;
}

}  // namespace mldsa
