// k_keygen.cu -- batched ML-DSA.KeyGen_internal, bytes out.
// Replaces src/ml_dsa.rs:57-134 (key_gen_internal) followed by SerDes::into_bytes
// (src/lib.rs:427-493).  The reference's struct pre-computation (ml_dsa.rs:108-127: L+3K extra
// NTTs) has no effect on the encoded keys and is not performed.
//
// Pipeline over a chunk of lanes (one lane = one key):
//   kg_seed      thread/lane   (rho, rho', K) = H(xi || k || l, 128)              ml_dsa.rs:68-74
//   rej_bounded  thread/stream s1, s2 = RejBoundedPoly(rho' || LE16(r))            hashing.rs:158-213, 252-272
//   kg_fused     (k_fused.cu)  t = INTT(A-hat o NTT(s1)) + s2 with A-hat consumed as it is squeezed, Power2Round,
//                              pkEncode / skEncode                                 hashing.rs:225-239, ml_dsa.rs:85-92
//   -- or, when the caller keeps A-hat (key sets, the chained pipeline): expand_a (thread/stream) then kg_arith
//      (warp/lane: the same arithmetic reading A-hat back), encodings.rs:18-152
//   shake rows   thread/lane   tr = H(pk, 64) written into sk[64..128]             ml_dsa.rs:99-101
#include "engine.h"
#include "keccak.cuh"
#include "ntt.cuh"
#include "pack.cuh"

namespace mldsa {

__global__ void __launch_bounds__(128) kg_seed_kernel(const uint8_t* __restrict__ xi, size_t n, int K, int L,
                                                      uint8_t* __restrict__ seeds)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Sponge<SHAKE256_RATE> sp;
    sp.init();
    sp.absorb(xi + 32 * i, 32);
    sp.absorb_byte((uint32_t)K);
    sp.absorb_byte((uint32_t)L);
    sp.finish();
    sp.out_words(reinterpret_cast<uint64_t*>(seeds + 128 * i), 16);
}

// RejBoundedPoly: one thread per SHAKE256 stream; nibbles are parsed out of the state registers
// (CoeffFromHalfByte, src/conversion.rs:80-111) and the accepted values appended to the thread's own row of a
// shared tile; after every squeeze block the thread moves its complete groups of sixteen to their final place
// with 128-bit stores and carries the rest to the head of the row (the scheme of expand_a: no warp-cooperative
// flush, whose 32 x (2 shuffles + copy loop) per block cost this ALU-bound kernel 30 % of its instructions).
// Output: int8 coefficients in [-eta, eta].
constexpr int RB_THREADS = 128;
// Bytes per thread: <= 15 carried + 272 nibbles of a block, as whole int4s = 288, padded to 304 = 76 words = 12 (mod 32):
// the 128-bit flush loads of a quarter-warp then cover all 32 banks, the byte stores are 4-way conflicts (288 = 8 (mod
// 32) words made them 8-way and the loads 2-way; see EA_ROW in k_common.cu).
constexpr int RB_ROW = 304;

template <int ETA>
__global__ void __launch_bounds__(RB_THREADS, 4)
rej_bounded_kernel(const uint8_t* __restrict__ seeds, size_t nstreams, int KL /* K + L */, int8_t* __restrict__ out)
{
    __shared__ __align__(16) int8_t tile_all[RB_THREADS * RB_ROW];
    int8_t* row = tile_all + threadIdx.x * RB_ROW;  // private to this thread: no barriers needed
    const size_t q = (size_t)blockIdx.x * RB_THREADS + threadIdx.x;
    const bool valid = q < nstreams;

    uint64_t a[25];
#pragma unroll
    for (int i = 0; i < 25; i++) a[i] = 0;
    if (valid) {
        const size_t key = q / KL;
        const uint32_t idx = (uint32_t)(q % KL);  // r for s1, r + L for s2 (hashing.rs:260,266)
        const uint64_t* rp = reinterpret_cast<const uint64_t*>(seeds + 128 * key + 32);
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = rp[i];
        a[8] = (uint64_t)(idx | (0x1Fu << 16));  // LE16(idx) then SHAKE pad
        a[16] = 0x8000000000000000ULL;           // byte 135 of the 136-byte rate
    }
    int4* dst = reinterpret_cast<int4*>(out + (valid ? q : 0) * 256);  // next group of sixteen to write
    int groups_left = valid ? 16 : 0;
    int have = 0;  // coefficients waiting at the head of the row (0..15 between blocks)
    while (__any_sync(0xffffffffu, groups_left > 0)) {
        keccak_f1600(a);
#pragma unroll
        for (int wi = 0; wi < 34; wi++) {
            const uint32_t w = (wi & 1) ? (uint32_t)(a[wi >> 1] >> 32) : (uint32_t)a[wi >> 1];
#pragma unroll
            for (int nb = 0; nb < 8; nb++) {
                const int b = (int)((w >> (4 * nb)) & 15);
                if (ETA == 2) {
                    if (b < 15) row[have++] = (int8_t)(2 - (b - 5 * ((b * 205) >> 10)));
                } else {
                    if (b < 9) row[have++] = (int8_t)(4 - b);
                }
            }
        }
        const int ng = min(have >> 4, groups_left);
        const int4* src = reinterpret_cast<const int4*>(row);
#pragma unroll
        for (int k = 0; k < (15 + 272) / 16; k++)  // at most 17 complete groups per block
            if (k < ng) dst[k] = src[k];
        const int4 tail = src[have >> 4];  // the incomplete group (a finished stream parks garbage there)
        dst += ng;
        groups_left -= ng;
        have &= 15;
        *reinterpret_cast<int4*>(row) = tail;
    }
    // the row held secret coefficients: leave nothing of them in shared memory
#pragma unroll
    for (int k = 0; k < RB_ROW / 16; k++) reinterpret_cast<int4*>(row)[k] = make_int4(0, 0, 0, 0);
}

constexpr int KG_WARPS = 4;

template <class P>
__global__ void __launch_bounds__(KG_WARPS * 32)
kg_arith_kernel(size_t n, const uint8_t* __restrict__ seeds, const int8_t* __restrict__ sall,
                const int32_t* __restrict__ Aall, uint8_t* __restrict__ pk, uint8_t* __restrict__ sk)
{
    using LO = Layout<P>;
    constexpr int K = P::K, L = P::L;
    __shared__ __align__(16) int32_t tiles[KG_WARPS][NTT_TILE_WORDS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t* tile = tiles[warp];
    const size_t stride = (size_t)gridDim.x * KG_WARPS;
    for (size_t ln = (size_t)blockIdx.x * KG_WARPS + warp; ln < n; ln += stride) {
        const uint8_t* sd = seeds + 128 * ln;
        const int8_t* sp = sall + ln * (size_t)(K + L) * 256;
        const int32_t* A = Aall + ln * (size_t)(K * L) * 256;
        uint8_t* pko = pk + ln * (size_t)P::PK_LEN;
        uint8_t* sko = sk + ln * (size_t)P::SK_LEN;
        // rho -> pk[0..32], sk[0..32]; K -> sk[32..64]          (encodings.rs:27, :112-113)
        pko[lane] = sd[lane];
        sko[lane] = sd[lane];
        sko[32 + lane] = sd[96 + lane];

        warp_prefetch_l2(A, 2 * L * 1024, lane);
        int32_t s1h[L][8];
#pragma unroll
        for (int l = 0; l < L; l++) {
            const uint2 raw = *reinterpret_cast<const uint2*>(sp + l * 256 + 8 * lane);
            uint32_t f[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int32_t c = (int32_t)(int8_t)(((i < 4 ? raw.x : raw.y) >> (8 * (i & 3))) & 0xFF);
                s1h[l][i] = c;
                f[i] = (uint32_t)(P::ETA - c);  // BitPack(s1, eta, eta): field = eta - s   (conversion.rs:164)
            }
            uint32_t w[(P::ETABITS + 3) / 4];
            pack_fields<P::ETABITS>(f, w);
            store_span<P::ETABITS>(sko + LO::SK_S1 + l * 32 * P::ETABITS + P::ETABITS * lane, w);
            sntt_fwd_B2B(s1h[l], tile, lane);
        }
#pragma unroll 1
        for (int k = 0; k < K; k++) {
            if (k + 2 < K) warp_prefetch_l2(A + (size_t)(k + 2) * L * 256, L * 1024, lane);
            int64_t acc[8];
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = 0;
#pragma unroll
            for (int l = 0; l < L; l++) {
                const Row8 av = ldg_row8(A + (size_t)(k * L + l) * 256 + 8 * lane);  // read once: one 256-bit load
#pragma unroll
                for (int i = 0; i < 8; i++) acc[i] += (int64_t)av.v[i] * s1h[l][i];
            }
            int32_t t[8];
#pragma unroll
            for (int i = 0; i < 8; i++) t[i] = mont_reduce(acc[i]);  // A-hat is prescaled by 2^32 / 256
            sntt_inv_B2B(t, tile, lane);

            const uint2 raw = *reinterpret_cast<const uint2*>(sp + (L + k) * 256 + 8 * lane);
            uint32_t fs[8], f1[8], f0[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int32_t c = (int32_t)(int8_t)(((i < 4 ? raw.x : raw.y) >> (8 * (i & 3))) & 0xFF);
                fs[i] = (uint32_t)(P::ETA - c);
                const int32_t tt = freeze(t[i] + c);                    // ml_dsa.rs:88-91
                const int32_t r1 = (tt + (1 << (D - 1)) - 1) >> D;      // high_low.rs:26-27
                const int32_t r0 = tt - (r1 << D);                      // high_low.rs:30-31
                f1[i] = (uint32_t)r1;
                f0[i] = (uint32_t)((1 << (D - 1)) - r0);                // BitPack(t0, 2^12 - 1, 2^12)
            }
            uint32_t ws[(P::ETABITS + 3) / 4], w1[3], w0[4];
            pack_fields<P::ETABITS>(fs, ws);
            store_span<P::ETABITS>(sko + LO::SK_S2 + k * 32 * P::ETABITS + P::ETABITS * lane, ws);
            pack_fields<10>(f1, w1);
            store_span<10>(pko + 32 + 320 * k + 10 * lane, w1);
            pack_fields<13>(f0, w0);
            store_span<13>(sko + LO::SK_T0 + 416 * k + 13 * lane, w0);
        }
    }
}

template <class P>
static cudaError_t keygen_run_t(cudaStream_t st, size_t n, const uint8_t* xi, uint8_t* pk, uint8_t* sk,
                                const KeygenWs& ws)
{
    constexpr int K = P::K, L = P::L;
    {
        KLaunch kl(KID_KG_SEED, st);
        kg_seed_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(xi, n, K, L, ws.seeds);
    }
    const size_t ns = n * (size_t)(K + L);
    {
        KLaunch kl(KID_REJ_BOUNDED, st);
        rej_bounded_kernel<P::ETA><<<(unsigned)((ns + RB_THREADS - 1) / RB_THREADS), RB_THREADS, 0, st>>>(
            ws.seeds, ns, K + L, ws.s);
    }
    cudaError_t e;
    if (ws.A == nullptr) {  // A-hat consumed as it is squeezed (k_fused.cu)
        e = launch_kg_fused(P::PS, st, n, ws.seeds, ws.s, pk, sk);
        if (e != cudaSuccess) return e;
    } else {                // A-hat kept for the caller (key sets, chained pipeline)
        e = launch_expand_a(st, K, L, ws.seeds, 128, n, ws.A);
        if (e != cudaSuccess) return e;
        const unsigned grid = capped_grid((n + KG_WARPS - 1) / KG_WARPS, arith_cap());
        KLaunch kl(KID_KG_ARITH, st);
        kg_arith_kernel<P><<<grid, KG_WARPS * 32, 0, st>>>(n, ws.seeds, ws.s, ws.A, pk, sk);
    }
    e = launch_shake256_rows(st, pk, P::PK_LEN, P::PK_LEN, n, sk + 64, P::SK_LEN, 64);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

cudaError_t keygen_run(int ps, cudaStream_t st, size_t n, const uint8_t* xi, uint8_t* pk, uint8_t* sk,
                       const KeygenWs& ws)
{
    if (n == 0) return cudaSuccess;
    switch (ps) {
    case 44: return keygen_run_t<Params<44>>(st, n, xi, pk, sk, ws);
    case 65: return keygen_run_t<Params<65>>(st, n, xi, pk, sk, ws);
    case 87: return keygen_run_t<Params<87>>(st, n, xi, pk, sk, ws);
    }
    return cudaErrorInvalidValue;
}

}  // namespace mldsa
