// k_common.cu -- kernels shared by keygen / sign / verify and the microbench entry points:
//   * expand_a       : thread-per-stream SHAKE128 + RejNTTPoly   (src/hashing.rs:111-146, :225-239)
//   * ntt / inv_ntt  : warp-per-polynomial transforms on int32   (src/ntt.rs:14-161)
//   * shake rows     : thread-per-message SHAKE128/256 (tr = H(pk); also the Keccak test hook)
#include <cstdlib>

#include "engine.h"
#include "keccak.cuh"
#include "ntt.cuh"
#include "ntt16.cuh"
#include "pack.cuh"
#include "sampling.cuh"
#include "sha2.cuh"
#include "ctilde.cuh"

namespace mldsa {

// ------------------------------------------------------------------------------------------------
// expand_a.  One thread owns one SHAKE128 stream (one polynomial of A-hat).  Per squeeze block the
// thread parses its 56 three-byte candidates straight out of the state registers
// (CoeffFromThreeBytes, src/conversion.rs:40-61) and appends the accepted ones to its own row of a
// shared-memory tile; after the block it moves the complete groups of four to their final position
// with 128-bit stores and keeps the 0..3 left-over coefficients at the head of the row for the next
// block, so a stream's coefficient j is its j-th accepted candidate.  The kernel is ALU-pipe bound
// (Keccak); the staging costs that pipe ~1 predicate per candidate and per store, everything else
// runs on the LSU / FMA pipes.  (Tried and measured worse: a warp-cooperative coalesced flush -- 15 %
// more ALU instructions -- and 32-bit stores straight from registers -- LSU bound, ncu long_scoreboard 54 %.)
// The loop runs until every stream of the warp has 256 coefficients: 5 blocks in practice, a 6th with
// p ~ 1e-39 -- handled exactly, not assumed away.
// ------------------------------------------------------------------------------------------------
constexpr int EA_THREADS = 128;
// Words per thread: 3 carried + 56 candidates, rounded up to whole int4s = 60, then padded to 68 = 4 (mod 32).  With
// rows 64 words apart every thread's word j sat in the SAME bank: each of the 56 candidate stores of a block was a
// 32-way conflict and each 128-bit flush load an 8-way one (ncu: 9.5e9 bank conflicts per launch, the L1 data pipe at
// 98 % of its wavefront peak -- the kernel was bound by it as much as by the ALU pipe, which sat at 91 %).  At 68 the
// 128-bit loads of a quarter-warp cover all 32 banks and the 32-bit stores are 4-way conflicts (the minimum for rows of
// contiguous int4s): ~280 instead of ~2 300 wavefronts per block and warp.
#ifndef EA_ROW_WORDS
#define EA_ROW_WORDS 68
#endif
constexpr int EA_ROW = EA_ROW_WORDS;

// SYNTH (test hook, mldsa_debug_rej_ntt_stream): the squeeze blocks come from a caller-supplied byte stream
// (rho_stride / 168 blocks of 168 bytes per polynomial) instead of the permutation, so that a test can force the
// parser through more blocks than SHAKE128 ever will (a 6th block has p ~ 1e-39).
template <bool SEED34, bool SYNTH = false>
__global__ void __launch_bounds__(EA_THREADS, 4)
expand_a_kernel(const uint8_t* __restrict__ rho, size_t rho_stride, size_t nstreams, int KL, int L,
                int32_t* __restrict__ out)
{
    __shared__ __align__(16) int32_t tile_all[EA_THREADS * EA_ROW];
    int32_t* row = tile_all + threadIdx.x * EA_ROW;  // private to this thread: no barriers needed
    // grid-stride over groups of EA_THREADS streams
    for (size_t q0 = (size_t)blockIdx.x * EA_THREADS; q0 < nstreams; q0 += (size_t)gridDim.x * EA_THREADS) {
    const size_t q = q0 + threadIdx.x;
    const bool valid = q < nstreams;

    uint64_t a[25];
#pragma unroll
    for (int i = 0; i < 25; i++) a[i] = 0;
    if (valid && !SYNTH) {
        const size_t key = SEED34 ? q : q / KL;
        const int idx = SEED34 ? 0 : (int)(q % KL);
        uint32_t w[9];
        load_span<SEED34 ? 34 : 32>(rho + key * rho_stride, reinterpret_cast<uint32_t(&)[SEED34 ? 9 : 8]>(w));
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
        // rho || IntegerToBytes(s,1) || IntegerToBytes(r,1)   (hashing.rs:236), then SHAKE pad 0x1F
        const uint32_t sr = SEED34 ? (w[8] & 0xFFFF) : (uint32_t)((idx % L) | ((idx / L) << 8));
        a[4] = (uint64_t)(sr | (0x1Fu << 16));
        a[20] = 0x8000000000000000ULL;  // byte 167 of the 168-byte rate
    }

    int4* dst = reinterpret_cast<int4*>(out + (valid ? q : 0) * 256);  // next group of four to write
    int quads_left = valid ? 64 : 0;
    int have = 0;  // coefficients waiting at the head of the row (0..3 between blocks)
    size_t synth_blk = 0;
    while (__any_sync(0xffffffffu, quads_left > 0)) {
        if (SYNTH) {
            if (valid && (synth_blk + 1) * 168 <= rho_stride) {
                const uint8_t* bp = rho + q * rho_stride + synth_blk * 168;
#pragma unroll
                for (int i = 0; i < 21; i++) {
                    uint64_t w = 0;
                    for (int k = 0; k < 8; k++) w |= (uint64_t)bp[8 * i + k] << (8 * k);
                    a[i] = w;
                }
            } else {
                quads_left = 0;  // stream exhausted: whatever was not produced stays unwritten
            }
            synth_blk++;
        } else {
            keccak_f1600(a);
        }
        // The A-hat that feeds the fused kernels is stored prescaled by 2^32 / 256 mod q (the Montgomery factor of
        // the matrix-vector accumulator and the 1/256 of the un-scaled inverse NTT; see ntt.cuh).  These multiplies
        // sit on the FMA pipe, which this ALU-bound kernel leaves idle.  The microbench entry point (SEED34)
        // returns the plain RejNTTPoly output.
#define EA_TAKE(c)                                                                                        \
        if ((c) < (uint32_t)Q)                                                                            \
            row[have++] = SEED34 ? (int32_t)(c) : shoup_mul((int32_t)(c), SC_MONT_INV256, SC_MONT_INV256S);
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
            EA_TAKE(c0) EA_TAKE(c1) EA_TAKE(c2) EA_TAKE(c3)
        }
#undef EA_TAKE
        const int nq = min(have >> 2, quads_left);
        const int4* src = reinterpret_cast<const int4*>(row);
#pragma unroll
        for (int k = 0; k < (3 + 56) / 4; k++)
            if (k < nq) dst[k] = src[k];
        // carry the incomplete group to the head of the row (a finished stream just parks garbage there)
        const int4 tail = src[have >> 2];
        dst += nq;
        quads_left -= nq;
        have &= 3;
        *reinterpret_cast<int4*>(row) = tail;
    }
    }
}

cudaError_t launch_expand_a(cudaStream_t st, int K, int L, const uint8_t* rho, size_t rho_stride, size_t nkeys,
                            int32_t* A)
{
    const size_t ns = nkeys * (size_t)(K * L);
    if (ns == 0) return cudaSuccess;
    const unsigned grid = capped_grid((ns + EA_THREADS - 1) / EA_THREADS, tuning().kcap);
    {
        KLaunch kl(KID_EXPAND_A, st);
        expand_a_kernel<false><<<grid, EA_THREADS, 0, st>>>(rho, rho_stride, ns, K * L, L, A);
    }
    return cudaGetLastError();
}
cudaError_t launch_expand_a_seed34(cudaStream_t st, const uint8_t* seeds, size_t n, int32_t* out)
{
    if (n == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n + EA_THREADS - 1) / EA_THREADS);
    {
        KLaunch kl(KID_EXPAND_A, st);
        expand_a_kernel<true><<<grid, EA_THREADS, 0, st>>>(seeds, 34, n, 1, 1, out);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Standalone NTT / inverse NTT over n polynomials in place (BASELINE config 2).  HBM-bound:
// 1 KiB in + 1 KiB out per polynomial.  Forward loads rows (layout A, 128 B per warp request) and
// stores 2 x 128-bit per lane (layout B); the inverse mirrors it.
// ------------------------------------------------------------------------------------------------
constexpr int NTT_WARPS = 8;

// streaming accesses: every polynomial is touched exactly once, keep it out of L1 and mark it evict-first
__device__ __forceinline__ int32_t ldg_stream(const int32_t* p)
{
    int32_t v;
    asm volatile("ld.global.cs.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int4 ldg_stream4(const int4* p)
{
    int4 v;
    asm volatile("ld.global.cs.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream(int32_t* p, int32_t v)
{
    asm volatile("st.global.cs.s32 [%0], %1;" ::"l"(p), "r"(v));
}
__device__ __forceinline__ void stg_stream4(int4* p, int4 v)
{
    asm volatile("st.global.cs.v4.s32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}
// a lane's 8 coefficients (one 32-byte sector) in one 256-bit streaming access (pack.cuh: Row8)
__device__ __forceinline__ void ldg_stream8(const int32_t* p, int32_t (&v)[8])
{
#ifndef NO_V8
    asm volatile("ld.global.cs.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(p));
#else
    const int4 x = ldg_stream4(reinterpret_cast<const int4*>(p)), y = ldg_stream4(reinterpret_cast<const int4*>(p) + 1);
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
#endif
}
__device__ __forceinline__ void stg_stream8(int32_t* p, const int32_t (&v)[8])
{
#ifndef NO_V8
    asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]),
                 "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
#else
    stg_stream4(reinterpret_cast<int4*>(p), make_int4(v[0], v[1], v[2], v[3]));
    stg_stream4(reinterpret_cast<int4*>(p) + 1, make_int4(v[4], v[5], v[6], v[7]));
#endif
}

// Two polynomials per warp (ntt16.cuh: a half-warp per polynomial, 16 coefficients per lane): a transform between the two
// global-memory-friendly layouts needs ONE shared-memory exchange (16 wavefronts per polynomial; the 8-per-lane kernel
// needed two, and sat at 83 % of the L1 data pipe beside 77 % of the FMA-heavy pipe and 66 % of DRAM).
template <bool INVERSE>
__global__ void __launch_bounds__(NTT_WARPS * 32) ntt_kernel(int32_t* __restrict__ polys, size_t n)
{
    __shared__ __align__(16) int32_t tiles[NTT_WARPS][T16_WORDS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, h = lane >> 4, t = lane & 15;
    int32_t* tile = tiles[warp] + h * T16_POLY;
    const size_t npairs = (n + 1) / 2;
    const size_t stride = (size_t)gridDim.x * NTT_WARPS;
    const size_t first = (size_t)blockIdx.x * NTT_WARPS + warp;
    // the warp's next two pairs are pulled into L2 while it transforms the current one (2 KiB per pair: 16 lines)
    auto prefetch = [&](size_t pair) {
        if (lane < 16 && 2 * pair + (lane >> 3) < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(polys + pair * 512 + 32 * lane));
    };
    if (first + stride < npairs) prefetch(first + stride);
    for (size_t pair = first; pair < npairs; pair += stride) {
        if (pair + 2 * stride < npairs) prefetch(pair + 2 * stride);
        const size_t p = 2 * pair + h;
        const bool live = p < n;  // an odd count leaves the second half-warp of the last pair without a polynomial
        int32_t* g = polys + (live ? p : 2 * pair) * 256;
        int32_t v[16];
        if (!INVERSE) {
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] = ldg_stream(g + 16 * i + t);  // layout X: 64 contiguous bytes per half-warp
            ntt16_fwd_mont_X2Y(v, tile, t);
            if (live) {
                const int32_t lo[8] = {v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]};
                const int32_t hi[8] = {v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]};
                stg_stream8(g + 16 * t, lo);  // layout Y: 64 contiguous bytes per lane
                stg_stream8(g + 16 * t + 8, hi);
            }
        } else {
            int32_t lo[8], hi[8];
            ldg_stream8(g + 16 * t, lo);
            ldg_stream8(g + 16 * t + 8, hi);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                v[i] = lo[i];
                v[8 + i] = hi[i];
            }
            sntt16_inv_scaled_Y2X(v, tile, t);
            if (live) {
#pragma unroll
                for (int i = 0; i < 16; i++) stg_stream(g + 16 * i + t, v[i]);
            }
        }
    }
}

// test hook: RejNTTPoly's parser over caller-supplied squeeze blocks (nblocks x 168 bytes per polynomial)
cudaError_t launch_rej_ntt_stream(cudaStream_t st, const uint8_t* blocks, size_t nblocks, size_t n, int32_t* out)
{
    if (n == 0) return cudaSuccess;
    expand_a_kernel<true, true><<<(unsigned)((n + EA_THREADS - 1) / EA_THREADS), EA_THREADS, 0, st>>>(blocks, nblocks * 168, n, 1, 1, out);
    return cudaGetLastError();
}

// test hook: SampleInBall's walk (sampling.cuh: SibWalk, src/hashing.rs:43-100) over caller-supplied squeeze blocks of
// H(c~) (nblocks x 136 bytes per challenge): the first 8 bytes are the sign bits, the rest the rejection-sampled indices
__global__ void __launch_bounds__(128) sib_stream_kernel(const uint8_t* __restrict__ blocks, size_t nblocks, size_t n, int tau,
                                                         uint8_t* __restrict__ cmask)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    SibWalk walk;
    walk.init(tau);
    uint64_t blk[17];
    for (size_t b = 0; b < nblocks; b++) {
        const uint8_t* bp = blocks + (i * nblocks + b) * SIB_BYTES;
        for (int k = 0; k < 17; k++) {
            uint64_t w = 0;
            for (int j = 0; j < 8; j++) w |= (uint64_t)bp[8 * k + j] << (8 * j);
            blk[k] = w;
        }
        if (walk.consume(blk, tau)) break;
    }
    walk.store(cmask + i * (size_t)CMASK_STRIDE);
}
cudaError_t launch_sib_stream(cudaStream_t st, const uint8_t* blocks, size_t nblocks, size_t n, int tau, uint8_t* cmask)
{
    if (n == 0) return cudaSuccess;
    sib_stream_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(blocks, nblocks, n, tau, cmask);
    return cudaGetLastError();
}

cudaError_t launch_ntt(cudaStream_t st, int32_t* polys, size_t n, bool inverse)
{
    if (n == 0) return cudaSuccess;
    // CTAs per SM of the grid-stride loop (measured, 16 coefficients per lane: 16 / 32 -> 0.376 ms, 64 -> 0.372 ms per 2^20
    // polynomials).
    // Tried and measured slower (0.585 against 0.486 ms at the time): moving the polynomials with the TMA unit (one
    // cp.async.bulk in, one out, two input slots per warp): the kernel is bound by instruction issue, not by its loads.
    static const int ncap = getenv("MLDSA_NCAP") ? atoi(getenv("MLDSA_NCAP")) : 64;
    const unsigned grid = capped_grid(((n + 1) / 2 + NTT_WARPS - 1) / NTT_WARPS, ncap);  // a warp takes two polynomials at a time
    {
        KLaunch kl(inverse ? KID_INTT : KID_NTT, st);
        if (inverse) ntt_kernel<true><<<grid, NTT_WARPS * 32, 0, st>>>(polys, n);
        else ntt_kernel<false><<<grid, NTT_WARPS * 32, 0, st>>>(polys, n);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Thread-per-row SHAKE over fixed-length rows via the flexible sponge.
// ------------------------------------------------------------------------------------------------
template <int RATE>
__global__ void __launch_bounds__(128) shake_rows_kernel(const uint8_t* __restrict__ in, size_t in_stride,
                                                         size_t in_len, size_t n, uint8_t* __restrict__ out,
                                                         size_t out_stride, int out_len)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Sponge<RATE> sp;
    sp.init();
    sp.absorb(in + i * in_stride, in_len);
    sp.finish();
    uint8_t* o = out + i * out_stride;
    int pos = 0;
    for (int k = 0; k < out_len; k++) {
        if (pos == RATE) {
            sp.next_block();
            pos = 0;
        }
        o[k] = (uint8_t)(sp.s[pos >> 3] >> (8 * (pos & 7)));
        pos++;
    }
}

cudaError_t launch_shake256_rows(cudaStream_t st, const uint8_t* in, size_t in_stride, size_t in_len, size_t n,
                                 uint8_t* out, size_t out_stride, int out_len)
{
    if (n == 0) return cudaSuccess;
    {
        KLaunch kl(KID_SHAKE_ROWS, st);
        shake_rows_kernel<SHAKE256_RATE><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(in, in_stride, in_len, n, out,
                                                                                    out_stride, out_len);
    }
    return cudaGetLastError();
}
cudaError_t launch_shake128_rows(cudaStream_t st, const uint8_t* in, size_t in_stride, size_t in_len, size_t n,
                                 uint8_t* out, size_t out_stride, int out_len)
{
    if (n == 0) return cudaSuccess;
    {
        KLaunch kl(KID_SHAKE_ROWS, st);
        shake_rows_kernel<SHAKE128_RATE><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(in, in_stride, in_len, n, out,
                                                                                    out_stride, out_len);
    }
    return cudaGetLastError();
}

// Commitment hash c~ = H(mu || w1Encode(w1), lambda/4): the per-row body lives in ctilde.cuh (the sign
// round kernels of k_sign.cu run it next to an arithmetic stage).
template <class P, bool SIGN>
__global__ void __launch_bounds__(128, 4)
ctilde_kernel(size_t n, const uint32_t* __restrict__ n_dev, const uint32_t* __restrict__ act, int depth,
              const uint8_t* __restrict__ mw1, const uint8_t* __restrict__ sig, const uint8_t* __restrict__ pre_ok,
              uint8_t* __restrict__ ok, uint8_t* __restrict__ ctilde_out, uint8_t* __restrict__ cmask_out, size_t row_cap)
{
    if (n_dev) n = (size_t)*n_dev * depth;  // sign rounds: live (active lane, candidate) count; n only bounds the launch
    if (n > row_cap) n = row_cap - row_cap % depth;  // tail rounds: whole lanes inside the shared tail region
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t row = act ? (size_t)act[idx / depth] * depth + idx % depth : idx;  // row = lane * depth + candidate
        ctilde_row<P, SIGN>(row, mw1, sig, pre_ok, ok, ctilde_out, cmask_out);
    }
}

template <class P>
static cudaError_t launch_ctilde_t(cudaStream_t st, size_t n, const uint32_t* n_dev, const uint32_t* act, int depth,
                                   const uint8_t* mw1, const uint8_t* sig, const uint8_t* pre_ok, uint8_t* ok,
                                   uint8_t* ctilde_out, uint8_t* sib_out, size_t row_cap)
{
    const unsigned grid = capped_grid((n + 127) / 128, tuning().kcap);
    {
        KLaunch kl(KID_CTILDE, st);
        if (ctilde_out)
            ctilde_kernel<P, true><<<grid, 128, 0, st>>>(n, n_dev, act, depth, mw1, sig, pre_ok, ok, ctilde_out, sib_out, row_cap);
        else
            ctilde_kernel<P, false><<<grid, 128, 0, st>>>(n, n_dev, act, depth, mw1, sig, pre_ok, ok, ctilde_out, sib_out, row_cap);
    }
    return cudaGetLastError();
}

cudaError_t launch_ctilde(int ps, cudaStream_t st, size_t n, const uint32_t* n_dev, const uint32_t* act, int depth,
                          const uint8_t* mw1, const uint8_t* sig, const uint8_t* pre_ok, uint8_t* ok,
                          uint8_t* ctilde_out, uint8_t* sib_out, size_t row_cap)
{
    if (n == 0) return cudaSuccess;
    switch (ps) {
    case 44: return launch_ctilde_t<Params<44>>(st, n, n_dev, act, depth, mw1, sig, pre_ok, ok, ctilde_out, sib_out, row_cap);
    case 65: return launch_ctilde_t<Params<65>>(st, n, n_dev, act, depth, mw1, sig, pre_ok, ok, ctilde_out, sib_out, row_cap);
    case 87: return launch_ctilde_t<Params<87>>(st, n, n_dev, act, depth, mw1, sig, pre_ok, ok, ctilde_out, sib_out, row_cap);
    }
    return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------------------------------
// HashML-DSA pre-hash (src/hashing.rs:317-354): out[i] = OID(11 bytes) || PH(M_i), rows packed back to back; ph 0 = SHA-256 (32 B), 1 = SHA-512 (64 B), 2 = SHAKE128 (32 B).  One thread per message.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
prehash_kernel(int ph, size_t n, const uint8_t* __restrict__ msgs, const uint64_t* __restrict__ msg_off,
               uint8_t* __restrict__ out, uint64_t* __restrict__ out_off)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const uint32_t rowlen = 11 + (ph == 1 ? 64 : 32);  // rows are packed back to back (<= PREHASH_STRIDE each)
    out_off[i] = (uint64_t)i * rowlen;                 // n + 1 offsets
    if (i == n) return;
    uint8_t* o = out + i * rowlen;
    const uint8_t oid[11] = {0x06, 0x09, 0x60, 0x86, 0x48, 0x01, 0x65, 0x03, 0x04, 0x02,
                             (uint8_t)(ph == 0 ? 0x01 : (ph == 1 ? 0x03 : 0x0B))};  // hashing.rs:321, 332, 343
    for (int k = 0; k < 11; k++) o[k] = oid[k];
    const uint8_t* m = msgs + msg_off[i];
    const size_t len = (size_t)(msg_off[i + 1] - msg_off[i]);
    uint8_t dg[64];
    if (ph == 0) {
        sha256_msg(m, len, dg);
    } else if (ph == 1) {
        sha512_msg(m, len, dg);
    } else {
        Sponge<SHAKE128_RATE> sp;
        sp.init();
        sp.absorb(m, len);
        sp.finish();
        for (int k = 0; k < 32; k++) dg[k] = (uint8_t)(sp.s[k >> 3] >> (8 * (k & 7)));
    }
    const int dl = ph == 1 ? 64 : 32;
    for (int k = 0; k < dl; k++) o[11 + k] = dg[k];
}

cudaError_t launch_prehash(cudaStream_t st, int ph, size_t n, const uint8_t* msgs, const uint64_t* msg_off,
                           uint8_t* out, uint64_t* out_off)
{
    KLaunch kl(KID_PREHASH, st);
    prehash_kernel<<<(unsigned)((n + 1 + 127) / 128), 128, 0, st>>>(ph, n, msgs, msg_off, out, out_off);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Integer issue-rate probe: the roofline denominator for this path ("measure it").
// mode 0: LOP3 + SHF streams (ALU pipe, what Keccak runs on); mode 1: IMAD + IMAD.HI streams (FMA
// pipe, the Montgomery multiplies); mode 2: both interleaved; modes 3..8: IMAD, IMAD.HI, IMAD.WIDE(+LOP3),
// SHF, LOP3, IADD3 alone.  16 independent chains per thread.
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) int_probe_kernel(uint32_t* out, int iters, uint32_t seed, uint32_t m)
{
    uint32_t x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u + blockIdx.x;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const uint32_t y = x[(i + 5) & 15];
                if (MODE == 0 || MODE == 2 || MODE == 7) x[i] = x[i] ^ (x[(i + 1) & 15] & m);   // LOP3
                if (MODE == 0 || MODE == 2 || MODE == 6) x[i] = __funnelshift_l(x[i], y, 7);    // SHF
                if (MODE == 1 || MODE == 2 || MODE == 3) x[i] = x[i] * m + x[(i + 3) & 15];     // IMAD
                if (MODE == 1 || MODE == 2 || MODE == 4)                                        // IMAD.HI
                    x[i] = (uint32_t)__mulhi((int32_t)x[i], (int32_t)m) + x[(i + 7) & 15];
                if (MODE == 5) {                                                                // IMAD.WIDE
                    const uint64_t w = (uint64_t)x[i] * m + (((uint64_t)y << 32) | x[(i + 9) & 15]);
                    x[i] = (uint32_t)w ^ (uint32_t)(w >> 32);                                   // (+1 LOP3)
                }
                if (MODE == 8) x[i] = x[i] + y + m;                                             // IADD3
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) acc ^= x[i];
    if (acc == 0x12345678u) out[0] = acc;  // keep the chains alive
}

static const int PROBE_OPS_PER_ELEM[9] = {2, 2, 4, 1, 1, 2, 1, 1, 1};

// returns thread-level integer instructions issued (ops) for the launched grid
cudaError_t launch_int_probe(cudaStream_t st, int mode, int blocks, int iters, uint32_t* scratch, double* ops)
{
    if (mode < 0 || mode > 8) return cudaErrorInvalidValue;
    const double per_thread = (double)iters * 4 * 16 * PROBE_OPS_PER_ELEM[mode];
    *ops = per_thread * 256.0 * blocks;
    KLaunch kl(KID_PROBE, st);
    switch (mode) {
    case 0: int_probe_kernel<0><<<blocks, 256, 0, st>>>(scratch, iters, 12345u, 0x9E3779B1u); break;
    case 1: int_probe_kernel<1><<<blocks, 256, 0, st>>>(scratch, iters, 12345u, 0x9E3779B1u); break;
    case 2: int_probe_kernel<2><<<blocks, 256, 0, st>>>(scratch, iters, 12345u, 0x9E3779B1u); break;
    case 3: int_probe_kernel<3><<<blocks, 256, 0, st>>>(scratch, iters, 12345u, 0x9E3779B1u); break;
    case 4: int_probe_kernel<4><<<blocks, 256, 0, st>>>(scratch, iters, 12345u, 0x9E3779B1u); break;
    case 5: int_probe_kernel<5><<<blocks, 256, 0, st>>>(scratch, iters, 12345u, 0x9E3779B1u); break;
    case 6: int_probe_kernel<6><<<blocks, 256, 0, st>>>(scratch, iters, 12345u, 0x9E3779B1u); break;
    case 7: int_probe_kernel<7><<<blocks, 256, 0, st>>>(scratch, iters, 12345u, 0x9E3779B1u); break;
    case 8: int_probe_kernel<8><<<blocks, 256, 0, st>>>(scratch, iters, 12345u, 0x9E3779B1u); break;
    }
    return cudaGetLastError();
}

}  // namespace mldsa
