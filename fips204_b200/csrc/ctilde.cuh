// ctilde.cuh -- commitment hash c~ = H(mu || w1Encode(w1), lambda/4) for ONE row (one thread), state in
// registers, one permutation call site.  The input is 64 + W1_LEN bytes, always a whole number of 64-bit
// words, read straight from the row's mu||w1 bytes.
// Verify mode: ok[row] = pre_ok[row] && c~' == c~ of the signature (src/ml_dsa.rs:427-436).  Sign mode:
// ctilde_out[row] = c~ and cmask_out[row] = SampleInBall(c~) as two 256-bit masks (src/ml_dsa.rs:231-237,
// src/hashing.rs:43-100).
#pragma once
#include "engine.h"
#include "keccak.cuh"
#include "pack.cuh"
#include "sampling.cuh"

namespace mldsa {

template <class P, bool SIGN>
__device__ __forceinline__ void ctilde_row(size_t ln, const uint8_t* __restrict__ mw1, const uint8_t* __restrict__ sig,
                                           const uint8_t* __restrict__ pre_ok, uint8_t* __restrict__ ok,
                                           uint8_t* __restrict__ ctilde_out, uint8_t* __restrict__ cmask_out)
{
    constexpr int NW = (MU_BYTES + P::W1_LEN) / 8;  // 104 or 136 words
    constexpr int NBLK = NW / 17 + 1;               // 7 or 9 absorb blocks (the last holds the padding)
    constexpr int CW = P::CTILDE / 8;               // c~ words
    const uint64_t* in = reinterpret_cast<const uint64_t*>(mw1 + ln * (size_t)(MU_BYTES + P::W1_LEN));
    uint64_t a[25];
#pragma unroll
    for (int i = 0; i < 25; i++) a[i] = 0;
    SibWalk walk;
    uint64_t blk[17];  // squeeze block of H(c~) in local memory: the walk indexes it by byte
    if (SIGN) walk.init(P::TAU);
    // One permutation call site for: the NBLK absorb blocks of H(mu || w1), and (sign) every squeeze
    // block of the second sponge H(c~) that SampleInBall consumes (one, except with negligible probability).
#pragma unroll 1
    for (int b = 0;; b++) {
        if (b < NBLK) {
#pragma unroll
            for (int k = 0; k < 17; k++)
                if (b * 17 + k < NW) a[k] ^= in[b * 17 + k];
            if (b == NBLK - 1) {
                a[NW % 17] ^= 0x1F;
                a[16] ^= 0x8000000000000000ULL;
            }
        } else if (b == NBLK) {  // sign: c~ is out; start the second sponge H(c~)
            uint64_t* co = reinterpret_cast<uint64_t*>(ctilde_out + ln * 64);
#pragma unroll
            for (int k = 0; k < CW; k++) co[k] = a[k];
#pragma unroll
            for (int k = CW; k < 25; k++) a[k] = 0;
            a[CW] = 0x1F;
            a[16] ^= 0x8000000000000000ULL;
        }
        keccak_f1600(a);
        if (!SIGN) {
            if (b == NBLK - 1) break;
        } else if (b >= NBLK) {
#pragma unroll
            for (int k = 0; k < 17; k++) blk[k] = a[k];
            if (walk.consume(blk, P::TAU)) break;
        }
    }
    if (SIGN) {
        walk.store(cmask_out + ln * (size_t)CMASK_STRIDE);
    } else {
        uint32_t w[P::CTILDE / 4];
        load_span<P::CTILDE>(sig + ln * (size_t)P::SIG_LEN, w);
        bool eq = true;
#pragma unroll
        for (int k = 0; k < CW; k++) eq &= a[k] == ((uint64_t)w[2 * k] | ((uint64_t)w[2 * k + 1] << 32));
        ok[ln] = (eq && pre_ok[ln]) ? 1 : 0;
    }
}

}  // namespace mldsa
