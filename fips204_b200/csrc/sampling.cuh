// sampling.cuh -- message-representative framing and SampleInBall for the fused kernels.
#pragma once
#include "engine.h"
#include "keccak.cuh"
#include "ntt.cuh"

namespace mldsa {

// mu = H(tr || M', 64) with the framings of src/ml_dsa.rs:185-196 / :386-397.  Leaves the sponge
// finished (first squeeze block ready).  Returns false when |ctx| > 255 (src/lib.rs:274, :368-370).
__device__ __forceinline__ bool absorb_mu(Sponge<SHAKE256_RATE>& sp, const uint8_t* tr, const MsgDesc& md, size_t i)
{
    sp.init();
    sp.absorb(tr, 64);
    const uint64_t m0 = md.msg_off[i], m1 = md.msg_off[i + 1];
    uint32_t clen = 0;
    const uint8_t* c = nullptr;
    if (md.ctx_off) {
        c = md.ctx + md.ctx_off[i];
        clen = md.ctx_off[i + 1] - md.ctx_off[i];
    }
    const bool ok = clen <= 255;
    if (md.framing != 1 && ok) {
        sp.absorb_byte(md.framing == 2 ? 1u : 0u);
        sp.absorb_byte(clen);
        sp.absorb(c, clen);
    }
    sp.absorb(md.msgs + m0, (size_t)(m1 - m0));
    sp.finish();
    return ok;
}

// Slow path of SampleInBall: block `blk` (>= 1) of SHAKE256(c~).  One SHAKE256 block covers the 8 sign
// bytes plus 128 index draws; tau <= 60 draws with acceptance >= 0.77 practically never needs more,
// but the reference's reader is unbounded (hashing.rs:68-77), so this exists for exactness.
static __device__ __noinline__ void sib_refill(const uint8_t* ctilde, int ctlen, int blk, uint32_t* sibs)
{
    Sponge<SHAKE256_RATE> sp;
    sp.init();
    for (int k = 0; k < ctlen; k++) sp.absorb_byte(ctilde[k]);
    sp.finish();
    for (int b = 0; b < blk; b++) sp.next_block();
    for (int k = 0; k < SIB_BYTES / 8; k++) {
        sibs[2 * k] = (uint32_t)sp.s[k];
        sibs[2 * k + 1] = (uint32_t)(sp.s[k] >> 32);
    }
}

// SampleInBall (src/hashing.rs:43-100).  The Fisher-Yates walk is inherently sequential (tau dependent
// swaps on public data), so lane 0 performs it on the warp's tile; the other lanes only help to stage
// the hash block and clear the tile.  Result: c in register layout A, ready for ntt_fwd_A2B.
template <class P>
__device__ __forceinline__ void sample_in_ball_warp(const uint8_t* sib_block, const uint8_t* ctilde, uint32_t* sibs,
                                                    int32_t* tile, int lane, int32_t (&c)[8])
{
    __syncwarp();
    for (int w = lane; w < SIB_BYTES / 4; w += 32) sibs[w] = reinterpret_cast<const uint32_t*>(sib_block)[w];
#pragma unroll
    for (int i = 0; i < 8; i++) tile[NTT_ROW * i + lane] = 0;
    __syncwarp();
    if (lane == 0) {
        const uint8_t* sb = reinterpret_cast<const uint8_t*>(sibs);
        const uint64_t signs = (uint64_t)sibs[0] | ((uint64_t)sibs[1] << 32);  // hashing.rs:53-55
        int pos = 8, blk = 0;
        for (int i = 256 - P::TAU; i < 256; i++) {
            int j;
            do {
                if (pos == SIB_BYTES) {
                    sib_refill(ctilde, P::CTILDE, ++blk, sibs);
                    pos = 0;
                }
                j = sb[pos++];
            } while (j > i);                                                     // hashing.rs:68-77
            const int ai = (i >> 5) * NTT_ROW + (i & 31), aj = (j >> 5) * NTT_ROW + (j & 31);
            tile[ai] = tile[aj];                                                 // hashing.rs:80
            tile[aj] = 1 - 2 * (int32_t)((signs >> (i + P::TAU - 256)) & 1);     // hashing.rs:83-86
        }
    }
    __syncwarp();
    tile_load_A(tile, lane, c);
}

}  // namespace mldsa
