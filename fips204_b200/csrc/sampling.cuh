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

// SampleInBall (src/hashing.rs:43-100) as a per-THREAD walk over the SHAKE256(c~) stream.  It runs
// inside the thread-per-lane hash kernels (32 lanes of a warp walk 32 different challenges at once);
// the result is c as two 256-bit masks (nonzero, negative) = CMASK_BYTES, which the warp-per-lane
// arithmetic kernels expand straight into registers.  The Fisher-Yates walk is on public data.
// The reader is unbounded like the reference's (hashing.rs:68-77): consume() is fed one 136-byte
// block at a time and reports whether it needs another one (practically never: tau <= 60 draws with
// acceptance >= 0.77 fit 128 bytes with overwhelming probability).
constexpr int CMASK_BYTES = 64;

struct SibWalk {
    uint32_t nz[8], sg[8];
    uint64_t signs;
    int i;
    bool have_signs;
    __device__ __forceinline__ void init(int tau)
    {
#pragma unroll
        for (int k = 0; k < 8; k++) nz[k] = sg[k] = 0;
        i = 256 - tau;
        have_signs = false;
        signs = 0;
    }
    // blk: 17 little-endian 64-bit words of one squeeze block (in local/shared memory)
    __device__ __forceinline__ bool consume(const uint64_t* blk, int tau)
    {
        int pos = 0;
        if (!have_signs) {
            signs = blk[0];  // hashing.rs:53-55
            have_signs = true;
            pos = 8;
        }
        const uint8_t* b = reinterpret_cast<const uint8_t*>(blk);
        while (i < 256) {
            if (pos == SIB_BYTES) return false;
            const int j = b[pos++];
            if (j > i) continue;                                            // hashing.rs:68-77
            const uint32_t nzj = (nz[j >> 5] >> (j & 31)) & 1u, sgj = (sg[j >> 5] >> (j & 31)) & 1u;
            nz[i >> 5] |= nzj << (i & 31);                                  // c_i <- c_j (bit i was still 0)
            sg[i >> 5] |= sgj << (i & 31);
            const uint32_t neg = (uint32_t)(signs >> (i + tau - 256)) & 1u; // hashing.rs:83-86
            nz[j >> 5] |= 1u << (j & 31);
            sg[j >> 5] = (sg[j >> 5] & ~(1u << (j & 31))) | (neg << (j & 31));
            i++;
        }
        return true;
    }
    __device__ __forceinline__ void store(uint8_t* cmask) const
    {
        uint32_t* o = reinterpret_cast<uint32_t*>(cmask);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            o[k] = nz[k];
            o[8 + k] = sg[k];
        }
    }
};

// c in register layout A (idx = 32*i + lane) from the challenge masks
__device__ __forceinline__ void load_c_layoutA(const uint8_t* cmask, int lane, int32_t (&c)[8])
{
    const uint4* p = reinterpret_cast<const uint4*>(cmask);
    const uint4 n0 = __ldg(p), n1 = __ldg(p + 1), s0 = __ldg(p + 2), s1 = __ldg(p + 3);
    const uint32_t nz[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
    const uint32_t sg[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int32_t on = (int32_t)((nz[i] >> lane) & 1u), neg = (int32_t)((sg[i] >> lane) & 1u);
        c[i] = on - 2 * (on & neg);
    }
}

}  // namespace mldsa
