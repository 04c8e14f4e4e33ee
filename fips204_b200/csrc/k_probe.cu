// k_probe.cu -- Keccak-f[1600] throughput probe: one thread per state, REPS permutations, for several
// splits of the 29 rotations per round between the ALU pipe (SHF) and the FMA pipe (IMAD.WIDE/IMAD/IMAD.HI).
// Used by tools/keccak_probe.py to choose KECCAK_FMA_SITES with a measurement instead of a guess.
#include "engine.h"
#include "keccak.cuh"

namespace mldsa {

template <uint32_t SITES>
__global__ void __launch_bounds__(128, 4) keccak_probe_kernel(uint64_t* st, int reps)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t s[25];
#pragma unroll
    for (int k = 0; k < 25; k++) s[k] = st[(size_t)k * gridDim.x * blockDim.x + i];
#pragma unroll 1
    for (int r = 0; r < reps; r++) keccak_f1600<SITES>(s);
#pragma unroll
    for (int k = 0; k < 25; k++) st[(size_t)k * gridDim.x * blockDim.x + i] = s[k];
}

static const uint32_t PROBE_MASKS[] = {0x0u, 0x11084210u, 0x14924924u, 0x15555554u, 0x1ad6ad6au, 0x1db6edb6u, 0x1eeeeeeeu, 0x1f7df7deu, 0x1ff7fdfeu, 0x1fffffffu};
int keccak_probe_variants() { return (int)(sizeof(PROBE_MASKS) / sizeof(PROBE_MASKS[0])); }
uint32_t keccak_probe_mask(int v) { return PROBE_MASKS[v]; }

cudaError_t launch_keccak_probe(cudaStream_t st, int variant, int blocks, int reps, uint64_t* state)
{
    KLaunch kl(KID_PROBE, st);
    switch (variant) {
    case 0: keccak_probe_kernel<0x0u><<<blocks, 128, 0, st>>>(state, reps); break;
    case 1: keccak_probe_kernel<0x11084210u><<<blocks, 128, 0, st>>>(state, reps); break;
    case 2: keccak_probe_kernel<0x14924924u><<<blocks, 128, 0, st>>>(state, reps); break;
    case 3: keccak_probe_kernel<0x15555554u><<<blocks, 128, 0, st>>>(state, reps); break;
    case 4: keccak_probe_kernel<0x1ad6ad6au><<<blocks, 128, 0, st>>>(state, reps); break;
    case 5: keccak_probe_kernel<0x1db6edb6u><<<blocks, 128, 0, st>>>(state, reps); break;
    case 6: keccak_probe_kernel<0x1eeeeeeeu><<<blocks, 128, 0, st>>>(state, reps); break;
    case 7: keccak_probe_kernel<0x1f7df7deu><<<blocks, 128, 0, st>>>(state, reps); break;
    case 8: keccak_probe_kernel<0x1ff7fdfeu><<<blocks, 128, 0, st>>>(state, reps); break;
    case 9: keccak_probe_kernel<0x1fffffffu><<<blocks, 128, 0, st>>>(state, reps); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace mldsa
