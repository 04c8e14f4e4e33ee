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

// Test hook: count the non-zero bytes of a device range (the ranges that held secrets must be all zero after a call).
__global__ void count_nonzero_kernel(const uint8_t* __restrict__ p, size_t bytes, unsigned long long* __restrict__ count)
{
    unsigned long long local = 0;
    const size_t head = (16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15;  // bytes before the first aligned int4
    const size_t nvec = bytes > head ? (bytes - head) / 16 : 0;
    const int4* v = reinterpret_cast<const int4*>(p + head);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
        const int4 x = v[i];
        if (x.x | x.y | x.z | x.w) {
            const uint32_t w[4] = {(uint32_t)x.x, (uint32_t)x.y, (uint32_t)x.z, (uint32_t)x.w};
            for (int k = 0; k < 4; k++)
                for (int b = 0; b < 4; b++) local += ((w[k] >> (8 * b)) & 0xFF) != 0;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (size_t i = 0; i < (head < bytes ? head : bytes); i++) local += p[i] != 0;
        for (size_t i = head + nvec * 16; i < bytes; i++) local += p[i] != 0;
    }
    for (int d = 16; d; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
}
cudaError_t launch_count_nonzero(cudaStream_t st, const uint8_t* p, size_t bytes, unsigned long long* count)
{
    if (bytes == 0) return cudaSuccess;
    count_nonzero_kernel<<<148 * 8, 256, 0, st>>>(p, bytes, count);
    return cudaGetLastError();
}

// ctx longer than 255 bytes anywhere in the batch (src/lib.rs:274): checked before a device-pointer sign call starts
__global__ void __launch_bounds__(256) ctx_check_kernel(size_t n, const uint32_t* __restrict__ ctx_off, volatile uint32_t* flag)
{
    bool bad = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        bad |= ctx_off[i + 1] - ctx_off[i] > 255u;
    if (bad) {
        *flag = 1;
        __threadfence_system();
    }
}
cudaError_t launch_ctx_check(cudaStream_t st, size_t n, const uint32_t* ctx_off, uint32_t* flag)
{
    if (n == 0) return cudaSuccess;
    ctx_check_kernel<<<(unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, st>>>(n, ctx_off, flag);
    return cudaGetLastError();
}

}  // namespace mldsa
