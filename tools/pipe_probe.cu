// pipe_probe.cu -- do the ALU pipe (LOP3/SHF/IADD3) and the FMA pipe (IMAD/IMAD.HI/IMAD.WIDE) of an sm_100a SMSP
// issue concurrently?  Independent register chains, instruction choice pinned with inline PTX.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu && ./pipe_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int KIND> __device__ __forceinline__ void fop(uint32_t& x, uint32_t& y, uint32_t m)
{
    if (KIND == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(m), "r"(y));
    if (KIND == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(m), "r"(y));
    if (KIND == 2) {
        uint64_t w = ((uint64_t)y << 32) | x;   // 64-bit accumulator chain: both halves stay live
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w) : "r"(x), "r"(m));
        x = (uint32_t)w;
        y = (uint32_t)(w >> 32);
    }
    if (KIND == 5) asm volatile("mad.hi.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(m), "r"(y));
    if (KIND == 6) asm volatile("mul.hi.s32 %0, %0, %1;" : "+r"(x) : "r"(m));
    if (KIND == 7) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x) : "r"(m));
    if (KIND == 8) {
        uint64_t w = ((uint64_t)y << 32) | x;
        asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w) : "r"(x), "r"(m));
        x = (uint32_t)w;
        y = (uint32_t)(w >> 32);
    }
    if (KIND == 9) {   // no addend; both halves live (hi becomes the next multiplier)
        uint64_t w;
        asm volatile("mul.wide.s32 %0, %1, %2;" : "=l"(w) : "r"(x), "r"(y));
        x = (uint32_t)w | 1u;
        y = (uint32_t)(w >> 32);
    }
    if (KIND == 10) asm volatile("mad.lo.s32 %0, %0, 8380417, %1;" : "+r"(x) : "r"(y));
    if (KIND == 11) asm volatile("mul.hi.s32 %0, %0, 8380417;" : "+r"(x));
    if (KIND == 3) asm volatile("shl.b32 %0, %0, 3;" : "+r"(x));   // may become IMAD.SHL or SHF
    if (KIND == 4) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y));
}
template <int AK> __device__ __forceinline__ void aop(uint32_t& x, uint32_t y, uint32_t z)
{
    if (AK == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(y), "r"(z));
    if (AK == 1) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(x) : "r"(y));
    if (AK == 2) asm volatile("lop3.b32 %0, %0, %1, 0x0f0f0f0f, 0x96;" : "+r"(x) : "r"(y));
}

template <int NA, int NF, int KIND, int AK>
__global__ void __launch_bounds__(256) probe(uint32_t* out, int iters, uint32_t m)
{
    uint32_t a[8], f[8], g[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        a[i] = threadIdx.x * 77 + i + m;
        f[i] = threadIdx.x * 31 + i * 5 + m;
        g[i] = threadIdx.x * 13 + i * 7 + m;
    }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (i < NA) aop<AK>(a[i], a[(i + 3) & 7], a[(i + 5) & 7]);
                if (i < NF) fop<KIND>(f[i], g[i], m);
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= a[i] ^ f[i] ^ g[i];
    if (acc == 0x12345678u) out[0] = acc;
}

template <int NA, int NF, int KIND, int AK> double run(uint32_t* d, int sms, double mhz)
{
    const int blocks = sms * 8, iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        probe<NA, NF, KIND, AK><<<blocks, 256>>>(d, iters, 0x9E3779B1u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    // warps per SMSP = 8 blocks * 8 warps / 4 = 16; loop bodies per SMSP = 16 * iters * 4
    const double cycles = best * 1e-3 * mhz * 1e6;
    return cycles / (16.0 * iters * 4);   // SMSP cycles per (NA alu + NF fma) group of warp-instructions
}

int main()
{
    int sms = 148, khz = 1965000;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    uint32_t* d;
    cudaMalloc(&d, 256);
    const char* kn[12] = {"IMAD", "IMAD.HI.U", "WIDE.U+acc", "SHL", "IADD", "HI.S+add", "HI.S", "HI.U", "WIDE.S+acc", "WIDE.S", "IMAD.imm", "HI.S.imm"};
    printf("sms %d clock %.0f MHz; cycles per SMSP for one group of (NA ALU + NF FMA-kind) warp-instr\n", sms, mhz);
#define ROW(K, AK)                                                                                          \
    printf("%-9s vs %s: A8F0 %.2f  A0F8 %.2f  A8F8 %.2f  A8F4 %.2f  A4F8 %.2f\n", kn[K], AK == 1 ? "SHF " : (AK == 2 ? "LOP3i" : "LOP3"), \
           run<8, 0, K, AK>(d, sms, mhz), run<0, 8, K, AK>(d, sms, mhz), run<8, 8, K, AK>(d, sms, mhz),       \
           run<8, 4, K, AK>(d, sms, mhz), run<4, 8, K, AK>(d, sms, mhz));
    ROW(0, 1) ROW(1, 1) ROW(2, 1) ROW(5, 1) ROW(6, 1) ROW(7, 1) ROW(8, 1) ROW(9, 1) ROW(10, 1) ROW(11, 1) ROW(0, 0) ROW(6, 0)
    return 0;
}
