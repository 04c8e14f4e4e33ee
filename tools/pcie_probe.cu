// pcie_probe.cu -- host<->device copy bandwidth of a multi-GPU box, measured the way the engine uses it: pinned host
// buffers, cudaMemcpyAsync on dedicated copy streams, all GPUs at once.
//
//   nvcc -O2 -o tools/pcie_probe tools/pcie_probe.cu
//   tools/pcie_probe [--gpus N] [--mode proc|thread] [--dir h2d|d2h|bidir] [--mb M] [--reps R] [--chunk-mb C]
//
// mode proc: one PROCESS per GPU (what `torchrun bench.py --gpus N` does); mode thread: one host THREAD per GPU inside
// one process (what mldsa_*_batch_multi does).  All workers start together (shared-memory barrier) and the aggregate rate
// is total bytes / (latest end - earliest start).  One JSON line per run on stdout.
// The question it answers: is the end-to-end 1 -> 8 GPU curve of the packed-key sign
// (8 x (4.30 GB in + 3.47 GB out) per step) bounded by the host fabric of the box, and at what GB/s?
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

struct Shared {
    std::atomic<int> ready, go, failed;
    double t0[64], t1[64], gbs[64];
};

static double now()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

#define CK(x)                                                                                    \
    do {                                                                                         \
        cudaError_t e_ = (x);                                                                    \
        if (e_ != cudaSuccess) {                                                                 \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                             \
            sh->failed.fetch_add(1);                                                             \
            sh->ready.fetch_add(1);                                                              \
            return;                                                                              \
        }                                                                                        \
    } while (0)

// dir: 0 h2d, 1 d2h, 2 both at once (two streams)
static void worker(Shared* sh, int g, int n, int dir, size_t bytes, size_t chunk, int reps)
{
    CK(cudaSetDevice(g));
    uint8_t *h_in = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr;
    cudaStream_t s_in, s_out;
    CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    if (dir != 1) {
        CK(cudaHostAlloc(reinterpret_cast<void**>(&h_in), bytes, cudaHostAllocPortable));
        CK(cudaMalloc(reinterpret_cast<void**>(&d_in), bytes));
        memset(h_in, 0x5a, bytes);  // touch: the pages exist before the clock starts
    }
    if (dir != 0) {
        CK(cudaHostAlloc(reinterpret_cast<void**>(&h_out), bytes, cudaHostAllocPortable));
        CK(cudaMalloc(reinterpret_cast<void**>(&d_out), bytes));
        memset(h_out, 0, bytes);
        CK(cudaMemset(d_out, 0xa5, bytes));
    }
    // one untimed pass
    for (size_t o = 0; o < bytes; o += chunk) {
        const size_t m = bytes - o < chunk ? bytes - o : chunk;
        if (dir != 1) CK(cudaMemcpyAsync(d_in + o, h_in + o, m, cudaMemcpyHostToDevice, s_in));
        if (dir != 0) CK(cudaMemcpyAsync(h_out + o, d_out + o, m, cudaMemcpyDeviceToHost, s_out));
    }
    CK(cudaDeviceSynchronize());
    sh->ready.fetch_add(1);
    while (sh->go.load() == 0) {
    }
    const double t0 = now();
    for (int r = 0; r < reps; r++)
        for (size_t o = 0; o < bytes; o += chunk) {
            const size_t m = bytes - o < chunk ? bytes - o : chunk;
            if (dir != 1) cudaMemcpyAsync(d_in + o, h_in + o, m, cudaMemcpyHostToDevice, s_in);
            if (dir != 0) cudaMemcpyAsync(h_out + o, d_out + o, m, cudaMemcpyDeviceToHost, s_out);
        }
    cudaError_t e = cudaStreamSynchronize(s_in);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s_out);
    const double t1 = now();
    if (e != cudaSuccess) sh->failed.fetch_add(1);
    sh->t0[g] = t0;
    sh->t1[g] = t1;
    sh->gbs[g] = (dir == 2 ? 2.0 : 1.0) * bytes * reps / (t1 - t0) / 1e9;
    (void)n;
}

int main(int argc, char** argv)
{
    int gpus = 1, reps = 8;
    size_t mb = 2048, chunk_mb = 256;
    std::string mode = "proc", dir = "h2d";
    for (int i = 1; i + 1 < argc; i += 2) {
        const std::string k = argv[i];
        if (k == "--gpus") gpus = atoi(argv[i + 1]);
        else if (k == "--mode") mode = argv[i + 1];
        else if (k == "--dir") dir = argv[i + 1];
        else if (k == "--mb") mb = strtoull(argv[i + 1], nullptr, 10);
        else if (k == "--reps") reps = atoi(argv[i + 1]);
        else if (k == "--chunk-mb") chunk_mb = strtoull(argv[i + 1], nullptr, 10);
    }
    if (gpus < 1 || gpus > 64) return 2;
    const int d = dir == "h2d" ? 0 : (dir == "d2h" ? 1 : 2);
    const size_t bytes = mb << 20, chunk = chunk_mb << 20;
    Shared* sh = static_cast<Shared*>(mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0));
    if (sh == MAP_FAILED) return 3;
    new (sh) Shared();
    sh->ready = 0;
    sh->go = 0;
    sh->failed = 0;
    std::vector<pid_t> kids;
    std::vector<std::thread> th;
    if (mode == "proc") {
        for (int g = 0; g < gpus; g++) {  // fork BEFORE any CUDA call: every child makes its own context
            const pid_t p = fork();
            if (p == 0) {
                worker(sh, g, gpus, d, bytes, chunk, reps);
                _exit(0);
            }
            kids.push_back(p);
        }
    } else {
        for (int g = 0; g < gpus; g++) th.emplace_back(worker, sh, g, gpus, d, bytes, chunk, reps);
    }
    while (sh->ready.load() < gpus) usleep(1000);
    sh->go = 1;
    for (pid_t p : kids) waitpid(p, nullptr, 0);
    for (auto& t : th) t.join();
    double first = 1e300, last = 0, slow = 1e300, fast = 0;
    for (int g = 0; g < gpus; g++) {
        first = sh->t0[g] < first ? sh->t0[g] : first;
        last = sh->t1[g] > last ? sh->t1[g] : last;
        slow = sh->gbs[g] < slow ? sh->gbs[g] : slow;
        fast = sh->gbs[g] > fast ? sh->gbs[g] : fast;
    }
    const double total = (d == 2 ? 2.0 : 1.0) * bytes * reps * gpus;
    printf("{\"gpus\": %d, \"mode\": \"%s\", \"dir\": \"%s\", \"mb_per_gpu_per_dir\": %zu, \"chunk_mb\": %zu, \"reps\": %d, "
           "\"aggregate_gbs\": %.1f, \"per_gpu_gbs_min\": %.1f, \"per_gpu_gbs_max\": %.1f, \"failed\": %d}\n",
           gpus, mode.c_str(), dir.c_str(), mb, chunk_mb, reps, sh->failed.load() ? 0.0 : total / (last - first) / 1e9, slow, fast,
           sh->failed.load());
    return sh->failed.load() ? 1 : 0;
}
