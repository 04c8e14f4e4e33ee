// engine.cu -- host runtime and the extern "C" boundary (include/mldsa_b200.h).
//
// One lazily created context per CUDA device (mutex-guarded): a work stream, a cached device arena
// that is carved into the per-call scratch (and, for the host-pointer entry points, the staged
// inputs/outputs), and a small pinned mailbox.  Batches are cut into chunks that fit the workspace
// limit; every chunk runs the kernel pipeline of k_keygen / k_sign / k_verify.  Calls are
// synchronous and re-entrant; calls on different devices run concurrently from different threads.
#include "engine.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mldsa_b200.h"
#include "params.h"

namespace mldsa {

static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_profiling{0};
static std::mutex g_prof_mu;
static double g_prof_ms[KID_COUNT];
static uint64_t g_prof_n[KID_COUNT];
static const char* const KERNEL_NAMES[KID_COUNT] = {
    "expand_a", "ntt", "intt", "shake_rows", "ctilde", "kg_seed", "rej_bounded", "kg_arith", "verify_pre",
    "t1hat", "verify_arith", "sign_prep", "sign_key", "mask", "sign_w", "sign_z", "compact", "int_probe", "prehash",
    "sk_to_pk"};

struct EvPair {
    int id;
    cudaEvent_t a, b;
};
static thread_local std::vector<EvPair> t_events;      // launched, not yet collected
static thread_local std::vector<cudaEvent_t> t_evpool;  // recycled events

static cudaEvent_t ev_get()
{
    if (!t_evpool.empty()) {
        cudaEvent_t e = t_evpool.back();
        t_evpool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

KLaunch::KLaunch(int id_, cudaStream_t st_) : id(id_), st(st_)
{
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (g_profiling.load(std::memory_order_relaxed)) {
        EvPair p{id, ev_get(), ev_get()};
        cudaEventRecord(p.a, st);
        t_events.push_back(p);
    }
}
KLaunch::~KLaunch()
{
    if (g_profiling.load(std::memory_order_relaxed) && !t_events.empty() && t_events.back().id == id)
        cudaEventRecord(t_events.back().b, st);
}
// Called at the end of every entry point.  The event pairs recorded on this thread are folded into the totals LAZILY --
// when a reader asks (mldsa_profile_get / _reset, same thread), when the thread is a worker about to exit, or when the
// queue grows large -- so that a caller timing a call with its own events does not also time 200 cudaEventElapsedTime
// calls made after the device has finished.
void prof_collect()
{
    if (t_events.size() < 8192) return;
    prof_collect_now();
}
void prof_collect_now()
{
    if (t_events.empty()) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    std::vector<EvPair> pending;  // launched on a stream that is still running (pipelined chunks): keep for later
    for (auto& p : t_events) {
        if (cudaEventQuery(p.b) == cudaErrorNotReady) {
            pending.push_back(p);
            continue;
        }
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            g_prof_ms[p.id] += ms;
            g_prof_n[p.id] += 1;
        } else {
            cudaGetLastError();
        }
        t_evpool.push_back(p.a);
        t_evpool.push_back(p.b);
    }
    t_events.swap(pending);
}

static std::atomic<int> g_kcap{-1}, g_acap{-1}, g_depth{-1}, g_tail_depth{-1}, g_tail_slots{-1}, g_max_rounds{0};
static int env_int(const char* name, int dflt)
{
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}
Tuning tuning()
{
    int k = g_kcap.load(), a = g_acap.load();
    if (k < 0) g_kcap.store(k = env_int("MLDSA_KCAP", 0));
    if (a < 0) g_acap.store(a = env_int("MLDSA_ACAP", 0));
    int d = g_depth.load();
    if (d < 0) g_depth.store(d = env_int("MLDSA_DEPTH", 2));
    if (d < 1) d = 1;
    if (d > 8) d = 8;
    int td = g_tail_depth.load(), ts = g_tail_slots.load();
    if (td < 0) g_tail_depth.store(td = env_int("MLDSA_TAILDEPTH", 8));
    if (ts < 0) g_tail_slots.store(ts = env_int("MLDSA_TAILSLOTS", 32768));
    if (td > 16) td = 16;
    if (ts < 0) ts = 0;
    return Tuning{k, a, d, td, ts, g_max_rounds.load()};
}
unsigned capped_grid(size_t want_ctas, int cap_per_sm)
{
    static std::atomic<int> sm_count[64];  // per device: a box may mix parts
    int dev = 0;
    cudaGetDevice(&dev);
    int sms = dev >= 0 && dev < 64 ? sm_count[dev].load(std::memory_order_relaxed) : 0;
    if (sms == 0) {
        sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (dev >= 0 && dev < 64) sm_count[dev].store(sms, std::memory_order_relaxed);
    }
    if (want_ctas < 1) want_ctas = 1;
    if (cap_per_sm > 0 && want_ctas > (size_t)sms * cap_per_sm) want_ctas = (size_t)sms * cap_per_sm;
    if (want_ctas > 0x7fffffffu) want_ctas = 0x7fffffffu;
    return (unsigned)want_ctas;
}

static thread_local std::string t_cuda_err;
static thread_local int t_last_rounds = 0;
static std::atomic<size_t> g_ws_limit{0};

static int cuda_fail(cudaError_t e, const char* what)
{
    char buf[256];
    snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    t_cuda_err = buf;
    cudaGetLastError();  // clear the sticky-less error state
    return e == cudaErrorMemoryAllocation ? E_NOMEM : E_CUDA;
}
#define CK(call)                                          \
    do {                                                  \
        cudaError_t e__ = (call);                         \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

struct PSInfo {
    int K, L, zbits, w1len;
    size_t pk, sk, sig;
};
static bool ps_info(int ps, PSInfo& o)
{
    switch (ps) {
    case 44: o = {4, 4, 18, 768, 1312, 2560, 2420}; return true;
    case 65: o = {6, 5, 20, 768, 1952, 4032, 3309}; return true;
    case 87: o = {8, 7, 20, 1024, 2592, 4896, 4627}; return true;
    }
    return false;
}

constexpr int SIGN_SETS = 3;  // depth of the host entry points' chunk pipeline (staging sets, sign workspaces)

struct DeviceCtx {
    std::mutex mu;
    int dev = -1;
    cudaStream_t stream = nullptr;
    cudaStream_t s_tail = nullptr;                 // high priority: the latency-bound tail rounds of a sign chunk
    cudaStream_t s_in = nullptr, s_out = nullptr;  // H2D / D2H streams of the chunk pipeline
    cudaEvent_t ev_in[SIGN_SETS] = {}, ev_out[SIGN_SETS] = {}, ev_bulk[SIGN_SETS] = {}, ev_tail[SIGN_SETS] = {};
    uint8_t* arena = nullptr;
    size_t arena_bytes = 0;
    uint32_t* mailbox = nullptr;  // mapped pinned memory, 16 control words per sign workspace
    // device ranges that held secret-derived data during the most recent keygen / sign call on this device; all of
    // them must read as zero once the call has returned (mldsa_debug_secret_residue)
    std::vector<std::pair<const uint8_t*, size_t>> secret_ranges;
    std::vector<int> secret_slot;  // pipeline slot of each range: a later chunk in the same slot replaces the entries
    int cur_slot = 0;
    void begin_slot(int slot)
    {
        cur_slot = slot;
        for (size_t i = secret_ranges.size(); i-- > 0;)
            if (secret_slot[i] == slot) {
                secret_ranges.erase(secret_ranges.begin() + (long)i);
                secret_slot.erase(secret_slot.begin() + (long)i);
            }
    }
    void clear_secrets()
    {
        secret_ranges.clear();
        secret_slot.clear();
        cur_slot = 0;
    }
    void note_secret(const void* p, size_t bytes)
    {
        if (p && bytes) {
            secret_ranges.emplace_back(static_cast<const uint8_t*>(p), bytes);
            secret_slot.push_back(cur_slot);
        }
    }
    // the secret-derived arrays of a sign workspace, for the n lanes a chunk used (K, L, zbits: parameter set)
    void note_sign_ws(const SignWs& w, size_t n, bool shared, int K, int L, int zbits)
    {
        const size_t nk = shared ? 1 : n, D = (size_t)w.depth;
        note_secret(w.sh, nk * (size_t)(2 * K + L) * 1024);
        note_secret(w.w, n * D * K * 1024);                 // w - c*s2
        note_secret(w.yraw, n * D * L * 32 * (size_t)zbits);  // y
        note_secret(w.rhopp, n * 64);
    }

    int init(int d)
    {
        if (dev == d) return OK;
        CK(cudaSetDevice(d));
        CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        int prio_lo = 0, prio_hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CK(cudaStreamCreateWithPriority(&s_tail, cudaStreamNonBlocking, env_int("MLDSA_TAILPRIO", 1) ? prio_hi : prio_lo));
        CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
        for (int i = 0; i < SIGN_SETS; i++) {
            CK(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ev_bulk[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ev_tail[i], cudaEventDisableTiming));
        }
        CK(cudaHostAlloc(reinterpret_cast<void**>(&mailbox), 512, cudaHostAllocMapped | cudaHostAllocPortable));
        dev = d;
        return OK;
    }
    // Grow the arena to at least `bytes`.  `keep` bytes at its start (inputs already staged by a host
    // entry point) survive a reallocation.
    int reserve(size_t bytes, size_t keep = 0)
    {
        if (bytes <= arena_bytes) return OK;
        const size_t want = (bytes + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
        uint8_t* fresh = nullptr;
        // Scratch that held secrets is zeroed by every entry point before it returns; the old arena is wiped once more
        // before it goes back to the driver, so that nothing depends on that invariant (ZeroizeOnDrop, src/types.rs:19).
        if (keep == 0 && arena) {
            cudaMemset(arena, 0, arena_bytes);
            cudaFree(arena);
            arena = nullptr;
            arena_bytes = 0;
        }
        CK(cudaMalloc(reinterpret_cast<void**>(&fresh), want));
        if (arena) {
            cudaError_t e = cudaMemcpy(fresh, arena, keep, cudaMemcpyDeviceToDevice);
            cudaMemset(arena, 0, arena_bytes);
            cudaFree(arena);
            if (e != cudaSuccess) {
                cudaFree(fresh);
                arena = nullptr;
                arena_bytes = 0;
                return cuda_fail(e, "arena grow");
            }
        }
        arena = fresh;
        arena_bytes = want;
        return OK;
    }
};

constexpr int MAX_DEV = 64;
static DeviceCtx g_ctx[MAX_DEV];

// A set of secret keys expanded once and kept in HBM (mldsa_sk_load): what the reference's PrivateKey struct caches
// (s1-hat, s2-hat, t0-hat, K, tr; src/types.rs:19-28, built by expand_private, src/ml_dsa.rs:445-469) plus A-hat, which
// the reference re-derives from rho on every sign (src/ml_dsa.rs:181).  Own allocation, not part of the arena.
struct KeySet {
    uint32_t magic = 0x4B534554;  // "KSET"
    int ps = 0, dev = 0;
    size_t n = 0, bytes = 0;
    uint8_t* base = nullptr;
    bool is_public = false;   // false: secret keys (sign), true: public keys (verify)
    int32_t *A = nullptr, *sh = nullptr;  // sh: [n][2K+L][256] s2-hat, s1-hat, t0-hat (SignWs::sh)
    uint8_t* mini = nullptr;  // secret keys: [n][128] rho || K || tr
    int32_t* t1h = nullptr;   // public keys: [n][K][256] NTT(t1 * 2^d), the reference's t1_d2_hat_mont (src/types.rs:40)
    uint8_t* tr = nullptr;    // public keys: [n][64]
};

// bump allocator over the arena, 256-byte granules
struct Carver {
    uint8_t* base;
    size_t off = 0;
    explicit Carver(uint8_t* b) : base(b) {}
    template <class T> T* take(size_t count)
    {
        T* p = reinterpret_cast<T*>(base + off);
        off += (count * sizeof(T) + 255) & ~size_t(255);
        return p;
    }
};
static size_t pad256(size_t b) { return (b + 255) & ~size_t(255); }

static size_t workspace_budget(size_t arena_bytes)
{
    size_t lim = g_ws_limit.load();
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) free_b = size_t(8) << 30;
    free_b += arena_bytes;  // the cached arena is ours to reuse
    size_t def = free_b / 10 * 6;
    const size_t cap = size_t(96) << 30;
    if (def > cap) def = cap;
    if (lim == 0 || lim > free_b / 10 * 9) lim = lim == 0 ? def : free_b / 10 * 9;
    if (lim < (size_t(64) << 20)) lim = size_t(64) << 20;
    return lim;
}

// ---- workspace carving per operation ------------------------------------------------------------
// keygen keeps no A-hat (it is consumed as it is squeezed, k_fused.cu): seeds and the sampled s1 / s2 only
static size_t keygen_ws_bytes(const PSInfo& p, size_t n)
{
    return pad256(n * 128) + pad256(n * (size_t)(p.K + p.L) * 256);
}
static KeygenWs keygen_carve(Carver& c, const PSInfo& p, size_t n)
{
    KeygenWs w;
    w.seeds = c.take<uint8_t>(n * 128);
    w.s = c.take<int8_t>(n * (size_t)(p.K + p.L) * 256);
    w.A = nullptr;
    return w;
}
// A-hat / t1-hat / tr of ONE key (the shared-key path); per-lane packed keys never materialise A-hat (k_fused.cu) and
// resident key sets bring their own
static size_t verify_ws_bytes(const PSInfo& p, size_t n)
{
    return pad256((size_t)p.K * p.L * 1024) + pad256((size_t)p.K * 1024) + pad256(64) +
           pad256(n * (size_t)(64 + p.w1len)) + pad256(n * (size_t)SIB_BYTES) + pad256(n);
}
static VerifyWs verify_carve(Carver& c, const PSInfo& p, size_t n)
{
    VerifyWs w;
    w.A = c.take<int32_t>((size_t)p.K * p.L * 256);
    w.t1h = c.take<int32_t>((size_t)p.K * 256);
    w.tr = c.take<uint8_t>(64);
    w.mw1 = c.take<uint8_t>(n * (size_t)(64 + p.w1len));
    w.cmask = c.take<uint8_t>(n * (size_t)SIB_BYTES);
    w.pre_ok = c.take<uint8_t>(n);
    return w;
}
// resident: A-hat and the s-hat rows live in a loaded key set (mldsa_sk_load), not in the workspace
static size_t sign_ws_bytes(const PSInfo& p, size_t n, bool shared, bool resident = false)
{
    const size_t nk = resident ? 0 : (shared ? 1 : n), D = (size_t)tuning().depth;
    return pad256(nk * (size_t)p.K * p.L * 1024) + pad256(nk * (size_t)(2 * p.K + p.L) * 1024) +
           pad256(n * D * p.K * 1024) + pad256(n * D * p.L * 32 * p.zbits) + pad256(n * D * (64 + p.w1len)) +
           2 * pad256(n * 64) + 2 * pad256(n * D * 64) + pad256(n * 4) + pad256(n) + 2 * pad256(n * 4) + pad256(64);
}
static SignWs sign_carve(Carver& c, const PSInfo& p, size_t n, bool shared, uint32_t* mailbox, bool resident = false)
{
    const size_t nk = resident ? 0 : (shared ? 1 : n), D = (size_t)tuning().depth;
    SignWs w;
    w.depth = (int)D;
    w.A = c.take<int32_t>(nk * (size_t)p.K * p.L * 256);
    w.sh = c.take<int32_t>(nk * (size_t)(2 * p.K + p.L) * 256);
    w.w = c.take<int32_t>(n * D * p.K * 256);
    w.yraw = c.take<uint8_t>(n * D * p.L * 32 * p.zbits);
    w.mw1 = c.take<uint8_t>(n * D * (64 + p.w1len));
    w.mu = c.take<uint8_t>(n * 64);
    w.rhopp = c.take<uint8_t>(n * 64);
    w.ctilde = c.take<uint8_t>(n * D * 64);
    w.cmask = c.take<uint8_t>(n * D * 64);
    w.kappa = c.take<uint32_t>(n);
    w.done = c.take<uint8_t>(n);
    w.act[0] = c.take<uint32_t>(n);
    w.act[1] = c.take<uint32_t>(n);
    w.nact = c.take<uint32_t>(16);
    w.h_nact = mailbox;
    return w;
}

static size_t pick_chunk(size_t arena_bytes, size_t n, size_t fixed_bytes, size_t per_lane_bytes)
{
    // Steady state: the cached arena already holds the whole batch -- no driver call.  (cudaMemGetInfo takes
    // 0.1-5 ms, more while another process of the box allocates: measured as the run-to-run spread of the 11 ms
    // shared-key verify.)  An explicit workspace limit is always honoured.
    if (g_ws_limit.load() == 0 && fixed_bytes + n * (per_lane_bytes + 64) <= arena_bytes) return n;
    const size_t budget = workspace_budget(arena_bytes);
    size_t room = budget > fixed_bytes ? budget - fixed_bytes : 0;
    size_t chunk = room / (per_lane_bytes + 64);
    if (chunk < 256) chunk = 256;
    if (chunk > n) chunk = n;
    return chunk;
}

// =================================================================================================
// device-pointer cores (current device must already be set; ctx locked by the caller)
// =================================================================================================
static int keygen_core(DeviceCtx& cx, cudaStream_t st, int ps, size_t n, const uint8_t* d_xi, uint8_t* d_pk,
                       uint8_t* d_sk, size_t arena_skip)
{
    PSInfo p;
    if (!ps_info(ps, p)) return E_PARAM;
    const size_t per_lane = keygen_ws_bytes(p, 1024) / 1024 + 1;
    const size_t chunk = pick_chunk(cx.arena_bytes, n, arena_skip, per_lane);
    int rc = cx.reserve(arena_skip + keygen_ws_bytes(p, chunk), arena_skip);
    if (rc) return rc;
    cx.clear_secrets();
    for (size_t lo = 0; lo < n; lo += chunk) {
        const size_t m = n - lo < chunk ? n - lo : chunk;
        Carver c(cx.arena + arena_skip);
        KeygenWs ws = keygen_carve(c, p, chunk);  // same layout for every chunk
        cx.begin_slot(0);
        cx.note_secret(ws.seeds, m * 128);
        cx.note_secret(ws.s, m * (size_t)(p.K + p.L) * 256);
        CK(keygen_run(ps, st, m, d_xi + 32 * lo, d_pk + p.pk * lo, d_sk + p.sk * lo, ws));
        // rho', K, s1, s2 are secret: zero the scratch (mirrors ZeroizeOnDrop, src/types.rs:19)
        CK(cudaMemsetAsync(ws.seeds, 0, m * 128, st));
        CK(cudaMemsetAsync(ws.s, 0, m * (size_t)(p.K + p.L) * 256, st));
    }
    CK(cudaStreamSynchronize(st));
    prof_collect();
    return OK;
}

static int verify_core(DeviceCtx& cx, cudaStream_t st, int ps, int framing, size_t n, const uint8_t* d_pk,
                       size_t pk_stride, const uint8_t* d_msgs, const uint64_t* d_msg_off, const uint8_t* d_ctx,
                       const uint32_t* d_ctx_off, const uint8_t* d_sig, uint8_t* d_ok, size_t arena_skip)
{
    PSInfo p;
    if (!ps_info(ps, p)) return E_PARAM;
    const bool trace = env_int("MLDSA_TRACE", 0) != 0;
    const auto t_start = std::chrono::steady_clock::now();
    auto now_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(); };
    const bool shared = pk_stride == 0;
    const size_t per_lane = (verify_ws_bytes(p, 1024) - verify_ws_bytes(p, 0)) / 1024 + 1;
    const size_t chunk = pick_chunk(cx.arena_bytes, n, arena_skip + verify_ws_bytes(p, 0), per_lane);
    int rc = cx.reserve(arena_skip + verify_ws_bytes(p, chunk), arena_skip);
    if (rc) return rc;
    const double t_res = now_ms();
    for (size_t lo = 0; lo < n; lo += chunk) {
        const size_t m = n - lo < chunk ? n - lo : chunk;
        Carver c(cx.arena + arena_skip);
        VerifyWs ws = verify_carve(c, p, m);
        MsgDesc md{d_msgs, d_msg_off + lo, d_ctx, d_ctx_off ? d_ctx_off + lo : nullptr, framing};
        CK(verify_run(ps, st, m, d_pk + pk_stride * lo, pk_stride, md, d_sig + p.sig * lo, d_ok + lo, ws));
    }
    const double t_enq = now_ms();
    CK(cudaStreamSynchronize(st));
    if (trace)
        fprintf(stderr, "[mldsa trace] verify %zu lanes in chunks of %zu: workspace %.2f ms, enqueued %.2f ms, done %.2f ms\n", n,
                chunk, t_res, t_enq, now_ms());
    prof_collect();
    return OK;
}

// secrets of a sign workspace that the kernels do not zero themselves: the shared key's s-hat rows, and
// everything when a run stopped early (the s-hat rows .. yraw are contiguous in the carve)
static void sign_wipe(const SignWs& ws, size_t m, bool shared, bool failed, cudaStream_t st)
{
    if (failed) {
        cudaMemsetAsync(ws.sh, 0, (size_t)((uint8_t*)ws.mw1 - (uint8_t*)ws.sh), st);
        cudaMemsetAsync(ws.rhopp, 0, m * 64, st);
    } else if (shared) {
        cudaMemsetAsync(ws.sh, 0, (size_t)((uint8_t*)ws.w - (uint8_t*)ws.sh), st);
    }
}

static int sign_core(DeviceCtx& cx, cudaStream_t st, int ps, int framing, size_t n, const uint8_t* d_sk,
                     size_t sk_stride, const uint8_t* d_msgs, const uint64_t* d_msg_off, const uint8_t* d_ctx,
                     const uint32_t* d_ctx_off, const uint8_t* d_rnd, uint8_t* d_sig, uint32_t* d_iters,
                     size_t arena_skip)
{
    PSInfo p;
    if (!ps_info(ps, p)) return E_PARAM;
    const bool shared = sk_stride == 0;
    const size_t fixed = sign_ws_bytes(p, 0, shared);
    const size_t per_lane = (sign_ws_bytes(p, 1024, shared) - (shared ? fixed : 0)) / 1024 + 1;
    const size_t chunk = pick_chunk(cx.arena_bytes, n, arena_skip + fixed + 4096, per_lane);
    int rc = cx.reserve(arena_skip + sign_ws_bytes(p, chunk, shared), arena_skip);
    if (rc) return rc;
    t_last_rounds = 0;
    cx.clear_secrets();
    if (d_ctx_off) {
        // ctx longer than 255 bytes fails the whole call (src/lib.rs:274) BEFORE any candidate is computed: the rounds
        // write z of not-yet-accepted candidates into d_sig, and mu of such a lane would lack the ctx framing
        uint32_t* flag = cx.mailbox + 64;
        *flag = 0;
        CK(launch_ctx_check(st, n, d_ctx_off, flag));
        CK(cudaStreamSynchronize(st));
        if (*flag) return E_CTX;
    }
    for (size_t lo = 0; lo < n; lo += chunk) {
        const size_t m = n - lo < chunk ? n - lo : chunk;
        Carver c(cx.arena + arena_skip);
        SignWs ws = sign_carve(c, p, chunk, shared, cx.mailbox);  // same layout for every chunk
        cx.begin_slot(0);
        cx.note_sign_ws(ws, m, shared, p.K, p.L, p.zbits);
        MsgDesc md{d_msgs, d_msg_off + lo, d_ctx, d_ctx_off ? d_ctx_off + lo : nullptr, framing};
        SignRun run;
        const bool trace = env_int("MLDSA_TRACE", 0) != 0;
        const auto t0 = std::chrono::steady_clock::now();
        auto ms_since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
        cudaError_t e = sign_begin(ps, st, m, d_sk + sk_stride * lo, sk_stride, md, d_rnd + 32 * lo, d_sig + p.sig * lo,
                                   d_iters ? d_iters + lo : nullptr, ws, 0, &run);
        const double t_begin = ms_since();
        if (e == cudaSuccess) e = sign_tail(st, &run);
        const double t_tail = ms_since();
        const int planned = run.rounds;
        if (e == cudaSuccess) e = sign_finish(&run);
        if (trace)
            fprintf(stderr, "[mldsa trace] sign chunk of %zu lanes: bulk enqueued at %.2f ms, tail at %.2f ms, finished at %.2f ms; "
                    "%d planned rounds + %d straggler rounds\n", m, t_begin, t_tail, ms_since(), planned, run.rounds - planned);
        // Everything derived from the secret key or the mask (src/types.rs:19 ZeroizeOnDrop) is zeroed by
        // sign_z as each lane finishes; A-hat, mu||w1, c~ and the challenge are public.
        sign_wipe(ws, m, shared, e != cudaSuccess, st);
        if (e != cudaSuccess) cudaMemsetAsync(d_sig, 0, n * p.sig, st);  // no partial output next to an error code
        cudaStreamSynchronize(st);
        prof_collect();
        t_last_rounds += run.rounds;
        if (e != cudaSuccess) return run.ctx_err ? E_CTX : (run.incomplete ? E_UNSUPPORTED : cuda_fail(e, "sign"));
    }
    CK(cudaStreamSynchronize(st));
    prof_collect();
    return OK;
}

static DeviceCtx* lock_ctx(int dev, std::unique_lock<std::mutex>& lk, int& rc)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        t_cuda_err = "no CUDA device available (this library has no CPU fallback)";
        rc = E_CUDA;
        return nullptr;
    }
    if (dev < 0 || dev >= ndev || dev >= MAX_DEV) {
        rc = E_PARAM;
        return nullptr;
    }
    DeviceCtx& cx = g_ctx[dev];
    lk = std::unique_lock<std::mutex>(cx.mu);
    if (cudaSetDevice(dev) != cudaSuccess) {
        rc = E_CUDA;
        return nullptr;
    }
    rc = cx.init(dev);
    return rc ? nullptr : &cx;
}
static DeviceCtx* lock_current(std::unique_lock<std::mutex>& lk, int& rc)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        rc = E_CUDA;
        return nullptr;
    }
    return lock_ctx(dev, lk, rc);
}

static bool check_msgs(size_t n, const uint8_t* msgs, const uint64_t* msg_off, const uint8_t* ctx,
                       const uint32_t* ctx_off)
{
    if (!msg_off) return false;
    // offsets must not decrease: a non-monotone pair would underflow the unsigned length arithmetic below
    for (size_t i = 0; i < n; i++)
        if (msg_off[i + 1] < msg_off[i]) return false;
    if (ctx_off)
        for (size_t i = 0; i < n; i++)
            if (ctx_off[i + 1] < ctx_off[i]) return false;
    if (!msgs && n && msg_off[n] != msg_off[0]) return false;
    if ((ctx == nullptr) != (ctx_off == nullptr)) return ctx_off == nullptr || ctx_off[n] == ctx_off[0];
    return true;
}

}  // namespace mldsa

using namespace mldsa;

// Lanes per chunk of a host-pipeline call over n lanes when at most `cap` lanes fit the workspace; empty = no plan
// (the caller falls back to chunks of `cap`).
static std::vector<size_t> plan_sizes(size_t n, size_t cap)
{
    // Chunks ramp up and down in units of 32 768 lanes (1, 2, 4, then at most 6 units = 196 608 lanes, mirrored at the
    // end): the exposed pipeline fill (first copy-in) and drain (last tail + copy-out) are one unit's worth, the bulk
    // rounds of a 32 768-lane chunk are as efficient as those of a large one (161 ns per lane, timeline in
    // profiles/round1/r06_e2e_timeline.md), and a chunk's latency-bound tail hides under the next chunk's bulk.  2^20 lanes:
    // 1-2-4-6-6-6-4-2-1; 2^17 lanes: 1-2-1; below two units: one chunk.  MLDSA_PLAN overrides the shape with explicit
    // proportions ("1,3,4,4,3,1" = sixteenths).
    constexpr size_t UNIT = 32768;
    std::vector<size_t> sizes;  // lanes per chunk
    if (const char* pe = getenv("MLDSA_PLAN"); pe && *pe) {  // (an empty value counts as unset)
        std::vector<size_t> f;
        size_t total = 0;
        for (const char* q = pe; *q;) {
            char* end = nullptr;
            const long v = strtol(q, &end, 10);
            if (end == q) break;
            if (v > 0 && v < 4096) {
                f.push_back((size_t)v);
                total += (size_t)v;
            }
            q = *end ? end + 1 : end;
        }
        if (!f.empty() && n >= 2 * UNIT) {
            size_t acc = 0, prev = 0;
            for (size_t v : f) {
                acc += v;
                sizes.push_back(n * acc / total - prev);
                prev = n * acc / total;
            }
        }
    } else if (n >= 2 * UNIT && cap >= 2 * UNIT) {
        const size_t units = n / UNIT, rest = n % UNIT;
        const size_t step_cap = std::min<size_t>(6, (cap - (UNIT - 1)) / UNIT);  // the largest chunk also takes `rest`
        std::vector<size_t> left, right;
        size_t rem = units, step = 1;
        while (rem > 0) {
            const size_t a = std::min(step, rem);
            left.push_back(a);
            rem -= a;
            if (rem == 0) break;
            const size_t b = std::min(step, rem);
            right.insert(right.begin(), b);
            rem -= b;
            step = std::min(step * 2, step_cap);
        }
        left.insert(left.end(), right.begin(), right.end());
        size_t big = left.size() / 2;  // the remainder goes to the largest chunk, the middle one among equals
        for (size_t i = 0; i < left.size(); i++)
            if (left[i] > left[big]) big = i;
        for (size_t i = 0; i < left.size(); i++) sizes.push_back(left[i] * UNIT + (i == big ? rest : 0));
    }
    return sizes;
}

extern "C" {

int mldsa_device_count(void)
{
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

const char* mldsa_strerror(int code)
{
    switch (code) {
    case MLDSA_OK: return "ok";
    case MLDSA_E_PARAM: return "bad parameter";
    case MLDSA_E_CTX: return "ML-DSA.Sign: ctx too long";
    case MLDSA_E_CUDA: return "CUDA error";
    case MLDSA_E_NOMEM: return "out of device memory";
    case MLDSA_E_UNSUPPORTED: return "not supported";
    }
    return "unknown error";
}
const char* mldsa_last_cuda_error(void) { return t_cuda_err.c_str(); }

int mldsa_sizes(int ps, size_t* pk_len, size_t* sk_len, size_t* sig_len)
{
    PSInfo p;
    if (!ps_info(ps, p)) return E_PARAM;
    if (pk_len) *pk_len = p.pk;
    if (sk_len) *sk_len = p.sk;
    if (sig_len) *sig_len = p.sig;
    return OK;
}

void* mldsa_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void mldsa_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}
void mldsa_set_workspace_limit(size_t bytes) { g_ws_limit.store(bytes); }
void mldsa_set_depth(int depth) { g_depth.store(depth < 1 ? 1 : (depth > 8 ? 8 : depth)); }
void mldsa_debug_set_max_rounds(int rounds) { g_max_rounds.store(rounds < 0 ? 0 : rounds); }
void mldsa_set_tail(int tail_depth, int tail_slots)
{
    g_tail_depth.store(tail_depth < 0 ? 0 : (tail_depth > 16 ? 16 : tail_depth));
    g_tail_slots.store(tail_slots < 0 ? 0 : tail_slots);
}
void mldsa_set_tuning(int ways, int kcap, int acap)
{
    (void)ways;
    g_kcap.store(kcap < 0 ? 0 : kcap);
    g_acap.store(acap < 0 ? 0 : acap);
}
uint64_t mldsa_kernel_launches(void) { return g_launches.load(); }

void mldsa_profile_enable(int on) { g_profiling.store(on ? 1 : 0); }
void mldsa_profile_reset(void)
{
    prof_collect_now();
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int i = 0; i < KID_COUNT; i++) {
        g_prof_ms[i] = 0;
        g_prof_n[i] = 0;
    }
}
int mldsa_profile_count(void) { return KID_COUNT; }
const char* mldsa_profile_name(int id) { return id >= 0 && id < KID_COUNT ? KERNEL_NAMES[id] : ""; }
int mldsa_profile_get(int id, double* ms, uint64_t* launches)
{
    if (id < 0 || id >= KID_COUNT) return E_PARAM;
    prof_collect_now();
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (ms) *ms = g_prof_ms[id];
    if (launches) *launches = g_prof_n[id];
    return OK;
}

// Keccak-f[1600] permutations per second for rotation-split variant `variant` (see k_probe.cu); *mask
// receives the variant's KECCAK_FMA_SITES value.  Returns -1 when variant is out of range.
int mldsa_keccak_probe(int variant, double* perms_per_sec, uint32_t* mask)
{
    if (variant < 0 || variant >= keccak_probe_variants() || !perms_per_sec) return E_PARAM;
    if (mask) *mask = keccak_probe_mask(variant);
    int dev = 0, sms = 148;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 4 * 4, reps = 64;
    uint64_t* state = nullptr;
    CK(cudaMalloc(reinterpret_cast<void**>(&state), (size_t)blocks * 128 * 25 * 8));
    CK(cudaMemset(state, 0x5a, (size_t)blocks * 128 * 25 * 8));
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(a, 0);
        cudaError_t e = launch_keccak_probe(0, variant, blocks, reps, state);
        cudaEventRecord(b, 0);
        if (e != cudaSuccess || cudaEventSynchronize(b) != cudaSuccess) {
            cudaFree(state);
            return cuda_fail(e != cudaSuccess ? e : cudaGetLastError(), "keccak probe");
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        const double rate = (double)blocks * 128 * reps / (ms * 1e-3);
        if (rep > 0 && rate > best) best = rate;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(state);
    *perms_per_sec = best;
    return OK;
}

// Integer issue-rate probe on the current device: mode 0 = LOP3+SHF (ALU pipe), 1 = IMAD+IMAD.HI (FMA
// pipe), 2 = both interleaved.  Returns thread-level integer instructions per second.
int mldsa_int32_peak_probe(int mode, double* ops_per_sec)
{
    if (mode < 0 || mode > 8 || !ops_per_sec) return E_PARAM;
    int dev = 0, sms = 148;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    uint32_t* scratch = nullptr;
    CK(cudaMalloc(reinterpret_cast<void**>(&scratch), 256));
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    double ops = 0, best = 0;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(a, 0);
        cudaError_t e = launch_int_probe(0, mode, sms * 8, 2048, scratch, &ops);
        cudaEventRecord(b, 0);
        if (e != cudaSuccess || cudaEventSynchronize(b) != cudaSuccess) {
            cudaFree(scratch);
            return cuda_fail(e != cudaSuccess ? e : cudaGetLastError(), "int probe");
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms > 0 && ops / (ms * 1e-3) > best) best = ops / (ms * 1e-3);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(scratch);
    *ops_per_sec = best;
    return OK;
}
int mldsa_last_sign_rounds(void) { return t_last_rounds; }

// Test hook: the chunk sizes the host pipelines would use for n lanes when `cap` lanes fit the workspace (no device
// needed).  Returns the number of chunks (0: no ramp plan, the caller cuts chunks of `cap`), at most max_out are written.
size_t mldsa_debug_chunk_plan(size_t n, size_t cap, size_t* out, size_t max_out)
{
    const std::vector<size_t> sizes = plan_sizes(n, cap);
    bool fits = !sizes.empty();
    for (size_t m : sizes) fits = fits && m > 0 && m <= cap;
    if (!fits) return 0;
    for (size_t i = 0; i < sizes.size() && i < max_out; i++) out[i] = sizes[i];
    return sizes.size();
}

// Test hook for the ZeroizeOnDrop analogue (src/types.rs:19): bytes that are not zero in the device ranges which held
// secret-derived data (staged sk / xi / rnd, rho', K, s1, s2, their NTT images, y, w - c*s2, rho'') during the most
// recent keygen or sign call on `dev`.  Must be 0 after every call, successful or not.
int mldsa_debug_secret_residue(int dev, uint64_t* nonzero_bytes, uint64_t* scanned_bytes)
{
    if (!nonzero_bytes) return E_PARAM;
    std::unique_lock<std::mutex> lk;
    int rc = OK;
    DeviceCtx* cx = lock_ctx(dev, lk, rc);
    if (!cx) return rc;
    unsigned long long* d_cnt = nullptr;
    CK(cudaMalloc(reinterpret_cast<void**>(&d_cnt), 8));
    CK(cudaMemsetAsync(d_cnt, 0, 8, cx->stream));
    uint64_t scanned = 0;
    for (auto& r : cx->secret_ranges) {
        if (r.first < cx->arena || r.first + r.second > cx->arena + cx->arena_bytes) continue;  // arena was re-made
        cudaError_t e = launch_count_nonzero(cx->stream, r.first, r.second, d_cnt);
        if (e != cudaSuccess) {
            cudaFree(d_cnt);
            return cuda_fail(e, "count_nonzero");
        }
        scanned += r.second;
    }
    unsigned long long h = 0;
    cudaError_t e = cudaMemcpyAsync(&h, d_cnt, 8, cudaMemcpyDeviceToHost, cx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cx->stream);
    cudaFree(d_cnt);
    if (e != cudaSuccess) return cuda_fail(e, "secret residue");
    *nonzero_bytes = h;
    if (scanned_bytes) *scanned_bytes = scanned;
    return OK;
}

// ---------------------------------------------------------------- device-pointer entry points
int mldsa_keygen_batch_dev(int ps, void* stream, size_t n, const uint8_t* d_xi, uint8_t* d_pk, uint8_t* d_sk)
{
    PSInfo p;
    if (!ps_info(ps, p) || (n && (!d_xi || !d_pk || !d_sk))) return E_PARAM;
    if (n == 0) return OK;
    std::unique_lock<std::mutex> lk;
    int rc = OK;
    DeviceCtx* cx = lock_current(lk, rc);
    if (!cx) return rc;
    return keygen_core(*cx, stream ? (cudaStream_t)stream : cx->stream, ps, n, d_xi, d_pk, d_sk, 0);
}

int mldsa_sign_batch_dev(int ps, void* stream, int framing, size_t n, const uint8_t* d_sk, size_t sk_stride,
                         const uint8_t* d_msgs, const uint64_t* d_msg_off, const uint8_t* d_ctx,
                         const uint32_t* d_ctx_off, const uint8_t* d_rnd, uint8_t* d_sig, uint32_t* d_iters)
{
    PSInfo p;
    if (!ps_info(ps, p) || (framing != 0 && framing != 1)) return framing == 2 ? E_UNSUPPORTED : E_PARAM;
    if (sk_stride != 0 && sk_stride != p.sk) return E_PARAM;
    if (n && (!d_sk || !d_msg_off || !d_rnd || !d_sig)) return E_PARAM;
    if (n == 0) return OK;
    std::unique_lock<std::mutex> lk;
    int rc = OK;
    DeviceCtx* cx = lock_current(lk, rc);
    if (!cx) return rc;
    return sign_core(*cx, stream ? (cudaStream_t)stream : cx->stream, ps, framing, n, d_sk, sk_stride, d_msgs,
                     d_msg_off, d_ctx, d_ctx_off, d_rnd, d_sig, d_iters, 0);
}

int mldsa_verify_batch_dev(int ps, void* stream, int framing, size_t n, const uint8_t* d_pk, size_t pk_stride,
                           const uint8_t* d_msgs, const uint64_t* d_msg_off, const uint8_t* d_ctx,
                           const uint32_t* d_ctx_off, const uint8_t* d_sig, uint8_t* d_ok)
{
    PSInfo p;
    if (!ps_info(ps, p) || (framing != 0 && framing != 1)) return framing == 2 ? E_UNSUPPORTED : E_PARAM;
    if (pk_stride != 0 && pk_stride != p.pk) return E_PARAM;
    if (n && (!d_pk || !d_msg_off || !d_sig || !d_ok)) return E_PARAM;
    if (n == 0) return OK;
    std::unique_lock<std::mutex> lk;
    int rc = OK;
    DeviceCtx* cx = lock_current(lk, rc);
    if (!cx) return rc;
    return verify_core(*cx, stream ? (cudaStream_t)stream : cx->stream, ps, framing, n, d_pk, pk_stride, d_msgs,
                       d_msg_off, d_ctx, d_ctx_off, d_sig, d_ok, 0);
}

int mldsa_ntt_batch_dev(void* stream, size_t n, int32_t* d_polys)
{
    if (n && (!d_polys || (reinterpret_cast<uintptr_t>(d_polys) & 31))) return E_PARAM;  // rows move as 256-bit vectors
    CK(launch_ntt((cudaStream_t)stream, d_polys, n, false));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    prof_collect();
    return OK;
}
int mldsa_intt_batch_dev(void* stream, size_t n, int32_t* d_polys)
{
    if (n && (!d_polys || (reinterpret_cast<uintptr_t>(d_polys) & 31))) return E_PARAM;
    CK(launch_ntt((cudaStream_t)stream, d_polys, n, true));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    prof_collect();
    return OK;
}
int mldsa_expand_a_batch_dev(void* stream, size_t n, const uint8_t* d_seed34, int32_t* d_out)
{
    if (n && (!d_seed34 || !d_out)) return E_PARAM;
    CK(launch_expand_a_seed34((cudaStream_t)stream, d_seed34, n, d_out));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    prof_collect();
    return OK;
}

// ---------------------------------------------------------------- host-pointer entry points
// Inputs are staged into the device arena chunk by chunk, results copied back.  (Pinned caller
// buffers -- mldsa_host_alloc -- make these copies true async DMA.)

int mldsa_keygen_batch(int ps, int dev, size_t n, const uint8_t* xi, uint8_t* pk, uint8_t* sk)
{
    PSInfo p;
    if (!ps_info(ps, p) || (n && (!xi || !pk || !sk))) return E_PARAM;
    if (n == 0) return OK;
    std::unique_lock<std::mutex> lk;
    int rc = OK;
    DeviceCtx* cx = lock_ctx(dev, lk, rc);
    if (!cx) return rc;
    // Keygen is copy-OUT bound (32 bytes in, pk + sk = 3.9 / 6.0 / 7.5 KB out per key against ~60 ns of compute),
    // so the chunks' kernels run ahead on the work stream and the copy-out stream never idles: NS output staging
    // sets, one workspace (kernels are stream-ordered), no host wait until the end.
    constexpr size_t NS = SIGN_SETS;
    const size_t io_lane = NS * (32 + p.pk + p.sk);
    size_t chunk = pick_chunk(cx->arena_bytes, n, 65536, keygen_ws_bytes(p, 1024) / 1024 + io_lane + 1);
    if (n >= (size_t(1) << 17) && chunk > (n + 15) / 16) chunk = (n + 15) / 16;  // >= 16 chunks: short fill and drain
    const size_t set_bytes = pad256(chunk * 32) + pad256(chunk * p.pk) + pad256(chunk * p.sk);
    rc = cx->reserve(NS * set_bytes + keygen_ws_bytes(p, chunk));
    if (rc) return rc;
    struct Set {
        uint8_t *xi, *pk, *sk;
    } set[NS];
    Carver c(cx->arena);
    for (auto& s : set) {
        s.xi = c.take<uint8_t>(chunk * 32);
        s.pk = c.take<uint8_t>(chunk * p.pk);
        s.sk = c.take<uint8_t>(chunk * p.sk);
    }
    uint8_t* const ws_base = cx->arena + c.off;
    cx->clear_secrets();
    size_t used[NS] = {};  // most lanes any chunk put into each set (only what this call touched is registered)
    auto note_used = [&]() {
        size_t most = 0;
        for (size_t k = 0; k < NS; k++) {
            cx->note_secret(set[k].xi, used[k] * 32);
            cx->note_secret(set[k].sk, used[k] * p.sk);
            if (used[k] > most) most = used[k];
        }
        Carver wc(ws_base);
        const KeygenWs w0 = keygen_carve(wc, p, chunk);
        cx->note_secret(w0.seeds, most * 128);
        cx->note_secret(w0.s, most * (size_t)(p.K + p.L) * 256);
    };
    cudaStream_t st = cx->stream;
    cudaEvent_t* ev_cmp = cx->ev_bulk;
    auto drain = [&](int code) -> int {
        const std::string err = t_cuda_err;
        cudaStreamSynchronize(st);
        cudaStreamSynchronize(cx->s_out);
        for (auto& s : set) {
            cudaMemsetAsync(s.xi, 0, chunk * 32, st);
            cudaMemsetAsync(s.sk, 0, chunk * p.sk, st);
        }
        cudaMemsetAsync(ws_base, 0, keygen_ws_bytes(p, chunk), st);
        cudaStreamSynchronize(st);
        cudaGetLastError();
        prof_collect();
        note_used();
        t_cuda_err = err;
        return code;
    };
    auto enqueue = [&](size_t i, size_t lo, size_t m) -> int {
        Set& s = set[i % NS];
        if (m > used[i % NS]) used[i % NS] = m;
        if (i >= NS) CK(cudaStreamWaitEvent(st, cx->ev_out[i % NS], 0));  // the set's previous keys are on the host
        CK(cudaMemcpyAsync(s.xi, xi + 32 * lo, m * 32, cudaMemcpyHostToDevice, st));
        Carver wc(ws_base);
        const KeygenWs ws = keygen_carve(wc, p, chunk);  // same layout for every chunk
        CK(keygen_run(ps, st, m, s.xi, s.pk, s.sk, ws));
        // xi, rho', K, s1, s2 are secret: zero the scratch (mirrors ZeroizeOnDrop, src/types.rs:19)
        CK(cudaMemsetAsync(s.xi, 0, m * 32, st));
        CK(cudaMemsetAsync(ws.seeds, 0, m * 128, st));
        CK(cudaMemsetAsync(ws.s, 0, m * (size_t)(p.K + p.L) * 256, st));
        CK(cudaEventRecord(ev_cmp[i % NS], st));
        CK(cudaStreamWaitEvent(cx->s_out, ev_cmp[i % NS], 0));
        CK(cudaMemcpyAsync(pk + p.pk * lo, s.pk, m * p.pk, cudaMemcpyDeviceToHost, cx->s_out));
        CK(cudaMemcpyAsync(sk + p.sk * lo, s.sk, m * p.sk, cudaMemcpyDeviceToHost, cx->s_out));
        CK(cudaMemsetAsync(s.sk, 0, m * p.sk, cx->s_out));
        CK(cudaEventRecord(cx->ev_out[i % NS], cx->s_out));
        return OK;
    };
    CK(cudaStreamSynchronize(cx->s_out));
    size_t i = 0;
    for (size_t lo = 0; lo < n; lo += chunk, i++)
        if ((rc = enqueue(i, lo, n - lo < chunk ? n - lo : chunk))) return drain(rc);
    cudaError_t e = cudaStreamSynchronize(cx->s_out);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return drain(cuda_fail(e, "keygen pipeline"));
    prof_collect();
    note_used();
    return OK;
}

// ---- chunk pipeline of the host-pointer entry points ---------------------------------------------
// A batch is cut into chunks; chunk i+1's inputs are copied in (stream s_in) and chunk i-1's results are
// copied out (stream s_out) while chunk i computes, with two sets of staging buffers.  With pinned caller
// buffers the copies are true async DMA in both directions, so a step costs ~max(compute, H2D, D2H) instead
// of their sum.  Per-chunk inputs are the packed bytes only (SURVEY.md section 7 "PCIe can beat compute").
struct ChunkPlan {
    std::vector<size_t> lo;  // chunk i covers lanes [lo[i], lo[i+1])
    size_t max_lanes = 0, max_mbytes = 0, max_cbytes = 0;
    size_t nchunks() const { return lo.size() - 1; }
};
static ChunkPlan plan_chunks(size_t arena_bytes, size_t n, size_t fixed, size_t per_lane, const uint64_t* msg_off,
                             const uint32_t* ctx_off)
{
    ChunkPlan pl;
    const size_t cap = pick_chunk(arena_bytes, n, fixed, per_lane);
    pl.lo.push_back(0);
    std::vector<size_t> sizes = plan_sizes(n, cap);
    bool fits = !sizes.empty();
    for (size_t m : sizes) fits = fits && m > 0 && m <= cap;
    if (fits) {
        size_t acc = 0;
        for (size_t m : sizes) {
            acc += m;
            pl.lo.push_back(acc);
        }
    } else {
        for (size_t lo = cap; lo < n; lo += cap) pl.lo.push_back(lo);
        pl.lo.push_back(n);
    }
    for (size_t i = 0; i + 1 < pl.lo.size(); i++) {
        const size_t lo = pl.lo[i], m = pl.lo[i + 1] - lo;
        const size_t mb = (size_t)(msg_off[lo + m] - msg_off[lo]);
        const size_t cb = ctx_off ? (size_t)(ctx_off[lo + m] - ctx_off[lo]) : 0;
        if (m > pl.max_lanes) pl.max_lanes = m;
        if (mb > pl.max_mbytes) pl.max_mbytes = mb;
        if (cb > pl.max_cbytes) pl.max_cbytes = cb;
    }
    return pl;
}

}  // extern "C"

// ph < 0: plain ML-DSA with `framing`; ph = 0/1/2: HashML-DSA with SHA-256 / SHA-512 / SHAKE128 pre-hash
// ks != nullptr: the keys are a resident key set; lane i signs with key key_idx[i] (nullptr: key i), sk is unused
static int sign_host(int ps, int dev, int framing, int ph, size_t n, const uint8_t* sk, size_t sk_stride,
                     const uint8_t* msgs, const uint64_t* msg_off, const uint8_t* ctx, const uint32_t* ctx_off,
                     const uint8_t* rnd, uint8_t* sig, uint32_t* iters, const KeySet* ks = nullptr,
                     const uint32_t* key_idx = nullptr)
{
    PSInfo p;
    if (!ps_info(ps, p) || (framing != 0 && framing != 1) || ph > 2) return E_PARAM;
    if (ph >= 0) framing = 2;
    if (!ks && sk_stride != 0 && sk_stride != p.sk) return E_PARAM;
    if (n && ((!ks && !sk) || !rnd || !sig || !check_msgs(n, msgs, msg_off, ctx, ctx_off))) return E_PARAM;
    if (n == 0) return OK;
    if (ks) {
        if (ks->ps != ps || ks->dev != dev || ks->is_public) return E_PARAM;
        if (!key_idx && n > ks->n) return E_PARAM;
        if (key_idx)
            for (size_t i = 0; i < n; i++)
                if (key_idx[i] >= ks->n) return E_PARAM;
        sk_stride = 128;
    }
    if (ctx_off)
        for (size_t i = 0; i < n; i++)
            if (ctx_off[i + 1] - ctx_off[i] > 255) return E_CTX;  // src/lib.rs:274
    std::unique_lock<std::mutex> lk;
    int rc = OK;
    DeviceCtx* cx = lock_ctx(dev, lk, rc);
    if (!cx) return rc;
    const bool shared = sk_stride == 0, resident = ks != nullptr;
    const size_t sk_lane = resident ? (key_idx ? 4 : 0) : p.sk;  // staged bytes per lane: packed sk, or a key index
    const size_t avg_msg = (size_t)((msg_off[n] - msg_off[0]) / n) + 1;
    const size_t avg_ctx = ctx_off ? (size_t)((ctx_off[n] - ctx_off[0]) / n) + 1 : 0;
    // NS sets of staged inputs / outputs AND rejection-loop workspaces: chunk i's tail rounds and its copy-out run
    // under the bulk rounds of chunks i+1 and i+2
    constexpr size_t NS = SIGN_SETS;
    const size_t fixed = NS * (sign_ws_bytes(p, 0, shared, resident) + 4096) + (shared ? NS * pad256(p.sk) : 0) + 65536;
    const size_t ws_lane =
        NS * ((sign_ws_bytes(p, 1024, shared, resident) - (shared ? sign_ws_bytes(p, 0, true) : sign_ws_bytes(p, 0, false, resident))) / 1024 + 1);
    const size_t io_lane = NS * ((shared ? 0 : sk_lane) + 32 + p.sig + 4 + 8 + 4 + avg_msg + avg_ctx + PREHASH_STRIDE + 8);
    const ChunkPlan pl = plan_chunks(cx->arena_bytes, n, fixed, ws_lane + io_lane, msg_off, ctx_off);
    const size_t C = pl.max_lanes, N = pl.nchunks();
    struct Set {
        uint8_t *sk, *rnd, *sig, *m, *c, *pm;
        uint32_t *it, *co;
        uint64_t *mo, *po;
    } set[NS];
    const size_t sk_set = shared ? p.sk : C * sk_lane + 16;
    const size_t set_bytes = pad256(sk_set) + pad256(C * 32) + pad256(C * p.sig) + pad256(C * 4) +
                             pad256((C + 1) * 8) + pad256((C + 1) * 4) + pad256(pl.max_mbytes + 8) +
                             pad256(pl.max_cbytes + 8) + pad256(C * PREHASH_STRIDE) + pad256((C + 1) * 8);
    const size_t ws_bytes = pad256(sign_ws_bytes(p, C, shared, resident)) + 4096;
    rc = cx->reserve(NS * (set_bytes + ws_bytes));
    if (rc) return rc;
    Carver c(cx->arena);
    for (auto& s : set) {
        s.sk = c.take<uint8_t>(sk_set);  // packed secret keys -- or, for a resident key set, the lanes' key indices
        s.rnd = c.take<uint8_t>(C * 32);
        s.sig = c.take<uint8_t>(C * p.sig);
        s.it = c.take<uint32_t>(C);
        s.mo = c.take<uint64_t>(C + 1);
        s.co = c.take<uint32_t>(C + 1);
        s.m = c.take<uint8_t>(pl.max_mbytes + 8);
        s.c = c.take<uint8_t>(pl.max_cbytes + 8);
        s.pm = c.take<uint8_t>(C * PREHASH_STRIDE);
        s.po = c.take<uint64_t>(C + 1);
    }
    uint8_t* ws_base[NS];
    for (size_t k = 0; k < NS; k++) ws_base[k] = cx->arena + c.off + k * ws_bytes;
    cx->clear_secrets();
    cudaStream_t B = cx->stream, T = cx->s_tail;
    // the bulk rounds stop when about this many lanes are expected to be live: below it the kernels no longer fill
    // the device and a round costs its (fixed) latency
    const size_t tail_lanes = N > 1 ? (size_t)env_int("MLDSA_TAIL", 16384) : 0;
    SignRun run[NS];
    size_t lanes_of[NS] = {};
    bool started[NS] = {};
    const bool trace = env_int("MLDSA_TRACE", 0) != 0;
    const auto t_start = std::chrono::steady_clock::now();
    auto now_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(); };
    // MLDSA_TRACE=1: device-side timeline of the pipeline (timing events on every stream, printed at the end)
    enum { TR_IN0, TR_IN1, TR_BULK0, TR_BULK1, TR_TAIL1, TR_OUT0, TR_OUT1, TR_N };
    std::vector<cudaEvent_t> tr;
    cudaEvent_t tr_origin = nullptr;
    auto mark = [&](size_t i, int what, cudaStream_t st) {
        if (trace) cudaEventRecord(tr[i * TR_N + what], st);
    };
    if (trace) {
        tr.resize(N * TR_N);
        for (auto& e : tr) cudaEventCreate(&e);
        cudaEventCreate(&tr_origin);
    }

    auto copy_in = [&](size_t i) -> int {
        const size_t lo = pl.lo[i], m = pl.lo[i + 1] - lo;
        Set& s = set[i % NS];
        const size_t mbytes = (size_t)(msg_off[lo + m] - msg_off[lo]);
        const size_t cbytes = ctx_off ? (size_t)(ctx_off[lo + m] - ctx_off[lo]) : 0;
        if (i >= NS) CK(cudaStreamWaitEvent(cx->s_in, cx->ev_bulk[i % NS], 0));  // the previous user of this set has consumed its inputs
        mark(i, TR_IN0, cx->s_in);
        if (!resident)
            CK(cudaMemcpyAsync(s.sk, sk + sk_stride * lo, shared ? p.sk : m * p.sk, cudaMemcpyHostToDevice, cx->s_in));
        else if (key_idx)
            CK(cudaMemcpyAsync(s.sk, key_idx + lo, m * 4, cudaMemcpyHostToDevice, cx->s_in));
        CK(cudaMemcpyAsync(s.rnd, rnd + 32 * lo, m * 32, cudaMemcpyHostToDevice, cx->s_in));
        CK(cudaMemcpyAsync(s.mo, msg_off + lo, (m + 1) * 8, cudaMemcpyHostToDevice, cx->s_in));
        if (mbytes) CK(cudaMemcpyAsync(s.m, msgs + msg_off[lo], mbytes, cudaMemcpyHostToDevice, cx->s_in));
        if (ctx_off) {
            CK(cudaMemcpyAsync(s.co, ctx_off + lo, (m + 1) * 4, cudaMemcpyHostToDevice, cx->s_in));
            if (cbytes) CK(cudaMemcpyAsync(s.c, ctx + ctx_off[lo], cbytes, cudaMemcpyHostToDevice, cx->s_in));
        }
        CK(cudaEventRecord(cx->ev_in[i % NS], cx->s_in));
        mark(i, TR_IN1, cx->s_in);
        return OK;
    };
    // set-up and bulk rounds of chunk i on the work stream; nothing here waits for the device
    auto enqueue_bulk = [&](size_t i) -> int {
        const size_t lo = pl.lo[i], m = pl.lo[i + 1] - lo;
        Set& s = set[i % NS];
        CK(cudaStreamWaitEvent(B, cx->ev_in[i % NS], 0));
        if (i >= NS) {
            CK(cudaStreamWaitEvent(B, cx->ev_out[i % NS], 0));   // the copy-out of the set's previous chunk has drained its outputs
            CK(cudaStreamWaitEvent(B, cx->ev_tail[i % NS], 0));  // ... and that chunk is done with the workspace
        }
        mark(i, TR_BULK0, B);
        const uint8_t* cm = s.m - msg_off[lo];
        const uint64_t* cmo = s.mo;
        if (ph >= 0) {  // HashML-DSA: M' carries OID || PH(M) (src/lib.rs:310-342)
            CK(launch_prehash(B, ph, m, cm, cmo, s.pm, s.po));
            cm = s.pm;
            cmo = s.po;
        }
        Carver wc(ws_base[i % NS]);
        SignWs ws = sign_carve(wc, p, C, shared, cx->mailbox + 16 * (i % NS), resident);  // same layout for every chunk
        const MsgDesc md{cm, cmo, ctx_off ? s.c - ctx_off[lo] : nullptr, ctx_off ? s.co : nullptr, framing};
        lanes_of[i % NS] = m;
        started[i % NS] = true;
        cx->begin_slot((int)(i % NS));  // (a later chunk re-carves the slot: only its ranges are meaningful)
        const uint8_t* key_rows = s.sk;
        if (resident) {
            // lanes [lo, lo + m) use keys key_idx[lo ..] -- or keys lo .. lo + m - 1 of the set
            const size_t k0 = key_idx ? 0 : lo;
            ws.keys_resident = true;
            ws.kidx = key_idx ? reinterpret_cast<const uint32_t*>(s.sk) : nullptr;
            ws.A = ks->A + k0 * (size_t)p.K * p.L * 256;
            ws.sh = ks->sh + k0 * (size_t)(2 * p.K + p.L) * 256;
            key_rows = ks->mini + k0 * 128;
            cx->note_secret(ws.w, m * (size_t)ws.depth * p.K * 1024);
            cx->note_secret(ws.yraw, m * (size_t)ws.depth * p.L * 32 * p.zbits);
            cx->note_secret(ws.rhopp, m * 64);
        } else {
            cx->note_sign_ws(ws, m, shared, p.K, p.L, p.zbits);
            cx->note_secret(s.sk, shared ? p.sk : m * p.sk);
        }
        cx->note_secret(s.rnd, m * 32);
        cudaError_t e = sign_begin(ps, B, m, key_rows, sk_stride, md, s.rnd, s.sig, s.it, ws, tail_lanes, &run[i % NS]);
        if (e != cudaSuccess) return run[i % NS].ctx_err ? E_CTX : cuda_fail(e, "sign_begin");
        if (!resident) CK(cudaMemsetAsync(s.sk, 0, shared ? p.sk : m * p.sk, B));  // secrets do not linger in the staging set
        CK(cudaMemsetAsync(s.rnd, 0, m * 32, B));
        CK(cudaEventRecord(cx->ev_bulk[i % NS], B));
        mark(i, TR_BULK1, B);
        return OK;
    };
    auto enqueue_tail = [&](size_t i) -> int {
        CK(cudaStreamWaitEvent(T, cx->ev_bulk[i % NS], 0));
        cudaError_t e = sign_tail(T, &run[i % NS]);
        return e == cudaSuccess ? OK : cuda_fail(e, "sign_tail");
    };
    auto finish = [&](size_t i) -> int {
        SignRun& r = run[i % NS];
        cudaError_t e = sign_finish(&r);  // the host waits here, with chunk i+1's bulk already queued behind
        if (e != cudaSuccess) return r.ctx_err ? E_CTX : (r.incomplete ? E_UNSUPPORTED : cuda_fail(e, "sign_finish"));
        if (!resident) sign_wipe(r.ws, lanes_of[i % NS], shared, false, T);
        CK(cudaEventRecord(cx->ev_tail[i % NS], T));
        mark(i, TR_TAIL1, T);
        started[i % NS] = false;
        t_last_rounds += r.rounds;
        prof_collect();
        return OK;
    };
    auto bail = [&](int code) -> int {  // stop everything, leave no secrets behind
        const std::string err = t_cuda_err;
        cudaStreamSynchronize(cx->s_in);
        cudaStreamSynchronize(B);
        cudaStreamSynchronize(T);
        cudaStreamSynchronize(cx->s_out);
        for (size_t k = 0; k < NS; k++) {
            if (started[k] && !resident) sign_wipe(run[k].ws, lanes_of[k], shared, true, B);
            if (started[k] && resident) {  // the key set stays; y, w - c*s2 and rho'' of the unfinished lanes go
                cudaMemsetAsync(run[k].ws.w, 0, (size_t)((uint8_t*)run[k].ws.mw1 - (uint8_t*)run[k].ws.w), B);
                cudaMemsetAsync(run[k].ws.rhopp, 0, lanes_of[k] * 64, B);
            }
            if (!resident) cudaMemsetAsync(set[k].sk, 0, sk_set, B);
            cudaMemsetAsync(set[k].rnd, 0, C * 32, B);
            cudaMemsetAsync(set[k].sig, 0, C * p.sig, B);  // z of candidates that were never accepted
        }
        cudaStreamSynchronize(B);
        cudaGetLastError();
        prof_collect();
        t_cuda_err = err;
        return code;
    };

    // make sure nothing of a previous call is still in flight on the copy streams
    CK(cudaStreamSynchronize(cx->s_in));
    CK(cudaStreamSynchronize(cx->s_out));
    t_last_rounds = 0;
    if (trace) cudaEventRecord(tr_origin, cx->s_in);
    if ((rc = copy_in(0))) return bail(rc);
    if (N > 1 && (rc = copy_in(1))) return bail(rc);
    if ((rc = enqueue_bulk(0))) return bail(rc);
    for (size_t i = 0; i < N; i++) {
        const size_t lo = pl.lo[i], m = pl.lo[i + 1] - lo;
        Set& s = set[i % NS];
        if ((rc = enqueue_tail(i))) return bail(rc);
        if (i + 1 < N && (rc = enqueue_bulk(i + 1))) return bail(rc);
        if (i + 2 < N && (rc = copy_in(i + 2))) return bail(rc);
        if ((rc = finish(i))) return bail(rc);
        if (trace) fprintf(stderr, "[mldsa trace] chunk %zu (%zu lanes): signed at %.2f ms, %d rounds\n", i, m, now_ms(), run[i % NS].rounds);
        mark(i, TR_OUT0, cx->s_out);
        if (cudaMemcpyAsync(sig + p.sig * lo, s.sig, m * p.sig, cudaMemcpyDeviceToHost, cx->s_out) != cudaSuccess ||
            (iters && cudaMemcpyAsync(iters + lo, s.it, m * 4, cudaMemcpyDeviceToHost, cx->s_out) != cudaSuccess) ||
            cudaEventRecord(cx->ev_out[i % NS], cx->s_out) != cudaSuccess)
            return bail(cuda_fail(cudaGetLastError(), "copy out"));
        mark(i, TR_OUT1, cx->s_out);
    }
    cudaError_t e = cudaStreamSynchronize(cx->s_out);
    if (e == cudaSuccess) e = cudaStreamSynchronize(T);
    if (e == cudaSuccess) e = cudaStreamSynchronize(B);
    if (e != cudaSuccess) return bail(cuda_fail(e, "sign pipeline"));
    prof_collect();
    if (trace) {
        fprintf(stderr, "[mldsa trace] all signatures on the host at %.2f ms\n", now_ms());
        fprintf(stderr, "[mldsa trace] chunk   lanes |  H2D start..end | bulk start..end | tail end | D2H start..end   (device ms)\n");
        for (size_t i = 0; i < N; i++) {
            float t[TR_N];
            for (int k = 0; k < TR_N; k++) cudaEventElapsedTime(&t[k], tr_origin, tr[i * TR_N + k]);
            fprintf(stderr, "[mldsa trace] %5zu %7zu | %7.2f %7.2f | %7.2f %7.2f | %7.2f | %7.2f %7.2f\n", i, pl.lo[i + 1] - pl.lo[i],
                    t[TR_IN0], t[TR_IN1], t[TR_BULK0], t[TR_BULK1], t[TR_TAIL1], t[TR_OUT0], t[TR_OUT1]);
        }
        for (auto& e : tr) cudaEventDestroy(e);
        cudaEventDestroy(tr_origin);
        cudaGetLastError();
    }
    return OK;
}

// ks != nullptr: the keys are a resident public-key set; lane i verifies under key key_idx[i] (nullptr: key i)
static int verify_host(int ps, int dev, int framing, int ph, size_t n, const uint8_t* pk, size_t pk_stride,
                       const uint8_t* msgs, const uint64_t* msg_off, const uint8_t* ctx, const uint32_t* ctx_off,
                       const uint8_t* sig, uint8_t* ok, const KeySet* ks = nullptr, const uint32_t* key_idx = nullptr)
{
    PSInfo p;
    if (!ps_info(ps, p) || (framing != 0 && framing != 1) || ph > 2) return E_PARAM;
    if (ph >= 0) framing = 2;
    if (!ks && pk_stride != 0 && pk_stride != p.pk) return E_PARAM;
    if (n && ((!ks && !pk) || !sig || !ok || !check_msgs(n, msgs, msg_off, ctx, ctx_off))) return E_PARAM;
    if (n == 0) return OK;
    if (ks) {
        if (ks->ps != ps || ks->dev != dev || !ks->is_public) return E_PARAM;
        if (!key_idx && n > ks->n) return E_PARAM;
        if (key_idx)
            for (size_t i = 0; i < n; i++)
                if (key_idx[i] >= ks->n) return E_PARAM;
        pk_stride = p.pk;
    }
    std::unique_lock<std::mutex> lk;
    int rc = OK;
    DeviceCtx* cx = lock_ctx(dev, lk, rc);
    if (!cx) return rc;
    // Verify has no data-dependent control flow on the host, so the whole pipeline is enqueued without a single host
    // wait: NS staging sets, copy-in on s_in, kernels on the work stream (one workspace: the chunks' kernels are
    // stream-ordered), verdicts out on s_out, ordered by events.  The packed pk + sig bytes make this path PCIe-bound
    // (5.3 KB per lane against 56 ns of compute), so what matters is that the copy-in stream never idles.
    constexpr size_t NS = SIGN_SETS;
    const bool resident = ks != nullptr, shared = !resident && pk_stride == 0;
    const size_t pk_lane = resident ? (key_idx ? 4 : 0) : p.pk;  // staged bytes per lane: packed pk, or a key index
    const size_t avg_msg = (size_t)((msg_off[n] - msg_off[0]) / n) + 1;
    const size_t avg_ctx = ctx_off ? (size_t)((ctx_off[n] - ctx_off[0]) / n) + 1 : 0;
    const size_t fixed = verify_ws_bytes(p, 0) + (shared ? NS * pad256(p.pk) : 0) + 65536;
    const size_t ws_lane = (verify_ws_bytes(p, 1024) - verify_ws_bytes(p, 0)) / 1024 + 1;
    const size_t io_lane = NS * ((shared ? 0 : pk_lane) + p.sig + 1 + 8 + 4 + avg_msg + avg_ctx + PREHASH_STRIDE + 8);
    const ChunkPlan pl = plan_chunks(cx->arena_bytes, n, fixed, ws_lane + io_lane, msg_off, ctx_off);
    const size_t C = pl.max_lanes, N = pl.nchunks();
    struct Set {
        uint8_t *pk, *sig, *ok, *m, *c, *pm;
        uint32_t* co;
        uint64_t *mo, *po;
    } set[NS];
    const size_t pk_set = shared ? p.pk : C * pk_lane + 16;
    const size_t set_bytes = pad256(pk_set) + pad256(C * p.sig) + pad256(C) + pad256((C + 1) * 8) +
                             pad256((C + 1) * 4) + pad256(pl.max_mbytes + 8) + pad256(pl.max_cbytes + 8) +
                             pad256(C * PREHASH_STRIDE) + pad256((C + 1) * 8);
    rc = cx->reserve(NS * set_bytes + verify_ws_bytes(p, C) + 4096);
    if (rc) return rc;
    Carver c(cx->arena);
    for (auto& s : set) {
        s.pk = c.take<uint8_t>(pk_set);  // packed public keys -- or, for a resident key set, the lanes' key indices
        s.sig = c.take<uint8_t>(C * p.sig);
        s.ok = c.take<uint8_t>(C);
        s.mo = c.take<uint64_t>(C + 1);
        s.co = c.take<uint32_t>(C + 1);
        s.m = c.take<uint8_t>(pl.max_mbytes + 8);
        s.c = c.take<uint8_t>(pl.max_cbytes + 8);
        s.pm = c.take<uint8_t>(C * PREHASH_STRIDE);
        s.po = c.take<uint64_t>(C + 1);
    }
    uint8_t* const ws_base = cx->arena + c.off;
    cudaStream_t st = cx->stream;
    cudaEvent_t* ev_cmp = cx->ev_bulk;  // "chunk computed"
    auto drain = [&](int code) -> int {
        const std::string err = t_cuda_err;
        cudaStreamSynchronize(cx->s_in);
        cudaStreamSynchronize(st);
        cudaStreamSynchronize(cx->s_out);
        cudaGetLastError();
        prof_collect();
        t_cuda_err = err;
        return code;
    };
    auto enqueue = [&](size_t i) -> int {
        const size_t lo = pl.lo[i], m = pl.lo[i + 1] - lo;
        Set& s = set[i % NS];
        const size_t mbytes = (size_t)(msg_off[lo + m] - msg_off[lo]);
        const size_t cbytes = ctx_off ? (size_t)(ctx_off[lo + m] - ctx_off[lo]) : 0;
        // copy-in
        if (i >= NS) CK(cudaStreamWaitEvent(cx->s_in, ev_cmp[i % NS], 0));  // the set's previous chunk is computed
        if (!resident)
            CK(cudaMemcpyAsync(s.pk, pk + pk_stride * lo, shared ? p.pk : m * p.pk, cudaMemcpyHostToDevice, cx->s_in));
        else if (key_idx)
            CK(cudaMemcpyAsync(s.pk, key_idx + lo, m * 4, cudaMemcpyHostToDevice, cx->s_in));
        CK(cudaMemcpyAsync(s.sig, sig + p.sig * lo, m * p.sig, cudaMemcpyHostToDevice, cx->s_in));
        CK(cudaMemcpyAsync(s.mo, msg_off + lo, (m + 1) * 8, cudaMemcpyHostToDevice, cx->s_in));
        if (mbytes) CK(cudaMemcpyAsync(s.m, msgs + msg_off[lo], mbytes, cudaMemcpyHostToDevice, cx->s_in));
        if (ctx_off) {
            CK(cudaMemcpyAsync(s.co, ctx_off + lo, (m + 1) * 4, cudaMemcpyHostToDevice, cx->s_in));
            if (cbytes) CK(cudaMemcpyAsync(s.c, ctx + ctx_off[lo], cbytes, cudaMemcpyHostToDevice, cx->s_in));
        }
        CK(cudaEventRecord(cx->ev_in[i % NS], cx->s_in));
        // kernels
        CK(cudaStreamWaitEvent(st, cx->ev_in[i % NS], 0));
        if (i >= NS) CK(cudaStreamWaitEvent(st, cx->ev_out[i % NS], 0));  // ... and its verdicts are on the host
        const uint8_t* cm = s.m - msg_off[lo];
        const uint64_t* cmo = s.mo;
        if (ph >= 0) {  // HashML-DSA (src/lib.rs:391-411)
            CK(launch_prehash(st, ph, m, cm, cmo, s.pm, s.po));
            cm = s.pm;
            cmo = s.po;
        }
        Carver wc(ws_base);
        VerifyWs ws = verify_carve(wc, p, m);
        if (resident) {  // lanes [lo, lo + m) use keys key_idx[lo ..] -- or keys lo .. lo + m - 1 of the set
            const size_t k0 = key_idx ? 0 : lo;
            ws.keys_resident = true;
            ws.kidx = key_idx ? reinterpret_cast<const uint32_t*>(s.pk) : nullptr;
            ws.A = ks->A + k0 * (size_t)p.K * p.L * 256;
            ws.t1h = ks->t1h + k0 * (size_t)p.K * 256;
            ws.tr = ks->tr + k0 * 64;
        }
        const MsgDesc md{cm, cmo, ctx_off ? s.c - ctx_off[lo] : nullptr, ctx_off ? s.co : nullptr, framing};
        CK(verify_run(ps, st, m, s.pk, pk_stride, md, s.sig, s.ok, ws));
        CK(cudaEventRecord(ev_cmp[i % NS], st));
        // verdicts out
        CK(cudaStreamWaitEvent(cx->s_out, ev_cmp[i % NS], 0));
        CK(cudaMemcpyAsync(ok + lo, s.ok, m, cudaMemcpyDeviceToHost, cx->s_out));
        CK(cudaEventRecord(cx->ev_out[i % NS], cx->s_out));
        return OK;
    };
    CK(cudaStreamSynchronize(cx->s_in));
    CK(cudaStreamSynchronize(cx->s_out));
    for (size_t i = 0; i < N; i++)
        if ((rc = enqueue(i))) return drain(rc);
    cudaError_t e = cudaStreamSynchronize(cx->s_out);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return drain(cuda_fail(e, "verify pipeline"));
    prof_collect();
    return OK;
}

extern "C" {

int mldsa_sign_batch(int ps, int dev, int framing, size_t n, const uint8_t* sk, size_t sk_stride, const uint8_t* msgs,
                     const uint64_t* msg_off, const uint8_t* ctx, const uint32_t* ctx_off, const uint8_t* rnd,
                     uint8_t* sig, uint32_t* iters)
{
    return sign_host(ps, dev, framing, -1, n, sk, sk_stride, msgs, msg_off, ctx, ctx_off, rnd, sig, iters);
}
int mldsa_verify_batch(int ps, int dev, int framing, size_t n, const uint8_t* pk, size_t pk_stride,
                       const uint8_t* msgs, const uint64_t* msg_off, const uint8_t* ctx, const uint32_t* ctx_off,
                       const uint8_t* sig, uint8_t* ok)
{
    return verify_host(ps, dev, framing, -1, n, pk, pk_stride, msgs, msg_off, ctx, ctx_off, sig, ok);
}
int mldsa_hash_sign_batch(int ps, int dev, int ph, size_t n, const uint8_t* sk, size_t sk_stride, const uint8_t* msgs,
                          const uint64_t* msg_off, const uint8_t* ctx, const uint32_t* ctx_off, const uint8_t* rnd,
                          uint8_t* sig, uint32_t* iters)
{
    if (ph < 0 || ph > 2) return E_PARAM;
    return sign_host(ps, dev, 0, ph, n, sk, sk_stride, msgs, msg_off, ctx, ctx_off, rnd, sig, iters);
}
int mldsa_hash_verify_batch(int ps, int dev, int ph, size_t n, const uint8_t* pk, size_t pk_stride,
                            const uint8_t* msgs, const uint64_t* msg_off, const uint8_t* ctx, const uint32_t* ctx_off,
                            const uint8_t* sig, uint8_t* ok)
{
    if (ph < 0 || ph > 2) return E_PARAM;
    return verify_host(ps, dev, 0, ph, n, pk, pk_stride, msgs, msg_off, ctx, ctx_off, sig, ok);
}

// ---------------------------------------------------------------- resident key sets
// PrivateKey::try_from_bytes + expand_private (src/lib.rs:462-468, src/ml_dsa.rs:445-469) for n keys at once, kept on
// the device: later sign calls skip the H2D copy of the packed keys, sk_decode, the L + 2K forward NTTs and ExpandA.
int mldsa_sk_load(int ps, int dev, size_t n, const uint8_t* sk, mldsa_keyset** out)
{
    PSInfo p;
    if (!ps_info(ps, p) || !out || n == 0 || !sk) return E_PARAM;
    *out = nullptr;
    std::unique_lock<std::mutex> lk;
    int rc = OK;
    DeviceCtx* cx = lock_ctx(dev, lk, rc);
    if (!cx) return rc;
    KeySet* ks = new (std::nothrow) KeySet;
    if (!ks) return E_NOMEM;
    ks->ps = ps;
    ks->dev = dev;
    ks->n = n;
    const size_t a_b = pad256(n * (size_t)p.K * p.L * 1024), sh_b = pad256(n * (size_t)(2 * p.K + p.L) * 1024), mini_b = pad256(n * 128);
    ks->bytes = a_b + sh_b + mini_b;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ks->base), ks->bytes);
    if (e != cudaSuccess) {
        delete ks;
        return cuda_fail(e, "key set allocation");
    }
    Carver c(ks->base);
    ks->A = c.take<int32_t>(n * (size_t)p.K * p.L * 256);
    ks->sh = c.take<int32_t>(n * (size_t)(2 * p.K + p.L) * 256);
    ks->mini = c.take<uint8_t>(n * 128);
    // packed keys pass through the arena in chunks (two staging buffers: copy-in of chunk i+1 under the kernels of i)
    size_t chunk = pick_chunk(cx->arena_bytes, n, 65536, 2 * p.sk);
    if (chunk > (size_t(1) << 17)) chunk = size_t(1) << 17;
    rc = cx->reserve(2 * pad256(chunk * p.sk));
    cudaStream_t st = cx->stream;
    auto fail = [&](int code) {
        cudaStreamSynchronize(cx->s_in);
        cudaStreamSynchronize(st);
        cudaMemset(ks->base, 0, ks->bytes);
        cudaFree(ks->base);
        if (cx->arena) cudaMemset(cx->arena, 0, 2 * pad256(chunk * p.sk) < cx->arena_bytes ? 2 * pad256(chunk * p.sk) : cx->arena_bytes);
        cudaGetLastError();
        delete ks;
        return code;
    };
    if (rc) return fail(rc);
    uint8_t* stage[2] = {cx->arena, cx->arena + pad256(chunk * p.sk)};
    cx->clear_secrets();
    cx->note_secret(stage[0], chunk < n ? chunk * p.sk : n * p.sk);
    if (n > chunk) cx->note_secret(stage[1], (n - chunk < chunk ? n - chunk : chunk) * p.sk);
    size_t i = 0;
    for (size_t lo = 0; lo < n; lo += chunk, i++) {
        const size_t m = n - lo < chunk ? n - lo : chunk;
        uint8_t* d_sk = stage[i & 1];
        if (i >= 2 && cudaStreamWaitEvent(cx->s_in, cx->ev_bulk[i & 1], 0) != cudaSuccess) return fail(E_CUDA);
        e = cudaMemcpyAsync(d_sk, sk + p.sk * lo, m * p.sk, cudaMemcpyHostToDevice, cx->s_in);
        if (e == cudaSuccess) e = cudaEventRecord(cx->ev_in[i & 1], cx->s_in);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, cx->ev_in[i & 1], 0);
        if (e == cudaSuccess)
            e = sign_expand_keys(ps, st, m, d_sk, ks->A + lo * (size_t)p.K * p.L * 256,
                                 ks->sh + lo * (size_t)(2 * p.K + p.L) * 256);
        if (e == cudaSuccess)  // rho || K || tr: the first 128 bytes of every packed key (src/encodings.rs:168-176)
            e = cudaMemcpy2DAsync(ks->mini + 128 * lo, 128, d_sk, p.sk, 128, m, cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_sk, 0, m * p.sk, st);
        if (e == cudaSuccess) e = cudaEventRecord(cx->ev_bulk[i & 1], st);
        if (e != cudaSuccess) return fail(cuda_fail(e, "key set load"));
    }
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(cuda_fail(e, "key set load"));
    prof_collect();
    *out = reinterpret_cast<mldsa_keyset*>(ks);
    return OK;
}

static KeySet* as_keyset(const mldsa_keyset* h);
// HashML-DSA on expanded keys (src/lib.rs:310-342, :391-411): ph = 0 / 1 / 2 = SHA-256 / SHA-512 / SHAKE128 pre-hash
int mldsa_hash_sign_batch_keys(const mldsa_keyset* h, int ph, size_t n, const uint32_t* key_idx, const uint8_t* msgs,
                               const uint64_t* msg_off, const uint8_t* ctx, const uint32_t* ctx_off, const uint8_t* rnd,
                               uint8_t* sig, uint32_t* iters)
{
    const KeySet* ks = as_keyset(h);
    if (!ks || ph < 0 || ph > 2) return E_PARAM;
    return sign_host(ks->ps, ks->dev, 0, ph, n, nullptr, 128, msgs, msg_off, ctx, ctx_off, rnd, sig, iters, ks, key_idx);
}
int mldsa_hash_verify_batch_keys(const mldsa_keyset* h, int ph, size_t n, const uint32_t* key_idx, const uint8_t* msgs,
                                 const uint64_t* msg_off, const uint8_t* ctx, const uint32_t* ctx_off, const uint8_t* sig,
                                 uint8_t* ok)
{
    const KeySet* ks = as_keyset(h);
    if (!ks || ph < 0 || ph > 2) return E_PARAM;
    return verify_host(ks->ps, ks->dev, 0, ph, n, nullptr, 0, msgs, msg_off, ctx, ctx_off, sig, ok, ks, key_idx);
}

static KeySet* as_keyset(const mldsa_keyset* h)
{
    KeySet* ks = reinterpret_cast<KeySet*>(const_cast<mldsa_keyset*>(h));
    return ks && ks->magic == 0x4B534554 ? ks : nullptr;
}

size_t mldsa_keyset_size(const mldsa_keyset* h)
{
    const KeySet* ks = as_keyset(h);
    return ks ? ks->n : 0;
}
size_t mldsa_keyset_device_bytes(const mldsa_keyset* h)
{
    const KeySet* ks = as_keyset(h);
    return ks ? ks->bytes : 0;
}

// Zeroize + ZeroizeOnDrop of the expanded keys (src/types.rs:19)
void mldsa_keyset_free(mldsa_keyset* h)
{
    KeySet* ks = as_keyset(h);
    if (!ks) return;
    std::unique_lock<std::mutex> lk;
    int rc = OK;
    if (lock_ctx(ks->dev, lk, rc)) {
        cudaMemset(ks->base, 0, ks->bytes);
        cudaFree(ks->base);
        cudaGetLastError();
    }
    ks->magic = 0;
    delete ks;
}

// Signer::try_sign_with_rng on an expanded PrivateKey (src/lib.rs:268-296): lane i signs msgs[i] with key key_idx[i]
// of the set (key_idx == NULL: key i; then n <= number of keys)
int mldsa_sign_batch_keys(const mldsa_keyset* h, int framing, size_t n, const uint32_t* key_idx, const uint8_t* msgs,
                          const uint64_t* msg_off, const uint8_t* ctx, const uint32_t* ctx_off, const uint8_t* rnd,
                          uint8_t* sig, uint32_t* iters)
{
    const KeySet* ks = as_keyset(h);
    if (!ks) return E_PARAM;
    return sign_host(ks->ps, ks->dev, framing, -1, n, nullptr, 128, msgs, msg_off, ctx, ctx_off, rnd, sig, iters, ks, key_idx);
}

// PublicKey::try_from_bytes + expand_public (src/lib.rs:469-474, src/ml_dsa.rs:477-498) for n keys, kept on the device:
// rho-derived A-hat, tr = H(pk) and NTT(t1 * 2^d) -- the fields of the reference's PublicKey (src/types.rs:35-41) plus
// A-hat.  Later verify calls send only signatures and messages.
int mldsa_pk_load(int ps, int dev, size_t n, const uint8_t* pk, mldsa_keyset** out)
{
    PSInfo p;
    if (!ps_info(ps, p) || !out || n == 0 || !pk) return E_PARAM;
    *out = nullptr;
    std::unique_lock<std::mutex> lk;
    int rc = OK;
    DeviceCtx* cx = lock_ctx(dev, lk, rc);
    if (!cx) return rc;
    KeySet* ks = new (std::nothrow) KeySet;
    if (!ks) return E_NOMEM;
    ks->ps = ps;
    ks->dev = dev;
    ks->n = n;
    ks->is_public = true;
    ks->bytes = pad256(n * (size_t)p.K * p.L * 1024) + pad256(n * (size_t)p.K * 1024) + pad256(n * 64);
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ks->base), ks->bytes);
    if (e != cudaSuccess) {
        delete ks;
        return cuda_fail(e, "key set allocation");
    }
    Carver c(ks->base);
    ks->A = c.take<int32_t>(n * (size_t)p.K * p.L * 256);
    ks->t1h = c.take<int32_t>(n * (size_t)p.K * 256);
    ks->tr = c.take<uint8_t>(n * 64);
    size_t chunk = pick_chunk(cx->arena_bytes, n, 65536, 2 * p.pk);
    if (chunk > (size_t(1) << 17)) chunk = size_t(1) << 17;
    rc = cx->reserve(2 * pad256(chunk * p.pk));
    cudaStream_t st = cx->stream;
    auto fail = [&](int code) {
        cudaStreamSynchronize(cx->s_in);
        cudaStreamSynchronize(st);
        cudaFree(ks->base);
        cudaGetLastError();
        delete ks;
        return code;
    };
    if (rc) return fail(rc);
    uint8_t* stage[2] = {cx->arena, cx->arena + pad256(chunk * p.pk)};
    size_t i = 0;
    for (size_t lo = 0; lo < n; lo += chunk, i++) {
        const size_t m = n - lo < chunk ? n - lo : chunk;
        uint8_t* d_pk = stage[i & 1];
        if (i >= 2 && cudaStreamWaitEvent(cx->s_in, cx->ev_bulk[i & 1], 0) != cudaSuccess) return fail(E_CUDA);
        e = cudaMemcpyAsync(d_pk, pk + p.pk * lo, m * p.pk, cudaMemcpyHostToDevice, cx->s_in);
        if (e == cudaSuccess) e = cudaEventRecord(cx->ev_in[i & 1], cx->s_in);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, cx->ev_in[i & 1], 0);
        if (e == cudaSuccess)
            e = verify_expand_keys(ps, st, m, d_pk, ks->A + lo * (size_t)p.K * p.L * 256, ks->t1h + lo * (size_t)p.K * 256,
                                   ks->tr + lo * 64);
        if (e == cudaSuccess) e = cudaEventRecord(cx->ev_bulk[i & 1], st);
        if (e != cudaSuccess) return fail(cuda_fail(e, "key set load"));
    }
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(cuda_fail(e, "key set load"));
    prof_collect();
    *out = reinterpret_cast<mldsa_keyset*>(ks);
    return OK;
}

// Verifier::verify on expanded PublicKeys (src/lib.rs:364-380): lane i is checked under key key_idx[i] of the set
// (key_idx == NULL: key i; then n <= number of keys)
int mldsa_verify_batch_keys(const mldsa_keyset* h, int framing, size_t n, const uint32_t* key_idx, const uint8_t* msgs,
                            const uint64_t* msg_off, const uint8_t* ctx, const uint32_t* ctx_off, const uint8_t* sig,
                            uint8_t* ok)
{
    const KeySet* ks = as_keyset(h);
    if (!ks) return E_PARAM;
    return verify_host(ks->ps, ks->dev, framing, -1, n, nullptr, 0, msgs, msg_off, ctx, ctx_off, sig, ok, ks, key_idx);
}

}  // extern "C"

// ---------------------------------------------------------------- device-side key provisioning and the chained pipeline
static KeySet* keyset_new(const PSInfo& p, int ps, int dev, size_t n, bool is_public, int* rc)
{
    KeySet* ks = new (std::nothrow) KeySet;
    if (!ks) {
        *rc = E_NOMEM;
        return nullptr;
    }
    ks->ps = ps;
    ks->dev = dev;
    ks->n = n;
    ks->is_public = is_public;
    const size_t a_b = pad256(n * (size_t)p.K * p.L * 1024), k_b = pad256(n * (size_t)p.K * 1024);
    ks->bytes = is_public ? a_b + k_b + pad256(n * 64) : a_b + pad256(n * (size_t)(2 * p.K + p.L) * 1024) + pad256(n * 128);
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ks->base), ks->bytes);
    if (e != cudaSuccess) {
        delete ks;
        *rc = cuda_fail(e, "key set allocation");
        return nullptr;
    }
    Carver c(ks->base);
    ks->A = c.take<int32_t>(n * (size_t)p.K * p.L * 256);
    if (is_public) {
        ks->t1h = c.take<int32_t>(n * (size_t)p.K * 256);
        ks->tr = c.take<uint8_t>(n * 64);
    } else {
        ks->sh = c.take<int32_t>(n * (size_t)(2 * p.K + p.L) * 256);
        ks->mini = c.take<uint8_t>(n * 128);
    }
    return ks;
}
static void keyset_drop(KeySet* ks)  // zeroize + free (the caller holds the device's lock)
{
    if (!ks) return;
    cudaMemset(ks->base, 0, ks->bytes);
    cudaFree(ks->base);
    cudaGetLastError();
    ks->magic = 0;
    delete ks;
}

extern "C" {

// KG::keygen_from_seed (src/lib.rs:247-251, src/ml_dsa.rs:57-134) straight into EXPANDED keys that stay on the device
// (expand_private / expand_public, src/ml_dsa.rs:445-498): a signing service provisions its keys from seeds and never
// moves a packed secret key over PCIe.  Only xi goes in; the packed public keys come back if `pk` is given.
int mldsa_keygen_to_keyset(int ps, int dev, size_t n, const uint8_t* xi, uint8_t* pk, mldsa_keyset** sk_set,
                           mldsa_keyset** pk_set)
{
    PSInfo p;
    if (!ps_info(ps, p) || n == 0 || !xi || (!sk_set && !pk_set)) return E_PARAM;
    if (sk_set) *sk_set = nullptr;
    if (pk_set) *pk_set = nullptr;
    std::unique_lock<std::mutex> lk;
    int rc = OK;
    DeviceCtx* cx = lock_ctx(dev, lk, rc);
    if (!cx) return rc;
    KeySet* sks = sk_set ? keyset_new(p, ps, dev, n, false, &rc) : nullptr;
    if (rc) return rc;
    KeySet* pks = pk_set ? keyset_new(p, ps, dev, n, true, &rc) : nullptr;
    if (rc) {
        keyset_drop(sks);
        return rc;
    }
    constexpr size_t NS = 2;
    const size_t KL = (size_t)p.K * p.L, KpL = (size_t)(p.K + p.L);
    size_t chunk = pick_chunk(cx->arena_bytes, n, 65536, NS * (32 + p.pk + p.sk) + 128 + KpL * 256 + 64);
    if (chunk > (size_t(1) << 17)) chunk = size_t(1) << 17;  // the pk copy-out of chunk i runs under the kernels of i + 1
    const size_t set_bytes = pad256(chunk * 32) + pad256(chunk * p.pk) + pad256(chunk * p.sk);
    struct Set {
        uint8_t *xi, *pk, *sk;
    } set[NS];
    uint8_t* seeds = nullptr;
    int8_t* sbuf = nullptr;
    cudaStream_t st = cx->stream;
    auto fail = [&](int code) {
        const std::string err = t_cuda_err;
        cudaStreamSynchronize(st);
        cudaStreamSynchronize(cx->s_out);
        if (cx->arena && seeds) cudaMemset(cx->arena, 0, NS * set_bytes + pad256(chunk * 128) + pad256(chunk * KpL * 256));
        keyset_drop(sks);
        keyset_drop(pks);
        cudaGetLastError();
        t_cuda_err = err;
        return code;
    };
    rc = cx->reserve(NS * set_bytes + pad256(chunk * 128) + pad256(chunk * KpL * 256));
    if (rc) return fail(rc);
    Carver c(cx->arena);
    for (auto& s : set) {
        s.xi = c.take<uint8_t>(chunk * 32);
        s.pk = c.take<uint8_t>(chunk * p.pk);
        s.sk = c.take<uint8_t>(chunk * p.sk);
    }
    seeds = c.take<uint8_t>(chunk * 128);
    sbuf = c.take<int8_t>(chunk * KpL * 256);
    cx->clear_secrets();
    const size_t first = n < chunk ? n : chunk;
    for (size_t k = 0; k < NS; k++) {
        const size_t used = k == 0 ? first : (n > chunk ? (n - chunk < chunk ? n - chunk : chunk) : 0);
        cx->note_secret(set[k].xi, used * 32);
        cx->note_secret(set[k].sk, used * p.sk);
    }
    cx->note_secret(seeds, first * 128);
    cx->note_secret(sbuf, first * KpL * 256);
    int32_t* const A_all = sks ? sks->A : pks->A;
    {
        const cudaError_t e0 = cudaStreamSynchronize(cx->s_out);
        if (e0 != cudaSuccess) return fail(cuda_fail(e0, "keygen_to_keyset"));
    }
    size_t i = 0;
    for (size_t lo = 0; lo < n; lo += chunk, i++) {
        const size_t m = n - lo < chunk ? n - lo : chunk;
        Set& s = set[i % NS];
        cudaError_t e = cudaSuccess;
        if (i >= NS) e = cudaStreamWaitEvent(st, cx->ev_out[i % NS], 0);  // the set's previous public keys are on the host
        if (e == cudaSuccess) e = cudaMemcpyAsync(s.xi, xi + 32 * lo, m * 32, cudaMemcpyHostToDevice, st);
        KeygenWs ws{seeds, sbuf, A_all + lo * KL * 256};  // A-hat is written once, straight into the key set
        if (e == cudaSuccess) e = keygen_run(ps, st, m, s.xi, s.pk, s.sk, ws);
        if (e == cudaSuccess && sks) {
            e = sign_expand_keys(ps, st, m, s.sk, nullptr, sks->sh + lo * (size_t)(2 * p.K + p.L) * 256, /*with_A*/ false);
            if (e == cudaSuccess)  // rho || K || tr rows (src/encodings.rs:168-176)
                e = cudaMemcpy2DAsync(sks->mini + 128 * lo, 128, s.sk, p.sk, 128, m, cudaMemcpyDeviceToDevice, st);
        }
        if (e == cudaSuccess && pks) {
            if (sks)
                e = cudaMemcpyAsync(pks->A + lo * KL * 256, sks->A + lo * KL * 256, m * KL * 1024, cudaMemcpyDeviceToDevice, st);
            if (e == cudaSuccess)
                e = verify_expand_keys(ps, st, m, s.pk, nullptr, pks->t1h + lo * (size_t)p.K * 256, nullptr, false, false);
            if (e == cudaSuccess)  // tr = H(pk) is bytes 64..127 of the secret key keygen just wrote (src/ml_dsa.rs:99-101)
                e = cudaMemcpy2DAsync(pks->tr + 64 * lo, 64, s.sk + 64, p.sk, 64, m, cudaMemcpyDeviceToDevice, st);
        }
        // xi, rho', K, s1, s2 and the packed secret key are secret: nothing of them outlives the chunk
        if (e == cudaSuccess) e = cudaMemsetAsync(s.xi, 0, m * 32, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(seeds, 0, m * 128, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(sbuf, 0, m * KpL * 256, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(s.sk, 0, m * p.sk, st);
        if (e == cudaSuccess) e = cudaEventRecord(cx->ev_bulk[i % NS], st);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(cx->s_out, cx->ev_bulk[i % NS], 0);
        if (e == cudaSuccess && pk) e = cudaMemcpyAsync(pk + p.pk * lo, s.pk, m * p.pk, cudaMemcpyDeviceToHost, cx->s_out);
        if (e == cudaSuccess) e = cudaEventRecord(cx->ev_out[i % NS], cx->s_out);
        if (e != cudaSuccess) return fail(cuda_fail(e, "keygen_to_keyset"));
    }
    cudaError_t e = cudaStreamSynchronize(cx->s_out);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(cuda_fail(e, "keygen_to_keyset"));
    prof_collect();
    if (sk_set) *sk_set = reinterpret_cast<mldsa_keyset*>(sks);
    if (pk_set) *pk_set = reinterpret_cast<mldsa_keyset*>(pks);
    return OK;
}

}  // extern "C"

// keygen -> sign -> verify chained on one device, keys and signatures staying in HBM between the stages (the smoke flow
// of src/lib.rs:503-551 as a batch).  A-hat is expanded ONCE per key, by keygen, and reused by sign and verify (the
// reference re-derives it in all three, src/ml_dsa.rs:78, :181, :360).  stage_ms (optional, 3 doubles): device time of the
// keygen / sign / verify stage summed over the chunks.
static int pipeline_host(int ps, int dev, int framing, size_t n, const uint8_t* xi, const uint8_t* msgs, const uint64_t* msg_off,
                         const uint8_t* ctx, const uint32_t* ctx_off, const uint8_t* rnd, uint8_t* pk, uint8_t* sk, uint8_t* sig,
                         uint8_t* ok, uint32_t* iters, double* stage_ms)
{
    PSInfo p;
    if (!ps_info(ps, p) || (framing != 0 && framing != 1)) return E_PARAM;
    if (n && (!xi || !rnd || !pk || !sig || !ok || !check_msgs(n, msgs, msg_off, ctx, ctx_off))) return E_PARAM;
    if (stage_ms) stage_ms[0] = stage_ms[1] = stage_ms[2] = 0;
    if (n == 0) return OK;
    if (ctx_off)
        for (size_t i = 0; i < n; i++)
            if (ctx_off[i + 1] - ctx_off[i] > 255) return E_CTX;  // src/lib.rs:274
    std::unique_lock<std::mutex> lk;
    int rc = OK;
    DeviceCtx* cx = lock_ctx(dev, lk, rc);
    if (!cx) return rc;
    constexpr size_t NS = 2;
    const size_t KL = (size_t)p.K * p.L, KpL = (size_t)(p.K + p.L);
    const size_t avg_msg = (size_t)((msg_off[n] - msg_off[0]) / n) + 1;
    const size_t avg_ctx = ctx_off ? (size_t)((ctx_off[n] - ctx_off[0]) / n) + 1 : 0;
    auto ws_bytes_for = [&](size_t C) {
        return pad256(C * 128) + pad256(C * KpL * 256) + pad256(C * KL * 1024) + pad256(C * (size_t)p.L * 1024) +
               2 * pad256(C * (size_t)p.K * 1024) + pad256(sign_ws_bytes(p, C, false, true)) + 4096 +
               verify_ws_bytes(p, C) + pad256(C * (size_t)p.K * 1024) + pad256(C * 64);
    };
    const size_t ws_lane = ws_bytes_for(1024) / 1024 + 1;
    const size_t io_lane = NS * (64 + p.pk + p.sk + p.sig + 1 + 4 + 8 + 4 + avg_msg + avg_ctx + 16);
    const size_t cap = pick_chunk(cx->arena_bytes, n, ws_bytes_for(0) + 65536, ws_lane + io_lane);
    // chunks of >= 32768 lanes (as efficient as large ones, profiles/round1/r06_e2e_timeline.md), at least ~8 per call so that
    // the copy-out of chunk i hides under the stages of chunk i + 1
    size_t C = n;
    if (n >= 65536) C = std::max<size_t>(32768, (n + 7) / 8);
    if (C > cap) C = cap;
    const size_t N = (n + C - 1) / C;
    size_t max_mbytes = 0, max_cbytes = 0;
    for (size_t lo = 0; lo < n; lo += C) {
        const size_t m = n - lo < C ? n - lo : C;
        max_mbytes = std::max(max_mbytes, (size_t)(msg_off[lo + m] - msg_off[lo]));
        if (ctx_off) max_cbytes = std::max(max_cbytes, (size_t)(ctx_off[lo + m] - ctx_off[lo]));
    }
    struct Set {
        uint8_t *xi, *rnd, *pk, *sk, *sig, *ok, *m, *c;
        uint32_t *it, *co;
        uint64_t* mo;
    } set[NS];
    const size_t set_bytes = 2 * pad256(C * 32) + pad256(C * p.pk) + pad256(C * p.sk) + pad256(C * p.sig) + pad256(C) +
                             pad256(C * 4) + pad256((C + 1) * 4) + pad256((C + 1) * 8) + pad256(max_mbytes + 8) +
                             pad256(max_cbytes + 8);
    rc = cx->reserve(NS * set_bytes + ws_bytes_for(C) + 4096);
    if (rc) return rc;
    Carver c(cx->arena);
    for (auto& s : set) {
        s.xi = c.take<uint8_t>(C * 32);
        s.rnd = c.take<uint8_t>(C * 32);
        s.pk = c.take<uint8_t>(C * p.pk);
        s.sk = c.take<uint8_t>(C * p.sk);
        s.sig = c.take<uint8_t>(C * p.sig);
        s.ok = c.take<uint8_t>(C);
        s.it = c.take<uint32_t>(C);
        s.co = c.take<uint32_t>(C + 1);
        s.mo = c.take<uint64_t>(C + 1);
        s.m = c.take<uint8_t>(max_mbytes + 8);
        s.c = c.take<uint8_t>(max_cbytes + 8);
    }
    uint8_t* const ws_base = cx->arena + c.off;
    KeygenWs kws;
    kws.seeds = c.take<uint8_t>(C * 128);
    kws.s = c.take<int8_t>(C * KpL * 256);
    kws.A = c.take<int32_t>(C * KL * 256);
    int32_t* sh = c.take<int32_t>(C * (size_t)(2 * p.K + p.L) * 256);
    SignWs sws = sign_carve(c, p, C, false, cx->mailbox, /*resident*/ true);
    sws.keys_resident = true;  // A-hat and the s-hat rows are this pipeline's own arrays, wiped after the stage
    sws.A = kws.A;
    sws.sh = sh;
    c.take<uint8_t>(4096);
    VerifyWs vws0 = verify_carve(c, p, C);
    int32_t* t1h = c.take<int32_t>(C * (size_t)p.K * 256);
    uint8_t* tr = c.take<uint8_t>(C * 64);
    uint8_t* const ws_end = cx->arena + c.off;
    cx->clear_secrets();
    cudaStream_t B = cx->stream;
    std::vector<cudaEvent_t> tev;
    if (stage_ms) {
        tev.resize(N * 4);
        for (auto& e : tev) cudaEventCreate(&e);
    }
    auto bail = [&](int code) -> int {  // stop everything, leave no secrets behind
        const std::string err = t_cuda_err;
        cudaStreamSynchronize(cx->s_in);
        cudaStreamSynchronize(B);
        cudaStreamSynchronize(cx->s_out);
        cudaMemsetAsync(ws_base, 0, (size_t)(ws_end - ws_base), B);
        for (auto& s : set) {
            cudaMemsetAsync(s.xi, 0, C * 32, B);
            cudaMemsetAsync(s.rnd, 0, C * 32, B);
            cudaMemsetAsync(s.sk, 0, C * p.sk, B);
            cudaMemsetAsync(s.sig, 0, C * p.sig, B);
        }
        cudaStreamSynchronize(B);
        for (auto& e : tev) cudaEventDestroy(e);
        cudaGetLastError();
        prof_collect();
        t_cuda_err = err;
        return code;
    };
    auto copy_in = [&](size_t i) -> int {
        const size_t lo = i * C, m = n - lo < C ? n - lo : C;
        Set& s = set[i % NS];
        const size_t mbytes = (size_t)(msg_off[lo + m] - msg_off[lo]);
        const size_t cbytes = ctx_off ? (size_t)(ctx_off[lo + m] - ctx_off[lo]) : 0;
        if (i >= NS) CK(cudaStreamWaitEvent(cx->s_in, cx->ev_bulk[i % NS], 0));  // the set's previous chunk is computed
        CK(cudaMemcpyAsync(s.xi, xi + 32 * lo, m * 32, cudaMemcpyHostToDevice, cx->s_in));
        CK(cudaMemcpyAsync(s.rnd, rnd + 32 * lo, m * 32, cudaMemcpyHostToDevice, cx->s_in));
        CK(cudaMemcpyAsync(s.mo, msg_off + lo, (m + 1) * 8, cudaMemcpyHostToDevice, cx->s_in));
        if (mbytes) CK(cudaMemcpyAsync(s.m, msgs + msg_off[lo], mbytes, cudaMemcpyHostToDevice, cx->s_in));
        if (ctx_off) {
            CK(cudaMemcpyAsync(s.co, ctx_off + lo, (m + 1) * 4, cudaMemcpyHostToDevice, cx->s_in));
            if (cbytes) CK(cudaMemcpyAsync(s.c, ctx + ctx_off[lo], cbytes, cudaMemcpyHostToDevice, cx->s_in));
        }
        CK(cudaEventRecord(cx->ev_in[i % NS], cx->s_in));
        return OK;
    };
    {
        cudaError_t e0 = cudaStreamSynchronize(cx->s_in);
        if (e0 == cudaSuccess) e0 = cudaStreamSynchronize(cx->s_out);
        if (e0 != cudaSuccess) return bail(cuda_fail(e0, "pipeline"));
    }
    t_last_rounds = 0;
    if ((rc = copy_in(0))) return bail(rc);
    for (size_t i = 0; i < N; i++) {
        const size_t lo = i * C, m = n - lo < C ? n - lo : C;
        Set& s = set[i % NS];
        auto mark = [&](int k) {
            if (stage_ms) cudaEventRecord(tev[i * 4 + k], B);
        };
        cudaError_t e = cudaStreamWaitEvent(B, cx->ev_in[i % NS], 0);
        if (e == cudaSuccess && i >= NS) e = cudaStreamWaitEvent(B, cx->ev_out[i % NS], 0);  // its previous outputs are on the host
        if (e != cudaSuccess) return bail(cuda_fail(e, "pipeline"));
        cx->begin_slot((int)(i % NS));
        cx->note_secret(s.xi, m * 32);
        cx->note_secret(s.rnd, m * 32);
        cx->note_secret(s.sk, m * p.sk);
        cx->note_secret(kws.seeds, m * 128);
        cx->note_secret(kws.s, m * KpL * 256);
        cx->note_secret(sh, m * (size_t)(2 * p.K + p.L) * 1024);
        cx->note_secret(sws.w, m * (size_t)sws.depth * p.K * 1024);
        cx->note_secret(sws.yraw, m * (size_t)sws.depth * p.L * 32 * p.zbits);
        cx->note_secret(sws.rhopp, m * 64);
        // ---- stage 1: keygen (src/ml_dsa.rs:57-134); A-hat stays in kws.A for the next two stages
        mark(0);
        e = keygen_run(ps, B, m, s.xi, s.pk, s.sk, kws);
        if (e == cudaSuccess) e = cudaMemsetAsync(s.xi, 0, m * 32, B);
        if (e == cudaSuccess) e = cudaMemsetAsync(kws.seeds, 0, m * 128, B);
        if (e == cudaSuccess) e = cudaMemsetAsync(kws.s, 0, m * KpL * 256, B);
        mark(1);
        // ---- stage 2: sign (src/ml_dsa.rs:153-337) with the packed keys where keygen left them
        if (e == cudaSuccess) e = sign_expand_keys(ps, B, m, s.sk, nullptr, sh, /*with_A*/ false);
        if (e != cudaSuccess) return bail(cuda_fail(e, "pipeline keygen"));
        const MsgDesc md{s.m - msg_off[lo], s.mo, ctx_off ? s.c - ctx_off[lo] : nullptr, ctx_off ? s.co : nullptr, framing};
        SignRun run;
        e = sign_begin(ps, B, m, s.sk, p.sk, md, s.rnd, s.sig, s.it, sws, 0, &run);
        if (e == cudaSuccess) e = sign_tail(B, &run);
        if (e != cudaSuccess) return bail(run.ctx_err ? E_CTX : cuda_fail(e, "pipeline sign"));
        if (i + 1 < N && (rc = copy_in(i + 1))) return bail(rc);  // the next chunk's inputs move while this one signs
        e = sign_finish(&run);  // the one host wait of a chunk
        if (e != cudaSuccess) return bail(run.ctx_err ? E_CTX : (run.incomplete ? E_UNSUPPORTED : cuda_fail(e, "pipeline sign")));
        t_last_rounds += run.rounds;
        e = cudaMemsetAsync(sh, 0, (size_t)((uint8_t*)sws.w - (uint8_t*)sh), B);  // s2-hat, s1-hat, t0-hat (contiguous)
        if (e == cudaSuccess) e = cudaMemsetAsync(s.rnd, 0, m * 32, B);
        mark(2);
        // ---- stage 3: verify (src/ml_dsa.rs:351-437) under the public keys keygen wrote
        if (e == cudaSuccess) e = verify_expand_keys(ps, B, m, s.pk, nullptr, t1h, nullptr, false, false);
        if (e == cudaSuccess) e = cudaMemcpy2DAsync(tr, 64, s.sk + 64, p.sk, 64, m, cudaMemcpyDeviceToDevice, B);  // tr = sk[64..128)
        if (e == cudaSuccess) {
            VerifyWs vws = vws0;
            vws.keys_resident = true;
            vws.kidx = nullptr;
            vws.A = kws.A;
            vws.t1h = t1h;
            vws.tr = tr;
            e = verify_run(ps, B, m, s.pk, p.pk, md, s.sig, s.ok, vws);
        }
        mark(3);
        if (e == cudaSuccess && !sk) e = cudaMemsetAsync(s.sk, 0, m * p.sk, B);
        if (e == cudaSuccess) e = cudaEventRecord(cx->ev_bulk[i % NS], B);
        // ---- results out: packed pk, sig, verdicts (and sk / trip counts when asked for)
        if (e == cudaSuccess) e = cudaStreamWaitEvent(cx->s_out, cx->ev_bulk[i % NS], 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(pk + p.pk * lo, s.pk, m * p.pk, cudaMemcpyDeviceToHost, cx->s_out);
        if (e == cudaSuccess) e = cudaMemcpyAsync(sig + p.sig * lo, s.sig, m * p.sig, cudaMemcpyDeviceToHost, cx->s_out);
        if (e == cudaSuccess) e = cudaMemcpyAsync(ok + lo, s.ok, m, cudaMemcpyDeviceToHost, cx->s_out);
        if (e == cudaSuccess && iters) e = cudaMemcpyAsync(iters + lo, s.it, m * 4, cudaMemcpyDeviceToHost, cx->s_out);
        if (e == cudaSuccess && sk) {
            e = cudaMemcpyAsync(sk + p.sk * lo, s.sk, m * p.sk, cudaMemcpyDeviceToHost, cx->s_out);
            if (e == cudaSuccess) e = cudaMemsetAsync(s.sk, 0, m * p.sk, cx->s_out);
        }
        if (e == cudaSuccess) e = cudaEventRecord(cx->ev_out[i % NS], cx->s_out);
        if (e != cudaSuccess) return bail(cuda_fail(e, "pipeline"));
        prof_collect();
    }
    cudaError_t e = cudaStreamSynchronize(cx->s_out);
    if (e == cudaSuccess) e = cudaStreamSynchronize(B);
    if (e != cudaSuccess) return bail(cuda_fail(e, "pipeline"));
    prof_collect();
    if (stage_ms) {
        for (size_t i = 0; i < N; i++)
            for (int k = 0; k < 3; k++) {
                float ms = 0;
                if (cudaEventElapsedTime(&ms, tev[i * 4 + k], tev[i * 4 + k + 1]) == cudaSuccess) stage_ms[k] += ms;
            }
        for (auto& ev : tev) cudaEventDestroy(ev);
        cudaGetLastError();
    }
    return OK;
}

extern "C" {

int mldsa_pipeline_batch(int ps, int dev, int framing, size_t n, const uint8_t* xi, const uint8_t* msgs, const uint64_t* msg_off,
                         const uint8_t* ctx, const uint32_t* ctx_off, const uint8_t* rnd, uint8_t* pk, uint8_t* sk, uint8_t* sig,
                         uint8_t* ok, uint32_t* iters, double* stage_ms)
{
    return pipeline_host(ps, dev, framing, n, xi, msgs, msg_off, ctx, ctx_off, rnd, pk, sk, sig, ok, iters, stage_ms);
}

// Signer::get_public_key (src/lib.rs:344-346 -> src/ml_dsa.rs:502-563): pk bytes from sk bytes
int mldsa_sk_to_pk_batch_dev(int ps, void* stream, size_t n, const uint8_t* d_sk, uint8_t* d_pk)
{
    PSInfo p;
    if (!ps_info(ps, p) || (n && (!d_sk || !d_pk))) return E_PARAM;
    if (n == 0) return OK;
    std::unique_lock<std::mutex> lk;
    int rc = OK;
    DeviceCtx* cx = lock_current(lk, rc);
    if (!cx) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : cx->stream;
    CK(launch_sk_to_pk_fused(ps, st, n, d_sk, d_pk));  // no scratch: A-hat is consumed as it is squeezed
    CK(cudaStreamSynchronize(st));
    prof_collect();
    return OK;
}
int mldsa_sk_to_pk_batch(int ps, int dev, size_t n, const uint8_t* sk, uint8_t* pk)
{
    PSInfo p;
    if (!ps_info(ps, p) || (n && (!sk || !pk))) return E_PARAM;
    if (n == 0) return OK;
    std::unique_lock<std::mutex> lk;
    int rc = OK;
    DeviceCtx* cx = lock_ctx(dev, lk, rc);
    if (!cx) return rc;
    cudaStream_t st = cx->stream;
    const size_t per_lane = p.sk + p.pk + 64;
    const size_t chunk = pick_chunk(cx->arena_bytes, n, 0, per_lane);
    rc = cx->reserve(pad256(chunk * p.sk) + pad256(chunk * p.pk));
    if (rc) return rc;
    Carver c(cx->arena);
    uint8_t* d_sk = c.take<uint8_t>(chunk * p.sk);
    uint8_t* d_pk = c.take<uint8_t>(chunk * p.pk);
    cx->clear_secrets();
    cx->note_secret(d_sk, (n < chunk ? n : chunk) * p.sk);
    auto wipe = [&](int code) {  // the staged secret keys never outlive the call, whatever happened
        const std::string err = t_cuda_err;
        cudaMemsetAsync(d_sk, 0, chunk * p.sk, st);
        cudaStreamSynchronize(st);
        cudaGetLastError();
        t_cuda_err = err;
        return code;
    };
    for (size_t lo = 0; lo < n; lo += chunk) {
        const size_t m = n - lo < chunk ? n - lo : chunk;
        cudaError_t e = cudaMemcpyAsync(d_sk, sk + p.sk * lo, m * p.sk, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = launch_sk_to_pk_fused(ps, st, m, d_sk, d_pk);
        if (e == cudaSuccess) e = cudaMemcpyAsync(pk + p.pk * lo, d_pk, m * p.pk, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_sk, 0, m * p.sk, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return wipe(cuda_fail(e, "sk_to_pk"));
    }
    prof_collect();
    return OK;
}

// ---------------------------------------------------------------- multi-GPU host entry points
}  // extern "C"
template <class Fn>
static int run_sharded(const int* devs, int ndev, size_t n, Fn&& fn)
{
    if (!devs || ndev <= 0) return E_PARAM;
    if (ndev == 1) return fn(devs[0], (size_t)0, n);
    std::vector<std::thread> th;
    std::vector<int> rcs((size_t)ndev, OK);
    for (int g = 0; g < ndev; g++) {
        const size_t lo = n * (size_t)g / (size_t)ndev, hi = n * (size_t)(g + 1) / (size_t)ndev;
        th.emplace_back([&, g, lo, hi] {
            rcs[(size_t)g] = hi > lo ? fn(devs[g], lo, hi - lo) : OK;
            prof_collect_now();  // this worker's event pairs die with the thread
        });
    }
    for (auto& t : th) t.join();
    for (int rc : rcs)
        if (rc) return rc;
    return OK;
}
extern "C" {

int mldsa_keygen_batch_multi(int ps, const int* devs, int ndev, size_t n, const uint8_t* xi, uint8_t* pk, uint8_t* sk)
{
    PSInfo p;
    if (!ps_info(ps, p)) return E_PARAM;
    return run_sharded(devs, ndev, n, [&](int dev, size_t lo, size_t m) {
        return mldsa_keygen_batch(ps, dev, m, xi + 32 * lo, pk + p.pk * lo, sk + p.sk * lo);
    });
}
int mldsa_sign_batch_multi(int ps, const int* devs, int ndev, int framing, size_t n, const uint8_t* sk,
                           size_t sk_stride, const uint8_t* msgs, const uint64_t* msg_off, const uint8_t* ctx,
                           const uint32_t* ctx_off, const uint8_t* rnd, uint8_t* sig, uint32_t* iters)
{
    PSInfo p;
    if (!ps_info(ps, p) || !msg_off) return E_PARAM;
    return run_sharded(devs, ndev, n, [&](int dev, size_t lo, size_t m) {
        return mldsa_sign_batch(ps, dev, framing, m, sk + sk_stride * lo, sk_stride, msgs, msg_off + lo, ctx,
                                ctx_off ? ctx_off + lo : nullptr, rnd + 32 * lo, sig + p.sig * lo,
                                iters ? iters + lo : nullptr);
    });
}
int mldsa_verify_batch_multi(int ps, const int* devs, int ndev, int framing, size_t n, const uint8_t* pk,
                             size_t pk_stride, const uint8_t* msgs, const uint64_t* msg_off, const uint8_t* ctx,
                             const uint32_t* ctx_off, const uint8_t* sig, uint8_t* ok)
{
    PSInfo p;
    if (!ps_info(ps, p) || !msg_off) return E_PARAM;
    return run_sharded(devs, ndev, n, [&](int dev, size_t lo, size_t m) {
        return mldsa_verify_batch(ps, dev, framing, m, pk + pk_stride * lo, pk_stride, msgs, msg_off + lo, ctx,
                                  ctx_off ? ctx_off + lo : nullptr, sig + p.sig * lo, ok + lo);
    });
}

// stage_ms: per stage, the slowest device's time
int mldsa_pipeline_batch_multi(int ps, const int* devs, int ndev, int framing, size_t n, const uint8_t* xi, const uint8_t* msgs,
                               const uint64_t* msg_off, const uint8_t* ctx, const uint32_t* ctx_off, const uint8_t* rnd,
                               uint8_t* pk, uint8_t* sk, uint8_t* sig, uint8_t* ok, uint32_t* iters, double* stage_ms)
{
    PSInfo p;
    if (!ps_info(ps, p) || !msg_off || ndev <= 0 || ndev > MAX_DEV) return E_PARAM;
    std::vector<double> ms((size_t)ndev * 3, 0.0);
    std::atomic<int> slot{0};
    const int rc = run_sharded(devs, ndev, n, [&](int dev, size_t lo, size_t m) {
        const int g = slot.fetch_add(1);
        return mldsa_pipeline_batch(ps, dev, framing, m, xi + 32 * lo, msgs, msg_off + lo, ctx, ctx_off ? ctx_off + lo : nullptr,
                                    rnd + 32 * lo, pk + p.pk * lo, sk ? sk + p.sk * lo : nullptr, sig + p.sig * lo, ok + lo,
                                    iters ? iters + lo : nullptr, stage_ms ? &ms[(size_t)g * 3] : nullptr);
    });
    if (stage_ms)
        for (int k = 0; k < 3; k++) {
            stage_ms[k] = 0;
            for (int g = 0; g < ndev; g++) stage_ms[k] = std::max(stage_ms[k], ms[(size_t)g * 3 + k]);
        }
    return rc;
}

static int micro_host(int dev, size_t n, const void* in, size_t in_bytes, void* out, size_t out_bytes, int which)
{
    std::unique_lock<std::mutex> lk;
    int rc = OK;
    DeviceCtx* cx = lock_ctx(dev, lk, rc);
    if (!cx) return rc;
    rc = cx->reserve(pad256(in_bytes) + pad256(out_bytes));
    if (rc) return rc;
    cudaStream_t st = cx->stream;
    Carver c(cx->arena);
    uint8_t* d_in = c.take<uint8_t>(in_bytes);
    uint8_t* d_out = which == 2 ? c.take<uint8_t>(out_bytes) : d_in;
    CK(cudaMemcpyAsync(d_in, in, in_bytes, cudaMemcpyHostToDevice, st));
    if (which == 0) CK(launch_ntt(st, reinterpret_cast<int32_t*>(d_in), n, false));
    if (which == 1) CK(launch_ntt(st, reinterpret_cast<int32_t*>(d_in), n, true));
    if (which == 2) CK(launch_expand_a_seed34(st, d_in, n, reinterpret_cast<int32_t*>(d_out)));
    CK(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    prof_collect();
    return OK;
}

int mldsa_ntt_batch(int dev, size_t n, int32_t* polys)
{
    if (n && !polys) return E_PARAM;
    return n ? micro_host(dev, n, polys, n * 1024, polys, n * 1024, 0) : OK;
}
int mldsa_intt_batch(int dev, size_t n, int32_t* polys)
{
    if (n && !polys) return E_PARAM;
    return n ? micro_host(dev, n, polys, n * 1024, polys, n * 1024, 1) : OK;
}
int mldsa_expand_a_batch(int dev, size_t n, const uint8_t* seed34, int32_t* out)
{
    if (n && (!seed34 || !out)) return E_PARAM;
    return n ? micro_host(dev, n, seed34, n * 34, out, n * 1024, 2) : OK;
}

// Test hooks for the rejection samplers' rare paths: the device parsers run over caller-supplied squeeze blocks.
// which = 0: RejNTTPoly (src/hashing.rs:111-146): nblocks x 168 bytes per polynomial -> out = n x 256 int32 (zero-filled
// first; coefficients the stream does not reach stay zero).  which = tau (39 / 49 / 60): SampleInBall (src/hashing.rs:
// 43-100): nblocks x 136 bytes per challenge -> out = n x 64 bytes (256 "non-zero" bits, then 256 "negative" bits).
int mldsa_debug_sampler_stream(int dev, int which, size_t n, const uint8_t* blocks, size_t nblocks, void* out)
{
    if (n == 0) return OK;
    if (!blocks || !out || nblocks == 0 || (which != 0 && which != 39 && which != 49 && which != 60)) return E_PARAM;
    std::unique_lock<std::mutex> lk;
    int rc = OK;
    DeviceCtx* cx = lock_ctx(dev, lk, rc);
    if (!cx) return rc;
    const size_t in_bytes = n * nblocks * (which == 0 ? 168 : SIB_BYTES), out_bytes = n * (which == 0 ? 1024 : CMASK_STRIDE);
    rc = cx->reserve(pad256(in_bytes) + pad256(out_bytes));
    if (rc) return rc;
    cudaStream_t st = cx->stream;
    Carver c(cx->arena);
    uint8_t* d_in = c.take<uint8_t>(in_bytes);
    uint8_t* d_out = c.take<uint8_t>(out_bytes);
    CK(cudaMemcpyAsync(d_in, blocks, in_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(d_out, 0, out_bytes, st));
    if (which == 0) CK(launch_rej_ntt_stream(st, d_in, nblocks, n, reinterpret_cast<int32_t*>(d_out)));
    else CK(launch_sib_stream(st, d_in, nblocks, n, which, d_out));
    CK(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return OK;
}

int mldsa_shake_batch(int dev, int bits, size_t n, const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len)
{
    if ((bits != 128 && bits != 256) || (n && (!out || (in_len && !in))) || out_len == 0 || out_len > (1u << 20))
        return E_PARAM;
    if (n == 0) return OK;
    std::unique_lock<std::mutex> lk;
    int rc = OK;
    DeviceCtx* cx = lock_ctx(dev, lk, rc);
    if (!cx) return rc;
    rc = cx->reserve(pad256(n * in_len + 8) + pad256(n * out_len));
    if (rc) return rc;
    cudaStream_t st = cx->stream;
    Carver c(cx->arena);
    uint8_t* d_in = c.take<uint8_t>(n * in_len + 8);
    uint8_t* d_out = c.take<uint8_t>(n * out_len);
    if (in_len) CK(cudaMemcpyAsync(d_in, in, n * in_len, cudaMemcpyHostToDevice, st));
    if (bits == 256) CK(launch_shake256_rows(st, d_in, in_len, in_len, n, d_out, out_len, (int)out_len));
    else CK(launch_shake128_rows(st, d_in, in_len, in_len, n, d_out, out_len, (int)out_len));
    CK(cudaMemcpyAsync(out, d_out, n * out_len, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return OK;
}

}  // extern "C"
