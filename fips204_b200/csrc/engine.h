// engine.h -- internal interface between the kernel translation units and the host runtime.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace mldsa {

// error codes of the C ABI (include/mldsa_b200.h)
enum : int { OK = 0, E_PARAM = -1, E_CTX = -2, E_CUDA = -3, E_NOMEM = -4, E_UNSUPPORTED = -5 };

constexpr int SIB_BYTES = 136;   // one SHAKE256 block of H(c~) for SampleInBall
constexpr int CMASK_STRIDE = 64;  // bytes per lane of the challenge masks
constexpr int MU_BYTES = 64;
constexpr int PREHASH_STRIDE = 80;  // OID (11) || PH(M) (<= 64), padded

// Kernel classes for launch telemetry (mldsa_kernel_launches) and the optional per-kernel device
// timing (mldsa_profile_*): CUDA events recorded around every launch on the launching stream.
enum KernelId : int {
    KID_EXPAND_A = 0, KID_NTT, KID_INTT, KID_SHAKE_ROWS, KID_CTILDE, KID_KG_SEED, KID_REJ_BOUNDED, KID_KG_ARITH,
    KID_VERIFY_PRE, KID_T1HAT, KID_VERIFY_ARITH, KID_SIGN_PREP, KID_SIGN_KEY, KID_MASK, KID_SIGN_W, KID_SIGN_Z,
    KID_COMPACT, KID_PROBE, KID_PREHASH, KID_SK2PK, KID_COUNT
};
struct KLaunch {  // scope one kernel launch: `{ KLaunch k(KID_X, st); kernel<<<...>>>(...); }`
    int id;
    cudaStream_t st;
    KLaunch(int id, cudaStream_t st);
    ~KLaunch();
};
void prof_collect();      // end of an entry point: bounds the queue of recorded event pairs (folding is lazy)
void prof_collect_now();  // fold finished event pairs into the per-kernel totals (after a stream sync)

// Launch tunables (mldsa_set_tuning / MLDSA_KCAP, MLDSA_ACAP): kcap / acap cap the CTAs per SM of the (grid-stride)
// Keccak / arithmetic kernels (0 = default).
struct Tuning {
    int kcap, acap;
    int depth;  // candidates tried per active lane and round of the sign loop (MLDSA_DEPTH, default 2)
    int tail_depth;  // sign: deepest speculation of the under-filled last rounds (MLDSA_TAILDEPTH, default 8; <= depth: off)
    int tail_slots;  // sign: candidate slots of the shared tail region (MLDSA_TAILSLOTS, default 32768)
    int max_rounds;  // test hook (mldsa_debug_set_max_rounds): stop a chunk's loop after this many rounds (0 = off)
};
Tuning tuning();
// CTAs per SM of the grid-stride arithmetic (warp-per-lane) kernels.  6-10 are resident at a time; with 16 the last
// wave ran a partly empty machine (measured, sign at 2^20 lanes: 16 -> 37.5 / 32.7 / 7.5 ms for sign_w / sign_z /
// sign_key, 120-400 -> 35.8 / 30.6 / 6.8 ms, one CTA per unit -> 36.3 / 34.5 / 6.8 ms).  MLDSA_ACAP overrides.
constexpr int ACAP_DEFAULT = 192;
inline int arith_cap() { const int a = tuning().acap; return a > 0 ? a : ACAP_DEFAULT; }
unsigned capped_grid(size_t want_ctas, int cap_per_sm);  // grid of a grid-stride kernel on the CURRENT device

// ---- k_common.cu ----
// Commitment hash.  Verify mode (ctilde_out == nullptr): ok[i] = pre_ok[i] && H(mu||w1')[0..CT) == sig_i[0..CT)
// (src/ml_dsa.rs:427-436).  Sign mode: ctilde_out[i] = H(mu||w1)[0..CT) and cmask_out[i] = SampleInBall(c~) as
// two 256-bit masks (src/ml_dsa.rs:231-237, src/hashing.rs:43-100).  act: optional list of lane ids (n entries), else lanes 0..n-1.
// n_dev: optional device-side lane count that overrides n inside the kernel (n then only sizes the grid).
// depth: rows per lane (sign: candidates per round; row = act[i / depth] * depth + i % depth), 1 for verify.
cudaError_t launch_ctilde(int ps, cudaStream_t st, size_t n, const uint32_t* n_dev, const uint32_t* act, int depth,
                          const uint8_t* mw1, const uint8_t* sig, const uint8_t* pre_ok, uint8_t* ok,
                          uint8_t* ctilde_out, uint8_t* cmask_out, size_t row_cap = ~size_t(0));

int keccak_probe_variants();
uint32_t keccak_probe_mask(int v);
cudaError_t launch_keccak_probe(cudaStream_t st, int variant, int blocks, int reps, uint64_t* state);
// HashML-DSA: out rows = OID || PH(M_i) with out_off[i] = i * PREHASH_STRIDE .. (message framing 2)
cudaError_t launch_prehash(cudaStream_t st, int ph, size_t n, const uint8_t* msgs, const uint64_t* msg_off,
                           uint8_t* out, uint64_t* out_off);
// *flag = 1 if any ctx_off[i + 1] - ctx_off[i] > 255 (src/lib.rs:274); flag may be mapped host memory
cudaError_t launch_ctx_check(cudaStream_t st, size_t n, const uint32_t* ctx_off, uint32_t* flag);
// *count += number of non-zero bytes in [p, p + bytes)   (test hook: secret residue scan)
cudaError_t launch_count_nonzero(cudaStream_t st, const uint8_t* p, size_t bytes, unsigned long long* count);
cudaError_t launch_int_probe(cudaStream_t st, int mode, int blocks, int iters, uint32_t* scratch, double* ops);
// A[key][r*L+s][256] = RejNTTPoly(rho_key || s || r)          (src/hashing.rs:225-239)
cudaError_t launch_expand_a(cudaStream_t st, int K, int L, const uint8_t* rho, size_t rho_stride, size_t nkeys,
                            int32_t* A);
// out[i][256] = RejNTTPoly(seed34[i])                          (src/hashing.rs:111-146)
cudaError_t launch_expand_a_seed34(cudaStream_t st, const uint8_t* seeds, size_t n, int32_t* out);
cudaError_t launch_ntt(cudaStream_t st, int32_t* polys, size_t n, bool inverse);
// test hooks: the rejection samplers' parsers over caller-supplied squeeze blocks (see k_common.cu)
cudaError_t launch_rej_ntt_stream(cudaStream_t st, const uint8_t* blocks, size_t nblocks, size_t n, int32_t* out);
cudaError_t launch_sib_stream(cudaStream_t st, const uint8_t* blocks, size_t nblocks, size_t n, int tau, uint8_t* cmask);
// generic thread-per-message SHAKE256(in_i) -> out_i[outlen<=136] (tr = H(pk), test hook for the sponge)
cudaError_t launch_shake256_rows(cudaStream_t st, const uint8_t* in, size_t in_stride, size_t in_len, size_t n,
                                 uint8_t* out, size_t out_stride, int out_len);
cudaError_t launch_shake128_rows(cudaStream_t st, const uint8_t* in, size_t in_stride, size_t in_len, size_t n,
                                 uint8_t* out, size_t out_stride, int out_len);

// message framing for mu (src/ml_dsa.rs:185-196, :386-397)
struct MsgDesc {
    const uint8_t* msgs;      // concatenated messages
    const uint64_t* msg_off;  // n + 1 offsets
    const uint8_t* ctx;       // concatenated contexts or nullptr
    const uint32_t* ctx_off;  // n + 1 offsets or nullptr (= empty ctx)
    int framing;              // 0 external, 1 internal, 2 pre-hash (msgs already hold OID || PH(M))
};

// ---- k_fused.cu: ExpandA fused with the matrix-vector product (A-hat never written) ----
// keygen arithmetic: pk and the sk fields rho, K, s1, s2, t0 from the seeds and the sampled s1 / s2 (src/ml_dsa.rs:78-98)
cudaError_t launch_kg_fused(int ps, cudaStream_t st, size_t n, const uint8_t* seeds, const int8_t* s, uint8_t* pk, uint8_t* sk);
// Signer::get_public_key (src/ml_dsa.rs:502-563): pk from packed sk
cudaError_t launch_sk_to_pk_fused(int ps, cudaStream_t st, size_t n, const uint8_t* sk, uint8_t* pk);
// verify under per-lane packed public keys: hints, z, w' = A z - c t1 2^d, UseHint, w1Encode (src/ml_dsa.rs:360-425)
cudaError_t launch_verify_fused(int ps, cudaStream_t st, size_t n, const uint8_t* pk, size_t pk_stride, const uint8_t* sig,
                                const uint8_t* cmask, uint8_t* mw1, uint8_t* pre_ok);

// ---- k_keygen.cu ----
struct KeygenWs {
    uint8_t* seeds;   // [n][128]  rho || rho' || K
    int8_t* s;        // [n][L+K][256]
    int32_t* A;       // [n][K*L][256], or nullptr: A-hat is not kept (fused kernel); callers that go on to sign / verify
                      // with the same keys (key sets, the chained pipeline) pass a buffer and reuse it
};
cudaError_t keygen_run(int ps, cudaStream_t st, size_t n, const uint8_t* xi, uint8_t* pk, uint8_t* sk,
                       const KeygenWs& ws);

// ---- k_verify.cu ----
struct VerifyWs {
    int32_t* A;        // [K*L][256] of a shared key (per-lane packed keys: A-hat is never materialised, k_fused.cu)
    int32_t* t1h;      // [nkeys][K][256]   NTT(t1 * 2^d)   (only materialised for a shared key)
    uint8_t* tr;       // [nkeys][64]
    uint8_t* mw1;      // [n][64 + W1_LEN]  mu || w1'
    uint8_t* cmask;    // [n][64]   c = SampleInBall(c~) as nonzero / negative bit masks
    uint8_t* pre_ok;   // [n]
    // Resident key set (mldsa_pk_load): A / t1h / tr point into the key set (per key), lane i uses key kidx[i] (nullptr: i)
    bool keys_resident = false;
    const uint32_t* kidx = nullptr;
};
cudaError_t verify_run(int ps, cudaStream_t st, size_t n, const uint8_t* pk, size_t pk_stride, const MsgDesc& md,
                       const uint8_t* sig, uint8_t* ok, const VerifyWs& ws);
// expand_public (src/ml_dsa.rs:477-498) for nkeys packed public keys: A-hat, NTT(t1 * 2^d), tr
// with_A / with_tr = false: the caller already holds A-hat (from keygen) / tr (bytes 64..127 of the secret key)
cudaError_t verify_expand_keys(int ps, cudaStream_t st, size_t nkeys, const uint8_t* pk, int32_t* A, int32_t* t1h, uint8_t* tr,
                               bool with_A = true, bool with_tr = true);

// ---- k_sign.cu ----
struct SignWs {
    int32_t* A;        // [nkeys][K*L][256]
    int32_t* sh;       // [nkeys][2K+L][256]: s2-hat (K rows), s1-hat (L rows), t0-hat (K rows) of a key, contiguous and in
                       // the order sign_z consumes them (one running row pointer per candidate)
    int depth;         // D = candidates per lane and round
    int32_t* w;        // [n][D][K][256]
    uint8_t* yraw;     // [n][D][L][32*ZBITS]
    uint8_t* mw1;      // [n][D][64 + W1_LEN]  mu || w1
    uint8_t* mu;       // [n][64]
    uint8_t* rhopp;    // [n][64]
    uint8_t* ctilde;   // [n][D][64]
    uint8_t* cmask;    // [n][D][64]
    uint32_t* kappa;   // [n]
    uint8_t* done;     // [n]
    uint32_t* act[2];  // [n] active lane lists (ping-pong)
    uint32_t* nact;    // [16] device control words: [buffer] active counts, [8] ctx error flag
    uint32_t* h_nact;  // mapped pinned host mirror of the 16 control words (written by publish_kernel)
    // Resident key set (mldsa_sk_load): A / sh point into the key set, the `sk` passed to sign_begin is
    // its 128-byte rho || K || tr rows, lane i uses key kidx[i] (nullptr: i), and nothing of the keys is wiped.
    bool keys_resident = false;
    const uint32_t* kidx = nullptr;
    static size_t bytes_per_lane(int K, int L, int zbits, int w1len, bool shared)
    {
        return (shared ? 0 : (size_t)(K * L + L + 2 * K) * 1024) + (size_t)K * 1024 + (size_t)L * 32 * zbits + 64 +
               w1len + 64 + 64 + SIB_BYTES + 4 + 1 + 8 + 16;
    }
};
// Host-side state of one chunk's rejection loop between its three phases (k_sign.cu).
struct SignRun {
    int ps = 0;
    size_t n = 0;
    bool shared = false;
    uint8_t* sig = nullptr;
    uint32_t* iters = nullptr;
    SignWs ws{};
    int cur = 0;            // which of ws.act[] / ws.nact[] holds the live list
    size_t bound = 0;       // host's upper bound on the live lanes
    double expect = 0;      // expected live lanes after the rounds enqueued so far
    int rounds = 0;         // rounds enqueued
    size_t tail_slots = 0;  // size of the shared tail region in candidate slots (0 = no deeper tail rounds)
    int tail_depth = 0;     // deepest speculation of the tail rounds
    size_t slot_w_bytes = 0, slot_y_bytes = 0;
    bool used_tail = false;
    bool finished = false;
    bool incomplete = false;  // stopped by the max_rounds test hook with lanes still live
    bool ctx_err = false;   // a lane's ctx was longer than 255 bytes (src/lib.rs:274)
    cudaStream_t tail_stream = nullptr;
};
// sign_begin: set-up kernels and the bulk rounds (until the expected live count drops to tail_lanes; 0 = all the
// rounds that almost surely suffice) on `st`, no host wait.  sign_tail: the remaining rounds on `st` (may differ
// from the bulk stream; the caller orders them).  sign_finish: waits for the tail stream, runs straggler rounds
// until no lane is live.  cudaErrorInvalidValue with run->ctx_err set = ctx too long.
cudaError_t sign_begin(int ps, cudaStream_t st, size_t n, const uint8_t* sk, size_t sk_stride, const MsgDesc& md,
                       const uint8_t* rnd, uint8_t* sig, uint32_t* iters, const SignWs& ws, size_t tail_lanes,
                       SignRun* run);
cudaError_t sign_tail(cudaStream_t st, SignRun* run);
// expand_private (src/ml_dsa.rs:445-469) for nkeys packed secret keys: A-hat, s1-hat, s2-hat, t0-hat
cudaError_t sign_expand_keys(int ps, cudaStream_t st, size_t nkeys, const uint8_t* sk, int32_t* A, int32_t* sh,
                             bool with_A = true);
cudaError_t sign_finish(SignRun* run);

}  // namespace mldsa
