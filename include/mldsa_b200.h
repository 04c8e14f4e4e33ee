/*
 * mldsa_b200.h -- C ABI of the B200-native batched ML-DSA (FIPS 204) engine.
 *
 * This is the drop-in boundary for the hot path of integritychain/fips204 (pure Rust, v0.4.6).
 * The reference has no FFI of its own; each entry point below names the reference interface it
 * replaces (file:line into the reference tree).  A Rust host crate binds these with `extern "C"`
 * (see INTEGRATION.md); the in-repo host mirrors are fips204_b200/host/fips204.hpp (C++) and
 * fips204_b200/__init__.py (ctypes).
 *
 * Conventions
 *   ps       parameter set: 44, 65 or 87            (ml_dsa_44/65/87, src/lib.rs:639,681,723)
 *   dev      CUDA device ordinal the batch runs on
 *   framing  0 = ML-DSA.Sign/Verify   mu = H(tr || 0x00 || |ctx| || ctx || M)   (src/ml_dsa.rs:190,391)
 *            1 = *_internal (ACVP)    mu = H(tr || M)                           (src/ml_dsa.rs:187,388)
 *            (HashML-DSA, mu = H(tr || 0x01 || |ctx| || ctx || OID || PH(M)), has its own entry points)
 *   All keys / signatures are the canonical FIPS 204 byte strings the reference's SerDes produces
 *   (src/lib.rs:417-494): pk PK_LEN, sk SK_LEN, sig SIG_LEN bytes, densely packed lane after lane.
 *   Randomness never originates on the device: xi / rnd are drawn by the caller (exactly one
 *   32-byte draw per keygen and per sign, src/ml_dsa.rs:40-41 and src/lib.rs:282-283), so results
 *   are byte-exact with the reference for identical seeds.
 *   Host entry points take host pointers (pageable or pinned) and return after the results are in
 *   the output buffers.  *_dev entry points take device pointers on the CURRENT device plus a
 *   cudaStream_t (as void*) and return after the work has completed on that stream.
 *   The library keeps no caller pointer after return.  Device scratch that held secrets
 *   (s1, s2, t0, y, rho'', K) is zeroed before it is released.
 *
 * Return codes: 0 ok, -1 bad parameter, -2 ctx longer than 255 bytes (sign only; verify yields
 * ok[i] = 0, src/lib.rs:368-370), -3 CUDA error (see mldsa_last_cuda_error), -4 out of device memory,
 * -5 not supported.  Malformed signatures are NOT errors: they verify as 0.
 */
#ifndef MLDSA_B200_H
#define MLDSA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLDSA_OK 0
#define MLDSA_E_PARAM (-1)
#define MLDSA_E_CTX (-2)
#define MLDSA_E_CUDA (-3)
#define MLDSA_E_NOMEM (-4)
#define MLDSA_E_UNSUPPORTED (-5)

/* ---- introspection ---------------------------------------------------------------------- */
int mldsa_device_count(void);
const char* mldsa_strerror(int code);
const char* mldsa_last_cuda_error(void); /* thread-local text of the last -3 */
/* PK_LEN / SK_LEN / SIG_LEN of a parameter set (src/lib.rs:649-653); 0 on success */
int mldsa_sizes(int ps, size_t* pk_len, size_t* sk_len, size_t* sig_len);

/* pinned host buffers for callers that want DMA straight from their own memory */
void* mldsa_host_alloc(size_t bytes);
void mldsa_host_free(void* p);

/* ---- keygen: KG::keygen_from_seed / try_keygen_with_rng + SerDes::into_bytes ------------
 * (src/lib.rs:230-251, src/ml_dsa.rs:57-134).  xi: n x 32 seed bytes; pk: n x PK_LEN; sk: n x SK_LEN. */
int mldsa_keygen_batch(int ps, int dev, size_t n, const uint8_t* xi, uint8_t* pk, uint8_t* sk);
int mldsa_keygen_batch_dev(int ps, void* stream, size_t n, const uint8_t* d_xi, uint8_t* d_pk, uint8_t* d_sk);

/* ---- sign: PrivateKey::try_sign_with_rng / try_sign_with_seed / _internal_sign ----------
 * (src/lib.rs:268-296, :586-594, src/ml_dsa.rs:153-337).
 * sk: n x SK_LEN with sk_stride = SK_LEN, or one key with sk_stride = 0 (shared key).
 * msgs + msg_off[n+1]: concatenated messages.  ctx + ctx_off[n+1]: contexts, or NULL,NULL = empty.
 * rnd: n x 32 (all-zero rows = deterministic variant).  sig: n x SIG_LEN.
 * iters: optional n trip counts of the rejection loop (src/ml_dsa.rs:212-330), may be NULL. */
int mldsa_sign_batch(int ps, int dev, int framing, size_t n, const uint8_t* sk, size_t sk_stride,
                     const uint8_t* msgs, const uint64_t* msg_off, const uint8_t* ctx, const uint32_t* ctx_off,
                     const uint8_t* rnd, uint8_t* sig, uint32_t* iters);
int mldsa_sign_batch_dev(int ps, void* stream, int framing, size_t n, const uint8_t* d_sk, size_t sk_stride,
                         const uint8_t* d_msgs, const uint64_t* d_msg_off, const uint8_t* d_ctx,
                         const uint32_t* d_ctx_off, const uint8_t* d_rnd, uint8_t* d_sig, uint32_t* d_iters);

/* ---- verify: PublicKey::verify / _internal_verify ----------------------------------------
 * (src/lib.rs:364-380, :605-612, src/ml_dsa.rs:351-437).
 * pk: n x PK_LEN with pk_stride = PK_LEN, or one key with pk_stride = 0 (A-hat, tr and
 * NTT(t1*2^d) are then expanded once and cached for the whole batch).  ok: n bytes, 1 = valid. */
int mldsa_verify_batch(int ps, int dev, int framing, size_t n, const uint8_t* pk, size_t pk_stride,
                       const uint8_t* msgs, const uint64_t* msg_off, const uint8_t* ctx, const uint32_t* ctx_off,
                       const uint8_t* sig, uint8_t* ok);
int mldsa_verify_batch_dev(int ps, void* stream, int framing, size_t n, const uint8_t* d_pk, size_t pk_stride,
                           const uint8_t* d_msgs, const uint64_t* d_msg_off, const uint8_t* d_ctx,
                           const uint32_t* d_ctx_off, const uint8_t* d_sig, uint8_t* d_ok);

/* ---- HashML-DSA: Signer::try_hash_sign_with_rng / Verifier::hash_verify -------------------
 * (src/lib.rs:310-342, :391-411; pre-hash + OID src/hashing.rs:317-354).  ph: 0 = SHA-256, 1 = SHA-512,
 * 2 = SHAKE128 (types::Ph).  The pre-hash runs on the device; mu = H(tr || 0x01 || |ctx| || ctx || OID || PH(M)).
 * Other arguments as mldsa_sign_batch / mldsa_verify_batch. */
int mldsa_hash_sign_batch(int ps, int dev, int ph, size_t n, const uint8_t* sk, size_t sk_stride, const uint8_t* msgs,
                          const uint64_t* msg_off, const uint8_t* ctx, const uint32_t* ctx_off, const uint8_t* rnd,
                          uint8_t* sig, uint32_t* iters);
int mldsa_hash_verify_batch(int ps, int dev, int ph, size_t n, const uint8_t* pk, size_t pk_stride,
                            const uint8_t* msgs, const uint64_t* msg_off, const uint8_t* ctx, const uint32_t* ctx_off,
                            const uint8_t* sig, uint8_t* ok);

/* ---- Signer::get_public_key (src/lib.rs:344-346, src/ml_dsa.rs:502-563): pk bytes from sk bytes ---- */
int mldsa_sk_to_pk_batch(int ps, int dev, size_t n, const uint8_t* sk, uint8_t* pk);
int mldsa_sk_to_pk_batch_dev(int ps, void* stream, size_t n, const uint8_t* d_sk, uint8_t* d_pk);

/* ---- multi-GPU: one host batch split contiguously over a device list -----------------------
 * Lanes are independent, so there is no collective: device g of ndev gets lanes
 * [n*g/ndev, n*(g+1)/ndev), runs on its own host thread, and writes its disjoint output slice.
 * Arguments as in the single-device host entry points. */
int mldsa_keygen_batch_multi(int ps, const int* devs, int ndev, size_t n, const uint8_t* xi, uint8_t* pk, uint8_t* sk);
int mldsa_sign_batch_multi(int ps, const int* devs, int ndev, int framing, size_t n, const uint8_t* sk,
                           size_t sk_stride, const uint8_t* msgs, const uint64_t* msg_off, const uint8_t* ctx,
                           const uint32_t* ctx_off, const uint8_t* rnd, uint8_t* sig, uint32_t* iters);
int mldsa_verify_batch_multi(int ps, const int* devs, int ndev, int framing, size_t n, const uint8_t* pk,
                             size_t pk_stride, const uint8_t* msgs, const uint64_t* msg_off, const uint8_t* ctx,
                             const uint32_t* ctx_off, const uint8_t* sig, uint8_t* ok);

/* ---- chained pipeline (BASELINE config 5): keygen -> sign -> verify with keys and signatures resident ------------
 * The reference's end-to-end flow (src/lib.rs:503-551: try_keygen, try_sign, verify) as ONE batch call: lane i derives
 * (pk_i, sk_i) from xi[i], signs msgs[i] (ctx[i], rnd[i]) with sk_i and verifies the signature under pk_i.  Only xi,
 * the messages / contexts and rnd cross PCIe on the way in; the packed pk, sig and verdicts on the way out (sk and the
 * rejection-loop trip counts only if the pointers are given).  Between the stages everything stays in HBM, and A-hat --
 * which the reference re-derives from rho in all three operations (src/ml_dsa.rs:78, :181, :360) -- is expanded once
 * per key.  Bytes are those of mldsa_keygen_batch / mldsa_sign_batch / mldsa_verify_batch called one after the other.
 * stage_ms: optional, 3 doubles: device time of the keygen / sign / verify stage (summed over the chunks; _multi: the
 * slowest device per stage).  Return codes as mldsa_sign_batch. */
int mldsa_pipeline_batch(int ps, int dev, int framing, size_t n, const uint8_t* xi, const uint8_t* msgs,
                         const uint64_t* msg_off, const uint8_t* ctx, const uint32_t* ctx_off, const uint8_t* rnd,
                         uint8_t* pk, uint8_t* sk /* optional */, uint8_t* sig, uint8_t* ok, uint32_t* iters /* optional */,
                         double* stage_ms /* optional */);
int mldsa_pipeline_batch_multi(int ps, const int* devs, int ndev, int framing, size_t n, const uint8_t* xi,
                               const uint8_t* msgs, const uint64_t* msg_off, const uint8_t* ctx, const uint32_t* ctx_off,
                               const uint8_t* rnd, uint8_t* pk, uint8_t* sk, uint8_t* sig, uint8_t* ok, uint32_t* iters,
                               double* stage_ms);

/* ---- microbench / leaf entry points (BASELINE config 2) ----------------------------------
 * ntt: src/ntt.rs:14-76, raw-bit-exact representatives; intt: src/ntt.rs:85-161, canonical [0,q).
 * expand_a: out[i] = RejNTTPoly(seed34[i]), src/hashing.rs:111-146.  In place on n x 256 int32.
 * Device pointers of the _dev forms: polynomials 32-byte aligned (-1 otherwise; cudaMalloc gives 256). */
int mldsa_ntt_batch(int dev, size_t n, int32_t* polys);
int mldsa_intt_batch(int dev, size_t n, int32_t* polys);
int mldsa_expand_a_batch(int dev, size_t n, const uint8_t* seed34, int32_t* out);
int mldsa_ntt_batch_dev(void* stream, size_t n, int32_t* d_polys);
int mldsa_intt_batch_dev(void* stream, size_t n, int32_t* d_polys);
int mldsa_expand_a_batch_dev(void* stream, size_t n, const uint8_t* d_seed34, int32_t* d_out);
/* SHAKE128 / SHAKE256 of n fixed-length rows (h256_xof / g128_xof, src/hashing.rs:13-27) */
int mldsa_shake_batch(int dev, int bits /*128|256*/, size_t n, const uint8_t* in, size_t in_len, uint8_t* out,
                      size_t out_len);

/* ---- resident key sets ----------------------------------------------------------------------
 * The reference's PrivateKey is the EXPANDED key: rho, K, tr and the NTT images of s1, s2, t0 (src/types.rs:19-28),
 * built once by try_from_bytes -> expand_private (src/lib.rs:462-468, src/ml_dsa.rs:445-469) and reused by every
 * sign call.  mldsa_sk_load is that step for n keys at once; the expanded keys (plus A-hat, which the reference
 * re-derives from rho on every sign, src/ml_dsa.rs:181) stay in HBM until mldsa_keyset_free, which zeroizes them
 * (ZeroizeOnDrop, src/types.rs:19).  Signing with a key set moves 64 bytes + the message per lane over PCIe
 * instead of the packed key as well, and skips sk_decode, L + 2K forward NTTs and ExpandA.
 * Device memory: 29 / 47 / 79 KiB per key for ML-DSA-44 / 65 / 87. */
typedef struct mldsa_keyset mldsa_keyset;
int mldsa_sk_load(int ps, int dev, size_t n, const uint8_t* sk /* n x SK_LEN, host */, mldsa_keyset** out);
size_t mldsa_keyset_size(const mldsa_keyset* ks);          /* number of keys */
size_t mldsa_keyset_device_bytes(const mldsa_keyset* ks);
void mldsa_keyset_free(mldsa_keyset* ks);
/* Signer::try_sign_with_rng on expanded keys (src/lib.rs:268-296): lane i signs with key key_idx[i] of the set
 * (key_idx == NULL: key i, n <= number of keys).  Same bytes as mldsa_sign_batch with those keys.  Other
 * arguments and return codes as mldsa_sign_batch; -1 if a key index is out of range. */
int mldsa_sign_batch_keys(const mldsa_keyset* ks, int framing, size_t n, const uint32_t* key_idx, const uint8_t* msgs,
                          const uint64_t* msg_off, const uint8_t* ctx, const uint32_t* ctx_off, const uint8_t* rnd,
                          uint8_t* sig, uint32_t* iters);

/* The same for public keys: PublicKey::try_from_bytes -> expand_public (src/lib.rs:469-474, src/ml_dsa.rs:477-498)
 * for n keys at once; rho-derived A-hat, tr = H(pk) and NTT(t1 * 2^d) (the reference's PublicKey, src/types.rs:35-41)
 * stay in HBM (36 KiB per ML-DSA-65 key) and verify calls send only signatures and messages.  Same handle type,
 * same mldsa_keyset_size / _device_bytes / _free; a secret-key set cannot verify and a public-key set cannot sign
 * (-1). */
int mldsa_pk_load(int ps, int dev, size_t n, const uint8_t* pk /* n x PK_LEN, host */, mldsa_keyset** out);
/* Verifier::verify on expanded keys (src/lib.rs:364-380): lane i is checked under key key_idx[i] (NULL: key i).
 * Same verdicts as mldsa_verify_batch with those keys. */
int mldsa_verify_batch_keys(const mldsa_keyset* ks, int framing, size_t n, const uint32_t* key_idx, const uint8_t* msgs,
                            const uint64_t* msg_off, const uint8_t* ctx, const uint32_t* ctx_off, const uint8_t* sig,
                            uint8_t* ok);

/* Device-side key provisioning: KG::keygen_from_seed (src/lib.rs:247-251, src/ml_dsa.rs:57-134) whose results are
 * kept as EXPANDED keys in HBM (expand_private / expand_public, src/ml_dsa.rs:445-498) instead of being packed and
 * shipped back: a signing service provisions n keys from n seeds and no packed secret key ever crosses PCIe (52 % of
 * the bytes of a packed-key mldsa_sign_batch call).  xi: n x 32 host bytes.  pk: optional host output, n x PK_LEN (the
 * public keys are what such a service publishes), NULL = none.  sk_set / pk_set: either may be NULL (not both); the
 * sets are the same objects mldsa_sk_load / mldsa_pk_load build from the packed keys of mldsa_keygen_batch(xi) --
 * signatures and verdicts are identical.  A-hat is expanded once and shared by construction of both sets. */
int mldsa_keygen_to_keyset(int ps, int dev, size_t n, const uint8_t* xi, uint8_t* pk, mldsa_keyset** sk_set,
                           mldsa_keyset** pk_set);

/* HashML-DSA on resident key sets (try_hash_sign_with_rng src/lib.rs:310-342, hash_verify :391-411): ph as in
 * mldsa_hash_sign_batch (0 SHA-256, 1 SHA-512, 2 SHAKE128), the pre-hash runs on the device. */
int mldsa_hash_sign_batch_keys(const mldsa_keyset* ks, int ph, size_t n, const uint32_t* key_idx, const uint8_t* msgs,
                               const uint64_t* msg_off, const uint8_t* ctx, const uint32_t* ctx_off, const uint8_t* rnd,
                               uint8_t* sig, uint32_t* iters);
int mldsa_hash_verify_batch_keys(const mldsa_keyset* ks, int ph, size_t n, const uint32_t* key_idx, const uint8_t* msgs,
                                 const uint64_t* msg_off, const uint8_t* ctx, const uint32_t* ctx_off, const uint8_t* sig,
                                 uint8_t* ok);

/* ---- tuning / telemetry ------------------------------------------------------------------
 * Environment knobs (read once; the setters below override them): MLDSA_DEPTH, MLDSA_TAILDEPTH, MLDSA_TAILSLOTS,
 * MLDSA_KCAP, MLDSA_ACAP as documented at the setters; host pipelines:
 * MLDSA_PLAN (chunk sizes of a large batch, e.g. "1,2,4,6,6,6,4,2,1"), MLDSA_TAIL (live lanes at which a sign chunk
 * hands over to the tail stream, default 16384), MLDSA_TAILPRIO (0: low-priority tail stream), MLDSA_NCAP (CTAs per
 * SM of the standalone NTT), MLDSA_TRACE=1 (device-side timeline of the host pipelines on stderr); test hook:
 * MLDSA_EST_SCALE (< 1 makes the host under-estimate the live-lane counts of the sign rounds). */
/* upper bound for the per-call device scratch, bytes (0 = default: 60% of free memory, <= 96 GiB) */
void mldsa_set_workspace_limit(size_t bytes);
/* launch tuning: kcap / acap cap the CTAs per SM of the (grid-stride) Keccak / arithmetic kernels (0 = default:
 * uncapped / 16).  Defaults come from MLDSA_KCAP / MLDSA_ACAP.  `ways` is accepted and ignored: running a chunk as
 * concurrent sub-batches on separate streams was measured to gain nothing and was removed. */
void mldsa_set_tuning(int ways, int kcap, int acap);
/* candidates (kappa, kappa + L, ...) every still-active lane tries per round of the sign rejection loop, 1..8
 * (default MLDSA_DEPTH, else 2).  Results never depend on it: candidates count in the reference's order
 * (src/ml_dsa.rs:212-330); a larger depth trades a little speculative work for fewer rounds and less A-hat traffic. */
void mldsa_set_depth(int depth);
/* the last, under-filled rounds of the sign rejection loop (a few thousand live lanes; a round then costs its fixed
 * latency) speculate deeper: once live_lanes * d candidates fit in `tail_slots` candidate slots, a round tries the
 * largest d in {4, 8, 16} with d <= tail_depth.  Results never depend on it (same argument as mldsa_set_depth).
 * Defaults MLDSA_TAILDEPTH / MLDSA_TAILSLOTS, else 8 / 32768; tail_depth <= depth switches it off. */
void mldsa_set_tail(int tail_depth, int tail_slots);
/* test hook for the ZeroizeOnDrop analogue (src/types.rs:19-41): number of non-zero bytes left in the device ranges
 * that held secret-derived data (staged xi / sk / rnd, rho', K, s1, s2 and their NTT images, y, w - c*s2, rho'')
 * during the most recent keygen or sign call on `dev`; *scanned_bytes (optional) = size of those ranges.  0 after
 * every call, successful or not. */
int mldsa_debug_secret_residue(int dev, uint64_t* nonzero_bytes, uint64_t* scanned_bytes);
/* test hook for the timing probe (tools/ct_probe.py): a sign call stops every chunk's rejection loop after `rounds`
 * rounds (0 = off, the default), wipes the unfinished lanes' state and returns -5; with rounds = 1 every kernel of
 * the loop runs exactly once over all lanes, so per-kernel device times can be compared between input classes. */
void mldsa_debug_set_max_rounds(int rounds);
/* test hook: chunk sizes (lanes) the host pipelines use for an n-lane call when `cap` lanes fit the workspace; returns
 * the number of chunks (0 = plain chunks of `cap`), writes at most max_out of them.  Needs no device. */
size_t mldsa_debug_chunk_plan(size_t n, size_t cap, size_t* out, size_t max_out);
/* test hook for the rare paths of the rejection samplers: the device parsers run over caller-supplied squeeze blocks
 * instead of SHAKE output.  which = 0: RejNTTPoly (src/hashing.rs:111-146, CoeffFromThreeBytes src/conversion.rs:40-61),
 * nblocks x 168 bytes per polynomial, out = n x 256 int32 (coefficients the stream does not reach stay 0) -- forces the
 * 6th, 7th ... squeeze block that SHAKE128 yields with p ~ 1e-39.  which = tau (39 / 49 / 60): SampleInBall
 * (src/hashing.rs:43-100), nblocks x 136 bytes per challenge, out = n x 64 bytes (256 "non-zero" bits, then 256
 * "negative" bits) -- forces the walk into a second squeeze block. */
int mldsa_debug_sampler_stream(int dev, int which, size_t n, const uint8_t* blocks, size_t nblocks, void* out);
/* kernels launched by this library since process start (all devices) */
uint64_t mldsa_kernel_launches(void);
/* rejection-loop rounds executed by the most recent sign call on this thread */
int mldsa_last_sign_rounds(void);
/* optional per-kernel device timing: CUDA events around every launch on the launching stream,
 * accumulated per kernel class.  Off by default (two event records per launch when on). */
void mldsa_profile_enable(int on);
void mldsa_profile_reset(void);
int mldsa_profile_count(void);
const char* mldsa_profile_name(int id);
int mldsa_profile_get(int id, double* ms, uint64_t* launches);
/* measured integer issue rate of the current device, thread-level instructions / s:
 * mode 0 LOP3+SHF (ALU pipe), 1 IMAD+IMAD.HI (FMA pipe), 2 both interleaved */
int mldsa_int32_peak_probe(int mode, double* ops_per_sec);
/* Keccak-f[1600] permutations/s (thread-per-state) for variant v of the ALU/FMA rotation split;
 * *mask = that variant's KECCAK_FMA_SITES.  -1 when v is out of range. */
int mldsa_keccak_probe(int variant, double* perms_per_sec, uint32_t* mask);

#ifdef __cplusplus
}
#endif
#endif
