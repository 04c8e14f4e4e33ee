/*
 * mldsa_oracle.h -- CPU ORACLE for the batched ML-DSA hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a from-scratch, scalar C restatement of the algorithm implemented by the reference crate
 * integritychain/fips204 v0.4.6 (pure Rust).  The reference cannot be compiled in this environment
 * (no rustc/cargo), so parity is pinned through the reference's OWN golden vectors instead:
 * all 180 NIST ACVP vectors (tests/nist_vectors), tests/messages.rs, tests/integration.rs::bad_sig,
 * the pk[0] pins of src/lib.rs:536-545 and the leaf KATs of src/helpers.rs:194-196 /
 * src/conversion.rs:423-488.  See tests/test_oracle_*.py.
 *
 * Keccak/SHAKE (crate sha3 ^0.10.2) and SHA-256/512 (crate sha2 ^0.10.8) live OUTSIDE the reference
 * tree; they are restated here from FIPS 202 / FIPS 180-4 and pinned against Python hashlib.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product path (fips204_b200/) never links, imports or calls it.
 */
#ifndef MLDSA_ORACLE_H
#define MLDSA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* parameter set ids: 44, 65, 87 (src/lib.rs:639-653, 681-695, 723-737) */
/* framing: 0 = external ML-DSA.Sign/Verify  (mu = H(tr || 0 || |ctx| || ctx || M),   ml_dsa.rs:190,391)
 *          1 = internal (ACVP)              (mu = H(tr || M),                         ml_dsa.rs:187,388)
 *          2 = HashML-DSA                   (mu = H(tr || 1 || |ctx| || ctx || OID || PH(M)), ml_dsa.rs:193,394)
 * ph (framing 2 only): 0 = SHA-256, 1 = SHA-512, 2 = SHAKE128 (hashing.rs:317-354) */

typedef struct {
    int ps, k, l, eta, tau, lambda_div4, gamma1, gamma2, omega, beta;
    int pk_len, sk_len, sig_len, w1_len, z_bits, w1_bits, eta_bits;
} oracle_params;

int oracle_get_params(int ps, oracle_params *out); /* 0 ok, -1 unknown ps */

/* FIPS 202 / FIPS 180-4 primitives */
void oracle_shake128(const uint8_t *in, size_t inlen, uint8_t *out, size_t outlen);
void oracle_shake256(const uint8_t *in, size_t inlen, uint8_t *out, size_t outlen);
void oracle_sha256(const uint8_t *in, size_t inlen, uint8_t out[32]);
void oracle_sha512(const uint8_t *in, size_t inlen, uint8_t out[64]);
void oracle_keccak_f1600(uint64_t st[25]);

/* field / transform leaves (helpers.rs, ntt.rs) */
int32_t oracle_mont_reduce(int64_t a);          /* helpers.rs:156-165 */
int32_t oracle_zeta_mont(int i);                /* helpers.rs:171-184: ZETA_TABLE_MONT[i] */
void oracle_ntt_raw(int32_t w[256]);            /* ntt.rs:14-76, representative-exact */
void oracle_inv_ntt(int32_t w[256]);            /* ntt.rs:85-161, output canonical [0,q) */
void oracle_ntt_batch(int32_t *w, size_t n);    /* n polynomials in place */
void oracle_inv_ntt_batch(int32_t *w, size_t n);

/* samplers (hashing.rs) */
void oracle_rej_ntt_poly(const uint8_t seed34[34], int32_t out[256]);           /* hashing.rs:111-146 */
void oracle_rej_bounded_poly(int eta, const uint8_t seed66[66], int32_t out[256]); /* hashing.rs:158-213 */
void oracle_expand_a_batch(const uint8_t *seed34, size_t n, int32_t *out);      /* n independent seeds */
void oracle_expand_mask_poly(int gamma1, const uint8_t rho64[64], uint16_t nonce, int32_t out[256]); /* hashing.rs:281-313 */
void oracle_sample_in_ball(int tau, const uint8_t *ctilde, size_t clen, int32_t out[256]); /* hashing.rs:43-100 */

/* high/low bits (high_low.rs) */
void oracle_power2round(int32_t r, int32_t *r1, int32_t *r0);   /* high_low.rs:15-48 */
void oracle_decompose(int gamma2, int32_t r, int32_t *r1, int32_t *r0); /* high_low.rs:66-96 */
int  oracle_make_hint(int gamma2, int32_t z, int32_t r);        /* high_low.rs:134-144 */
int32_t oracle_use_hint(int gamma2, int h, int32_t r);          /* high_low.rs:155-192 */

/* coding leaves (conversion.rs) */
void oracle_bit_pack(const int32_t w[256], int a, int b, uint8_t *out);    /* conversion.rs:143-182 */
void oracle_bit_unpack(const uint8_t *in, int a, int b, int32_t w[256]);   /* conversion.rs:227-262 */
int  oracle_hint_bit_unpack(int k, int omega, const uint8_t *y, int32_t *h /* k*256 */); /* 0 ok / -1 */

/* scheme (ml_dsa.rs) */
int oracle_keygen(int ps, const uint8_t xi[32], uint8_t *pk, uint8_t *sk);   /* ml_dsa.rs:57-134 + SerDes */
int oracle_sign(int ps, int framing, int ph, const uint8_t *sk,
                const uint8_t *msg, size_t mlen, const uint8_t *ctx, size_t ctxlen,
                const uint8_t rnd[32], uint8_t *sig, uint32_t *iters);       /* ml_dsa.rs:153-337 */
int oracle_verify(int ps, int framing, int ph, const uint8_t *pk,
                  const uint8_t *msg, size_t mlen, const uint8_t *ctx, size_t ctxlen,
                  const uint8_t *sig);                                        /* ml_dsa.rs:351-437; 1/0 */
int oracle_sk_to_pk(int ps, const uint8_t *sk, uint8_t *pk);                 /* ml_dsa.rs:502-563 */

/* multi-threaded batch runners (one scalar instance per thread) for the CPU baseline */
int oracle_keygen_batch(int ps, int nthreads, size_t n, const uint8_t *xi, uint8_t *pk, uint8_t *sk);
int oracle_sign_batch(int ps, int framing, int nthreads, size_t n,
                      const uint8_t *sk, size_t sk_stride,
                      const uint8_t *msgs, const uint64_t *msg_off,
                      const uint8_t *ctx, const uint32_t *ctx_off,
                      const uint8_t *rnd, uint8_t *sig, uint32_t *iters);
int oracle_verify_batch(int ps, int framing, int nthreads, size_t n,
                        const uint8_t *pk, size_t pk_stride,
                        const uint8_t *msgs, const uint64_t *msg_off,
                        const uint8_t *ctx, const uint32_t *ctx_off,
                        const uint8_t *sig, uint8_t *ok);

#ifdef __cplusplus
}
#endif
#endif
