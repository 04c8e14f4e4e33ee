//! `fips204-b200` -- the public surface of the `fips204` crate (integritychain/fips204 v0.4.6) on top of the
//! B200-native batched ML-DSA engine (`libmldsa_b200.so`, C ABI in `include/mldsa_b200.h`).
//!
//! Drop-in part: the modules [`ml_dsa_44`], [`ml_dsa_65`], [`ml_dsa_87`] with `KG`, `PrivateKey`, `PublicKey`,
//! `try_keygen`, `try_keygen_with_rng`, the constants `SK_LEN` / `PK_LEN` / `SIG_LEN`, and the traits
//! [`traits::KeyGen`], [`traits::Signer`], [`traits::Verifier`], [`traits::SerDes`] with the reference's method
//! names, argument order, byte layouts and error strings (reference: `src/lib.rs:116-614`, `src/traits.rs:8-424`).
//! A single operation is a batch of one lane.
//!
//! Batch part (what the engine exists for): [`batch::BatchKeyGen`], [`batch::BatchSigner`], [`batch::BatchVerifier`],
//! the resident key sets [`batch::PrivateKeySet`] / [`batch::PublicKeySet`] (the reference's *expanded* keys,
//! `src/types.rs:19-41`, kept in HBM; zeroized on `Drop` like `ZeroizeOnDrop`), device-side key provisioning
//! ([`batch::keygen_to_keysets`]) and the chained keygen -> sign -> verify pipeline ([`batch::pipeline`]).
//!
//! Randomness never originates on the device: exactly one 32-byte draw per keygen (`src/ml_dsa.rs:40-41`) and per
//! signature (`src/lib.rs:282-283`) is taken from the caller's RNG on the host, in lane order, so results are
//! byte-for-byte those of the reference for the same RNG stream.  There is no CPU fallback: without the CUDA
//! library and a device every call returns `Err`.
//!
//! NOTE: this crate is shipped source-only.  The image it was written in has no Rust toolchain, so it has not been
//! compiled; `tests/test_rust_crate.py` checks its `extern "C"` block against the header symbol by symbol.
#![deny(missing_docs)]
#![deny(unsafe_op_in_unsafe_fn)]

pub mod ffi;
pub mod traits;

pub use rand_core::{CryptoRng, CryptoRngCore, Error as RngError, RngCore};

/// Pre-hash function selector of HashML-DSA (`pub enum Ph`, reference `src/types.rs:5-12`).
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Ph {
    /// SHA-256 (OID ...04 02 01)
    SHA256,
    /// SHA-512 (OID ...04 02 03)
    SHA512,
    /// SHAKE128 with 256 bits of output (OID ...04 02 0B)
    SHAKE128,
}

impl Ph {
    pub(crate) fn code(self) -> core::ffi::c_int {
        match self {
            Ph::SHA256 => 0,
            Ph::SHA512 => 1,
            Ph::SHAKE128 => 2,
        }
    }
}

/// Device the single-operation trait methods run on (the batch functions take it explicitly).
pub const DEFAULT_DEVICE: i32 = 0;

pub(crate) mod glue {
    //! Ragged-buffer marshalling and error mapping shared by the per-parameter-set modules.
    use crate::ffi;
    use core::ffi::c_int;

    /// `mldsa_*` return code -> the reference's `&'static str` errors.
    pub fn check(rc: c_int, what_ctx: &'static str) -> Result<(), &'static str> {
        match rc {
            0 => Ok(()),
            -1 => Err("mldsa_b200: bad parameter"),
            -2 => Err(what_ctx), // src/lib.rs:274
            -4 => Err("mldsa_b200: out of device memory"),
            -5 => Err("mldsa_b200: not supported"),
            _ => Err("mldsa_b200: CUDA error (see mldsa_last_cuda_error)"),
        }
    }

    /// Text of the last CUDA error on this thread (`mldsa_last_cuda_error`).
    pub fn last_cuda_error() -> String {
        // SAFETY: the library returns a NUL-terminated thread-local string that stays valid until the next call.
        unsafe { core::ffi::CStr::from_ptr(ffi::mldsa_last_cuda_error()) }.to_string_lossy().into_owned()
    }

    /// Concatenated byte strings plus `n + 1` offsets, the shape `msgs` / `ctx` have in the C ABI.
    pub struct Ragged<O> {
        pub blob: Vec<u8>,
        pub off: Vec<O>,
    }

    pub fn ragged_u64(items: &[&[u8]]) -> Ragged<u64> {
        let mut blob = Vec::with_capacity(items.iter().map(|m| m.len()).sum::<usize>() + 1);
        let mut off = Vec::with_capacity(items.len() + 1);
        off.push(0u64);
        for m in items {
            blob.extend_from_slice(m);
            off.push(blob.len() as u64);
        }
        if blob.is_empty() {
            blob.push(0);
        }
        Ragged { blob, off }
    }

    pub fn ragged_u32(items: &[&[u8]]) -> Ragged<u32> {
        let mut blob = Vec::with_capacity(items.iter().map(|m| m.len()).sum::<usize>() + 1);
        let mut off = Vec::with_capacity(items.len() + 1);
        off.push(0u32);
        for m in items {
            blob.extend_from_slice(m);
            off.push(blob.len() as u32);
        }
        if blob.is_empty() {
            blob.push(0);
        }
        Ragged { blob, off }
    }

    /// One 32-byte draw per lane, in lane order (`src/ml_dsa.rs:40-41`, `src/lib.rs:282-283`).
    pub fn draw(rng: &mut impl rand_core::CryptoRngCore, n: usize, err: &'static str) -> Result<Vec<u8>, &'static str> {
        let mut out = vec![0u8; 32 * n];
        for lane in out.chunks_exact_mut(32) {
            rng.try_fill_bytes(lane).map_err(|_| err)?;
        }
        Ok(out)
    }
}

pub mod batch;

/// Everything one of the reference's parameter-set modules exports (`functionality!`, `src/lib.rs:116-614`),
/// instantiated for a parameter set `$ps` with its byte lengths.
macro_rules! functionality {
    ($ps:expr) => {
        use crate::batch::{BatchKeyGen, BatchSigner, BatchVerifier};
        use crate::traits::{KeyGen, SerDes, Signer, Verifier};
        use crate::{ffi, glue, Ph, DEFAULT_DEVICE};
        use rand_core::CryptoRngCore;
        use zeroize::{Zeroize, ZeroizeOnDrop};

        /// Parameter-set number passed to the C ABI (44, 65 or 87).
        pub const PS: i32 = $ps;

        /// Correctly sized private key specific to the target security parameter set.  Holds the canonical
        /// encoding (`skEncode`, `src/encodings.rs:94-152`); the expanded form lives on the device for the
        /// duration of a call, or in a [`crate::batch::PrivateKeySet`].  Zeroized on drop (`src/types.rs:19`).
        #[derive(Clone, Zeroize, ZeroizeOnDrop)]
        pub struct PrivateKey(pub(crate) [u8; SK_LEN]);

        /// Correctly sized public key specific to the target security parameter set (canonical `pkEncode` bytes).
        #[derive(Clone, Debug, PartialEq, Eq)]
        pub struct PublicKey(pub(crate) [u8; PK_LEN]);

        /// Empty struct to enable `KeyGen` trait objects across security parameter sets (`src/lib.rs:140`).
        pub struct KG();

        /// Signature bytes of this parameter set.
        pub type Signature = [u8; SIG_LEN];

        /// `KeyGen::try_keygen` with the operating system's RNG (`src/lib.rs:195`).
        #[cfg(feature = "default-rng")]
        pub fn try_keygen() -> Result<(PublicKey, PrivateKey), &'static str> {
            KG::try_keygen()
        }

        /// `KeyGen::try_keygen_with_rng` (`src/lib.rs:230`).
        pub fn try_keygen_with_rng(rng: &mut impl CryptoRngCore) -> Result<(PublicKey, PrivateKey), &'static str> {
            KG::try_keygen_with_rng(rng)
        }

        impl KeyGen for KG {
            type PrivateKey = PrivateKey;
            type PublicKey = PublicKey;

            fn try_keygen_with_rng(rng: &mut impl CryptoRngCore) -> Result<(PublicKey, PrivateKey), &'static str> {
                // one draw of xi (src/ml_dsa.rs:40-41), same error text
                let mut xi = [0u8; 32];
                rng.try_fill_bytes(&mut xi).map_err(|_| "KeyGen: Random number generator failed")?;
                let keys = Self::keygen_from_seed(&xi);
                xi.zeroize();
                Ok(keys)
            }

            fn keygen_from_seed(xi: &[u8; 32]) -> (PublicKey, PrivateKey) {
                let (mut pk, mut sk) = ([0u8; PK_LEN], [0u8; SK_LEN]);
                // SAFETY: xi is 32 bytes, pk / sk are PK_LEN / SK_LEN bytes: one lane of the batch call.
                let rc = unsafe {
                    ffi::mldsa_keygen_batch(PS, DEFAULT_DEVICE, 1, xi.as_ptr(), pk.as_mut_ptr(), sk.as_mut_ptr())
                };
                // the reference's signature is infallible; a device failure has no value to return
                assert!(rc == 0, "mldsa_keygen_batch failed: {} ({})", rc, glue::last_cuda_error());
                (PublicKey(pk), PrivateKey(sk))
            }
        }

        impl PrivateKey {
            fn sign_one(&self, framing_or_ph: Result<i32, Ph>, rnd: &[u8; 32], message: &[u8], ctx: &[u8],
                        err_ctx: &'static str) -> Result<Signature, &'static str> {
                if ctx.len() > 255 {
                    return Err(err_ctx); // src/lib.rs:274, :316
                }
                let (moff, coff) = ([0u64, message.len() as u64], [0u32, ctx.len() as u32]);
                let (m, c) = (if message.is_empty() { &[0u8][..] } else { message }, if ctx.is_empty() { &[0u8][..] } else { ctx });
                let mut sig = [0u8; SIG_LEN];
                // SAFETY: one lane; every pointer covers the length the ABI documents for n = 1.
                let rc = unsafe {
                    match framing_or_ph {
                        Ok(framing) => ffi::mldsa_sign_batch(PS, DEFAULT_DEVICE, framing, 1, self.0.as_ptr(), SK_LEN, m.as_ptr(),
                                                             moff.as_ptr(), c.as_ptr(), coff.as_ptr(), rnd.as_ptr(), sig.as_mut_ptr(),
                                                             core::ptr::null_mut()),
                        Err(ph) => ffi::mldsa_hash_sign_batch(PS, DEFAULT_DEVICE, ph.code(), 1, self.0.as_ptr(), SK_LEN, m.as_ptr(),
                                                              moff.as_ptr(), c.as_ptr(), coff.as_ptr(), rnd.as_ptr(),
                                                              sig.as_mut_ptr(), core::ptr::null_mut()),
                    }
                };
                glue::check(rc, err_ctx).map(|_| sig)
            }
        }

        impl Signer for PrivateKey {
            type PublicKey = PublicKey;
            type Signature = Signature;

            fn try_sign_with_rng(&self, rng: &mut impl CryptoRngCore, message: &[u8], ctx: &[u8]) -> Result<Signature, &'static str> {
                if ctx.len() > 255 {
                    return Err("ML-DSA.Sign: ctx too long"); // checked before the draw, like src/lib.rs:274-283
                }
                let mut rnd = [0u8; 32];
                rng.try_fill_bytes(&mut rnd).map_err(|_| "ML-DSA.Sign: random number generator failed")?;
                self.sign_one(Ok(0), &rnd, message, ctx, "ML-DSA.Sign: ctx too long")
            }

            fn try_hash_sign_with_rng(&self, rng: &mut impl CryptoRngCore, message: &[u8], ctx: &[u8], ph: &Ph)
                                      -> Result<Signature, &'static str> {
                if ctx.len() > 255 {
                    return Err("HashML-DSA.Sign: ctx too long"); // src/lib.rs:316
                }
                let mut rnd = [0u8; 32];
                rng.try_fill_bytes(&mut rnd).map_err(|_| "HashML-DSA.Sign: random number generator failed")?;
                self.sign_one(Err(*ph), &rnd, message, ctx, "HashML-DSA.Sign: ctx too long")
            }

            fn get_public_key(&self) -> PublicKey {
                let mut pk = [0u8; PK_LEN];
                // SAFETY: one lane, SK_LEN bytes in, PK_LEN bytes out (src/ml_dsa.rs:502-563).
                let rc = unsafe { ffi::mldsa_sk_to_pk_batch(PS, DEFAULT_DEVICE, 1, self.0.as_ptr(), pk.as_mut_ptr()) };
                assert!(rc == 0, "mldsa_sk_to_pk_batch failed: {} ({})", rc, glue::last_cuda_error());
                PublicKey(pk)
            }
        }

        impl PublicKey {
            fn verify_one(&self, framing_or_ph: Result<i32, Ph>, message: &[u8], sig: &Signature, ctx: &[u8]) -> bool {
                if ctx.len() > 255 {
                    return false; // src/lib.rs:368-370, :395-397
                }
                let (moff, coff) = ([0u64, message.len() as u64], [0u32, ctx.len() as u32]);
                let (m, c) = (if message.is_empty() { &[0u8][..] } else { message }, if ctx.is_empty() { &[0u8][..] } else { ctx });
                let mut ok = 0u8;
                // SAFETY: one lane; pk_stride = 0 selects the shared-key path (A-hat, tr, t1-hat expanded once).
                let rc = unsafe {
                    match framing_or_ph {
                        Ok(framing) => ffi::mldsa_verify_batch(PS, DEFAULT_DEVICE, framing, 1, self.0.as_ptr(), 0, m.as_ptr(),
                                                               moff.as_ptr(), c.as_ptr(), coff.as_ptr(), sig.as_ptr(), &mut ok),
                        Err(ph) => ffi::mldsa_hash_verify_batch(PS, DEFAULT_DEVICE, ph.code(), 1, self.0.as_ptr(), 0, m.as_ptr(),
                                                                moff.as_ptr(), c.as_ptr(), coff.as_ptr(), sig.as_ptr(), &mut ok),
                    }
                };
                rc == 0 && ok == 1 // malformed signatures are never errors (src/ml_dsa.rs:368-376)
            }
        }

        impl Verifier for PublicKey {
            type Signature = Signature;

            fn verify(&self, message: &[u8], sig: &Signature, ctx: &[u8]) -> bool {
                self.verify_one(Ok(0), message, sig, ctx)
            }

            fn hash_verify(&self, message: &[u8], sig: &Signature, ctx: &[u8], ph: &Ph) -> bool {
                self.verify_one(Err(*ph), message, sig, ctx)
            }
        }

        impl SerDes for PrivateKey {
            type ByteArray = [u8; SK_LEN];

            fn try_from_bytes(sk: [u8; SK_LEN]) -> Result<Self, &'static str> {
                // The reference's range checks on eta / t0 fields are tautological (src/conversion.rs:259-260):
                // every correctly sized byte string deserializes, and signing with it is the same mod-q computation.
                Ok(PrivateKey(sk))
            }

            fn into_bytes(self) -> [u8; SK_LEN] {
                self.0
            }
        }

        impl SerDes for PublicKey {
            type ByteArray = [u8; PK_LEN];

            fn try_from_bytes(pk: [u8; PK_LEN]) -> Result<Self, &'static str> {
                Ok(PublicKey(pk)) // t1 fields are 10 bits wide: every value is in range (src/encodings.rs:55-80)
            }

            fn into_bytes(self) -> [u8; PK_LEN] {
                self.0
            }
        }

        /// ACVP interface: sign with mu = H(tr || M) (`_internal_sign`, `src/lib.rs:586-594`).
        pub fn _internal_sign(sk: &PrivateKey, message: &[u8], ctx: &[u8], rnd: [u8; 32]) -> Result<Signature, &'static str> {
            sk.sign_one(Ok(1), &rnd, message, ctx, "_internal_sign: ctx too long")
        }

        /// ACVP interface: verify with mu = H(tr || M) (`_internal_verify`, `src/lib.rs:605-612`).
        #[must_use]
        pub fn _internal_verify(pk: &PublicKey, message: &[u8], sig: &Signature, ctx: &[u8]) -> bool {
            pk.verify_one(Ok(1), message, sig, ctx)
        }

        // ---- batch extension traits ------------------------------------------------------------------------
        impl BatchKeyGen for KG {
            type PrivateKey = PrivateKey;
            type PublicKey = PublicKey;

            fn keygen_batch_from_seeds(xis: &[[u8; 32]], devs: &[i32]) -> Result<Vec<(PublicKey, PrivateKey)>, &'static str> {
                let n = xis.len();
                let (mut pk, mut sk) = (vec![0u8; n * PK_LEN], vec![0u8; n * SK_LEN]);
                // SAFETY: n x 32 bytes in, n x PK_LEN / SK_LEN bytes out; devs lists ndev device ordinals.
                let rc = unsafe {
                    ffi::mldsa_keygen_batch_multi(PS, devs.as_ptr(), devs.len() as i32, n, xis.as_ptr().cast(), pk.as_mut_ptr(),
                                                  sk.as_mut_ptr())
                };
                glue::check(rc, "KeyGen: bad parameter")?;
                let out = pk.chunks_exact(PK_LEN).zip(sk.chunks_exact(SK_LEN))
                    .map(|(p, s)| (PublicKey(p.try_into().unwrap()), PrivateKey(s.try_into().unwrap()))).collect();
                sk.zeroize();
                Ok(out)
            }
        }

        impl BatchSigner for PrivateKey {
            type Signature = Signature;

            fn try_sign_batch_with_seeds(keys: &[&Self], msgs: &[&[u8]], ctxs: &[&[u8]], rnd: &[[u8; 32]], devs: &[i32])
                                         -> Result<Vec<Signature>, &'static str> {
                let n = msgs.len();
                let shared = keys.len() == 1 && n != 1;
                if (!shared && keys.len() != n) || ctxs.len() != n || rnd.len() != n {
                    return Err("ML-DSA.Sign: batch arguments differ in length");
                }
                if ctxs.iter().any(|c| c.len() > 255) {
                    return Err("ML-DSA.Sign: ctx too long");
                }
                let mut skb = Vec::with_capacity(keys.len() * SK_LEN);
                keys.iter().for_each(|k| skb.extend_from_slice(&k.0));
                let (m, c) = (glue::ragged_u64(msgs), glue::ragged_u32(ctxs));
                let mut sig = vec![0u8; n * SIG_LEN];
                // SAFETY: buffers sized as the ABI documents; sk_stride = 0 marks one key shared by all lanes.
                let rc = unsafe {
                    ffi::mldsa_sign_batch_multi(PS, devs.as_ptr(), devs.len() as i32, 0, n, skb.as_ptr(),
                                                if shared { 0 } else { SK_LEN }, m.blob.as_ptr(), m.off.as_ptr(), c.blob.as_ptr(),
                                                c.off.as_ptr(), rnd.as_ptr().cast(), sig.as_mut_ptr(), core::ptr::null_mut())
                };
                skb.zeroize();
                glue::check(rc, "ML-DSA.Sign: ctx too long")?;
                Ok(sig.chunks_exact(SIG_LEN).map(|s| s.try_into().unwrap()).collect())
            }
        }

        impl BatchVerifier for PublicKey {
            type Signature = Signature;

            fn verify_batch(keys: &[&Self], msgs: &[&[u8]], sigs: &[Signature], ctxs: &[&[u8]], devs: &[i32]) -> Vec<bool> {
                let n = msgs.len();
                let shared = keys.len() == 1 && n != 1;
                if (!shared && keys.len() != n) || ctxs.len() != n || sigs.len() != n {
                    return vec![false; n];
                }
                let mut pkb = Vec::with_capacity(keys.len() * PK_LEN);
                keys.iter().for_each(|k| pkb.extend_from_slice(&k.0));
                let (m, c) = (glue::ragged_u64(msgs), glue::ragged_u32(ctxs));
                let mut ok = vec![0u8; n];
                // SAFETY: buffers sized as the ABI documents; lanes with |ctx| > 255 come back as 0 (src/lib.rs:368-370).
                let rc = unsafe {
                    ffi::mldsa_verify_batch_multi(PS, devs.as_ptr(), devs.len() as i32, 0, n, pkb.as_ptr(),
                                                  if shared { 0 } else { PK_LEN }, m.blob.as_ptr(), m.off.as_ptr(), c.blob.as_ptr(),
                                                  c.off.as_ptr(), sigs.as_ptr().cast(), ok.as_mut_ptr())
                };
                ok.iter().map(|&v| rc == 0 && v == 1).collect()
            }
        }
    };
}

/// ML-DSA-44 (security category 2), reference `src/lib.rs:639-676`.
#[cfg(feature = "ml-dsa-44")]
pub mod ml_dsa_44 {
    /// Private (secret) key length in bytes.
    pub const SK_LEN: usize = 2560;
    /// Public key length in bytes.
    pub const PK_LEN: usize = 1312;
    /// Signature length in bytes.
    pub const SIG_LEN: usize = 2420;
    functionality!(44);
}

/// ML-DSA-65 (security category 3), reference `src/lib.rs:681-718`.
#[cfg(feature = "ml-dsa-65")]
pub mod ml_dsa_65 {
    /// Private (secret) key length in bytes.
    pub const SK_LEN: usize = 4032;
    /// Public key length in bytes.
    pub const PK_LEN: usize = 1952;
    /// Signature length in bytes.
    pub const SIG_LEN: usize = 3309;
    functionality!(65);
}

/// ML-DSA-87 (security category 5), reference `src/lib.rs:723-760`.
#[cfg(feature = "ml-dsa-87")]
pub mod ml_dsa_87 {
    /// Private (secret) key length in bytes.
    pub const SK_LEN: usize = 4896;
    /// Public key length in bytes.
    pub const PK_LEN: usize = 2592;
    /// Signature length in bytes.
    pub const SIG_LEN: usize = 4627;
    functionality!(87);
}
