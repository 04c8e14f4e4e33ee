//! Batch entry points -- what the engine exists for.  One lane = one key, or one (key, message[, signature]).
//! Lanes are independent; `devs` lists the CUDA devices a batch is split over (contiguous lane ranges, one host
//! thread per device inside the library, no collective).
use crate::{ffi, glue, Ph};
use rand_core::CryptoRngCore;

/// Batched `KeyGen` (reference `src/traits.rs:8-114`): n key pairs from n seeds / n RNG draws.
pub trait BatchKeyGen {
    /// A public key specific to the chosen security parameter set
    type PublicKey;
    /// A private (secret) key specific to the chosen security parameter set
    type PrivateKey;

    /// Key pair i is `keygen_from_seed(&xis[i])`.
    /// # Errors
    /// Device errors only.
    fn keygen_batch_from_seeds(xis: &[[u8; 32]], devs: &[i32]) -> Result<Vec<(Self::PublicKey, Self::PrivateKey)>, &'static str>;

    /// n key pairs; the RNG is drawn from n times, 32 bytes each, in lane order (`src/ml_dsa.rs:40-41` per lane).
    /// # Errors
    /// `"KeyGen: Random number generator failed"`, device errors.
    fn try_keygen_batch_with_rng(rng: &mut impl CryptoRngCore, n: usize, devs: &[i32])
                                 -> Result<Vec<(Self::PublicKey, Self::PrivateKey)>, &'static str> {
        let flat = glue::draw(rng, n, "KeyGen: Random number generator failed")?;
        let xis: Vec<[u8; 32]> = flat.chunks_exact(32).map(|c| c.try_into().unwrap()).collect();
        Self::keygen_batch_from_seeds(&xis, devs)
    }
}

/// Batched `Signer::try_sign_with_*` (reference `src/traits.rs:118-308`).
pub trait BatchSigner: Sized {
    /// The signature is specific to the chosen security parameter set
    type Signature;

    /// Lane i signs `msgs[i]` under `ctxs[i]` with `keys[i]` (or `keys[0]` for every lane when one key is given) and
    /// `rnd[i]` as the value of its draw.  Same bytes as n calls of `try_sign_with_seed`.
    /// # Errors
    /// `"ML-DSA.Sign: ctx too long"` if any context exceeds 255 bytes (the whole batch fails, like the scalar call).
    fn try_sign_batch_with_seeds(keys: &[&Self], msgs: &[&[u8]], ctxs: &[&[u8]], rnd: &[[u8; 32]], devs: &[i32])
                                 -> Result<Vec<Self::Signature>, &'static str>;

    /// As above, drawing one 32-byte value per lane from `rng`, in lane order (`src/lib.rs:282-283` per lane).
    /// # Errors
    /// `"ML-DSA.Sign: ctx too long"`, `"ML-DSA.Sign: random number generator failed"`.
    fn try_sign_batch_with_rng(rng: &mut impl CryptoRngCore, keys: &[&Self], msgs: &[&[u8]], ctxs: &[&[u8]], devs: &[i32])
                               -> Result<Vec<Self::Signature>, &'static str> {
        if ctxs.iter().any(|c| c.len() > 255) {
            return Err("ML-DSA.Sign: ctx too long");
        }
        let flat = glue::draw(rng, msgs.len(), "ML-DSA.Sign: random number generator failed")?;
        let rnd: Vec<[u8; 32]> = flat.chunks_exact(32).map(|c| c.try_into().unwrap()).collect();
        Self::try_sign_batch_with_seeds(keys, msgs, ctxs, &rnd, devs)
    }
}

/// Batched `Verifier::verify` (reference `src/traits.rs:330-362`).
pub trait BatchVerifier: Sized {
    /// The signature is specific to the chosen security parameter set
    type Signature;

    /// Verdict i is `keys[i].verify(msgs[i], &sigs[i], ctxs[i])` (`keys[0]` for every lane when one key is given: its
    /// A-hat, tr and NTT(t1 2^d) are then expanded once for the whole batch).  Never fails.
    fn verify_batch(keys: &[&Self], msgs: &[&[u8]], sigs: &[Self::Signature], ctxs: &[&[u8]], devs: &[i32]) -> Vec<bool>;
}

/// n expanded private keys resident in HBM: rho, K, tr, NTT(s1), NTT(s2), NTT(t0) -- the fields of the reference's
/// `PrivateKey` (`src/types.rs:19-28`, built by `expand_private`, `src/ml_dsa.rs:445-469`) -- plus A-hat.  Zeroized and
/// released on `Drop` (`ZeroizeOnDrop`, `src/types.rs:19`).
pub struct PrivateKeySet {
    handle: *mut ffi::mldsa_keyset,
    sig_len: usize,
}

/// n expanded public keys resident in HBM: A-hat, tr, NTT(t1 2^d) (`src/types.rs:35-41`, `src/ml_dsa.rs:477-498`).
pub struct PublicKeySet {
    handle: *mut ffi::mldsa_keyset,
    sig_len: usize,
}

// SAFETY: the library guards every key set with its device's mutex; the handle is an opaque token.
unsafe impl Send for PrivateKeySet {}
// SAFETY: as above.
unsafe impl Send for PublicKeySet {}

fn lens(ps: i32) -> Result<(usize, usize, usize), &'static str> {
    let (mut pk, mut sk, mut sig) = (0usize, 0usize, 0usize);
    // SAFETY: three valid out-pointers.
    let rc = unsafe { ffi::mldsa_sizes(ps, &mut pk, &mut sk, &mut sig) };
    glue::check(rc, "unknown parameter set").map(|_| (pk, sk, sig))
}

impl PrivateKeySet {
    /// Expands `n` packed secret keys (`sks` = n x SK_LEN bytes) of parameter set `ps` on device `dev`.
    /// # Errors
    /// Bad sizes, device errors.
    pub fn load(ps: i32, dev: i32, sks: &[u8]) -> Result<Self, &'static str> {
        let (_, sk_len, sig_len) = lens(ps)?;
        if sks.is_empty() || sks.len() % sk_len != 0 {
            return Err("incorrect private key length");
        }
        let mut handle = core::ptr::null_mut();
        // SAFETY: sks holds n x SK_LEN bytes; handle receives the set.
        let rc = unsafe { ffi::mldsa_sk_load(ps, dev, sks.len() / sk_len, sks.as_ptr(), &mut handle) };
        glue::check(rc, "bad parameter").map(|_| PrivateKeySet { handle, sig_len })
    }

    /// Number of keys in the set.
    #[must_use]
    pub fn len(&self) -> usize {
        // SAFETY: live handle.
        unsafe { ffi::mldsa_keyset_size(self.handle) }
    }

    /// `true` for an empty set (never constructed by this crate).
    #[must_use]
    pub fn is_empty(&self) -> bool {
        self.len() == 0
    }

    /// Lane i signs `msgs[i]` under `ctxs[i]` with key `key_idx[i]` of the set (`None`: key i) and `rnd[i]` as its draw;
    /// `ph = Some(..)` selects HashML-DSA.  Same bytes as signing with the packed keys.
    /// # Errors
    /// `"ML-DSA.Sign: ctx too long"`, index out of range, device errors.
    pub fn try_sign_batch_with_seeds(&self, key_idx: Option<&[u32]>, msgs: &[&[u8]], ctxs: &[&[u8]], rnd: &[[u8; 32]],
                                     ph: Option<Ph>) -> Result<Vec<Vec<u8>>, &'static str> {
        let n = msgs.len();
        if ctxs.len() != n || rnd.len() != n || key_idx.map_or(false, |k| k.len() != n) {
            return Err("ML-DSA.Sign: batch arguments differ in length");
        }
        if ctxs.iter().any(|c| c.len() > 255) {
            return Err("ML-DSA.Sign: ctx too long");
        }
        let (m, c) = (glue::ragged_u64(msgs), glue::ragged_u32(ctxs));
        let mut sig = vec![0u8; n * self.sig_len];
        let kp = key_idx.map_or(core::ptr::null(), <[u32]>::as_ptr);
        // SAFETY: buffers sized as the ABI documents for n lanes.
        let rc = unsafe {
            match ph {
                None => ffi::mldsa_sign_batch_keys(self.handle, 0, n, kp, m.blob.as_ptr(), m.off.as_ptr(), c.blob.as_ptr(),
                                                   c.off.as_ptr(), rnd.as_ptr().cast(), sig.as_mut_ptr(), core::ptr::null_mut()),
                Some(p) => ffi::mldsa_hash_sign_batch_keys(self.handle, p.code(), n, kp, m.blob.as_ptr(), m.off.as_ptr(),
                                                           c.blob.as_ptr(), c.off.as_ptr(), rnd.as_ptr().cast(), sig.as_mut_ptr(),
                                                           core::ptr::null_mut()),
            }
        };
        glue::check(rc, "ML-DSA.Sign: ctx too long")?;
        Ok(sig.chunks_exact(self.sig_len).map(<[u8]>::to_vec).collect())
    }
}

impl PublicKeySet {
    /// Expands `n` packed public keys (`pks` = n x PK_LEN bytes) of parameter set `ps` on device `dev`.
    /// # Errors
    /// Bad sizes, device errors.
    pub fn load(ps: i32, dev: i32, pks: &[u8]) -> Result<Self, &'static str> {
        let (pk_len, _, sig_len) = lens(ps)?;
        if pks.is_empty() || pks.len() % pk_len != 0 {
            return Err("incorrect public key length");
        }
        let mut handle = core::ptr::null_mut();
        // SAFETY: pks holds n x PK_LEN bytes; handle receives the set.
        let rc = unsafe { ffi::mldsa_pk_load(ps, dev, pks.len() / pk_len, pks.as_ptr(), &mut handle) };
        glue::check(rc, "bad parameter").map(|_| PublicKeySet { handle, sig_len })
    }

    /// Number of keys in the set.
    #[must_use]
    pub fn len(&self) -> usize {
        // SAFETY: live handle.
        unsafe { ffi::mldsa_keyset_size(self.handle) }
    }

    /// `true` for an empty set (never constructed by this crate).
    #[must_use]
    pub fn is_empty(&self) -> bool {
        self.len() == 0
    }

    /// Verdict i: `sigs[i]` (SIG_LEN bytes each, concatenated) on `msgs[i]` / `ctxs[i]` under key `key_idx[i]` (`None`:
    /// key i); `ph = Some(..)` selects HashML-DSA.  Never fails: any error yields `false` verdicts.
    #[must_use]
    pub fn verify_batch(&self, key_idx: Option<&[u32]>, msgs: &[&[u8]], sigs: &[u8], ctxs: &[&[u8]], ph: Option<Ph>) -> Vec<bool> {
        let n = msgs.len();
        if ctxs.len() != n || sigs.len() != n * self.sig_len || key_idx.map_or(false, |k| k.len() != n) {
            return vec![false; n];
        }
        let (m, c) = (glue::ragged_u64(msgs), glue::ragged_u32(ctxs));
        let mut ok = vec![0u8; n];
        let kp = key_idx.map_or(core::ptr::null(), <[u32]>::as_ptr);
        // SAFETY: buffers sized as the ABI documents for n lanes.
        let rc = unsafe {
            match ph {
                None => ffi::mldsa_verify_batch_keys(self.handle, 0, n, kp, m.blob.as_ptr(), m.off.as_ptr(), c.blob.as_ptr(),
                                                     c.off.as_ptr(), sigs.as_ptr(), ok.as_mut_ptr()),
                Some(p) => ffi::mldsa_hash_verify_batch_keys(self.handle, p.code(), n, kp, m.blob.as_ptr(), m.off.as_ptr(),
                                                             c.blob.as_ptr(), c.off.as_ptr(), sigs.as_ptr(), ok.as_mut_ptr()),
            }
        };
        ok.iter().map(|&v| rc == 0 && v == 1).collect()
    }
}

impl Drop for PrivateKeySet {
    fn drop(&mut self) {
        // SAFETY: the handle came from the library and is released exactly once; the library zeroizes the device memory.
        unsafe { ffi::mldsa_keyset_free(self.handle) }
    }
}

impl Drop for PublicKeySet {
    fn drop(&mut self) {
        // SAFETY: as above.
        unsafe { ffi::mldsa_keyset_free(self.handle) }
    }
}

/// `KeyGen::keygen_from_seed` for n seeds whose expanded keys stay on the device (`mldsa_keygen_to_keyset`): the packed
/// public keys come back, no packed secret key ever crosses PCIe.
/// # Errors
/// Device errors.
pub fn keygen_to_keysets(ps: i32, dev: i32, xis: &[[u8; 32]]) -> Result<(Vec<u8>, PrivateKeySet, PublicKeySet), &'static str> {
    let (pk_len, _, sig_len) = lens(ps)?;
    let mut pk = vec![0u8; xis.len() * pk_len];
    let (mut s, mut p) = (core::ptr::null_mut(), core::ptr::null_mut());
    // SAFETY: n x 32 bytes in, n x PK_LEN bytes out, two out-handles.
    let rc = unsafe { ffi::mldsa_keygen_to_keyset(ps, dev, xis.len(), xis.as_ptr().cast(), pk.as_mut_ptr(), &mut s, &mut p) };
    glue::check(rc, "bad parameter")?;
    Ok((pk, PrivateKeySet { handle: s, sig_len }, PublicKeySet { handle: p, sig_len }))
}

/// Result of [`pipeline`]: packed public keys and signatures (lane after lane), verdicts, and the device time of the
/// keygen / sign / verify stage in milliseconds.
pub struct PipelineOutput {
    /// n x PK_LEN bytes
    pub pk: Vec<u8>,
    /// n x SIG_LEN bytes
    pub sig: Vec<u8>,
    /// verdict per lane
    pub ok: Vec<bool>,
    /// keygen, sign, verify
    pub stage_ms: [f64; 3],
}

/// keygen -> sign -> verify chained on the device(s) with keys and signatures resident between the stages (the flow of
/// the reference's smoke test, `src/lib.rs:503-551`, as one batch call; BASELINE config 5).  Lane i derives its key pair
/// from `xis[i]`, signs `msgs[i]` with `rnd[i]` as its draw and verifies the result.
/// # Errors
/// `"ML-DSA.Sign: ctx too long"`, device errors.
pub fn pipeline(ps: i32, devs: &[i32], xis: &[[u8; 32]], msgs: &[&[u8]], ctxs: &[&[u8]], rnd: &[[u8; 32]])
                -> Result<PipelineOutput, &'static str> {
    let n = xis.len();
    if msgs.len() != n || ctxs.len() != n || rnd.len() != n {
        return Err("pipeline: batch arguments differ in length");
    }
    let (pk_len, _, sig_len) = lens(ps)?;
    let (m, c) = (glue::ragged_u64(msgs), glue::ragged_u32(ctxs));
    let (mut pk, mut sig, mut ok) = (vec![0u8; n * pk_len], vec![0u8; n * sig_len], vec![0u8; n]);
    let mut stage_ms = [0f64; 3];
    // SAFETY: buffers sized as the ABI documents for n lanes; sk and iters are not requested.
    let rc = unsafe {
        ffi::mldsa_pipeline_batch_multi(ps, devs.as_ptr(), devs.len() as i32, 0, n, xis.as_ptr().cast(), m.blob.as_ptr(),
                                        m.off.as_ptr(), c.blob.as_ptr(), c.off.as_ptr(), rnd.as_ptr().cast(), pk.as_mut_ptr(),
                                        core::ptr::null_mut(), sig.as_mut_ptr(), ok.as_mut_ptr(), core::ptr::null_mut(),
                                        stage_ms.as_mut_ptr())
    };
    glue::check(rc, "ML-DSA.Sign: ctx too long")?;
    Ok(PipelineOutput { pk, sig, ok: ok.iter().map(|&v| v == 1).collect(), stage_ms })
}
