//! The four traits of the reference crate (`src/traits.rs:8-424`) with the same names, associated types, method
//! names, argument order and error type, so that code written against `fips204::traits::*` compiles against this crate
//! by changing the `use` line.  Provided methods have the reference's semantics: `try_keygen` / `try_sign` /
//! `try_hash_sign` draw from the OS RNG (feature `default-rng`), the `*_with_seed` forms replay a fixed 32-byte value
//! as the one draw the operation takes.
use crate::Ph;
use rand_core::CryptoRngCore;
#[cfg(feature = "default-rng")]
use rand_core::OsRng;

/// Key generation for one security parameter set (reference `src/traits.rs:8-114`).
pub trait KeyGen {
    /// A public key specific to the chosen security parameter set
    type PublicKey;
    /// A private (secret) key specific to the chosen security parameter set
    type PrivateKey;

    /// Generates a key pair with the operating system's RNG.
    /// # Errors
    /// `"KeyGen: Random number generator failed"` when the RNG fails.
    #[cfg(feature = "default-rng")]
    fn try_keygen() -> Result<(Self::PublicKey, Self::PrivateKey), &'static str> {
        Self::try_keygen_with_rng(&mut OsRng)
    }

    /// Generates a key pair; takes exactly one 32-byte draw from `rng` (`src/ml_dsa.rs:40-41`).
    /// # Errors
    /// `"KeyGen: Random number generator failed"` when the RNG fails.
    fn try_keygen_with_rng(rng: &mut impl CryptoRngCore) -> Result<(Self::PublicKey, Self::PrivateKey), &'static str>;

    /// Generates the key pair of the seed `xi` (`ML-DSA.KeyGen_internal`).
    #[must_use]
    fn keygen_from_seed(xi: &[u8; 32]) -> (Self::PublicKey, Self::PrivateKey);
}

/// Replays a fixed 32-byte value as the single draw of a deterministic-seed operation (the reference's `DummyRng`,
/// `src/traits.rs:310-326`).
struct SeedRng([u8; 32]);

impl rand_core::RngCore for SeedRng {
    fn next_u32(&mut self) -> u32 {
        unimplemented!()
    }
    fn next_u64(&mut self) -> u64 {
        unimplemented!()
    }
    fn fill_bytes(&mut self, _out: &mut [u8]) {
        unimplemented!()
    }
    fn try_fill_bytes(&mut self, out: &mut [u8]) -> Result<(), rand_core::Error> {
        out.copy_from_slice(&self.0);
        Ok(())
    }
}
impl rand_core::CryptoRng for SeedRng {}

/// Signing with a private key (reference `src/traits.rs:118-308`).
pub trait Signer {
    /// The signature is specific to the chosen security parameter set
    type Signature;
    /// The public key that corresponds to the private/secret key
    type PublicKey;

    /// Signs `message` under `ctx` (at most 255 bytes) with the operating system's RNG.
    /// # Errors
    /// `"ML-DSA.Sign: ctx too long"`, `"ML-DSA.Sign: random number generator failed"`.
    #[cfg(feature = "default-rng")]
    fn try_sign(&self, message: &[u8], ctx: &[u8]) -> Result<Self::Signature, &'static str> {
        self.try_sign_with_rng(&mut OsRng, message, ctx)
    }

    /// Signs `message` under `ctx`; takes exactly one 32-byte draw from `rng` (`src/lib.rs:282-283`).
    /// # Errors
    /// As [`Signer::try_sign`].
    fn try_sign_with_rng(&self, rng: &mut impl CryptoRngCore, message: &[u8], ctx: &[u8]) -> Result<Self::Signature, &'static str>;

    /// Signs with `seed` as the value of the draw (all-zero = the deterministic variant).
    /// # Errors
    /// As [`Signer::try_sign`].
    fn try_sign_with_seed(&self, seed: &[u8; 32], message: &[u8], ctx: &[u8]) -> Result<Self::Signature, &'static str> {
        self.try_sign_with_rng(&mut SeedRng(*seed), message, ctx)
    }

    /// HashML-DSA: signs the pre-hash `ph` of `message` with the operating system's RNG.
    /// # Errors
    /// `"HashML-DSA.Sign: ctx too long"`, `"HashML-DSA.Sign: random number generator failed"`.
    #[cfg(feature = "default-rng")]
    fn try_hash_sign(&self, message: &[u8], ctx: &[u8], ph: &Ph) -> Result<Self::Signature, &'static str> {
        self.try_hash_sign_with_rng(&mut OsRng, message, ctx, ph)
    }

    /// HashML-DSA with a caller-supplied RNG (`src/lib.rs:310-342`).
    /// # Errors
    /// As [`Signer::try_hash_sign`].
    fn try_hash_sign_with_rng(&self, rng: &mut impl CryptoRngCore, message: &[u8], ctx: &[u8], ph: &Ph)
                              -> Result<Self::Signature, &'static str>;

    /// HashML-DSA with `seed` as the value of the draw.
    /// # Errors
    /// As [`Signer::try_hash_sign`].
    fn try_hash_sign_with_seed(&self, seed: &[u8; 32], message: &[u8], ctx: &[u8], ph: &Ph) -> Result<Self::Signature, &'static str> {
        self.try_hash_sign_with_rng(&mut SeedRng(*seed), message, ctx, ph)
    }

    /// The public key of this private key (`src/ml_dsa.rs:502-563`).
    fn get_public_key(&self) -> Self::PublicKey;
}

/// Verification with a public key (reference `src/traits.rs:330-362`).
pub trait Verifier {
    /// The signature is specific to the chosen security parameter set
    type Signature;

    /// `true` iff `signature` is valid for `message` under `ctx`.  Never fails: malformed signatures and contexts
    /// longer than 255 bytes yield `false` (`src/lib.rs:368-370`).
    fn verify(&self, message: &[u8], signature: &Self::Signature, ctx: &[u8]) -> bool;

    /// HashML-DSA verification with pre-hash function `ph` (`src/lib.rs:391-411`).
    fn hash_verify(&self, message: &[u8], sig: &Self::Signature, ctx: &[u8], ph: &Ph) -> bool;
}

/// Serialization of keys to and from their canonical byte arrays (reference `src/traits.rs:372-424`).
pub trait SerDes {
    /// Correctly sized byte array of the structure
    type ByteArray;

    /// The canonical encoding.
    fn into_bytes(self) -> Self::ByteArray;

    /// Deserializes a canonical encoding.
    /// # Errors
    /// Kept in the signature for source compatibility; every correctly sized array deserializes (the reference's
    /// range checks cannot fail, `src/conversion.rs:259-260`).
    fn try_from_bytes(ba: Self::ByteArray) -> Result<Self, &'static str>
    where
        Self: Sized;
}
