//! The reference's in-crate smoke test (`src/lib.rs:503-551`) and its external-interface KAT (`tests/messages.rs`)
//! against this crate.  Needs a CUDA device; NOT run in the image this crate was written in (no Rust toolchain).
use fips204_b200::traits::{KeyGen, SerDes, Signer, Verifier};
use fips204_b200::{batch, ml_dsa_44, ml_dsa_65, ml_dsa_87, Ph};
use rand_chacha::rand_core::SeedableRng;

#[test]
fn smoke_all_parameter_sets() {
    let mut rng = rand_chacha::ChaCha8Rng::seed_from_u64(123);
    let (m1, m2) = ([0u8, 1, 2, 3, 4, 5, 6, 7], [7u8, 7, 7, 7, 7, 7, 7, 7]);
    for _ in 0..8 {
        let (pk, sk) = ml_dsa_44::try_keygen_with_rng(&mut rng).unwrap();
        let sig = sk.try_sign_with_rng(&mut rng, &m1, &[]).unwrap();
        assert!(pk.verify(&m1, &sig, &[]));
        assert!(!pk.verify(&m2, &sig, &[]));
        assert!(sk.try_sign_with_rng(&mut rng, &m1, &[0u8; 256]).is_err());
        assert!(!pk.verify(&m1, &sig, &[0u8; 256]));
        for ph in [Ph::SHA256, Ph::SHA512, Ph::SHAKE128] {
            let hs = sk.try_hash_sign_with_rng(&mut rng, &m1, &[], &ph).unwrap();
            assert!(pk.hash_verify(&m1, &hs, &[], &ph));
        }
        assert_eq!(pk.clone().into_bytes(), sk.get_public_key().into_bytes());
    }
    // first byte of the public key of xi = 0x11.. (src/lib.rs:536-545)
    let xi = [0x11u8; 32];
    assert_eq!(ml_dsa_44::KG::keygen_from_seed(&xi).0.into_bytes()[0], 197);
    assert_eq!(ml_dsa_65::KG::keygen_from_seed(&xi).0.into_bytes()[0], 177);
    assert_eq!(ml_dsa_87::KG::keygen_from_seed(&xi).0.into_bytes()[0], 16);
}

#[test]
fn batch_equals_scalar_and_key_sets_equal_packed_keys() {
    use batch::{BatchKeyGen, BatchSigner, BatchVerifier};
    let n = 1000usize;
    let xis: Vec<[u8; 32]> = (0..n).map(|i| [i as u8; 32]).collect();
    let keys = ml_dsa_65::KG::keygen_batch_from_seeds(&xis, &[0]).unwrap();
    let msgs: Vec<Vec<u8>> = (0..n).map(|i| vec![i as u8; i % 40]).collect();
    let m: Vec<&[u8]> = msgs.iter().map(Vec::as_slice).collect();
    let c: Vec<&[u8]> = vec![&[]; n];
    let rnd = vec![[7u8; 32]; n];
    let sks: Vec<&ml_dsa_65::PrivateKey> = keys.iter().map(|k| &k.1).collect();
    let pks: Vec<&ml_dsa_65::PublicKey> = keys.iter().map(|k| &k.0).collect();
    let sigs = ml_dsa_65::PrivateKey::try_sign_batch_with_seeds(&sks, &m, &c, &rnd, &[0]).unwrap();
    assert_eq!(sigs[5], keys[5].1.try_sign_with_seed(&rnd[5], m[5], &[]).unwrap());
    assert!(ml_dsa_65::PublicKey::verify_batch(&pks, &m, &sigs, &c, &[0]).iter().all(|&v| v));
    let (pk_bytes, sset, pset) = batch::keygen_to_keysets(65, 0, &xis).unwrap();
    assert_eq!(&pk_bytes[..ml_dsa_65::PK_LEN], &keys[0].0.clone().into_bytes()[..]);
    let s2 = sset.try_sign_batch_with_seeds(None, &m, &c, &rnd, None).unwrap();
    assert!(s2.iter().zip(&sigs).all(|(a, b)| a[..] == b[..]));
    let flat: Vec<u8> = s2.concat();
    assert!(pset.verify_batch(None, &m, &flat, &c, None).iter().all(|&v| v));
    let out = batch::pipeline(65, &[0], &xis, &m, &c, &rnd).unwrap();
    assert!(out.ok.iter().all(|&v| v) && out.sig == flat && out.pk == pk_bytes);
}
