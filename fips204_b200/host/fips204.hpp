// fips204.hpp -- C++17 host mirror of the public surface of the Rust crate integritychain/fips204
// (src/lib.rs:116-614, src/traits.rs) on top of the C ABI of include/mldsa_b200.h.
//
// Same names, argument meaning and error behaviour as the reference, so code and tests written
// against the crate read the same here:
//
//     using namespace fips204;
//     auto [pk, sk] = ml_dsa_65::try_keygen_with_rng(rng);          // lib.rs:230
//     auto sig      = sk.try_sign_with_rng(rng, msg, ctx);          // lib.rs:268   (throws on |ctx| > 255)
//     bool ok       = pk.verify(msg, sig, ctx);                     // lib.rs:364   (never throws)
//     auto bytes    = sk.into_bytes();                              // lib.rs:427
//     auto sk2      = ml_dsa_65::PrivateKey::try_from_bytes(bytes); // lib.rs:420
//
// plus the batch entry points (keygen_batch / sign_batch / verify_batch) that the engine exists for.
// Result<T, &'static str> becomes an exception of type fips204::Error carrying the reference's message.
// Keys hold canonical bytes; the expanded-struct caches of the reference (src/types.rs:19-41) live on
// the device for the duration of a call.  Single operations are batches of one; there is no CPU path.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <random>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/mldsa_b200.h"

namespace fips204 {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// rand_core::CryptoRngCore analogue: anything with `bool try_fill_bytes(uint8_t*, size_t)`
struct OsRng {
    bool try_fill_bytes(uint8_t* out, size_t n)
    {
        std::random_device rd;
        for (size_t i = 0; i < n; i++) out[i] = static_cast<uint8_t>(rd());
        return true;
    }
};

inline void check(int rc)
{
    if (rc == MLDSA_OK) return;
    std::string m = mldsa_strerror(rc);
    if (rc == MLDSA_E_CUDA) m += std::string(": ") + mldsa_last_cuda_error();
    throw Error(rc, m);
}

template <int PS, size_t PK, size_t SK, size_t SIG>
struct ParamSet {
    static constexpr size_t PK_LEN = PK, SK_LEN = SK, SIG_LEN = SIG;  // lib.rs:649-653
    using Signature = std::array<uint8_t, SIG>;
    using Bytes = std::vector<uint8_t>;

    // ---- batch entry points ------------------------------------------------------------------
    static void keygen_batch(size_t n, const uint8_t* xi, uint8_t* pk, uint8_t* sk, int dev = 0)
    {
        check(mldsa_keygen_batch(PS, dev, n, xi, pk, sk));
    }
    // msgs: n byte strings; rnd: n x 32; sk: n x SK_LEN, or one key when shared_key
    static std::vector<Signature> sign_batch(const uint8_t* sk, bool shared_key, const std::vector<Bytes>& msgs,
                                             const std::vector<Bytes>* ctxs, const uint8_t* rnd, int framing = 0,
                                             int dev = 0, std::vector<uint32_t>* iters = nullptr)
    {
        const size_t n = msgs.size();
        std::vector<uint8_t> mb, cb;
        std::vector<uint64_t> mo(n + 1, 0);
        std::vector<uint32_t> co(n + 1, 0);
        for (size_t i = 0; i < n; i++) {
            mb.insert(mb.end(), msgs[i].begin(), msgs[i].end());
            mo[i + 1] = mb.size();
            if (ctxs) {
                if ((*ctxs)[i].size() > 255) throw Error(MLDSA_E_CTX, "ML-DSA.Sign: ctx too long");  // lib.rs:274
                cb.insert(cb.end(), (*ctxs)[i].begin(), (*ctxs)[i].end());
                co[i + 1] = static_cast<uint32_t>(cb.size());
            }
        }
        mb.push_back(0);
        cb.push_back(0);
        std::vector<Signature> sig(n);
        if (iters) iters->assign(n, 0);
        check(mldsa_sign_batch(PS, dev, framing, n, sk, shared_key ? 0 : SK, mb.data(), mo.data(),
                               ctxs ? cb.data() : nullptr, ctxs ? co.data() : nullptr, rnd,
                               reinterpret_cast<uint8_t*>(sig.data()), iters ? iters->data() : nullptr));
        return sig;
    }
    static std::vector<uint8_t> verify_batch(const uint8_t* pk, bool shared_key, const std::vector<Bytes>& msgs,
                                             const std::vector<Bytes>* ctxs, const uint8_t* sigs, int framing = 0,
                                             int dev = 0)
    {
        const size_t n = msgs.size();
        std::vector<uint8_t> mb, cb, ok(n, 0);
        std::vector<uint64_t> mo(n + 1, 0);
        std::vector<uint32_t> co(n + 1, 0);
        for (size_t i = 0; i < n; i++) {
            mb.insert(mb.end(), msgs[i].begin(), msgs[i].end());
            mo[i + 1] = mb.size();
            if (ctxs) {
                cb.insert(cb.end(), (*ctxs)[i].begin(), (*ctxs)[i].end());
                co[i + 1] = static_cast<uint32_t>(cb.size());
            }
        }
        mb.push_back(0);
        cb.push_back(0);
        check(mldsa_verify_batch(PS, dev, framing, n, pk, shared_key ? 0 : PK, mb.data(), mo.data(),
                                 ctxs ? cb.data() : nullptr, ctxs ? co.data() : nullptr, sigs, ok.data()));
        return ok;
    }

    // ---- expanded secret keys resident on the device ----------------------------------------------
    // What the reference's PrivateKey struct holds (rho, K, tr, NTT(s1), NTT(s2), NTT(t0): types.rs:19-28, built once by
    // try_from_bytes -> expand_private, lib.rs:462-468, ml_dsa.rs:445-469), for n keys, plus A-hat.  Zeroized on
    // destruction (ZeroizeOnDrop).  sign_batch: lane i signs msgs[i] with key key_idx[i] (nullptr: key i).
    struct PrivateKeySet {
        mldsa_keyset* handle = nullptr;
        PrivateKeySet(size_t n, const uint8_t* sk, int dev = 0) { check(mldsa_sk_load(PS, dev, n, sk, &handle)); }
        explicit PrivateKeySet(mldsa_keyset* adopted) : handle(adopted) {}  // from keygen_to_keysets
        PrivateKeySet(PrivateKeySet&& o) noexcept : handle(o.handle) { o.handle = nullptr; }
        ~PrivateKeySet() { mldsa_keyset_free(handle); }
        PrivateKeySet(const PrivateKeySet&) = delete;
        PrivateKeySet& operator=(const PrivateKeySet&) = delete;
        size_t size() const { return mldsa_keyset_size(handle); }
        std::vector<Signature> sign_batch(const std::vector<Bytes>& msgs, const std::vector<Bytes>* ctxs, const uint8_t* rnd,
                                          const uint32_t* key_idx = nullptr, int framing = 0,
                                          std::vector<uint32_t>* iters = nullptr) const
        {
            const size_t n = msgs.size();
            std::vector<uint8_t> mb, cb;
            std::vector<uint64_t> mo(n + 1, 0);
            std::vector<uint32_t> co(n + 1, 0);
            for (size_t i = 0; i < n; i++) {
                mb.insert(mb.end(), msgs[i].begin(), msgs[i].end());
                mo[i + 1] = mb.size();
                if (ctxs) {
                    if ((*ctxs)[i].size() > 255) throw Error(MLDSA_E_CTX, "ML-DSA.Sign: ctx too long");  // lib.rs:274
                    cb.insert(cb.end(), (*ctxs)[i].begin(), (*ctxs)[i].end());
                    co[i + 1] = static_cast<uint32_t>(cb.size());
                }
            }
            mb.push_back(0);
            cb.push_back(0);
            std::vector<Signature> sig(n);
            if (iters) iters->assign(n, 0);
            check(mldsa_sign_batch_keys(handle, framing, n, key_idx, mb.data(), mo.data(), ctxs ? cb.data() : nullptr,
                                        ctxs ? co.data() : nullptr, rnd, reinterpret_cast<uint8_t*>(sig.data()),
                                        iters ? iters->data() : nullptr));
            return sig;
        }
    };

    // Expanded public keys resident on the device: rho-derived A-hat, tr, NTT(t1 * 2^d) per key -- the fields of the
    // reference's PublicKey (types.rs:35-41, built by expand_public, ml_dsa.rs:477-498).  verify_batch: lane i is
    // checked under key key_idx[i] (nullptr: key i); never throws for malformed signatures (lib.rs:368-370).
    struct PublicKeySet {
        mldsa_keyset* handle = nullptr;
        PublicKeySet(size_t n, const uint8_t* pk, int dev = 0) { check(mldsa_pk_load(PS, dev, n, pk, &handle)); }
        explicit PublicKeySet(mldsa_keyset* adopted) : handle(adopted) {}  // from keygen_to_keysets
        PublicKeySet(PublicKeySet&& o) noexcept : handle(o.handle) { o.handle = nullptr; }
        ~PublicKeySet() { mldsa_keyset_free(handle); }
        PublicKeySet(const PublicKeySet&) = delete;
        PublicKeySet& operator=(const PublicKeySet&) = delete;
        size_t size() const { return mldsa_keyset_size(handle); }
        std::vector<uint8_t> verify_batch(const std::vector<Bytes>& msgs, const std::vector<Bytes>* ctxs, const uint8_t* sigs,
                                          const uint32_t* key_idx = nullptr, int framing = 0) const
        {
            const size_t n = msgs.size();
            std::vector<uint8_t> mb, cb, ok(n, 0);
            std::vector<uint64_t> mo(n + 1, 0);
            std::vector<uint32_t> co(n + 1, 0);
            for (size_t i = 0; i < n; i++) {
                mb.insert(mb.end(), msgs[i].begin(), msgs[i].end());
                mo[i + 1] = mb.size();
                if (ctxs) {
                    cb.insert(cb.end(), (*ctxs)[i].begin(), (*ctxs)[i].end());
                    co[i + 1] = static_cast<uint32_t>(cb.size());
                }
            }
            mb.push_back(0);
            cb.push_back(0);
            check(mldsa_verify_batch_keys(handle, framing, n, key_idx, mb.data(), mo.data(), ctxs ? cb.data() : nullptr,
                                          ctxs ? co.data() : nullptr, sigs, ok.data()));
            return ok;
        }
    };

    // ---- the reference's types ----------------------------------------------------------------
    // KG::keygen_from_seed for n seeds whose expanded keys stay on the device (mldsa_keygen_to_keyset): the packed
    // public keys come back (pk: n x PK_LEN, may be nullptr), no packed secret key ever crosses PCIe.
    static std::pair<PrivateKeySet, PublicKeySet> keygen_to_keysets(size_t n, const uint8_t* xi, uint8_t* pk, int dev = 0)
    {
        mldsa_keyset *s = nullptr, *p = nullptr;
        check(mldsa_keygen_to_keyset(PS, dev, n, xi, pk, &s, &p));
        return {PrivateKeySet(s), PublicKeySet(p)};
    }
    // keygen -> sign -> verify chained on the device (the flow of lib.rs:503-551 as one batch call, BASELINE config 5)
    struct PipelineResult {
        std::vector<uint8_t> pk, sig, ok;
        double stage_ms[3] = {0, 0, 0};
    };
    static PipelineResult pipeline_batch(size_t n, const uint8_t* xi, const std::vector<Bytes>& msgs, const uint8_t* rnd,
                                         const std::vector<int>& devs = {0}, int framing = 0)
    {
        if (msgs.size() != n) throw Error(MLDSA_E_PARAM, "pipeline_batch: msgs must have n entries");
        std::vector<uint8_t> mb;
        std::vector<uint64_t> mo(n + 1, 0);
        for (size_t i = 0; i < n; i++) {
            mb.insert(mb.end(), msgs[i].begin(), msgs[i].end());
            mo[i + 1] = mb.size();
        }
        if (mb.empty()) mb.push_back(0);
        PipelineResult r;
        r.pk.resize(n * PK);
        r.sig.resize(n * SIG);
        r.ok.assign(n, 0);
        check(mldsa_pipeline_batch_multi(PS, devs.data(), static_cast<int>(devs.size()), framing, n, xi, mb.data(), mo.data(),
                                         nullptr, nullptr, rnd, r.pk.data(), nullptr, r.sig.data(), r.ok.data(), nullptr,
                                         r.stage_ms));
        return r;
    }

    struct PublicKey {  // types.rs:35-41 (here: canonical bytes)
        std::array<uint8_t, PK> bytes{};
        static PublicKey try_from_bytes(const std::array<uint8_t, PK>& b) { return PublicKey{b}; }  // lib.rs:469
        std::array<uint8_t, PK> into_bytes() const { return bytes; }                                // lib.rs:477
        // Verifier::verify, lib.rs:364-380: |ctx| > 255 -> false, never an error
        bool verify(const Bytes& message, const Signature& sig, const Bytes& ctx = {}) const
        {
            if (ctx.size() > 255) return false;
            std::vector<Bytes> m{message}, c{ctx};
            return verify_batch(bytes.data(), true, m, &c, sig.data())[0] != 0;
        }
    };

    struct PrivateKey {  // types.rs:19-28 (here: canonical bytes, zeroed on destruction like ZeroizeOnDrop)
        std::array<uint8_t, SK> bytes{};
        ~PrivateKey()
        {
            volatile uint8_t* p = bytes.data();
            for (size_t i = 0; i < SK; i++) p[i] = 0;
        }
        static PrivateKey try_from_bytes(const std::array<uint8_t, SK>& b) { return PrivateKey{b}; }  // lib.rs:420
        std::array<uint8_t, SK> into_bytes() const { return bytes; }                                  // lib.rs:427
        // Signer::try_sign_with_rng, lib.rs:268-296: ctx check, ONE 32-byte draw, sign
        template <class Rng>
        Signature try_sign_with_rng(Rng& rng, const Bytes& message, const Bytes& ctx = {}) const
        {
            if (ctx.size() > 255) throw Error(MLDSA_E_CTX, "ML-DSA.Sign: ctx too long");
            std::array<uint8_t, 32> rnd{};
            if (!rng.try_fill_bytes(rnd.data(), 32))
                throw Error(MLDSA_E_PARAM, "ML-DSA.Sign: random number generator failed");  // lib.rs:283
            return try_sign_with_seed(rnd, message, ctx);
        }
        Signature try_sign(const Bytes& message, const Bytes& ctx = {}) const  // traits.rs:156
        {
            OsRng rng;
            return try_sign_with_rng(rng, message, ctx);
        }
        // traits.rs:228-232: rnd := seed (DummyRng); an all-zero seed is the deterministic variant
        Signature try_sign_with_seed(const std::array<uint8_t, 32>& seed, const Bytes& message,
                                     const Bytes& ctx = {}) const
        {
            if (ctx.size() > 255) throw Error(MLDSA_E_CTX, "ML-DSA.Sign: ctx too long");
            std::vector<Bytes> m{message}, c{ctx};
            return sign_batch(bytes.data(), true, m, &c, seed.data())[0];
        }
    };

    struct KG {  // KeyGen trait, traits.rs:8-114
        static std::pair<PublicKey, PrivateKey> keygen_from_seed(const std::array<uint8_t, 32>& xi)  // lib.rs:247
        {
            PublicKey pk;
            PrivateKey sk;
            keygen_batch(1, xi.data(), pk.bytes.data(), sk.bytes.data());
            return {pk, sk};
        }
        template <class Rng>
        static std::pair<PublicKey, PrivateKey> try_keygen_with_rng(Rng& rng)  // lib.rs:241, ml_dsa.rs:40-41
        {
            std::array<uint8_t, 32> xi{};
            if (!rng.try_fill_bytes(xi.data(), 32)) throw Error(MLDSA_E_PARAM, "KeyGen: Random number generator failed");
            return keygen_from_seed(xi);
        }
        static std::pair<PublicKey, PrivateKey> try_keygen()
        {
            OsRng rng;
            return try_keygen_with_rng(rng);
        }
    };
    template <class Rng> static std::pair<PublicKey, PrivateKey> try_keygen_with_rng(Rng& rng)  // lib.rs:230
    {
        return KG::try_keygen_with_rng(rng);
    }
    static std::pair<PublicKey, PrivateKey> try_keygen() { return KG::try_keygen(); }  // lib.rs:195

    // ACVP "internal" interface, lib.rs:586-612
    static Signature _internal_sign(const PrivateKey& sk, const Bytes& message, const Bytes& ctx,
                                    const std::array<uint8_t, 32>& rnd)
    {
        if (ctx.size() > 255) throw Error(MLDSA_E_CTX, "_internal_sign: ctx too long");
        std::vector<Bytes> m{message};
        return sign_batch(sk.bytes.data(), true, m, nullptr, rnd.data(), /*framing=*/1)[0];
    }
    static bool _internal_verify(const PublicKey& pk, const Bytes& message, const Signature& sig, const Bytes& ctx = {})
    {
        if (ctx.size() > 255) return false;
        std::vector<Bytes> m{message};
        return verify_batch(pk.bytes.data(), true, m, nullptr, sig.data(), /*framing=*/1)[0] != 0;
    }
};

using ml_dsa_44 = ParamSet<44, 1312, 2560, 2420>;  // lib.rs:639-653
using ml_dsa_65 = ParamSet<65, 1952, 4032, 3309>;  // lib.rs:681-695
using ml_dsa_87 = ParamSet<87, 2592, 4896, 4627>;  // lib.rs:723-737

}  // namespace fips204
