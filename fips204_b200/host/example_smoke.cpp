// Mirrors the reference's in-crate smoke test (src/lib.rs:503-551) through the C++ host mirror.
//   g++ -std=c++17 -O2 example_smoke.cpp -L.. -lmldsa_b200 -Wl,-rpath,'$ORIGIN/..' -o example_smoke
#include <algorithm>
#include <cstdio>

#include "fips204.hpp"

template <class M> static int run(const char* name, int want_pk0)
{
    using namespace fips204;
    std::array<uint8_t, 32> xi;
    xi.fill(0x11);
    auto [pk, sk] = M::KG::keygen_from_seed(xi);  // lib.rs:535
    std::vector<uint8_t> m1{0, 1, 2, 3, 4, 5, 6, 7}, m2{7, 7, 7, 7, 7, 7, 7, 7};
    std::array<uint8_t, 32> seed;
    seed.fill(12);
    auto sig = sk.try_sign_with_seed(seed, m1, {});
    bool ok = pk.verify(m1, sig, {}) && !pk.verify(m2, sig, {}) && !pk.verify(m1, sig, std::vector<uint8_t>(257, 0));
    bool threw = false;
    try {
        sk.try_sign_with_seed(seed, m1, std::vector<uint8_t>(257, 0));
    } catch (const Error& e) {
        threw = std::string(e.what()) == "ML-DSA.Sign: ctx too long";
    }
    ok = ok && threw && pk.into_bytes()[0] == want_pk0;  // lib.rs:543-545
    // the same key expanded once and kept on the device signs the same bytes
    typename M::PrivateKeySet set(1, sk.bytes.data());
    std::vector<std::vector<uint8_t>> msgs{m1, m2, m1};
    const uint32_t idx[3] = {0, 0, 0};
    std::vector<uint8_t> rnd(3 * 32, 12);
    auto sigs = set.sign_batch(msgs, nullptr, rnd.data(), idx);
    ok = ok && set.size() == 1 && sigs[0] == sig && sigs[2] == sig && pk.verify(m2, sigs[1], {});
    // ... and the expanded public key verifies them; the signature of m1 is not one of m2
    typename M::PublicKeySet pset(1, pk.bytes.data());
    std::vector<uint8_t> flat;
    for (auto& s : sigs) flat.insert(flat.end(), s.begin(), s.end());
    auto v = pset.verify_batch(msgs, nullptr, flat.data(), idx);
    std::vector<std::vector<uint8_t>> swapped{m2, m2, m1};
    auto w = pset.verify_batch(swapped, nullptr, flat.data(), idx);
    ok = ok && v[0] && v[1] && v[2] && !w[0] && w[1] && w[2];
    // keys provisioned from seeds straight into resident sets, and the chained keygen -> sign -> verify pipeline
    std::vector<uint8_t> xis(3 * 32, 0x11), pkb(3 * M::PK_LEN);
    auto sets = M::keygen_to_keysets(3, xis.data(), pkb.data());
    auto sigs2 = sets.first.sign_batch(msgs, nullptr, rnd.data());
    auto v2 = sets.second.verify_batch(msgs, nullptr, flat.data());
    ok = ok && std::equal(pk.bytes.begin(), pk.bytes.end(), pkb.begin()) && sigs2[0] == sig && sigs2[1] == sigs[1] && v2[0] && v2[1];
    auto pr = M::pipeline_batch(3, xis.data(), msgs, rnd.data());
    ok = ok && pr.ok[0] && pr.ok[1] && pr.ok[2] && pr.pk == pkb && std::equal(sig.begin(), sig.end(), pr.sig.begin());
    std::printf("%s: %s\n", name, ok ? "ok" : "FAILED");
    return ok ? 0 : 1;
}

int main()
{
    try {
        return run<fips204::ml_dsa_44>("ml_dsa_44", 197) | run<fips204::ml_dsa_65>("ml_dsa_65", 177) |
               run<fips204::ml_dsa_87>("ml_dsa_87", 16);
    } catch (const fips204::Error& e) {
        std::printf("error %d: %s\n", e.code, e.what());
        return 2;
    }
}
