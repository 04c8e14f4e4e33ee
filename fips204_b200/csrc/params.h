// params.h -- ML-DSA parameter sets as compile-time traits (host + device).
// Mirrors the per-module constants of the reference: src/lib.rs:639-653 (ml_dsa_44), :681-695
// (ml_dsa_65), :723-737 (ml_dsa_87) and the derived sizes of src/lib.rs:129-131.
#pragma once
#include <cstdint>

namespace mldsa {

constexpr int32_t Q = 8380417;        // src/lib.rs:109
constexpr int32_t QINV = 58728449;    // src/helpers.rs:157  (Q * QINV = 1 mod 2^32)
constexpr int D = 13;                 // src/lib.rs:111
constexpr int N = 256;

template <int PS> struct Params;

template <> struct Params<44> {
    static constexpr int PS = 44, K = 4, L = 4, ETA = 2, TAU = 39, CTILDE = 32;
    static constexpr int32_t GAMMA1 = 1 << 17, GAMMA2 = (Q - 1) / 88;
    static constexpr int OMEGA = 80, BETA = TAU * ETA;
    static constexpr int ZBITS = 18, W1BITS = 6, ETABITS = 3;
    static constexpr int PK_LEN = 1312, SK_LEN = 2560, SIG_LEN = 2420, W1_LEN = 768;
};
template <> struct Params<65> {
    static constexpr int PS = 65, K = 6, L = 5, ETA = 4, TAU = 49, CTILDE = 48;
    static constexpr int32_t GAMMA1 = 1 << 19, GAMMA2 = (Q - 1) / 32;
    static constexpr int OMEGA = 55, BETA = TAU * ETA;
    static constexpr int ZBITS = 20, W1BITS = 4, ETABITS = 4;
    static constexpr int PK_LEN = 1952, SK_LEN = 4032, SIG_LEN = 3309, W1_LEN = 768;
};
template <> struct Params<87> {
    static constexpr int PS = 87, K = 8, L = 7, ETA = 2, TAU = 60, CTILDE = 64;
    static constexpr int32_t GAMMA1 = 1 << 19, GAMMA2 = (Q - 1) / 32;
    static constexpr int OMEGA = 75, BETA = TAU * ETA;
    static constexpr int ZBITS = 20, W1BITS = 4, ETABITS = 3;
    static constexpr int PK_LEN = 2592, SK_LEN = 4896, SIG_LEN = 4627, W1_LEN = 1024;
};

// byte offsets inside the canonical encodings (src/encodings.rs:94-152, :238-276)
template <class P> struct Layout {
    static constexpr int SK_S1 = 128;
    static constexpr int SK_S2 = SK_S1 + P::L * 32 * P::ETABITS;
    static constexpr int SK_T0 = SK_S2 + P::K * 32 * P::ETABITS;
    static constexpr int SIG_Z = P::CTILDE;
    static constexpr int SIG_H = SIG_Z + P::L * 32 * P::ZBITS;
    static_assert(SK_T0 + P::K * 32 * D == P::SK_LEN, "sk layout");
    static_assert(SIG_H + P::OMEGA + P::K == P::SIG_LEN, "sig layout");
    static_assert(32 + P::K * 320 == P::PK_LEN, "pk layout");
    static_assert(P::K * 32 * P::W1BITS == P::W1_LEN, "w1 layout");
};

}  // namespace mldsa
