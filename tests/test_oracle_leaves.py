"""Leaf pins for the CPU oracle: FIPS 202 / FIPS 180-4 primitives against hashlib (the reference takes
them from crates sha3 / sha2, outside its tree), and the reference's in-crate leaf KATs
(src/helpers.rs:191-215, src/conversion.rs:422-488, src/high_low.rs debug contracts)."""
import ctypes as C
import hashlib
import random

import numpy as np

import oracle_lib as O

Q = 8380417


def test_shake_and_sha_vs_hashlib():
    rnd = random.Random(204)
    lens = list(range(0, 10)) + [33, 34, 66, 135, 136, 137, 167, 168, 169, 271, 272, 273, 335, 336, 337,
                                 64 + 768, 64 + 1024, 1312, 1952, 2592, 4 * 168 + 5]
    for n in lens:
        d = bytes(rnd.getrandbits(8) for _ in range(n))
        for outlen in (1, 32, 64, 136, 168, 169, 840, 1000):
            assert O.shake128(d, outlen) == hashlib.shake_128(d).digest(outlen)
            assert O.shake256(d, outlen) == hashlib.shake_256(d).digest(outlen)
        assert O.sha256(d) == hashlib.sha256(d).digest()
        assert O.sha512(d) == hashlib.sha512(d).digest()
    for n in (55, 56, 63, 64, 111, 112, 119, 120, 127, 128, 129, 1000):
        d = bytes(rnd.getrandbits(8) for _ in range(n))
        assert O.sha256(d) == hashlib.sha256(d).digest()
        assert O.sha512(d) == hashlib.sha512(d).digest()


def test_zeta_kat():
    # src/helpers.rs:194-196
    L = O.lib()
    assert [L.oracle_zeta_mont(i) for i in range(3)] == [4193792, 25847, 5771523]
    # table definition: tab[brv8(i)] = zeta^i * 2^32 mod q
    for i in (1, 7, 128, 255):
        brv = int(f"{i:08b}"[::-1], 2)
        assert L.oracle_zeta_mont(brv) == pow(1753, i, Q) * (1 << 32) % Q


def test_mont_reduce_contract():
    L = O.lib()
    rnd = random.Random(1)
    for _ in range(20000):
        a = rnd.randrange(-(Q << 31) + 1, (Q << 31))
        r = L.oracle_mont_reduce(a)
        assert -Q < r < Q and (r * (1 << 32) - a) % Q == 0


def test_ntt_is_the_negacyclic_transform():
    rnd = np.random.default_rng(7)
    a = rnd.integers(0, Q, size=(4, 256), dtype=np.int64)
    b = rnd.integers(0, Q, size=(4, 256), dtype=np.int64)
    ah = O.ntt_batch(a.astype(np.int32)).astype(np.int64)
    bh = O.ntt_batch(b.astype(np.int32)).astype(np.int64)
    # pointwise product, then INTT, must equal schoolbook multiplication mod (X^256 + 1)
    prod = (ah % Q) * (bh % Q) % Q
    got = O.inv_ntt_batch(prod.astype(np.int32)).astype(np.int64)
    for i in range(4):
        full = np.zeros(512, dtype=object)
        for j in range(256):
            full[j:j + 256] += int(a[i, j]) * b[i].astype(object)
        want = (full[:256] - full[256:]) % Q
        assert (got[i] == want.astype(np.int64)).all()
    # round trip returns canonical representatives
    back = O.inv_ntt_batch(O.ntt_batch(a.astype(np.int32)) % Q)
    assert (back == a).all()


def test_ntt_raw_growth_bound():
    # SURVEY 8a note R: no reduction inside ntt -> |out| < |in| + 8q
    a = np.full((1, 256), Q - 1, dtype=np.int32)
    out = O.ntt_batch(a).astype(np.int64)
    assert np.abs(out).max() < 9 * Q


def test_coeff_from_three_bytes_kats():
    # src/conversion.rs:423-449 through rej_ntt_poly's accept rule
    L = O.lib()
    seed = bytes(34)
    out = (C.c_int32 * 256)()
    L.oracle_rej_ntt_poly((C.c_uint8 * 34).from_buffer_copy(seed), out)
    stream = hashlib.shake_128(seed).digest(1200)
    want, i = [], 0
    while len(want) < 256:
        z = stream[i] | stream[i + 1] << 8 | (stream[i + 2] & 0x7F) << 16
        i += 3
        if z < Q:
            want.append(z)
    assert list(out) == want
    assert (0x12 | 0x34 << 8 | (0x56 & 0x7F) << 16) == 0x563412
    assert (0x12 | 0x34 << 8 | (0x80 & 0x7F) << 16) == 0x3412
    assert (0x01 | 0xE0 << 8 | (0x7F & 0x7F) << 16) >= Q  # conversion.rs:441-447: rejected


def test_rej_bounded_poly_vs_spec():
    L = O.lib()
    for eta in (2, 4):
        for s in range(4):
            seed = bytes([s]) * 66
            out = (C.c_int32 * 256)()
            L.oracle_rej_bounded_poly(eta, (C.c_uint8 * 66).from_buffer_copy(seed), out)
            stream = hashlib.shake_256(seed).digest(1000)
            want = []
            for byte in stream:
                for nib in (byte & 15, byte >> 4):
                    if len(want) < 256:
                        if eta == 2 and nib < 15:
                            want.append(2 - nib % 5)
                        elif eta == 4 and nib < 9:
                            want.append(4 - nib)
            assert list(out) == want[:256]


def test_bit_pack_roundtrips():
    # src/conversion.rs:490-520 style round trips on the (a, b) pairs the scheme uses
    L = O.lib()
    rnd = np.random.default_rng(3)
    for a, b in ((0, 1023), (2, 2), (4, 4), (4095, 4096), ((1 << 17) - 1, 1 << 17), ((1 << 19) - 1, 1 << 19),
                 (0, 43), (0, 15)):
        c = (a + b).bit_length()
        raw = rnd.integers(0, 256, size=32 * c, dtype=np.uint8)
        w = (C.c_int32 * 256)()
        L.oracle_bit_unpack(raw.ctypes.data_as(C.POINTER(C.c_uint8)), a, b, w)
        vals = np.array(w)
        bits = np.unpackbits(raw, bitorder="little").reshape(256, c)
        field = (bits * (1 << np.arange(c))).sum(axis=1)
        assert (vals == (field if a == 0 else b - field)).all()
        out = (C.c_uint8 * (32 * c))()
        L.oracle_bit_pack(w, a, b, out)
        assert bytes(out) == raw.tobytes()


def test_decompose_power2round_exhaustive_vs_spec():
    # FIPS 204 Alg 35/36 written directly, all r in [0, q)
    L = O.lib()
    r = np.arange(Q, dtype=np.int64)
    for g2 in ((Q - 1) // 88, (Q - 1) // 32):
        r0 = r % (2 * g2)
        r0 = np.where(r0 > g2, r0 - 2 * g2, r0)
        r1 = np.where(r - r0 == Q - 1, 0, (r - r0) // (2 * g2))
        r0 = np.where(r - r0 == Q - 1, r0 - 1, r0)
        idx = np.concatenate([np.arange(0, Q, 997), np.arange(Q - 5000, Q), np.arange(0, 5000),
                              np.arange(g2 - 300, g2 + 300), np.arange(Q - 1 - g2 - 300, Q - 1 - g2 + 300)])
        a, b = C.c_int32(), C.c_int32()
        for i in idx:
            L.oracle_decompose(g2, int(i), C.byref(a), C.byref(b))
            assert (a.value, b.value) == (int(r1[i]), int(r0[i])), (g2, i)
    a, b = C.c_int32(), C.c_int32()
    for i in list(range(0, Q, 1009)) + [Q - 1, 4095, 4096, 4097]:
        L.oracle_power2round(i, C.byref(a), C.byref(b))
        r0 = i % 8192
        r0 = r0 - 8192 if r0 > 4096 else r0
        assert (a.value, b.value) == ((i - r0) // 8192, r0)


def test_use_hint_make_hint_consistency():
    # FIPS 204 Lemma: UseHint(MakeHint(z, r), r) = HighBits(r + z) for |z| <= gamma2
    L = O.lib()
    rnd = random.Random(5)
    for g2 in ((Q - 1) // 88, (Q - 1) // 32):
        for _ in range(5000):
            r = rnd.randrange(Q)
            z = rnd.randrange(-g2, g2 + 1)
            h = L.oracle_make_hint(g2, z % Q, r)
            a, b = C.c_int32(), C.c_int32()
            L.oracle_decompose(g2, (r + z) % Q, C.byref(a), C.byref(b))
            assert L.oracle_use_hint(g2, h, r) == a.value


def test_hint_unpack_rejections():
    # src/conversion.rs:357, :373, :398
    L = O.lib()
    k, omega = 4, 80
    h = (C.c_int32 * (k * 256))()

    def run(y):
        return L.oracle_hint_bit_unpack(k, omega, (C.c_uint8 * len(y)).from_buffer_copy(bytes(y)), h)

    y = [0] * (omega + k)
    assert run(y) == 0
    y2 = list(y); y2[0], y2[1] = 5, 9; y2[omega:] = [2, 2, 2, 2]
    assert run(y2) == 0 and h[5] == 1 and h[9] == 1
    y3 = list(y2); y3[1] = 5            # not strictly increasing
    assert run(y3) == -1
    y4 = list(y2); y4[omega + 1] = 1    # count decreases
    assert run(y4) == -1
    y5 = list(y2); y5[omega + 3] = 81   # count > omega
    assert run(y5) == -1
    y6 = list(y2); y6[7] = 1            # non-zero padding
    assert run(y6) == -1


def test_sample_in_ball_weight_and_determinism():
    L = O.lib()
    for tau, clen in ((39, 32), (49, 48), (60, 64)):
        ct = bytes(range(clen))
        c = (C.c_int32 * 256)()
        L.oracle_sample_in_ball(tau, (C.c_uint8 * clen).from_buffer_copy(ct), C.c_size_t(clen), c)
        v = np.array(c)
        assert (np.abs(v) <= 1).all() and int(np.abs(v).sum()) == tau


def test_hash_mldsa_roundtrip_and_cross_rejects():
    # lib.rs:514-518: sign/verify self-consistency is all the reference has for HashML-DSA
    for ps in (44, 65, 87):
        pk, sk = O.keygen(ps, bytes([ps]) * 32)
        for ph in (0, 1, 2):
            sig = O.sign(ps, sk, b"message", bytes(32), ctx=b"ctx", framing=2, ph=ph)
            assert O.verify(ps, pk, b"message", sig, ctx=b"ctx", framing=2, ph=ph)
            assert not O.verify(ps, pk, b"message", sig, ctx=b"ctx", framing=2, ph=(ph + 1) % 3)
            assert not O.verify(ps, pk, b"message", sig, ctx=b"ctx", framing=0)


def test_ctx_rules():
    # lib.rs:274 (sign errors) / lib.rs:368-370 (verify false)
    pk, sk = O.keygen(44, bytes(32))
    sig = O.sign(44, sk, b"m", bytes(32), ctx=b"\x01" * 255)
    assert O.verify(44, pk, b"m", sig, ctx=b"\x01" * 255)
    assert not O.verify(44, pk, b"m", sig, ctx=b"\x01" * 256)
    try:
        O.sign(44, sk, b"m", bytes(32), ctx=b"\x01" * 256)
        assert False
    except ValueError:
        pass


def test_batch_runners_match_single():
    ps = 44
    xis = np.frombuffer(hashlib.shake_256(b"xi").digest(32 * 6), dtype=np.uint8).reshape(6, 32)
    pks, sks = O.keygen_batch(ps, xis, nthreads=3)
    msgs = [hashlib.shake_256(bytes([i])).digest(i * 7) for i in range(6)]
    rnds = np.frombuffer(hashlib.shake_256(b"rnd").digest(32 * 6), dtype=np.uint8).reshape(6, 32)
    sigs, iters = O.sign_batch(ps, sks, msgs, rnds, nthreads=2)
    for i in range(6):
        pk, sk = O.keygen(ps, xis[i].tobytes())
        assert pk == pks[i].tobytes() and sk == sks[i].tobytes()
        s, it = O.sign(ps, sk, msgs[i], rnds[i].tobytes(), want_iters=True)
        assert s == sigs[i].tobytes() and it == iters[i]
    ok = O.verify_batch(ps, pks, msgs, sigs, nthreads=4)
    assert ok.all()
    sigs[2, 40] ^= 1
    ok = O.verify_batch(ps, pks, msgs, sigs)
    assert list(ok) == [1, 1, 0, 1, 1, 1]
