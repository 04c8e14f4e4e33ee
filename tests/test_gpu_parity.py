"""GPU parity tests: the CUDA engine, called through the C ABI, against the CPU oracle and the
reference's golden vectors.  Bit-exact everywhere (integer / byte work)."""
import hashlib

import numpy as np
import pytest

import oracle_lib as O
from conftest import PS_OF
from synth import synth, synth_block, synth_rows

pytestmark = pytest.mark.gpu
Q = 8380417


@pytest.fixture(scope="module")
def F():
    import fips204_b200 as F
    assert F.device_count() >= 1, "no CUDA device: GPU tests need the B200 box"
    return F


# ---------------------------------------------------------------- leaves
def test_shake_rows_vs_hashlib(F):
    for bits, fn in ((128, hashlib.shake_128), (256, hashlib.shake_256)):
        for in_len in (1, 7, 8, 33, 34, 66, 135, 136, 137, 167, 168, 169, 272, 1312, 1952):
            rows = synth_block(f"shake{bits}-{in_len}", 37, in_len)
            for out_len in (32, 64, 136, 200, 840):
                got = F.micro.shake(bits, rows, out_len)
                for i in range(rows.shape[0]):
                    assert got[i].tobytes() == fn(rows[i].tobytes()).digest(out_len), (bits, in_len, out_len, i)


def test_ntt_raw_bit_exact(F):
    n = 4096
    raw = synth_block("ntt", n, 1024).view(np.uint32).astype(np.int64) % Q
    polys = raw.astype(np.int32).reshape(n, 256)
    got = F.micro.ntt(polys)
    want = O.ntt_batch(polys)
    assert (got == want).all(), f"{(got != want).sum()} coefficient mismatches"
    # extreme inputs: q-1 everywhere and alternating signs near the reference's documented growth bound
    edge = np.stack([np.full(256, Q - 1), np.zeros(256), np.arange(256), -np.arange(256),
                     np.where(np.arange(256) % 2 == 0, Q - 1, -(Q - 1))]).astype(np.int32)
    assert (F.micro.ntt(edge) == O.ntt_batch(edge)).all()


def test_intt_canonical(F):
    n = 4097  # odd: the kernels take two polynomials per warp, the last pair is half empty
    raw = synth_block("intt", n, 1024).view(np.uint32).astype(np.int64) % Q
    polys = raw.astype(np.int32).reshape(n, 256)
    got = F.micro.inv_ntt(polys)
    want = O.inv_ntt_batch(polys)
    assert (got == want).all()
    assert got.min() >= 0 and got.max() < Q
    # round trip through both device transforms
    back = F.micro.inv_ntt((F.micro.ntt(polys).astype(np.int64) % Q).astype(np.int32))
    assert (back == polys).all()
    one = polys[:1]
    assert (F.micro.inv_ntt(one) == O.inv_ntt_batch(one)).all() and (F.micro.ntt(one) == O.ntt_batch(one)).all()


def test_expand_a_exact(F):
    n = 2048 + 5  # ragged tail: not a multiple of the block size
    seeds = synth_block("rho34", n, 34)
    got = F.micro.expand_a(seeds)
    want = O.expand_a_batch(seeds)
    assert (got == want).all()


# ---------------------------------------------------------------- keygen
@pytest.mark.parametrize("ps", [44, 65, 87])
def test_keygen_acvp(F, acvp_keygen, ps):
    m = F.PARAM_SETS[ps]
    vec = [t for t in acvp_keygen if PS_OF[t["ps"]] == ps]
    assert len(vec) == 25
    xi = b"".join(bytes.fromhex(t["seed"]) for t in vec)
    pk, sk = m.keygen_batch(xi)
    for i, t in enumerate(vec):
        assert pk[i].tobytes().hex().upper() == t["pk"].upper(), ("pk", t["tc"])
        assert sk[i].tobytes().hex().upper() == t["sk"].upper(), ("sk", t["tc"])


@pytest.mark.parametrize("ps", [44, 65, 87])
def test_keygen_vs_oracle(F, ps):
    m = F.PARAM_SETS[ps]
    n = 1000 + 3
    xi = synth_block(f"xi{ps}", n, 32)
    pk, sk = m.keygen_batch(xi)
    opk, osk = O.keygen_batch(ps, xi)
    assert (pk == opk).all()
    assert (sk == osk).all()


def test_keygen_pk0_pins(F):
    # src/lib.rs:536-545
    for ps, want in ((44, 197), (65, 177), (87, 16)):
        pk, sk = F.PARAM_SETS[ps].KG.keygen_from_seed(b"\x11" * 32)
        assert pk.into_bytes()[0] == want


# ---------------------------------------------------------------- verify
@pytest.mark.parametrize("ps", [44, 65, 87])
def test_verify_acvp(F, acvp_sigver, ps):
    m = F.PARAM_SETS[ps]
    vec = [t for t in acvp_sigver if PS_OF[t["ps"]] == ps]
    assert len(vec) == 15
    pk = bytes.fromhex(vec[0]["pk"])
    msgs = [bytes.fromhex(t["message"]) for t in vec]
    sigs = [bytes.fromhex(t["signature"]) for t in vec]
    want = [t["passed"] for t in vec]
    shared = m.verify_batch(pk, msgs, sigs, shared_key=True, framing=1)
    assert [bool(x) for x in shared] == want, [t["reason"] for t in vec]
    distinct = m.verify_batch(pk * len(vec), msgs, sigs, framing=1)
    assert [bool(x) for x in distinct] == want


def test_verify_bad_sig(F, reference_kats):
    # tests/integration.rs:60-74
    b = reference_kats["bad_sig"]
    m = F.ml_dsa_44
    pk = m.PublicKey.try_from_bytes(bytes.fromhex(b["pk"]))
    assert m._internal_verify(pk, bytes.fromhex(b["msg"]), bytes.fromhex(b["good_sig"])) is True
    assert m._internal_verify(pk, bytes.fromhex(b["msg"]), bytes.fromhex(b["bad_sig"])) is False


def _mutate(ps, sig, cls, p):
    """The 8 corruption classes of SURVEY.md section 8d config 3."""
    s = bytearray(sig)
    zoff, hoff = p.lambda_div4, p.lambda_div4 + p.l * 32 * p.z_bits
    if cls == 0:
        s[3] ^= 0x10                      # c~
    elif cls == 1:
        s[zoff + 77] ^= 0x04              # z
    elif cls == 2:
        s[hoff + 1] ^= 0x01               # hint index
    elif cls == 3:
        s[hoff + p.omega + p.k - 1] = p.omega + 1   # hint count > omega
    elif cls == 4:
        s[hoff], s[hoff + 1] = 9, 9
        s[hoff + p.omega] = max(s[hoff + p.omega], 2)  # non-increasing hint index
    elif cls == 5:
        s[hoff + p.omega - 1] = 1 if s[hoff + p.omega - 1] == 0 else s[hoff + p.omega - 1]  # padding
    elif cls == 6:
        # force z[0][0] = gamma1 - beta exactly (the smallest rejected magnitude): field = gamma1 - z
        field = p.gamma1 - (p.gamma1 - p.beta)
        v = int.from_bytes(s[zoff:zoff + 4], "little")
        v = (v & ~((1 << p.z_bits) - 1)) | field
        s[zoff:zoff + 4] = v.to_bytes(4, "little")
    return bytes(s)


@pytest.mark.parametrize("ps", [44, 65, 87])
def test_verify_mutations_vs_oracle(F, ps):
    m = F.PARAM_SETS[ps]
    p = O.params(ps)
    n = 256
    pk, sk = O.keygen(ps, synth("xi", ps, 32))
    msgs = [synth("msg", i, 32 + (i % 5)) for i in range(n)]
    ctxs = [synth("ctx", i, i % 7) for i in range(n)]
    rnds = synth_rows("rnd", n, 32)
    sigs, _ = O.sign_batch(ps, np.frombuffer(sk, dtype=np.uint8), msgs, rnds, ctxs=ctxs, shared_key=True)
    sigs = [s.tobytes() for s in sigs]
    for i in range(0, n, 2):
        cls = (i // 2) % 8
        if cls == 7:
            msgs[i] = msgs[i] + b"!"       # wrong message
        else:
            sigs[i] = _mutate(ps, sigs[i], cls, p)
    want = O.verify_batch(ps, np.frombuffer(pk, dtype=np.uint8), msgs, np.frombuffer(b"".join(sigs), dtype=np.uint8),
                          ctxs=ctxs, shared_key=True)
    assert want[1::2].all() and not want[0::2].all()
    got_s = m.verify_batch(pk, msgs, sigs, ctxs=ctxs, shared_key=True)
    assert (got_s == want).all(), np.nonzero(got_s != want)
    got_d = m.verify_batch(pk * n, msgs, sigs, ctxs=ctxs)
    assert (got_d == want).all(), np.nonzero(got_d != want)


@pytest.mark.parametrize("ps", [44, 65, 87])
def test_verify_distinct_keys(F, ps):
    m = F.PARAM_SETS[ps]
    n = 96
    xi = synth_block(f"vk{ps}", n, 32)
    pks, sks = O.keygen_batch(ps, xi)
    msgs = [synth("vm", i, 1 + 3 * i) for i in range(n)]   # ragged, up to ~290 B: multi-block mu
    rnds = synth_rows("vr", n, 32)
    sigs, _ = O.sign_batch(ps, sks, msgs, rnds)
    ok = m.verify_batch(pks, msgs, sigs)
    assert ok.all()
    # swapping two public keys must break exactly those two lanes
    pks2 = pks.copy()
    pks2[[3, 4]] = pks2[[4, 3]]
    ok = m.verify_batch(pks2, msgs, sigs)
    assert list(np.nonzero(ok == 0)[0]) == [3, 4]


# ---------------------------------------------------------------- sign
@pytest.mark.parametrize("ps", [44, 65, 87])
def test_sign_acvp(F, acvp_siggen, ps):
    m = F.PARAM_SETS[ps]
    vec = [t for t in acvp_siggen if PS_OF[t["ps"]] == ps]
    assert len(vec) == 20
    sks = np.frombuffer(b"".join(bytes.fromhex(t["sk"]) for t in vec), dtype=np.uint8)
    msgs = [bytes.fromhex(t["message"]) for t in vec]
    rnds = b"".join(bytes.fromhex(t["rnd"]) if t["rnd"] else bytes(32) for t in vec)  # mod.rs:107-112
    sig, iters = m.sign_batch(sks, msgs, rnds, framing=1, want_iters=True)
    for i, t in enumerate(vec):
        assert sig[i].tobytes().hex().upper() == t["signature"].upper(), (t["tc"], int(iters[i]))


def test_sign_messages_rs_kat(F, reference_kats):
    # tests/messages.rs:9-21 through the single-op mirror API (external framing, ctx = [])
    k = reference_kats["messages_rs"]
    m = F.ml_dsa_44

    class Replay:
        def __init__(self, *chunks):
            self.q = list(chunks)

        def try_fill_bytes(self, n):
            return self.q.pop(0)

    rng = Replay(bytes.fromhex(k["xi"]), bytes.fromhex(k["rnd"]))
    pk, sk = m.KG.try_keygen_with_rng(rng)
    sig = sk.try_sign_with_rng(rng, bytes.fromhex(k["msg"]), b"")
    assert pk.verify(bytes.fromhex(k["msg"]), sig, b"")
    assert sk.into_bytes().hex() == k["sk"]
    assert sig.hex() == k["sig"]
    assert pk.into_bytes().hex() == k["pk"]


@pytest.mark.parametrize("ps", [44, 65, 87])
def test_sign_vs_oracle_distinct_keys(F, ps):
    m = F.PARAM_SETS[ps]
    n = 700 + 1
    xi = synth_block(f"sk{ps}", n, 32)
    pks, sks = O.keygen_batch(ps, xi)
    msgs = [synth("sm", i, (i * 13) % 300) for i in range(n)]   # includes the empty message
    ctxs = [synth("sc", i, (i * 5) % 256) for i in range(n)]    # includes |ctx| = 255
    rnds = synth_rows("sr", n, 32)
    sig, iters = m.sign_batch(sks, msgs, rnds, ctxs=ctxs, want_iters=True)
    osig, oiters = O.sign_batch(ps, sks, msgs, rnds, ctxs=ctxs)
    bad = np.nonzero((sig != osig).any(axis=1))[0]
    assert bad.size == 0, (bad[:10], iters[bad[:10]], oiters[bad[:10]])
    assert (iters == oiters).all()
    assert m.verify_batch(pks, msgs, sig, ctxs=ctxs).all()


@pytest.mark.parametrize("ps", [44, 65, 87])
def test_sign_shared_key_and_deterministic(F, ps):
    m = F.PARAM_SETS[ps]
    n = 300
    pk, sk = O.keygen(ps, synth("shk", ps, 32))
    msgs = [synth("shm", i, 32) for i in range(n)]
    rnds = np.zeros((n, 32), dtype=np.uint8)                    # deterministic variant
    sig = m.sign_batch(sk, msgs, rnds, shared_key=True)
    osig, _ = O.sign_batch(ps, np.frombuffer(sk, dtype=np.uint8), msgs, rnds, shared_key=True)
    assert (sig == osig).all()
    assert m.verify_batch(pk, msgs, sig, shared_key=True).all()


def test_sign_rejects_long_ctx(F):
    m = F.ml_dsa_44
    pk, sk = m.KG.keygen_from_seed(bytes(32))
    with pytest.raises(ValueError, match="ctx too long"):
        sk.try_sign_with_seed(bytes(32), b"m", b"\1" * 256)
    sig = sk.try_sign_with_seed(bytes(32), b"m", b"\1" * 255)
    assert pk.verify(b"m", sig, b"\1" * 255)
    assert not pk.verify(b"m", sig, b"\1" * 256)
    assert not pk.verify(b"m", sig, b"\1" * 254)


def test_sign_out_of_range_secret_key(F):
    # SURVEY 8a note T: sk_decode never fails; fuzz_all signs with arbitrary sk bytes.  Same mod-q result.
    for ps in (44, 65, 87):
        m = F.PARAM_SETS[ps]
        n = 24
        sks = synth_block(f"fuzzsk{ps}", n, m.SK_LEN)
        msgs = [synth("fm", i, 16) for i in range(n)]
        rnds = synth_rows("fr", n, 32)
        sig, iters = m.sign_batch(sks, msgs, rnds, want_iters=True)
        osig, oiters = O.sign_batch(ps, sks, msgs, rnds)
        assert (sig == osig).all() and (iters == oiters).all()


def test_host_chunk_pipeline_many_chunks(F):
    """Force the host entry points to cut a batch into many chunks (tiny workspace limit) so the
    double-buffered H2D / compute / D2H pipeline is exercised; bytes must not depend on the chunking."""
    from fips204_b200 import _lib
    L = _lib.lib()
    ps, n = 65, 2500
    m = F.PARAM_SETS[ps]
    xi = synth_block("chunk-xi", n, 32)
    pks, sks = O.keygen_batch(ps, xi)
    msgs = [synth("chunk-m", i, i % 97) for i in range(n)]
    ctxs = [synth("chunk-c", i, i % 11) for i in range(n)]
    rnds = synth_rows("chunk-r", n, 32)
    want, wit = O.sign_batch(ps, sks, msgs, rnds, ctxs=ctxs)
    try:
        L.mldsa_set_workspace_limit(64 << 20)   # -> chunks of a few hundred lanes
        for ways in (1, 2):
            L.mldsa_set_tuning(ways, 0, 0)
            sig, it = m.sign_batch(sks, msgs, rnds, ctxs=ctxs, want_iters=True)
            assert (sig == want).all() and (it == wit).all()
        bad = sig.copy()
        bad[[5, 700, 2499], 100] ^= 0x40
        ok = m.verify_batch(pks, msgs, bad, ctxs=ctxs)
        assert list(np.nonzero(ok == 0)[0]) == [5, 700, 2499]
        ok = m.verify_batch(pks[7].tobytes(), [msgs[7]] * n, np.repeat(sig[7:8], n, axis=0), ctxs=[ctxs[7]] * n,
                            shared_key=True)
        assert ok.all()
        pk2, sk2 = m.keygen_batch(xi)
        assert (pk2 == pks).all() and (sk2 == sks).all()
    finally:
        L.mldsa_set_workspace_limit(0)
        L.mldsa_set_tuning(1, 0, 0)


# ---------------------------------------------------------------- HashML-DSA and get_public_key (8f-3)
@pytest.mark.parametrize("ps", [44, 65, 87])
def test_hash_sign_verify_vs_oracle(F, ps):
    """try_hash_sign* / hash_verify (src/lib.rs:310-342, :391-411): device pre-hash (SHA-256 / SHA-512 /
    SHAKE128) + OID framing.  The reference has no KAT for these (sha2 is 'parity unpinned' there); the
    oracle's SHA-2 is pinned against hashlib (test_oracle_leaves) and the GPU must match it byte for byte."""
    m = F.PARAM_SETS[ps]
    n = 90
    pk, sk = O.keygen(ps, synth("hxi", ps, 32))
    # ragged messages straddling the SHA-256 (55/56/64) and SHA-512 (111/112/128) padding boundaries
    lens = [0, 1, 54, 55, 56, 63, 64, 65, 110, 111, 112, 119, 120, 127, 128, 129, 135, 136, 137, 167, 168, 169, 300, 1000]
    msgs = [synth("hm", i, lens[i % len(lens)] + (i // len(lens))) for i in range(n)]
    ctxs = [synth("hc", i, (i * 37) % 256) for i in range(n)]
    rnds = synth_rows("hr", n, 32)
    for ph in (F.Ph.SHA256, F.Ph.SHA512, F.Ph.SHAKE128):
        sig = m.hash_sign_batch(sk, msgs, rnds, ph, ctxs=ctxs, shared_key=True)
        for i in range(n):
            want = O.sign(ps, sk, msgs[i], rnds[i].tobytes(), ctx=ctxs[i], framing=2, ph=ph)
            assert sig[i].tobytes() == want, (ph, i, len(msgs[i]))
        assert m.hash_verify_batch(pk, msgs, sig, ph, ctxs=ctxs, shared_key=True).all()
        # a HashML-DSA signature is not a pure ML-DSA signature, nor valid under another pre-hash
        assert not m.verify_batch(pk, msgs, sig, ctxs=ctxs, shared_key=True).any()
        other = (ph + 1) % 3
        assert not m.hash_verify_batch(pk, msgs, sig, other, ctxs=ctxs, shared_key=True).any()
        for i in (0, 7, 33):
            assert O.verify(ps, pk, msgs[i], sig[i].tobytes(), ctx=ctxs[i], framing=2, ph=ph)


def test_hash_sign_single_op_api(F):
    # the self-consistency flow of src/lib.rs:514-533 through the mirrored Signer / Verifier methods
    for m in (F.ml_dsa_44, F.ml_dsa_65, F.ml_dsa_87):
        pk, sk = m.KG.keygen_from_seed(bytes(range(32)))
        msg, ctx = b"message to pre-hash", b"context"
        for ph in (F.Ph.SHA256, F.Ph.SHA512, F.Ph.SHAKE128):
            sig = sk.try_hash_sign_with_seed(bytes(32), msg, ctx, ph)
            assert pk.hash_verify(msg, sig, ctx, ph)
            assert not pk.hash_verify(msg + b"x", sig, ctx, ph)
            assert not pk.hash_verify(msg, sig, b"other", ph)
            assert not pk.verify(msg, sig, ctx)
        with pytest.raises(ValueError, match="HashML-DSA.Sign: ctx too long"):
            sk.try_hash_sign_with_seed(bytes(32), msg, bytes(256), F.Ph.SHA256)
        assert pk.hash_verify(msg, sig, bytes(256), F.Ph.SHA256) is False


@pytest.mark.parametrize("ps", [44, 65, 87])
def test_get_public_key(F, ps, acvp_keygen):
    """Signer::get_public_key (src/lib.rs:344-346, ml_dsa.rs:502-563): pk bytes recomputed from sk bytes."""
    m = F.PARAM_SETS[ps]
    vec = [t for t in acvp_keygen if PS_OF[t["ps"]] == ps]
    sks = np.frombuffer(b"".join(bytes.fromhex(t["sk"]) for t in vec), dtype=np.uint8)
    pk = m.sk_to_pk_batch(sks)
    for i, t in enumerate(vec):
        assert pk[i].tobytes().hex().upper() == t["pk"].upper(), t["tc"]
    n = 500 + 7
    pks, sks = O.keygen_batch(ps, synth_block(f"gpk{ps}", n, 32))
    assert (m.sk_to_pk_batch(sks) == pks).all()
    sk0 = m.PrivateKey.try_from_bytes(sks[0].tobytes())
    assert sk0.get_public_key().into_bytes() == pks[0].tobytes()
    # arbitrary (out-of-range) sk bytes: same mod-q computation as the oracle's private_to_public_key
    fz = synth_block(f"gpkfuzz{ps}", 16, m.SK_LEN)
    got = m.sk_to_pk_batch(fz)
    for i in range(16):
        assert got[i].tobytes() == O.sk_to_pk(ps, fz[i].tobytes())


@pytest.mark.parametrize("depth", [1, 2, 3, 4, 8])
def test_sign_speculation_depth_does_not_change_bytes(F, depth):
    """The sign loop tries `depth` candidates per lane and round (mldsa_set_depth); a candidate only counts
    if all earlier ones were rejected, so signatures AND iteration counts are those of the reference's
    sequential loop (src/ml_dsa.rs:212-330) for every depth."""
    from fips204_b200 import _lib
    L = _lib.lib()
    try:
        L.mldsa_set_depth(depth)
        for ps in (44, 65, 87):
            m = F.PARAM_SETS[ps]
            n = 257
            pks, sks = O.keygen_batch(ps, synth_block(f"dk{ps}", n, 32))
            msgs = [synth("dm", i, i % 50) for i in range(n)]
            rnds = synth_rows("dr", n, 32)
            sig, it = m.sign_batch(sks, msgs, rnds, want_iters=True)
            osig, oit = O.sign_batch(ps, sks, msgs, rnds)
            assert (sig == osig).all() and (it == oit).all(), (ps, depth)
            one = m.sign_batch(sks[5].tobytes(), msgs, rnds, shared_key=True)
            oone, _ = O.sign_batch(ps, sks[5], msgs, rnds, shared_key=True)
            assert (one == oone).all()
    finally:
        L.mldsa_set_depth(2)


@pytest.mark.parametrize("ps", [44, 65, 87])
def test_sign_tail_rounds_speculate_deeper(F, ps):
    """The last, under-filled rounds of the rejection loop try up to 16 candidates per lane in slots addressed by
    the lane's position in the active list (mldsa_set_tail).  Signatures and iteration counts must still be the
    reference's sequential loop (src/ml_dsa.rs:212-330), with the tail region small enough here that depths 4, 8
    and 16 are all used, through both entry points and for a shared key."""
    from fips204_b200 import _lib
    L = _lib.lib()
    m = F.PARAM_SETS[ps]
    n = 5000
    pks, sks = O.keygen_batch(ps, synth_block(f"tk{ps}", n, 32))
    msgs = [synth("tm", i, i % 40) for i in range(n)]
    rnds = synth_rows("tr", n, 32)
    osig, oit = O.sign_batch(ps, sks, msgs, rnds)
    try:
        for tail_depth, slots in ((16, 8192), (8, 32768), (4, 1024), (0, 0)):
            L.mldsa_set_tail(tail_depth, slots)
            sig, it = m.sign_batch(sks, msgs, rnds, want_iters=True)
            assert (sig == osig).all() and (it == oit).all(), (ps, tail_depth, slots)
        L.mldsa_set_tail(16, 8192)
        one = m.sign_batch(sks[9].tobytes(), msgs[:1500], rnds[:1500], shared_key=True)
        oone, _ = O.sign_batch(ps, sks[9], msgs[:1500], rnds[:1500], shared_key=True)
        assert (one == oone).all()
        assert m.verify_batch(pks, msgs, sig).all()
    finally:
        L.mldsa_set_tail(8, 32768)


@pytest.mark.parametrize("ps", [44, 65, 87])
def test_no_secret_residue_on_device(F, ps):
    """ZeroizeOnDrop analogue (src/types.rs:19-41): once keygen / sign has returned, every device range that held
    secret-derived data (staged xi / sk / rnd, rho', K, s1, s2 and their NTT images, y, w - c*s2, rho'') reads as
    zero -- for distinct and shared keys, one and many chunks, and with the deeper tail rounds in their shared slots."""
    import ctypes as C
    from fips204_b200 import _lib
    L = _lib.lib()
    m = F.PARAM_SETS[ps]

    def residue():
        nz, sc = C.c_uint64(), C.c_uint64()
        _lib.check(L.mldsa_debug_secret_residue(0, C.byref(nz), C.byref(sc)))
        assert sc.value > 0, "no secret ranges were registered"
        return nz.value

    n = 3000
    xi = synth_block(f"zk{ps}", n, 32)
    pks, sks = m.keygen_batch(xi)
    assert residue() == 0
    msgs = [synth("zm", i, i % 60) for i in range(n)]
    rnds = synth_rows("zr", n, 32)
    try:
        for tail_depth, slots, limit in ((8, 32768, 0), (16, 2048, 0), (0, 0, 0), (16, 1024, 48 << 20)):
            L.mldsa_set_tail(tail_depth, slots)
            L.mldsa_set_workspace_limit(limit)
            sig = m.sign_batch(sks, msgs, rnds)
            assert residue() == 0, (ps, tail_depth, slots, limit)
            one = m.sign_batch(sks[3].tobytes(), msgs, rnds, shared_key=True)
            assert residue() == 0, (ps, tail_depth, slots, limit, "shared")
        assert m.verify_batch(pks, msgs, sig).all() and m.verify_batch(pks[3].tobytes(), msgs, one, shared_key=True).all()
    finally:
        L.mldsa_set_tail(8, 32768)
        L.mldsa_set_workspace_limit(0)


@pytest.mark.parametrize("ps", [44, 65, 87])
def test_resident_key_set_signs_like_packed_keys(F, ps):
    """mldsa_sk_load / mldsa_sign_batch_keys: keys expanded once and kept on the device (the reference's expanded
    PrivateKey, src/types.rs:19-28, src/ml_dsa.rs:445-469).  Signatures and iteration counts equal the oracle's for
    the identity mapping and for an arbitrary key index per lane, over repeated calls (nothing of the set is wiped),
    in one chunk and in many."""
    import ctypes as C
    from fips204_b200 import _lib
    L = _lib.lib()
    m = F.PARAM_SETS[ps]
    nk, n = 700, 2600
    pks, sks = O.keygen_batch(ps, synth_block(f"rk{ps}", nk, 32))
    ks = m.load_private_keys(sks)
    assert len(ks) == nk and ks.device_bytes >= nk * (m.K * m.L + m.L + 2 * m.K) * 1024
    nz = C.c_uint64()
    _lib.check(L.mldsa_debug_secret_residue(0, C.byref(nz), None))
    assert nz.value == 0                                   # the staged packed keys are gone
    msgs = [synth("rm", i, i % 70) for i in range(n)]
    ctxs = [synth("rc", i, i % 9) for i in range(n)]
    rnds = synth_rows("rr", n, 32)
    idx = (np.arange(n, dtype=np.uint64) * 2654435761 % nk).astype(np.uint32)
    want, wit = O.sign_batch(ps, sks[idx], msgs, rnds, ctxs=ctxs)
    try:
        for limit in (0, 40 << 20):
            L.mldsa_set_workspace_limit(limit)
            for _ in range(2):
                sig, it = ks.sign_batch(msgs, rnds, ctxs=ctxs, key_idx=idx, want_iters=True)
                assert (sig == want).all() and (it == wit).all(), (ps, limit)
            sig1 = ks.sign_batch(msgs[:nk], rnds[:nk])         # identity mapping: lane i, key i
            want1, _ = O.sign_batch(ps, sks, msgs[:nk], rnds[:nk])
            assert (sig1 == want1).all()
            _lib.check(L.mldsa_debug_secret_residue(0, C.byref(nz), None))
            assert nz.value == 0
        assert m.verify_batch(pks[idx], msgs, sig, ctxs=ctxs).all()
        with pytest.raises(F.MldsaError):
            ks.sign_batch(msgs[:4], rnds[:4], key_idx=np.array([0, 1, nk, 2], dtype=np.uint32))
        with pytest.raises(ValueError, match="ctx too long"):
            ks.sign_batch(msgs[:1], rnds[:1], ctxs=[bytes(256)])
    finally:
        L.mldsa_set_workspace_limit(0)
        ks.free()
    assert len(ks) == 0


@pytest.mark.parametrize("ps", [44, 65, 87])
def test_edge_shapes_vs_oracle(F, ps):
    """Empty batch, single lane, empty and block-boundary message lengths (mu absorbs tr || 0 || |ctx| || ctx || M:
    the SHAKE256 rate boundary falls at |ctx| + |M| = 70, 206, ...), a 100 kB message, the longest legal ctx."""
    m = F.PARAM_SETS[ps]
    pk, sk = O.keygen(ps, synth("edge-xi", ps, 32))
    assert m.sign_batch(sk, [], np.zeros((0, 32), np.uint8), shared_key=True).shape == (0, m.SIG_LEN)
    assert m.verify_batch(pk, [], np.zeros((0, m.SIG_LEN), np.uint8), shared_key=True).shape == (0,)
    pk0, sk0 = m.keygen_batch(np.zeros((0, 32), np.uint8))
    assert pk0.shape == (0, m.PK_LEN) and sk0.shape == (0, m.SK_LEN)
    lens = [0, 1, 69, 70, 71, 72, 135, 136, 137, 205, 206, 207, 341, 342, 343, 5000, 100000]
    msgs = [synth("edge-m", i, n) for i, n in enumerate(lens)]
    ctxs = [b"" if i % 3 == 0 else synth("edge-c", i, 255 if i % 3 == 1 else i) for i in range(len(lens))]
    rnds = synth_rows("edge-r", len(lens), 32)
    sig, it = m.sign_batch(sk, msgs, rnds, ctxs=ctxs, shared_key=True, want_iters=True)
    for i in range(len(lens)):
        want, wit = O.sign(ps, sk, msgs[i], rnds[i].tobytes(), ctx=ctxs[i], want_iters=True)
        assert sig[i].tobytes() == want and it[i] == wit, (ps, lens[i], len(ctxs[i]))
        one = m.sign_batch(sk, [msgs[i]], rnds[i], ctxs=[ctxs[i]], shared_key=True)   # n = 1
        assert one[0].tobytes() == want
    assert m.verify_batch(pk, msgs, sig, ctxs=ctxs, shared_key=True).all()
    swapped = [ctxs[1], ctxs[0]] + ctxs[2:]
    assert list(m.verify_batch(pk, msgs, sig, ctxs=swapped, shared_key=True)[:2]) == [0, 0]


@pytest.mark.parametrize("ps", [44, 65, 87])
def test_resident_public_key_set_verifies_like_packed_keys(F, ps):
    """mldsa_pk_load / mldsa_verify_batch_keys: expand_public done once per key (the reference's PublicKey,
    src/types.rs:35-41, src/ml_dsa.rs:477-498).  Verdicts equal the oracle's on valid and mutated signatures, for
    the identity mapping and for a key index per lane, in one chunk and in many; wrong key -> false."""
    from fips204_b200 import _lib
    L = _lib.lib()
    m = F.PARAM_SETS[ps]
    nk, n = 300, 2100
    pks, sks = O.keygen_batch(ps, synth_block(f"pk{ps}", nk, 32))
    ks = m.load_public_keys(pks)
    assert len(ks) == nk
    msgs = [synth("pm", i, i % 45) for i in range(n)]
    ctxs = [synth("pc", i, i % 7) for i in range(n)]
    rnds = synth_rows("pr", n, 32)
    idx = (np.arange(n, dtype=np.uint64) * 40503 % nk).astype(np.uint32)
    sig, _ = O.sign_batch(ps, sks[idx], msgs, rnds, ctxs=ctxs)
    bad = sig.copy()
    hint_off = m.SIG_LEN - (O.params(ps).omega + m.K)
    bad[0::7, 3] ^= 1                       # c~
    bad[1::7, 80] ^= 4                      # z
    bad[2::7, m.SIG_LEN - 1] = 255          # hint count > omega
    bad[3::7, hint_off + 1] ^= 1            # hint index / padding
    want = O.verify_batch(ps, pks[idx], msgs, bad, ctxs=ctxs)
    assert 0 < want.sum() < n
    try:
        for limit in (0, 24 << 20):
            L.mldsa_set_workspace_limit(limit)
            assert ks.verify_batch(msgs, sig, ctxs=ctxs, key_idx=idx).all()
            assert (ks.verify_batch(msgs, bad, ctxs=ctxs, key_idx=idx) == want).all(), (ps, limit)
            wrong = (idx + 1) % nk
            assert not ks.verify_batch(msgs, sig, ctxs=ctxs, key_idx=wrong).any()
            s1, _ = O.sign_batch(ps, sks, msgs[:nk], rnds[:nk])
            assert ks.verify_batch(msgs[:nk], s1).all()                      # identity mapping
        with pytest.raises(ValueError):
            ks.sign_batch(msgs[:1], rnds[:1])
        with pytest.raises(F.MldsaError):
            ks.verify_batch(msgs[:2], sig[:2], key_idx=np.array([0, nk], dtype=np.uint32))
        sks_set = m.load_private_keys(sks[:4])
        with pytest.raises(F.MldsaError):                                    # a secret-key set cannot verify
            _lib.check(L.mldsa_verify_batch_keys(sks_set._h, 0, 1, None, None, np.zeros(2, np.uint64).ctypes.data, None, None,
                                                 sig[:1].ctypes.data, np.zeros(1, np.uint8).ctypes.data))
        sks_set.free()
    finally:
        L.mldsa_set_workspace_limit(0)
        ks.free()


def test_sign_survives_a_low_live_count_estimate():
    """The host sizes every round's grids (and picks the tail depth) from an ESTIMATE of the live-lane count.
    MLDSA_EST_SCALE=0.3 makes that estimate far too low: the kernels must still cover the true count (grid-stride
    loops, grid-stride compaction), lanes that find no slot in the shared tail region must sit the round out
    untouched, and the straggler rounds must finish the batch -- with the reference's bytes and iteration counts."""
    import os
    import subprocess
    import sys
    code = r'''
import sys
sys.path.insert(0, "tests")
import numpy as np
import fips204_b200 as F, oracle_lib as O
from fips204_b200 import _lib
from synth import synth, synth_block, synth_rows
L = _lib.lib()
for ps in (44, 65, 87):
    m = F.PARAM_SETS[ps]
    n = 4000
    pks, sks = O.keygen_batch(ps, synth_block(f"est{ps}", n, 32))
    msgs = [synth("em", i, i % 33) for i in range(n)]
    rnds = synth_rows("er", n, 32)
    want, wit = O.sign_batch(ps, sks, msgs, rnds)
    for td, slots in ((16, 1024), (8, 32768), (0, 0)):
        L.mldsa_set_tail(td, slots)
        sig, it = m.sign_batch(sks, msgs, rnds, want_iters=True)
        assert (sig == want).all() and (it == wit).all(), (ps, td, slots)
print("ok")
'''
    env = dict(os.environ, MLDSA_EST_SCALE="0.3")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env,
                       cwd=str(__import__("pathlib").Path(__file__).resolve().parent.parent), timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), (r.stdout[-300:], r.stderr[-1500:])


def test_multi_device_entry_points_split_contiguously(F):
    """mldsa_*_batch_multi: one host thread per listed device, contiguous lane ranges, disjoint output slices, ragged
    message / ctx blobs addressed by absolute offsets.  Three shards on device 0 (they serialise on the device context):
    this pins the splitting logic on any box; test_gpu_round2.py::test_multi_entry_points_on_distinct_devices runs the
    same entry points over every visible GPU and SKIPS when there is only one."""
    import ctypes as C
    from fips204_b200 import _lib
    L = _lib.lib()
    ps = 65
    m = F.PARAM_SETS[ps]
    devs = [0, 0, 0]
    darr = (C.c_int * len(devs))(*devs)
    n = 1000 + 7
    xi = synth_block("md-xi", n, 32)
    pk = np.empty((n, m.PK_LEN), np.uint8)
    sk = np.empty((n, m.SK_LEN), np.uint8)
    _lib.check(L.mldsa_keygen_batch_multi(ps, darr, len(devs), n, xi.ctypes.data, pk.ctypes.data, sk.ctypes.data))
    opk, osk = O.keygen_batch(ps, xi)
    assert (pk == opk).all() and (sk == osk).all()
    msgs = [synth("md-m", i, i % 53) for i in range(n)]
    ctxs = [synth("md-c", i, i % 5) for i in range(n)]
    rnd = synth_rows("md-r", n, 32)
    mb = np.frombuffer(b"".join(msgs) + b"\0", np.uint8).copy()
    mo = np.zeros(n + 1, np.uint64)
    mo[1:] = np.cumsum([len(x) for x in msgs])
    cb = np.frombuffer(b"".join(ctxs) + b"\0", np.uint8).copy()
    co = np.zeros(n + 1, np.uint32)
    co[1:] = np.cumsum([len(x) for x in ctxs])
    sig = np.empty((n, m.SIG_LEN), np.uint8)
    it = np.zeros(n, np.uint32)
    _lib.check(L.mldsa_sign_batch_multi(ps, darr, len(devs), 0, n, sk.ctypes.data, m.SK_LEN, mb.ctypes.data, mo.ctypes.data,
                                        cb.ctypes.data, co.ctypes.data, rnd.ctypes.data, sig.ctypes.data, it.ctypes.data))
    osig, oit = O.sign_batch(ps, sk, msgs, rnd, ctxs=ctxs)
    assert (sig == osig).all() and (it == oit).all()
    sig[[0, 500, n - 1], 9] ^= 2
    ok = np.zeros(n, np.uint8)
    _lib.check(L.mldsa_verify_batch_multi(ps, darr, len(devs), 0, n, pk.ctypes.data, m.PK_LEN, mb.ctypes.data,
                                          mo.ctypes.data, cb.ctypes.data, co.ctypes.data, sig.ctypes.data, ok.ctypes.data))
    assert list(np.nonzero(ok == 0)[0]) == [0, 500, n - 1]


def test_config1_reference_benchmark_inputs(F):
    """BASELINE configs[0]: the inputs of the reference's own benchmark (benches/benchmark.rs: counter RNG, keys from
    draws 0 / 1 / 2, message 00..07, sign with ctx = [], verify of a ctx = [0] signature under ctx = [] -> false)
    through the engine, byte for byte against the oracle (tools/bench_config1.py times the same lanes on the CPU)."""
    import importlib.util
    from pathlib import Path
    spec = importlib.util.spec_from_file_location("bench_config1", Path(__file__).resolve().parent.parent / "tools" / "bench_config1.py")
    c1 = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(c1)
    assert c1.draw(0x01020304)[:6] == bytes([1, 2, 3, 4, 4, 4]) and c1.draw(7) == bytes([0, 0, 0, 7] + [7] * 28)
    n = 96
    for ps in (44, 65, 87):
        m = F.PARAM_SETS[ps]
        xi, pk, sk, rnd, sig0 = c1.workload(ps, n)
        opk, osk = O.keygen_batch(ps, xi)
        gpk, gsk = m.keygen_batch(xi)
        assert (gpk == opk).all() and (gsk == osk).all()
        osig, oit = O.sign_batch(ps, np.frombuffer(sk, np.uint8), [c1.MSG] * n, rnd, ctxs=[b""] * n, shared_key=True)
        gsig, git = m.sign_batch(sk, [c1.MSG] * n, rnd, ctxs=[b""] * n, shared_key=True, want_iters=True)
        assert (gsig == osig).all() and (git == oit).all()
        sigs = np.tile(np.frombuffer(sig0, np.uint8), (n, 1))
        assert not m.verify_batch(pk, [c1.MSG] * n, sigs, ctxs=[b""] * n, shared_key=True).any()      # benchmark.rs:58
        assert m.verify_batch(pk, [c1.MSG] * n, sigs, ctxs=[b"\0"] * n, shared_key=True).all()


def test_hash_ml_dsa_on_resident_key_sets(F):
    """try_hash_sign* / hash_verify (src/lib.rs:310-342, :391-411) on expanded keys: same bytes as the packed-key
    HashML-DSA path and the oracle, for all three pre-hash functions and a key index per lane."""
    ps = 65
    m = F.PARAM_SETS[ps]
    nk, n = 40, 300
    pks, sks = O.keygen_batch(ps, synth_block("hk", nk, 32))
    sset, pset = m.load_private_keys(sks), m.load_public_keys(pks)
    msgs = [synth("hkm", i, (i * 13) % 400) for i in range(n)]
    ctxs = [synth("hkc", i, i % 12) for i in range(n)]
    rnds = synth_rows("hkr", n, 32)
    idx = (np.arange(n) * 7 % nk).astype(np.uint32)
    try:
        for ph in (F.Ph.SHA256, F.Ph.SHA512, F.Ph.SHAKE128):
            sig = sset.sign_batch(msgs, rnds, ctxs=ctxs, key_idx=idx, ph=ph)
            assert (sig == m.hash_sign_batch(sks[idx], msgs, rnds, ph, ctxs=ctxs)).all()
            for i in (0, 17, n - 1):
                assert sig[i].tobytes() == O.sign(ps, sks[idx[i]].tobytes(), msgs[i], rnds[i].tobytes(), ctx=ctxs[i], framing=2, ph=ph)
            assert pset.verify_batch(msgs, sig, ctxs=ctxs, key_idx=idx, ph=ph).all()
            assert not pset.verify_batch(msgs, sig, ctxs=ctxs, key_idx=idx).any()           # not a pure ML-DSA signature
            assert not pset.verify_batch(msgs, sig, ctxs=ctxs, key_idx=idx, ph=(ph + 1) % 3).any()
    finally:
        sset.free()
        pset.free()
