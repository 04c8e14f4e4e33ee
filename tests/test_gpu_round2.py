"""GPU parity tests of the entry points added after round 1: device-side key provisioning
(mldsa_keygen_to_keyset), the chained keygen -> sign -> verify pipeline (mldsa_pipeline_batch[_multi],
BASELINE config 5 / src/lib.rs:503-551), the in-process multi-GPU entry points on DISTINCT devices, and the
error-path guarantees (no partial signatures next to an error code, offsets validated).  Everything is compared
byte for byte with the CPU oracle."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from synth import synth, synth_block, synth_rows

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    import fips204_b200 as F
    assert F.device_count() >= 1, "no CUDA device: GPU tests need the B200 box"
    return F


def _residue(dev=0):
    from fips204_b200 import _lib
    nz, sc = C.c_uint64(), C.c_uint64()
    _lib.check(_lib.lib().mldsa_debug_secret_residue(dev, C.byref(nz), C.byref(sc)))
    assert sc.value > 0, "no secret ranges were registered"
    return nz.value


@pytest.mark.parametrize("ps", [44, 65, 87])
def test_keygen_to_keysets_equals_packed_key_path(F, ps):
    """KG::keygen_from_seed (src/lib.rs:247-251) with the expanded keys kept on the device: the public keys are the
    oracle's bytes, signing with the private set gives the oracle's signatures and trip counts, the public set gives
    the oracle's verdicts on valid and corrupted signatures; nothing secret is left in the staging area."""
    m = F.PARAM_SETS[ps]
    n = 700
    xi = synth_block(f"k2k-xi-{ps}", n, 32)
    pk, sks, pks = m.keygen_to_keysets(xi)
    assert _residue() == 0
    opk, osk = O.keygen_batch(ps, xi)
    assert (pk == opk).all()
    assert len(sks) == n and len(pks) == n and sks.device_bytes > 0
    msgs = [synth("k2k-m", i, i % 70) for i in range(n)]
    ctxs = [synth("k2k-c", i, i % 4) for i in range(n)]
    rnd = synth_rows("k2k-r", n, 32)
    sig, it = sks.sign_batch(msgs, rnd, ctxs=ctxs, want_iters=True)
    osig, oit = O.sign_batch(ps, osk, msgs, rnd, ctxs=ctxs)
    assert (sig == osig).all() and (it == oit).all()
    # a key index per lane (a pool of keys signing many messages)
    kidx = (np.arange(n, dtype=np.uint32) * 7919) % 13
    sig2 = sks.sign_batch(msgs, rnd, ctxs=ctxs, key_idx=kidx)
    osig2, _ = O.sign_batch(ps, osk[kidx], msgs, rnd, ctxs=ctxs)
    assert (sig2 == osig2).all()
    bad = sig.copy()
    bad[::3, 40] ^= 8
    got = pks.verify_batch(msgs, bad, ctxs=ctxs)
    want = O.verify_batch(ps, opk, msgs, bad, ctxs=ctxs)
    assert (got == want).all() and want[1].all() and not want[0]
    # one kind of set only
    pk_only, none_s, p2 = m.keygen_to_keysets(xi[:50], want_private=False)
    assert none_s is None and (pk_only == opk[:50]).all() and p2.verify_batch(msgs[:50], sig[:50], ctxs=ctxs[:50]).all()
    _, s2, none_p = m.keygen_to_keysets(xi[:50], want_pk=False, want_public=False)
    assert none_p is None and (s2.sign_batch(msgs[:50], rnd[:50], ctxs=ctxs[:50]) == osig[:50]).all()
    for h in (sks, pks, p2, s2):
        h.free()


@pytest.mark.parametrize("ps", [44, 65, 87])
def test_pipeline_batch_equals_three_calls(F, ps):
    """keygen -> sign -> verify chained on the device (the smoke flow of src/lib.rs:503-551): pk, sk, sig and the trip
    counts are the oracle's bytes for the same xi / message / ctx / rnd, every verdict is true, and the secret
    scratch of all three stages reads as zero afterwards -- for one chunk and for many small chunks."""
    from fips204_b200 import _lib
    L = _lib.lib()
    m = F.PARAM_SETS[ps]
    n = 1500
    xi = synth_block(f"pl-xi-{ps}", n, 32)
    msgs = [synth("pl-m", i, (i * 13) % 90) for i in range(n)]
    ctxs = [synth("pl-c", i, i % 6) for i in range(n)]
    rnd = synth_rows("pl-r", n, 32)
    opk, osk = O.keygen_batch(ps, xi)
    osig, oit = O.sign_batch(ps, osk, msgs, rnd, ctxs=ctxs)
    try:
        for limit in (0, 96 << 20):
            L.mldsa_set_workspace_limit(limit)
            out = m.pipeline_batch(xi, msgs, rnd, ctxs=ctxs, want_sk=True, want_iters=True, want_stage_ms=True)
            assert (out["pk"] == opk).all() and (out["sk"] == osk).all(), (ps, limit)
            assert (out["sig"] == osig).all() and (out["iters"] == oit).all(), (ps, limit)
            assert out["ok"].all()
            assert all(t > 0 for t in out["stage_ms"])
            assert _residue() == 0, (ps, limit)
    finally:
        L.mldsa_set_workspace_limit(0)
    # internal framing (ACVP interface), no ctx, sk not requested
    out = m.pipeline_batch(xi[:300], msgs[:300], rnd[:300], framing=1)
    isig, _ = O.sign_batch(ps, osk[:300], msgs[:300], rnd[:300], framing=1)
    assert (out["sig"] == isig).all() and out["ok"].all() and "sk" not in out
    assert _residue() == 0


def test_pipeline_rejects_long_ctx_and_bad_offsets(F):
    from fips204_b200 import _lib
    L = _lib.lib()
    m = F.ml_dsa_44
    n = 8
    xi = synth_block("plx", n, 32)
    msgs = [b"m"] * n
    rnd = synth_rows("plr", n, 32)
    with pytest.raises(ValueError, match="ctx too long"):
        m.pipeline_batch(xi, msgs, rnd, ctxs=[b""] * 7 + [b"x" * 256])
    # through the raw ABI: the length check is the library's, not only the wrapper's (src/lib.rs:274)
    pk = np.empty((n, m.PK_LEN), np.uint8)
    sig = np.full((n, m.SIG_LEN), 0xEE, np.uint8)
    ok = np.zeros(n, np.uint8)
    mb = np.frombuffer(b"m" * n, np.uint8).copy()
    mo = np.arange(n + 1, dtype=np.uint64)
    cb = np.zeros(300, np.uint8)
    co = np.zeros(n + 1, np.uint32)
    co[-1] = 256
    rc = L.mldsa_pipeline_batch(44, 0, 0, n, xi.ctypes.data, mb.ctypes.data, mo.ctypes.data, cb.ctypes.data, co.ctypes.data,
                                rnd.ctypes.data, pk.ctypes.data, None, sig.ctypes.data, ok.ctypes.data, None, None)
    assert rc == -2 and (sig == 0xEE).all()
    # non-monotone offsets are a parameter error, not an out-of-bounds read
    mo_bad = mo.copy()
    mo_bad[3] = 1
    for fn_rc in (
        L.mldsa_pipeline_batch(44, 0, 0, n, xi.ctypes.data, mb.ctypes.data, mo_bad.ctypes.data, None, None, rnd.ctypes.data,
                               pk.ctypes.data, None, sig.ctypes.data, ok.ctypes.data, None, None),
        L.mldsa_verify_batch(44, 0, 0, n, pk.ctypes.data, m.PK_LEN, mb.ctypes.data, mo_bad.ctypes.data, None, None,
                             sig.ctypes.data, ok.ctypes.data),
    ):
        assert fn_rc == -1
    sk = np.zeros((n, m.SK_LEN), np.uint8)
    co_bad = np.zeros(n + 1, np.uint32)
    co_bad[2], co_bad[3] = 5, 1
    rc = L.mldsa_sign_batch(44, 0, 0, n, sk.ctypes.data, m.SK_LEN, mb.ctypes.data, mo.ctypes.data, cb.ctypes.data,
                            co_bad.ctypes.data, rnd.ctypes.data, sig.ctypes.data, None)
    assert rc == -1


def test_device_pointer_sign_leaves_no_partial_output_on_ctx_error(F):
    """mldsa_sign_batch_dev with one ctx longer than 255 bytes: -2 (src/lib.rs:274) BEFORE any candidate is
    computed -- the caller's signature buffer is untouched and no internal-framing signature is ever produced."""
    torch = pytest.importorskip("torch")
    from fips204_b200 import _lib
    L = _lib.lib()
    torch.cuda.set_device(0)
    m = F.ml_dsa_65
    n = 300
    _, sk = O.keygen_batch(65, synth_block("dv-xi", n, 32))
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.array(a, copy=True)).to(dev)  # noqa: E731
    p = lambda x: C.c_void_p(x.data_ptr())  # noqa: E731
    d_sk, d_rnd = t(sk), t(synth_rows("dv-r", n, 32))
    msg = t(synth_block("dv-m", n, 32))
    moff = torch.arange(0, 32 * (n + 1), 32, dtype=torch.int64, device=dev)
    ctx_len = np.full(n, 3, np.uint32)
    ctx_len[200] = 256
    coff = np.zeros(n + 1, np.uint32)
    coff[1:] = np.cumsum(ctx_len)
    d_ctx = torch.zeros(int(coff[-1]) + 8, dtype=torch.uint8, device=dev)
    d_coff = t(coff.view(np.int32))
    d_sig = torch.full((n, m.SIG_LEN), 0xAB, dtype=torch.uint8, device=dev)
    sh = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = L.mldsa_sign_batch_dev(65, sh, 0, n, p(d_sk), m.SK_LEN, p(msg), p(moff), p(d_ctx), p(d_coff), p(d_rnd), p(d_sig), None)
    assert rc == -2
    assert bool((d_sig == 0xAB).all())
    # and with legal contexts the same call signs like the oracle
    ctx_len[200] = 255
    coff[1:] = np.cumsum(ctx_len)
    d_coff = t(coff.view(np.int32))
    rc = L.mldsa_sign_batch_dev(65, sh, 0, n, p(d_sk), m.SK_LEN, p(msg), p(moff), p(d_ctx), p(d_coff), p(d_rnd), p(d_sig), None)
    assert rc == 0
    msgs = [bytes(r) for r in msg.cpu().numpy()]
    ctxs = [bytes(int(c)) for c in ctx_len]
    want, _ = O.sign_batch(65, sk, msgs, d_rnd.cpu().numpy(), ctxs=ctxs)
    assert (d_sig.cpu().numpy() == want).all()


def test_max_rounds_hook_leaves_no_signature_bytes(F):
    """A sign call that stops early (the max_rounds test hook, -5) must not hand back z polynomials of candidates that
    were never accepted: the device-pointer path zeroes the caller's buffer."""
    torch = pytest.importorskip("torch")
    from fips204_b200 import _lib
    L = _lib.lib()
    torch.cuda.set_device(0)
    m = F.ml_dsa_44
    n = 2000
    _, sk = O.keygen_batch(44, synth_block("mr-xi", n, 32))
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.array(a, copy=True)).to(dev)  # noqa: E731
    p = lambda x: C.c_void_p(x.data_ptr())  # noqa: E731
    d_sk, d_rnd, msg = t(sk), t(synth_rows("mr-r", n, 32)), t(synth_block("mr-m", n, 32))
    moff = torch.arange(0, 32 * (n + 1), 32, dtype=torch.int64, device=dev)
    d_sig = torch.full((n, m.SIG_LEN), 0xCD, dtype=torch.uint8, device=dev)
    sh = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    try:
        L.mldsa_debug_set_max_rounds(1)
        rc = L.mldsa_sign_batch_dev(44, sh, 0, n, p(d_sk), m.SK_LEN, p(msg), p(moff), None, None, p(d_rnd), p(d_sig), None)
    finally:
        L.mldsa_debug_set_max_rounds(0)
    assert rc == -5
    assert int(d_sig.count_nonzero()) == 0
    assert _residue() == 0


def _distinct_devices(F, need=2):
    ndev = F.device_count()
    if ndev < need:
        pytest.skip(f"needs {need} visible GPUs (this box has {ndev}); the split logic on one device is covered by "
                    "test_multi_device_entry_points_split_contiguously")
    return list(range(ndev))


def test_multi_entry_points_on_distinct_devices(F):
    """mldsa_{keygen,sign,verify,pipeline}_batch_multi over every visible GPU from ONE process (one host thread per
    device): the oracle's bytes, lane for lane.  ML-DSA-87 at depth 1 is included on purpose: its per-warp staged
    sign_w needs > 48 KB of dynamic shared memory, a per-DEVICE function attribute (round-1 advisor finding)."""
    from fips204_b200 import _lib
    L = _lib.lib()
    devs = _distinct_devices(F)
    darr = (C.c_int * len(devs))(*devs)
    for ps, depth in ((87, 1), (65, 2), (44, 3)):
        m = F.PARAM_SETS[ps]
        n = 997 * len(devs) + 5
        xi = synth_block(f"mdd-xi-{ps}", n, 32)
        msgs = [synth("mdd-m", i, i % 41) for i in range(n)]
        ctxs = [synth("mdd-c", i, i % 3) for i in range(n)]
        rnd = synth_rows("mdd-r", n, 32)
        opk, osk = O.keygen_batch(ps, xi)
        osig, oit = O.sign_batch(ps, osk, msgs, rnd, ctxs=ctxs)
        mb = np.frombuffer(b"".join(msgs) + b"\0", np.uint8).copy()
        mo = np.zeros(n + 1, np.uint64)
        mo[1:] = np.cumsum([len(x) for x in msgs])
        cb = np.frombuffer(b"".join(ctxs) + b"\0", np.uint8).copy()
        co = np.zeros(n + 1, np.uint32)
        co[1:] = np.cumsum([len(x) for x in ctxs])
        try:
            L.mldsa_set_depth(depth)
            pk = np.empty((n, m.PK_LEN), np.uint8)
            sk = np.empty((n, m.SK_LEN), np.uint8)
            _lib.check(L.mldsa_keygen_batch_multi(ps, darr, len(devs), n, xi.ctypes.data, pk.ctypes.data, sk.ctypes.data))
            assert (pk == opk).all() and (sk == osk).all(), ps
            sig = np.empty((n, m.SIG_LEN), np.uint8)
            it = np.zeros(n, np.uint32)
            _lib.check(L.mldsa_sign_batch_multi(ps, darr, len(devs), 0, n, sk.ctypes.data, m.SK_LEN, mb.ctypes.data,
                                                mo.ctypes.data, cb.ctypes.data, co.ctypes.data, rnd.ctypes.data,
                                                sig.ctypes.data, it.ctypes.data))
            assert (sig == osig).all() and (it == oit).all(), (ps, depth)
            bad = sig.copy()
            flip = [0, n // 2, n - 1]
            bad[flip, 11] ^= 4
            ok = np.zeros(n, np.uint8)
            _lib.check(L.mldsa_verify_batch_multi(ps, darr, len(devs), 0, n, pk.ctypes.data, m.PK_LEN, mb.ctypes.data,
                                                  mo.ctypes.data, cb.ctypes.data, co.ctypes.data, bad.ctypes.data,
                                                  ok.ctypes.data))
            assert list(np.nonzero(ok == 0)[0]) == flip
            out = m.pipeline_batch(xi, msgs, rnd, ctxs=ctxs, devs=devs, want_sk=True, want_iters=True, want_stage_ms=True)
            assert (out["pk"] == opk).all() and (out["sk"] == osk).all() and (out["sig"] == osig).all(), ps
            assert (out["iters"] == oit).all() and out["ok"].all()
            for d in devs:
                assert _residue(d) == 0
        finally:
            L.mldsa_set_depth(2)


def test_key_sets_live_on_their_own_device(F):
    """Resident key sets on a device other than 0, used while device 0 runs its own calls."""
    devs = _distinct_devices(F)
    d = devs[-1]
    m = F.ml_dsa_65
    n = 400
    xi = synth_block("ksd-xi", n, 32)
    opk, osk = O.keygen_batch(65, xi)
    pk, sks, pks = m.keygen_to_keysets(xi, dev=d)
    msgs = [synth("ksd-m", i, 20) for i in range(n)]
    rnd = synth_rows("ksd-r", n, 32)
    sig0 = m.sign_batch(osk, msgs, rnd, dev=0)
    sig = sks.sign_batch(msgs, rnd)
    want, _ = O.sign_batch(65, osk, msgs, rnd)
    assert (pk == opk).all() and (sig == want).all() and (sig0 == want).all()
    assert pks.verify_batch(msgs, sig).all()
    sks.free()
    pks.free()


def test_ntt_device_entry_points_want_32_byte_aligned_polynomials(F):
    """The standalone transforms move a lane's coefficients as 256-bit vectors: a device pointer that is not 32-byte
    aligned is a parameter error (-1, include/mldsa_b200.h), not a misaligned-address fault; an aligned one works."""
    torch = pytest.importorskip("torch")
    from fips204_b200 import _lib
    L = _lib.lib()
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    polys = (synth_block("ntt-al", 3, 1024).view(np.uint32).astype(np.int64) % 8380417).astype(np.int32).reshape(3, 256)
    buf = torch.zeros(3 * 256 + 8, dtype=torch.int32, device=dev)
    sh = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for fn in (L.mldsa_ntt_batch_dev, L.mldsa_intt_batch_dev):
        for off in (1, 4):  # 4 and 16 bytes past an aligned address
            assert fn(sh, 3, C.c_void_p(buf.data_ptr() + 4 * off)) == -1
    buf[:768] = torch.from_numpy(polys.reshape(-1)).to(dev)
    assert L.mldsa_ntt_batch_dev(sh, 3, C.c_void_p(buf.data_ptr())) == 0
    assert (buf[:768].cpu().numpy().reshape(3, 256) == O.ntt_batch(polys)).all()
