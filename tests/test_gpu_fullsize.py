"""BASELINE.json's configurations at their FULL sizes (2^20 lanes on one GPU), checked through properties that do
not need the CPU oracle to run a million lanes: round trips (sign -> verify, NTT -> INTT), a known verdict pattern
for a structured mutation set, independence of the result from the execution path
(device-resident vs the chunk-pipelined host entry point), iteration statistics, plus byte-for-byte comparison
with the oracle on an evenly spaced sample of lanes.  Device buffers come from torch (plumbing only); every compute
call goes through the C ABI."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
Q = 8380417
N = 1 << 20


@pytest.fixture(scope="module")
def G():
    torch = pytest.importorskip("torch")
    import fips204_b200 as F
    from fips204_b200 import _lib
    assert torch.cuda.is_available() and F.device_count() >= 1, "no CUDA device: GPU tests need the B200 box"
    torch.cuda.set_device(0)

    class Ctx:
        pass

    g = Ctx()
    g.torch, g.F, g.L, g.check = torch, F, _lib.lib(), _lib.check
    g.dev = torch.device("cuda", 0)
    g.sh = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    g.p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    g.gen = torch.Generator(device=g.dev)
    g.gen.manual_seed(0xF1B5204)
    g.rand = lambda *shape: torch.randint(0, 256, shape, dtype=torch.uint8, device=g.dev, generator=g.gen)
    yield g
    torch.cuda.empty_cache()


def _keys(g, ps, n):
    m = g.F.PARAM_SETS[ps]
    xi = g.rand(n, 32)
    pk = g.torch.empty((n, m.PK_LEN), dtype=g.torch.uint8, device=g.dev)
    sk = g.torch.empty((n, m.SK_LEN), dtype=g.torch.uint8, device=g.dev)
    g.check(g.L.mldsa_keygen_batch_dev(ps, g.sh, n, g.p(xi), g.p(pk), g.p(sk)))
    return xi, pk, sk


def _sign(g, ps, sk, sk_stride, msg, moff, rnd, iters=None):
    m = g.F.PARAM_SETS[ps]
    n = msg.shape[0]
    sig = g.torch.empty((n, m.SIG_LEN), dtype=g.torch.uint8, device=g.dev)
    g.check(g.L.mldsa_sign_batch_dev(ps, g.sh, 0, n, g.p(sk), sk_stride, g.p(msg), g.p(moff), None, None, g.p(rnd),
                                     g.p(sig), g.p(iters)))
    return sig


def _verify(g, ps, pk, pk_stride, msg, moff, sig):
    n = msg.shape[0]
    ok = g.torch.zeros(n, dtype=g.torch.uint8, device=g.dev)
    g.check(g.L.mldsa_verify_batch_dev(ps, g.sh, 0, n, g.p(pk), pk_stride, g.p(msg), g.p(moff), None, None, g.p(sig),
                                       g.p(ok)))
    return ok


def test_config2_ntt_roundtrip_and_expand_a_full_size(G):
    """Config 2: 2^20 polynomials.  INTT(NTT(x) mod q) == x for every polynomial; NTT output equals the oracle's raw
    (non-canonical, src/ntt.rs:42-54) representatives on a sample; expand_a rows are in [0, q) and equal the
    oracle's RejNTTPoly on a sample."""
    g, t = G, G.torch
    x = t.randint(0, Q, (N, 256), dtype=t.int32, device=g.dev, generator=g.gen)
    y = x.clone()
    g.check(g.L.mldsa_ntt_batch_dev(g.sh, N, g.p(y)))
    idx = np.linspace(0, N - 1, 257).astype(np.int64)
    want = O.ntt_batch(x[idx].cpu().numpy())
    assert (y[idx].cpu().numpy() == want).all()
    y.remainder_(Q)
    g.check(g.L.mldsa_intt_batch_dev(g.sh, N, g.p(y)))
    assert bool(t.equal(x, y))
    del x
    seeds = g.rand(N, 34)
    g.check(g.L.mldsa_expand_a_batch_dev(g.sh, N, g.p(seeds), g.p(y)))
    assert int(y.min()) >= 0 and int(y.max()) < Q
    assert (y[idx].cpu().numpy() == O.expand_a_batch(seeds[idx].cpu().numpy())).all()
    # a stream's output depends on its own seed only: same seeds in another order, same rows
    perm = t.randperm(N, device=g.dev, generator=g.gen)
    z = t.empty_like(y)
    g.check(g.L.mldsa_expand_a_batch_dev(g.sh, N, g.p(seeds[perm].contiguous()), g.p(z)))
    assert bool(t.equal(z, y[perm]))


def test_config3_shared_key_verify_mutation_pattern_full_size(G):
    """Config 3: ML-DSA-44, 2^20 signatures under ONE public key; every 8th lane carries one of eight corruptions
    (SURVEY.md 8d): the verdict vector must be exactly `lane % 8 != 0`, and must equal the oracle's on a sample."""
    g, t, ps = G, G.torch, 44
    m = g.F.PARAM_SETS[ps]
    p = O.params(ps)
    _, pk, sk = _keys(g, ps, 1)
    msg = g.rand(N, 32)
    rnd = g.rand(N, 32)
    moff = t.arange(N + 1, dtype=t.int64, device=g.dev) * 32
    sig = _sign(g, ps, sk, 0, msg, moff, rnd)
    assert bool(_verify(g, ps, pk, 0, msg, moff, sig).all())
    ct, zlen = p.lambda_div4, 32 * p.z_bits
    hint = ct + p.l * zlen
    lanes = t.arange(0, N, 8, device=g.dev)
    cls = (lanes // 8) % 8
    sel = lambda c: lanes[cls == c]
    sig[sel(0), 0] ^= 1                                   # bit flip in c~
    sig[sel(1), ct + 5] ^= 0x10                           # bit flip in z
    sig[sel(2), hint] ^= 1                                # bit flip in the hint region
    sig[sel(3), m.SIG_LEN - 1] = p.omega + 1              # hint count > omega           (conversion.rs:357)
    s4 = sel(4)                                           # indices not strictly increasing (conversion.rs:373)
    sig[s4, hint:hint + p.omega] = 0
    sig[s4, hint] = 9
    sig[s4, hint + 1] = 3
    sig[s4, hint + p.omega:] = 2
    sig[sel(5), hint + p.omega - 1] ^= 1                  # non-zero padding (conversion.rs:398); with exactly omega
                                                          # hints (0.3 % of the lanes) it is the last index instead
    s6 = sel(6)                                           # z[0][0] = gamma1 - beta: field value beta in 18 bits
    sig[s6, ct] = p.beta & 0xFF
    sig[s6, ct + 1] = (p.beta >> 8) & 0xFF
    sig[s6, ct + 2] &= 0xFC
    msg[sel(7), 0] ^= 1                                   # wrong message
    ok = _verify(g, ps, pk, 0, msg, moff, sig)
    want = (t.arange(N, device=g.dev) % 8 != 0).to(t.uint8)
    if not bool(t.equal(ok, want)):
        d = t.nonzero(ok != want).flatten()[:12].cpu().numpy()
        o = O.verify_batch(ps, pk[0].cpu().numpy(), [bytes(x) for x in msg[d].cpu().numpy()], sig[d].cpu().numpy(), shared_key=True)
        raise AssertionError(f"{int((ok != want).sum())} verdicts differ from the mutation pattern: lanes {d.tolist()}, classes "
                             f"{[(int(i) // 8) % 8 for i in d]}, gpu {ok[d].cpu().tolist()}, oracle {o.tolist()}, "
                             f"hint tails {[sig[int(i), hint + p.omega - 2:].cpu().tolist() for i in d[:4]]}")
    idx = np.arange(0, 4096)                              # 512 corrupted lanes of all eight classes + 3584 valid ones
    o = O.verify_batch(ps, pk[0].cpu().numpy(), [bytes(x) for x in msg[idx].cpu().numpy()], sig[idx].cpu().numpy(),
                       shared_key=True)
    assert (o == ok[idx].cpu().numpy()).all()
    # through the host-pointer entry point (asynchronous chunk pipeline, pinned buffers) the verdicts are the same
    h_pk, h_msg, h_moff, h_sig = pk.cpu().pin_memory(), msg.cpu().pin_memory(), moff.cpu().pin_memory(), sig.cpu().pin_memory()
    h_ok = t.zeros(N, dtype=t.uint8).pin_memory()
    g.check(g.L.mldsa_verify_batch(ps, 0, 0, N, g.p(h_pk), 0, g.p(h_msg), g.p(h_moff), None, None, g.p(h_sig), g.p(h_ok)))
    assert bool(t.equal(h_ok, want.cpu()))


def test_config4_sign65_distinct_keys_full_size(G):
    """Config 4: ML-DSA-65, 2^20 distinct keys / messages.  Every signature verifies; a 512-lane sample is
    byte-identical to the oracle (signature and iteration count); the mean iteration count is FIPS 204's ~5.1;
    the chunk-pipelined host entry point returns the same bytes as the device-resident call."""
    g, t, ps = G, G.torch, 65
    m = g.F.PARAM_SETS[ps]
    xi, pk, sk = _keys(g, ps, N)
    # the keys themselves against the oracle on a 65 536-key sample (signatures that verify under the GPU's own public
    # keys cannot catch a wrong key pair): every 16th lane
    kidx = np.arange(0, N, 16)
    opk, osk = O.keygen_batch(ps, xi[kidx].cpu().numpy())
    assert (opk == pk[kidx].cpu().numpy()).all() and (osk == sk[kidx].cpu().numpy()).all()
    del opk, osk
    msg = g.rand(N, 32)
    rnd = g.rand(N, 32)
    moff = t.arange(N + 1, dtype=t.int64, device=g.dev) * 32
    it = t.zeros(N, dtype=t.int32, device=g.dev)
    sig = _sign(g, ps, sk, m.SK_LEN, msg, moff, rnd, it)
    assert bool(_verify(g, ps, pk, m.PK_LEN, msg, moff, sig).all())
    mean = float(it.to(t.float64).mean())
    assert 4.9 < mean < 5.3 and int(it.min()) >= 1, mean
    idx = np.linspace(0, N - 1, 512).astype(np.int64)
    osig, oit = O.sign_batch(ps, sk[idx].cpu().numpy(), [bytes(x) for x in msg[idx].cpu().numpy()], rnd[idx].cpu().numpy())
    assert (osig == sig[idx].cpu().numpy()).all() and (oit == it[idx].cpu().numpy().astype(np.uint32)).all()
    # host-pointer entry point, pinned buffers: nine chunks, overlapped tails -- same bytes
    h_sk, h_msg, h_rnd, h_moff = sk.cpu().pin_memory(), msg.cpu().pin_memory(), rnd.cpu().pin_memory(), moff.cpu().pin_memory()
    h_sig = t.empty((N, m.SIG_LEN), dtype=t.uint8).pin_memory()
    h_it = t.zeros(N, dtype=t.int32).pin_memory()
    g.check(g.L.mldsa_sign_batch(ps, 0, 0, N, g.p(h_sk), m.SK_LEN, g.p(h_msg), g.p(h_moff), None, None, g.p(h_rnd),
                                 g.p(h_sig), g.p(h_it)))
    assert bool(t.equal(h_sig, sig.cpu())) and bool(t.equal(h_it, it.cpu()))
    # and the host-pointer verify of those bytes, with one lane's signature swapped for its neighbour's
    h_pk = pk.cpu().pin_memory()
    h_ok = t.zeros(N, dtype=t.uint8).pin_memory()
    h_sig[123457] = h_sig[123456]
    g.check(g.L.mldsa_verify_batch(ps, 0, 0, N, g.p(h_pk), m.PK_LEN, g.p(h_msg), g.p(h_moff), None, None, g.p(h_sig),
                                   g.p(h_ok)))
    assert int(h_ok.sum()) == N - 1 and int(h_ok[123457]) == 0


def test_config5_mldsa87_keygen_sign_verify_shard(G):
    """Config 5: ML-DSA-87 keygen -> sign -> verify with the keys and signatures staying in HBM between the stages,
    on one GPU's shard of the 8-GPU batch (2^20 / 8 lanes).  Keys and signatures equal the oracle's on a sample,
    all verdicts are true, and a mutated subset is rejected."""
    g, t, ps = G, G.torch, 87
    m = g.F.PARAM_SETS[ps]
    n = N // 8
    xi, pk, sk = _keys(g, ps, n)
    msg = g.rand(n, 32)
    rnd = g.rand(n, 32)
    moff = t.arange(n + 1, dtype=t.int64, device=g.dev) * 32
    it = t.zeros(n, dtype=t.int32, device=g.dev)
    sig = _sign(g, ps, sk, m.SK_LEN, msg, moff, rnd, it)
    assert bool(_verify(g, ps, pk, m.PK_LEN, msg, moff, sig).all())
    idx = np.linspace(0, n - 1, 256).astype(np.int64)
    opk, osk = O.keygen_batch(ps, xi[idx].cpu().numpy())
    assert (opk == pk[idx].cpu().numpy()).all() and (osk == sk[idx].cpu().numpy()).all()
    osig, oit = O.sign_batch(ps, osk, [bytes(x) for x in msg[idx].cpu().numpy()], rnd[idx].cpu().numpy())
    assert (osig == sig[idx].cpu().numpy()).all() and (oit == it[idx].cpu().numpy().astype(np.uint32)).all()
    bad = t.arange(5, n, 1000, device=g.dev)
    sig[bad, 70] ^= 0x20
    ok = _verify(g, ps, pk, m.PK_LEN, msg, moff, sig)
    want = t.ones(n, dtype=t.uint8, device=g.dev)
    want[bad] = 0
    assert bool(t.equal(ok, want))
    # get_public_key over the whole shard gives back the generated public keys
    pk2 = t.empty_like(pk)
    g.check(g.L.mldsa_sk_to_pk_batch_dev(ps, g.sh, n, g.p(sk), g.p(pk2)))
    assert bool(t.equal(pk, pk2))


def test_config5_pipeline_full_size_over_all_gpus(G):
    """Config 5 through its entry point: mldsa_pipeline_batch_multi runs ML-DSA-87 keygen -> sign -> verify for ONE batch
    of 2^20 lanes split over every visible GPU (one GPU: the whole batch), only xi / msg / rnd going in and pk / sig /
    verdicts coming out.  All verdicts are true; a 256-lane sample equals the oracle's keys, signatures and trip counts;
    the per-stage device times are reported; a mutated subset of the returned signatures is rejected under the returned
    public keys."""
    g, t, ps = G, G.torch, 87
    m = g.F.PARAM_SETS[ps]
    ndev = g.F.device_count()
    devs = (C.c_int * ndev)(*range(ndev))
    rng = np.random.default_rng(87)
    xi = rng.integers(0, 256, size=(N, 32), dtype=np.uint8)
    msg = rng.integers(0, 256, size=(N, 32), dtype=np.uint8)
    rnd = rng.integers(0, 256, size=(N, 32), dtype=np.uint8)
    moff = np.arange(N + 1, dtype=np.uint64) * 32
    pk = np.empty((N, m.PK_LEN), np.uint8)
    sig = np.empty((N, m.SIG_LEN), np.uint8)
    ok = np.zeros(N, np.uint8)
    it = np.zeros(N, np.uint32)
    ms = (C.c_double * 3)()
    g.check(g.L.mldsa_pipeline_batch_multi(ps, devs, ndev, 0, N, xi.ctypes.data, msg.ctypes.data, moff.ctypes.data, None, None,
                                           rnd.ctypes.data, pk.ctypes.data, None, sig.ctypes.data, ok.ctypes.data, it.ctypes.data, ms))
    assert ok.all() and all(x > 0 for x in ms)
    assert 3.6 < it.mean() < 4.1, it.mean()  # FIPS 204: 3.85 expected iterations for ML-DSA-87
    idx = np.linspace(0, N - 1, 256).astype(np.int64)
    opk, osk = O.keygen_batch(ps, xi[idx])
    osig, oit = O.sign_batch(ps, osk, [bytes(x) for x in msg[idx]], rnd[idx])
    assert (opk == pk[idx]).all() and (osig == sig[idx]).all() and (oit == it[idx]).all()
    sub = np.arange(0, N, 997)
    bad = sig[sub].copy()
    bad[::2, 100] ^= 1
    v = g.F.PARAM_SETS[ps].verify_batch(pk[sub], [bytes(x) for x in msg[sub]], bad)
    assert (v[1::2] == 1).all() and (v[0::2] == 0).all()
