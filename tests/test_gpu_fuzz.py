"""Randomised negative / fuzz corpora: verify verdicts of the CUDA engine against the CPU oracle, lane by lane, on
>= 100 000 lanes per parameter set.  The corpus shapes follow the reference's fuzz target
(fuzz/fuzz_targets/fuzz_all.rs): a good signature XORed with arbitrary bytes (:14-15), arbitrary signature bytes
(:22), arbitrary public-key bytes (:25-26, :36-38), arbitrary secret-key bytes used for signing (:18-19, :29-33) --
plus corpora aimed at the decoder's branches (hint region, z at the norm bound), which random bytes alone almost
never reach.  Every corpus runs through the three verify paths: packed keys per lane, a resident public-key set, and
(for the lanes of one key) the shared-key path with its cached A-hat.  Seeds are fixed; the oracle is the judge."""
import numpy as np
import pytest

import oracle_lib as O
from synth import synth, synth_block, synth_rows

pytestmark = pytest.mark.gpu

NKEYS = 64
POOL = 2048          # good signatures the mutations start from
PER_CLASS = 20480    # lanes per corpus class; 6 classes -> 122 880 lanes per parameter set


@pytest.fixture(scope="module")
def F():
    import fips204_b200 as F
    assert F.device_count() >= 1, "no CUDA device: GPU tests need the B200 box"
    return F


def _sparse_xor(rng, sigs):
    """fuzz_all.rs:14-15 with masks of random weight: 1 .. 64 flipped bits at uniform positions, or a short run of
    random bytes at a random offset (what a byte-level mutator produces)."""
    n, ln = sigs.shape
    out = sigs.copy()
    w = rng.choice([1, 1, 1, 2, 2, 3, 4, 8, 16, 64], size=n)
    rows = np.arange(n)
    for k in range(64):
        act = rows[w > k]
        pos = rng.integers(0, ln * 8, size=act.size)
        out[act, pos >> 3] ^= (1 << (pos & 7)).astype(np.uint8)
    runs = rows[rng.random(n) < 0.25]
    for r in runs:
        a = int(rng.integers(0, ln - 8))
        out[r, a:a + int(rng.integers(1, 9))] ^= rng.integers(0, 256, size=1, dtype=np.uint8)
    return out


def _hint_region(rng, p, sigs):
    """Rebuild the hint section (conversion.rs:277-328 layout: omega index bytes, then k cumulative counts) in ways that
    walk every branch of HintBitUnpack (conversion.rs:340-414): well-formed but different hints, repeated or decreasing
    indices inside a polynomial, non-zero padding, counts that decrease or exceed omega."""
    n = sigs.shape[0]
    out = sigs.copy()
    hoff = p.lambda_div4 + p.l * 32 * p.z_bits
    for r in range(n):
        kind = r % 8
        total = int(rng.integers(0, p.omega + 1))
        cuts = np.sort(rng.integers(0, total + 1, size=p.k - 1))
        counts = np.concatenate([cuts, [total]]).astype(np.int64)
        idx = np.zeros(p.omega, np.int64)
        lo = 0
        for k in range(p.k):
            c = int(counts[k]) - lo
            if c > 256:
                c = 256
            idx[lo:lo + c] = np.sort(rng.choice(256, size=c, replace=False))
            lo += c
        if kind == 1 and total >= 2:      # equal neighbours inside one polynomial
            j = int(rng.integers(1, total))
            idx[j] = idx[j - 1]
        elif kind == 2 and total >= 2:    # decreasing index
            j = int(rng.integers(1, total))
            idx[j - 1], idx[j] = max(idx[j - 1], idx[j]) | 1, 0
        elif kind == 3 and total < p.omega:  # non-zero padding after the last index
            idx[int(rng.integers(total, p.omega))] = int(rng.integers(1, 256))
        elif kind == 4:                   # a count beyond omega
            counts[int(rng.integers(0, p.k))] = int(rng.integers(p.omega + 1, 256))
        elif kind == 5 and p.k >= 2:      # counts not monotone
            j = int(rng.integers(1, p.k))
            counts[j - 1], counts[j] = max(counts[j - 1], counts[j]), min(counts[j - 1], counts[j]) - (1 if counts[j] == counts[j - 1] and counts[j] > 0 else 0)
        elif kind == 6:                   # arbitrary bytes in the whole region
            idx = rng.integers(0, 256, size=p.omega)
            counts = rng.integers(0, 256, size=p.k)
        # kinds 0 and 7: well-formed, merely different from the signer's hints
        out[r, hoff:hoff + p.omega] = (idx & 0xFF).astype(np.uint8)
        out[r, hoff + p.omega:hoff + p.omega + p.k] = (np.clip(counts, 0, 255)).astype(np.uint8)
    return out


def _z_boundary(rng, p, sigs):
    """Force a few z coefficients to +-(gamma1 - beta) (rejected, src/ml_dsa.rs:434) or +-(gamma1 - beta - 1) (accepted
    by the norm check; the commitment hash then decides)."""
    n = sigs.shape[0]
    out = sigs.copy()
    zoff = p.lambda_div4
    for r in range(n):
        for _ in range(int(rng.integers(1, 4))):
            coef = int(rng.integers(0, p.l * 256))
            mag = p.gamma1 - p.beta - int(rng.integers(0, 2))
            z = mag if rng.random() < 0.5 else -mag
            field = p.gamma1 - z
            bit = coef * p.z_bits
            a = zoff + (bit >> 3)
            v = int.from_bytes(out[r, a:a + 4].tobytes(), "little")
            sh = bit & 7
            v = (v & ~(((1 << p.z_bits) - 1) << sh)) | (field << sh)
            out[r, a:a + 4] = np.frombuffer((v & 0xFFFFFFFF).to_bytes(4, "little"), np.uint8)
    return out


@pytest.fixture(scope="module")
def corpus_factory():
    cache = {}

    def make(ps):
        if ps in cache:
            return cache[ps]
        p = O.params(ps)
        rng = np.random.default_rng(0xF204_0000 + ps)
        pks, sks = O.keygen_batch(ps, synth_block(f"fz-xi-{ps}", NKEYS, 32))
        pool_key = np.arange(POOL) % NKEYS
        pool_msgs = [synth("fz-m", i, 4 + i % 29) for i in range(POOL)]
        pool_sig, _ = O.sign_batch(ps, sks[pool_key], pool_msgs, synth_rows("fz-r", POOL, 32))
        classes = []

        def from_pool(nl):
            j = rng.integers(0, POOL, size=nl)
            return j, pool_sig[j].copy()

        j, s = from_pool(PER_CLASS)
        classes.append(("xor-sparse", pool_key[j], j, _sparse_xor(rng, s), None))
        j, s = from_pool(PER_CLASS)
        classes.append(("hint-region", pool_key[j], j, _hint_region(rng, p, s), None))
        j, s = from_pool(PER_CLASS)
        classes.append(("z-boundary", pool_key[j], j, _z_boundary(rng, p, s), None))
        j = rng.integers(0, POOL, size=PER_CLASS)  # fuzz_all.rs:22: arbitrary signature bytes, good key
        classes.append(("random-sig", pool_key[j], j, rng.integers(0, 256, size=(PER_CLASS, p.sig_len), dtype=np.uint8), None))
        # fuzz_all.rs:25-26, :36-38: arbitrary public-key bytes, under a good signature (half) or arbitrary bytes (half)
        j, s = from_pool(PER_CLASS)
        s[PER_CLASS // 2:] = rng.integers(0, 256, size=(PER_CLASS - PER_CLASS // 2, p.sig_len), dtype=np.uint8)
        fuzz_pk = rng.integers(0, 256, size=(NKEYS, p.pk_len), dtype=np.uint8)
        classes.append(("random-pk", rng.integers(0, NKEYS, size=PER_CLASS), j, s, fuzz_pk))
        # untouched signatures and signatures moved to another key / message of the pool (must all be decided alike)
        j, s = from_pool(PER_CLASS)
        kk = pool_key[j].copy()
        swap = rng.random(PER_CLASS) < 0.5
        kk[swap] = (kk[swap] + 1 + rng.integers(0, NKEYS - 1, size=int(swap.sum()))) % NKEYS
        classes.append(("good-or-wrong-key", kk, j, s, None))
        cache[ps] = (p, pks, sks, pool_msgs, classes)
        return cache[ps]

    return make


@pytest.mark.parametrize("ps", [44, 65, 87])
def test_verify_fuzz_corpora_match_oracle(F, corpus_factory, ps):
    m = F.PARAM_SETS[ps]
    p, pks, sks, pool_msgs, classes = corpus_factory(ps)
    total = 0
    seen_true = seen_false = 0
    for name, kidx, midx, sigs, fuzz_pk in classes:
        keys = pks if fuzz_pk is None else fuzz_pk
        n = sigs.shape[0]
        msgs = [pool_msgs[i] for i in midx]
        lane_pk = np.ascontiguousarray(keys[kidx])
        want = O.verify_batch(ps, lane_pk, msgs, sigs)
        # 1. packed public keys, one per lane
        got = m.verify_batch(lane_pk, msgs, sigs)
        bad = np.nonzero(got != want)[0]
        assert bad.size == 0, (ps, name, "distinct", bad[:8], want[bad[:8]])
        # 2. resident public-key set with a key index per lane
        kset = m.load_public_keys(keys)
        got = kset.verify_batch(msgs, sigs, key_idx=kidx.astype(np.uint32))
        kset.free()
        bad = np.nonzero(got != want)[0]
        assert bad.size == 0, (ps, name, "resident", bad[:8], want[bad[:8]])
        # 3. shared-key path (A-hat, tr, t1-hat cached once) for the lanes of two of the keys
        for key in (0, NKEYS - 1):
            sel = np.nonzero(kidx == key)[0]
            if sel.size == 0:
                continue
            got = m.verify_batch(keys[key].tobytes(), [msgs[i] for i in sel], np.ascontiguousarray(sigs[sel]), shared_key=True)
            assert (got == want[sel]).all(), (ps, name, "shared", key)
        total += n
        seen_true += int(want.sum())
        seen_false += int(n - want.sum())
    assert total >= 100_000
    assert seen_true > 1000 and seen_false > 50_000  # the corpus exercises both verdicts


@pytest.mark.parametrize("ps", [44, 65, 87])
def test_shared_key_fuzz_corpus_match_oracle(F, corpus_factory, ps):
    """32 768 mutated signatures under ONE public key through the shared-key path (BASELINE config 3's shape)."""
    m = F.PARAM_SETS[ps]
    p, pks, sks, pool_msgs, _ = corpus_factory(ps)
    rng = np.random.default_rng(0x5AED_0000 + ps)
    nb = 512
    msgs0 = [synth("fzs-m", i, 16) for i in range(nb)]
    base, _ = O.sign_batch(ps, sks[7], msgs0, synth_rows("fzs-r", nb, 32), shared_key=True)
    n = 32768
    j = rng.integers(0, nb, size=n)
    sigs = base[j].copy()
    third = n // 4
    sigs[:third] = _sparse_xor(rng, sigs[:third])
    sigs[third:2 * third] = _hint_region(rng, p, sigs[third:2 * third])
    sigs[2 * third:3 * third] = _z_boundary(rng, p, sigs[2 * third:3 * third])
    msgs = [msgs0[i] for i in j]
    want = O.verify_batch(ps, pks[7], msgs, sigs, shared_key=True)
    got = m.verify_batch(pks[7].tobytes(), msgs, sigs, shared_key=True)
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, (ps, bad[:8], want[bad[:8]])
    assert want[3 * third:].all() and not want[:third].all()


@pytest.mark.parametrize("ps", [44, 65, 87])
def test_sign_with_fuzzed_secret_keys_matches_oracle(F, ps):
    """fuzz_all.rs:18-19, :29-33: PrivateKey::try_from_bytes accepts any correctly sized byte string (the range checks
    of the reference are tautological, conversion.rs:259-260), and signing with it is the same mod-q computation.
    2 048 arbitrary secret keys: signatures and trip counts equal the oracle's; packed keys and a resident set."""
    m = F.PARAM_SETS[ps]
    n = 2048
    rng = np.random.default_rng(0x5EC0_0000 + ps)
    sks = rng.integers(0, 256, size=(n, m.SK_LEN), dtype=np.uint8)
    msgs = [synth("fsk-m", i, i % 17) for i in range(n)]
    rnd = synth_rows("fsk-r", n, 32)
    want, wit = O.sign_batch(ps, sks, msgs, rnd)
    got, it = m.sign_batch(sks, msgs, rnd, want_iters=True)
    assert (got == want).all() and (it == wit).all()
    kset = m.load_private_keys(sks)
    assert (kset.sign_batch(msgs, rnd) == want).all()
    kset.free()
    # the public keys such secret keys imply (Signer::get_public_key, src/ml_dsa.rs:502-563) accept the signatures
    # whenever the oracle says so
    pk = m.sk_to_pk_batch(sks)
    wv = O.verify_batch(ps, pk, msgs, want)
    assert (m.verify_batch(pk, msgs, got) == wv).all()
