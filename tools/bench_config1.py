#!/usr/bin/env python3
"""BASELINE.json configs[0]: the reference's own CPU benchmark (benches/benchmark.rs), reproduced with the CPU
oracle because the crate cannot be built here (no rustc / cargo).

Inputs exactly as benches/benchmark.rs builds them:
  * counter RNG (:20-25): the v-th 32-byte draw is all (v & 0xff), bytes 0..3 = BE32(v); v = 0, 1, 2, ...
  * keys: ML-DSA-44 / 65 / 87 from draws 0 / 1 / 2 (:35-42); message 00 01 .. 07 (:33)
  * "keygen": try_keygen_with_rng with the running counter (:44-46)
  * "sk sign": try_sign_with_rng(rng, msg, ctx = []) with the running counter as rnd (:48-56)
  * "pk verify": verify(msg, sig, ctx = []) of a signature made with ctx = [0] -- a FAILING verification (:36, :58-60)
Prints microseconds per operation on ONE host thread beside the reference's published single-core figures
(benches/README.md:16-26, i7-7700K @ 4.2 GHz), and with --gpu runs the same lanes through the engine and checks
that keys, signatures, iteration counts and verdicts equal the oracle's byte for byte.
    python tools/bench_config1.py [--ops 400] [--gpu]
"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib as O  # noqa: E402

PUBLISHED_US = {44: (104.89, 226.32, 21.016), 65: (194.80, 352.89, 27.996), 87: (290.24, 385.05, 36.468)}
MSG = bytes(range(8))


def draw(v):
    b = bytearray([v & 0xFF] * 32)
    b[0:4] = (v & 0xFFFFFFFF).to_bytes(4, "big")
    return bytes(b)


def draws(v0, n):
    return np.frombuffer(b"".join(draw(v0 + i) for i in range(n)), dtype=np.uint8).reshape(n, 32)


def workload(ps, n):
    """(xi for the timed keygens, sk / pk of the bench's key, rnd for the timed signs, the ctx=[0] signature)"""
    pk, sk = O.keygen(ps, draw({44: 0, 65: 1, 87: 2}[ps]))
    sig0 = O.sign(ps, sk, MSG, draw(1 << 20), ctx=b"\0")
    return draws(3, n), pk, sk, draws(3 + n, n), sig0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", type=int, default=400)
    ap.add_argument("--gpu", action="store_true")
    a = ap.parse_args()
    n = a.ops
    print(f"config 1 (benches/benchmark.rs inputs), {n} operations each, CPU oracle on 1 thread")
    print(f"{'':10s} {'keygen us':>10s} {'(ref)':>8s} {'sign us':>10s} {'(ref)':>8s} {'verify us':>10s} {'(ref)':>8s}  mean sign iters")
    for ps in (44, 65, 87):
        xi, pk, sk, rnd, sig0 = workload(ps, n)
        t0 = time.perf_counter()
        pks, sks = O.keygen_batch(ps, xi, nthreads=1)
        t1 = time.perf_counter()
        sig, it = O.sign_batch(ps, np.frombuffer(sk, np.uint8), [MSG] * n, rnd, ctxs=[b""] * n, nthreads=1, shared_key=True)
        t2 = time.perf_counter()
        ok = O.verify_batch(ps, np.frombuffer(pk, np.uint8), [MSG] * n, np.tile(np.frombuffer(sig0, np.uint8), (n, 1)),
                            ctxs=[b""] * n, nthreads=1, shared_key=True)
        t3 = time.perf_counter()
        assert not ok.any(), "the bench's verify is a failing verification (ctx mismatch)"
        ref = PUBLISHED_US[ps]
        print(f"ML-DSA-{ps:<3d} {1e6 * (t1 - t0) / n:10.1f} {ref[0]:8.1f} {1e6 * (t2 - t1) / n:10.1f} {ref[1]:8.1f} "
              f"{1e6 * (t3 - t2) / n:10.1f} {ref[2]:8.1f}  {it.mean():.2f}")
        if a.gpu:
            import fips204_b200 as F
            m = F.PARAM_SETS[ps]
            gpk, gsk = m.keygen_batch(xi)
            gsig, git = m.sign_batch(sk, [MSG] * n, rnd, ctxs=[b""] * n, shared_key=True, want_iters=True)
            gok = m.verify_batch(pk, [MSG] * n, np.tile(np.frombuffer(sig0, np.uint8), (n, 1)), ctxs=[b""] * n, shared_key=True)
            same = (gpk == pks).all() and (gsk == sks).all() and (gsig == sig).all() and (git == it).all() and not gok.any()
            print(f"           GPU engine on the same {n} lanes: keys, signatures, iteration counts, verdicts "
                  f"{'identical' if same else 'DIFFER'}")
            assert same
            assert m.verify_batch(pk, [MSG] * n, np.tile(np.frombuffer(sig0, np.uint8), (n, 1)), ctxs=[b"\0"] * n,
                                  shared_key=True).all()


if __name__ == "__main__":
    main()
