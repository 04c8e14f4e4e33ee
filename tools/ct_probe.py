#!/usr/bin/env python3
"""dudect-style two-class timing of the sign kernels on the GPU (the reference's harness is
dudect/src/main.rs:30-56: keygen + sign timed for two fixed RNG seeds, Welch's t-test, |t| > 5 = not constant time).

A GPU kernel launch covers a whole batch, so the unit of measurement is one LAUNCH over n lanes whose secret inputs
all belong to one class, and the observable is that kernel's device time (CUDA events around the launch,
mldsa_profile_*).  mldsa_debug_set_max_rounds(1) stops the rejection loop after its first round, so every kernel
runs exactly once over all n lanes per call and the work is the same for every class:

  sign_key   sk_decode + NTT of s1, s2, t0                 touches the secret key directly
  sign_prep  mu, rho'' = H(K || rnd || mu)                 touches K
  mask       y = ExpandMask(rho'', kappa)                  touches rho''
  sign_w     w = INTT(A o NTT(y)), w1 = HighBits(w)        touches y
  ctilde     c~ = H(mu || w1), SampleInBall                public data (w1 is part of what the signature reveals)
  sign_z     c*s1, c*s2, c*t0, norm checks, hints          touches s-hat; EARLY EXIT at the first failed check --
                                                           the rejection pattern is the documented carve-out
                                                           (loop trip count, src/lib.rs:85-89), reported separately

Classes (messages and rnd are the same in every class; only the keys differ):
  A: all n lanes sign with the key generated from xi = 0xAA..AA      (the reference's "left" seed)
  B: all n lanes sign with the key generated from xi = 0x55..55      (the reference's "right" seed)
  R: n distinct random keys (fixed-vs-random)
  A': the bytes of class A once more, in a different device allocation (CONTROL: what separates A from A' is not data)
Calls of the classes are interleaved in random order.  Output: per kernel, mean device time per class and
Welch's t for A vs B, A vs R and the control A vs A'.

--op keygen does the same for key generation, the other half of the reference's harness (dudect/src/main.rs:48-52 times
keygen + sign together).  The classes are the seeds themselves (A: every lane xi = 0xAA.., B: 0x55.., R: random xi):

  kg_seed      (rho, rho', K) = H(xi || k || l)              touches xi
  rej_bounded  s1, s2 = RejBoundedPoly(rho' || r)            touches rho'; the number of SHAKE256 blocks a stream needs
                                                             depends on it -- the reference's own expand_s has the
                                                             same data-dependent loop (hashing.rs:158-213): carve-out
  kg_arith     NTT(s1), A-hat o s1-hat (A-hat squeezed from the PUBLIC rho), INTT, + s2, Power2Round, packing
                                                             touches s1, s2, t0
  shake_rows   tr = H(pk)                                    public data

    python tools/ct_probe.py [--op sign|keygen] [--lanes 16384] [--reps 300] [--ps 65] [--out profiles/ct_probe.md] [--json f]
"""
import argparse
import ctypes as C
import math
import random
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import fips204_b200 as F  # noqa: E402
from fips204_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--lanes", type=int, default=16384)
ap.add_argument("--reps", type=int, default=300)
ap.add_argument("--ps", type=int, default=65)
ap.add_argument("--out", default="")
ap.add_argument("--op", default="sign", choices=["sign", "keygen"])
ap.add_argument("--json", default="")
a = ap.parse_args()
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
L = _lib.lib()
m = F.PARAM_SETS[a.ps]
n = a.lanes
sh = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
g = torch.Generator(device=dev)
g.manual_seed(0xD0DEC7)


def keys_from(xi):
    pk = torch.empty((xi.shape[0], m.PK_LEN), dtype=torch.uint8, device=dev)
    sk = torch.empty((xi.shape[0], m.SK_LEN), dtype=torch.uint8, device=dev)
    _lib.check(L.mldsa_keygen_batch_dev(a.ps, sh, xi.shape[0], p(xi), p(pk), p(sk)))
    return sk


xi_a = torch.full((n, 32), 0xAA, dtype=torch.uint8, device=dev)
xi_b = torch.full((n, 32), 0x55, dtype=torch.uint8, device=dev)
xi_r = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
if a.op == "sign":
    classes = {"A": keys_from(xi_a[:1]).repeat(n, 1).contiguous(), "B": keys_from(xi_b[:1]).repeat(n, 1).contiguous(),
               "R": keys_from(xi_r)}
else:
    classes = {"A": xi_a, "B": xi_b, "R": xi_r}
# CONTROL: the SAME bytes as class A in a different allocation.  Whatever separates A from A' is not the data: buffer
# placement (DRAM pages / L2 sets), or drift of the box -- the floor any A-vs-B figure has to be read against.
classes["A'"] = classes["A"].clone()
kg_pk = torch.empty((n, m.PK_LEN), dtype=torch.uint8, device=dev)
kg_sk = torch.empty((n, m.SK_LEN), dtype=torch.uint8, device=dev)
msg = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
rnd = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
moff = torch.arange(n + 1, dtype=torch.int64, device=dev) * 32
sig = torch.empty((n, m.SIG_LEN), dtype=torch.uint8, device=dev)
KERNELS = ["sign_key", "sign_prep", "mask", "sign_w", "ctilde", "sign_z"] if a.op == "sign" else ["kg_seed", "rej_bounded", "kg_arith", "shake_rows"]


def one(sk):
    L.mldsa_profile_reset()
    if a.op == "sign":
        rc = L.mldsa_sign_batch_dev(a.ps, sh, 0, n, p(sk), m.SK_LEN, p(msg), p(moff), None, None, p(rnd), p(sig), None)
        assert rc in (0, -5), rc  # -5: stopped after one round with lanes still live (the hook's documented outcome)
    else:
        _lib.check(L.mldsa_keygen_batch_dev(a.ps, sh, n, p(sk), p(kg_pk), p(kg_sk)))
    out = {}
    for i in range(L.mldsa_profile_count()):
        ms, cnt = C.c_double(), C.c_uint64()
        L.mldsa_profile_get(i, C.byref(ms), C.byref(cnt))
        name = L.mldsa_profile_name(i).decode()
        if name in KERNELS:
            assert cnt.value == 1, (name, cnt.value)
            out[name] = ms.value * 1e3  # microseconds
    return out


if a.op == "sign":
    L.mldsa_debug_set_max_rounds(1)
L.mldsa_profile_enable(1)
try:
    for c in classes.values():  # warm-up
        for _ in range(5):
            one(c)
    order = [c for c in classes for _ in range(a.reps)]
    random.Random(204).shuffle(order)
    samples = {c: {k: [] for k in KERNELS} for c in classes}
    for c in order:
        r = one(classes[c])
        for k in KERNELS:
            samples[c][k].append(r[k])
finally:
    L.mldsa_profile_enable(0)
    L.mldsa_debug_set_max_rounds(0)


def crop(x):
    """dudect's pre-processing: discard the slowest 5 % of each class (interrupt / scheduling outliers)"""
    x = np.sort(np.asarray(x))
    return x[: max(2, int(len(x) * 0.95))]


def welch(x, y):
    x, y = crop(x), crop(y)
    return (x.mean() - y.mean()) / math.sqrt(x.var(ddof=1) / len(x) + y.var(ddof=1) / len(y))


if a.op == "sign":
    head = ["Device time of ONE launch over all lanes, first round of the rejection loop only; classes A / B = every lane signs",
            "with the key from xi = 0xAA.. / 0x55.., R = distinct random keys; same messages and rnd in all classes."]
    what = {"sign_key": "s1, s2, t0", "sign_prep": "K", "mask": "rho''", "sign_w": "y", "ctilde": "(public: w1)",
            "sign_z": "s-hat rows; early exit on the rejection pattern (carve-out)"}
    carve = ("sign_z", "ctilde")
else:
    head = ["Device time of ONE launch over all lanes; classes A / B = every lane generates the key of xi = 0xAA.. / 0x55..,",
            "R = distinct random seeds."]
    what = {"kg_seed": "xi", "rej_bounded": "rho'; SHAKE256 blocks per stream depend on it (carve-out, as in the reference)",
            "kg_arith": "s1, s2, t0 (A-hat from the public rho)", "shake_rows": "(public: pk)"}
    carve = ("rej_bounded", "shake_rows")
lines = [f"# GPU two-class timing probe: ML-DSA-{a.ps} {a.op}, {n} lanes per launch, {a.reps} launches per class", "",
         "`python tools/ct_probe.py` (method in its docstring; dudect convention: |t| > 5 rejects constant time)."] + head + ["",
         "| kernel | secret input | mean A (us) | mean B (us) | mean R (us) | t (A vs B) | t (A vs R) | control: t (A vs A') |",
         "|---|---|---:|---:|---:|---:|---:|---:|"]
res = {}
for k in KERNELS:
    ma, mb, mr = (float(crop(samples[c][k]).mean()) for c in "ABR")
    tab, tar = welch(samples["A"][k], samples["B"][k]), welch(samples["A"][k], samples["R"][k])
    tctl = welch(samples["A"][k], samples["A'"][k])
    rel = max(abs(mb - ma), abs(mr - ma)) / ma
    res[k] = (tab, tar, tctl, rel)
    lines.append(f"| `{k}` | {what[k]} | {ma:.2f} | {mb:.2f} | {mr:.2f} | {tab:+.2f} | {tar:+.2f} | {tctl:+.2f} |")
fixed = [k for k in KERNELS if k not in carve]
worst = max(abs(t) for k in fixed for t in res[k][:2])
worst_ctl = max(abs(res[k][2]) for k in fixed)
worst_rel = max(res[k][3] for k in fixed)
lines += ["", f"Largest |t| over the fixed-work kernels that touch secret data ({', '.join(fixed)}): **{worst:.2f}**; "
              f"largest |t| of the same kernels between class A and its byte-identical copy A': **{worst_ctl:.2f}**; "
              f"largest relative difference of their class means: **{100 * worst_rel:.3f} %**."]
if a.op == "sign":
    lines += ["`ctilde` works on public data only: its SampleInBall walk rejects candidates j > i drawn from H(c~) exactly like",
              "the reference's loop (src/hashing.rs:43-100), so its time follows the challenge bytes, which are part of the",
              "signature.  `sign_z` stops at the first failed norm check, so its time follows the rejection pattern of the",
              "batch, which differs between keys; that is the reference's documented exception (the loop trip count is not",
              "constant time, src/lib.rs:85-89, traits.rs:128-132), not a data-dependent instruction or address inside the",
              "arithmetic."]
else:
    lines += ["`rej_bounded` squeezes SHAKE256(rho' || r) until it has 256 accepted nibbles: how many blocks that takes depends",
              "on rho', exactly like the reference's loop (src/hashing.rs:158-213); with every lane on the same seed (classes",
              "A, B) all warps take the same number of blocks, with random seeds (R) the mix differs.  `shake_rows` hashes the",
              "public key."]
text = "\n".join(lines) + "\n"
print(text)
if a.out:
    Path(a.out).parent.mkdir(parents=True, exist_ok=True)
    with open(a.out, "a") as f:
        f.write(text + "\n")
if a.json:
    import json
    Path(a.json).write_text(json.dumps({"op": a.op, "ps": a.ps, "lanes": n, "reps": a.reps, "fixed": fixed, "worst_abs_t_fixed": worst, "worst_abs_t_control": worst_ctl, "worst_rel_diff_fixed": worst_rel,
                                        "t": {k: list(v) for k, v in res.items()}}))
