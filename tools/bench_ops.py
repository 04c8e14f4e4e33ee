#!/usr/bin/env python3
"""Device-resident throughput of every operation on one GPU:
NTT / INTT / expand_a microbenches (config 2), keygen, verify (shared + distinct keys), sign, for
ML-DSA-44/65/87.  CUDA events on the launching stream, inputs >> L2.  Writes one JSON object.
    python tools/bench_ops.py [--lanes 1048576] [--reps 3] [--out profiles/ops.json]"""
import argparse
import ctypes as C
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import fips204_b200 as F  # noqa: E402
from fips204_b200 import _lib  # noqa: E402

KECCAK_F, NTT_OPS, INTT_OPS, MAC_OPS = 4320, 6144, 7168, 5
PAR = {44: (4, 4, 7, 10, 4.25), 65: (6, 5, 7, 15, 5.1), 87: (8, 7, 9, 20, 3.85)}  # K, L, ct blocks, tr blocks, E[iters]

ap = argparse.ArgumentParser()
ap.add_argument("--lanes", type=int, default=1 << 20)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--out", default="")
ap.add_argument("--sets", default="44,65,87")
a = ap.parse_args()
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
L = _lib.lib()
st = torch.cuda.current_stream()
sh = C.c_void_p(st.cuda_stream)
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
peaks = json.load(open(ROOT / "MEASURED_PEAKS.json")) if (ROOT / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0}
alu = C.c_double()
L.mldsa_int32_peak_probe(0, C.byref(alu))
res = {"lanes": a.lanes, "hbm_peak_gbs": peaks["hbm_gbs"], "alu_pipe_peak_tops": alu.value / 1e12,
       "dual_pipe_peak_tops": 2 * alu.value / 1e12, "ops": {}}


def timed(fn, reps=a.reps, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    # every call timed on its own, the MEDIAN reported: at ~850 W the board's power regulation stretches single calls of the
    # ALU-saturating kernels by 5-40 % now and then, and a mean over three calls then says more about that than about the code
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        fn()
        e1.record(st)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def put(name, n, ms, ops=None, byts=None, **kw):
    r = {"n": n, "ms": ms, "per_s": n / (ms * 1e-3)}
    if ops:
        r["int32_tops"] = ops / (ms * 1e-3) / 1e12
    if byts:
        r["hbm_gbs"] = byts / (ms * 1e-3) / 1e9
        r["hbm_frac"] = r["hbm_gbs"] / peaks["hbm_gbs"]
    r.update(kw)
    res["ops"][name] = r
    print(name, json.dumps(r), flush=True)


n = a.lanes
g = torch.Generator(device=dev)
g.manual_seed(1)
# ---- config 2 microbenches
polys = torch.randint(0, 8380417, (n, 256), dtype=torch.int32, device=dev, generator=g)
ms = timed(lambda: _lib.check(L.mldsa_ntt_batch_dev(sh, n, p(polys))))
put("ntt", n, ms, n * NTT_OPS, n * 2048)
polys.remainder_(8380417)
ms = timed(lambda: _lib.check(L.mldsa_intt_batch_dev(sh, n, p(polys))))
put("intt", n, ms, n * INTT_OPS, n * 2048)
seeds = torch.randint(0, 256, (n, 34), dtype=torch.uint8, device=dev, generator=g)
ms = timed(lambda: _lib.check(L.mldsa_expand_a_batch_dev(sh, n, p(seeds), p(polys))))
put("expand_a", n, ms, n * 5 * KECCAK_F, n * (34 + 1024), alu_pipe_frac=n * 5 * KECCAK_F / (ms * 1e-3) / alu.value)
del polys, seeds

for ps in [int(x) for x in a.sets.split(",")]:
    m = F.PARAM_SETS[ps]
    K, Lp, ctb, trb, eit = PAR[ps]
    xi = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    msg = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    rnd = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    moff = torch.arange(n + 1, dtype=torch.int64, device=dev) * 32
    pk = torch.empty((n, m.PK_LEN), dtype=torch.uint8, device=dev)
    sk = torch.empty((n, m.SK_LEN), dtype=torch.uint8, device=dev)
    sig = torch.empty((n, m.SIG_LEN), dtype=torch.uint8, device=dev)
    it = torch.zeros(n, dtype=torch.int32, device=dev)
    ok = torch.zeros(n, dtype=torch.uint8, device=dev)
    ms = timed(lambda: _lib.check(L.mldsa_keygen_batch_dev(ps, sh, n, p(xi), p(pk), p(sk))))
    eta2 = 1.53 if ps != 65 else 2.0
    kg_ops = n * ((1 + (K + Lp) * eta2 + K * Lp * 5 + trb) * KECCAK_F + Lp * NTT_OPS + K * INTT_OPS + K * Lp * 256 * MAC_OPS)
    put(f"keygen{ps}", n, ms, kg_ops, n * (32 + m.PK_LEN + m.SK_LEN))
    ms = timed(lambda: _lib.check(L.mldsa_sign_batch_dev(ps, sh, 0, n, p(sk), m.SK_LEN, p(msg), p(moff), None, None,
                                                         p(rnd), p(sig), p(it))))
    iters = float(it.to(torch.float64).mean().item())
    per_iter = (Lp * 5 + ctb + 1) * KECCAK_F + (Lp + 1) * NTT_OPS + (2 * K + Lp) * INTT_OPS + (K * Lp + Lp + K) * 256 * MAC_OPS
    once = (K * Lp * 5 + 2) * KECCAK_F + (Lp + 2 * K) * NTT_OPS + K * (INTT_OPS + 256 * MAC_OPS)
    put(f"sign{ps}", n, ms, n * (once + iters * per_iter), n * (m.SK_LEN + 64 + m.SIG_LEN), mean_iters=iters)
    for shared in (False, True):
        if shared:
            _lib.check(L.mldsa_sign_batch_dev(ps, sh, 0, n, p(sk), 0, p(msg), p(moff), None, None, p(rnd), p(sig), None))
        stride = 0 if shared else m.PK_LEN
        ms = timed(lambda: _lib.check(L.mldsa_verify_batch_dev(ps, sh, 0, n, p(pk), stride, p(msg), p(moff), None, None,
                                                               p(sig), p(ok))))
        assert bool(ok.all().item())
        kec = 2 + ctb + (0 if shared else K * Lp * 5 + trb)
        vops = n * (kec * KECCAK_F + (Lp + 1 + (0 if shared else K)) * NTT_OPS + K * INTT_OPS + (K * Lp + K) * 256 * MAC_OPS)
        put(f"verify{ps}_{'shared' if shared else 'distinct'}", n, ms, vops,
            n * ((0 if shared else m.PK_LEN) + 32 + m.SIG_LEN + 1))
    del xi, msg, rnd, moff, pk, sk, sig, it, ok
    torch.cuda.empty_cache()

if a.out:
    Path(a.out).write_text(json.dumps(res, indent=1))
print("JSON", json.dumps(res))
