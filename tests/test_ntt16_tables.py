"""CPU checks of the 16-coefficients-per-lane transforms (fips204_b200/csrc/ntt16.cuh): the index maps and twiddle
indices the kernels use, modelled in Python against a direct transform mod q (src/ntt.rs:14-161), the per-lane packed
tables of zetas.inc against those indices, and zetas.inc against its generator."""
import re
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "fips204_b200" / "csrc"
Q, ZETA = 8380417, 1753


def brv8(x):
    return int(f"{x:08b}"[::-1], 2)


Z = [pow(ZETA, brv8(m), Q) for m in range(256)]  # ZETA[m] of the reference's table without the Montgomery factor


def ntt_ref(a):
    a, m, ln = list(a), 0, 128
    while ln >= 1:
        for st in range(0, 256, 2 * ln):
            m += 1
            for j in range(st, st + ln):
                t = Z[m] * a[j + ln] % Q
                a[j + ln], a[j] = (a[j] - t) % Q, (a[j] + t) % Q
        ln //= 2
    return a


def intt_ref_unscaled(a):
    a, m, ln = list(a), 256, 1
    while ln < 256:
        for st in range(0, 256, 2 * ln):
            m -= 1
            for j in range(st, st + ln):
                t = a[j]
                a[j], a[j + ln] = (t + a[j + ln]) % Q, (t - a[j + ln]) * (-Z[m]) % Q
        ln *= 2
    return a


# twiddle index of level s (len = 2^s) for register i of lane t, as ntt16.cuh uses them
FWD_X = {7: lambda i, t: 1, 6: lambda i, t: 2 + (i >> 3), 5: lambda i, t: 4 + (i >> 2), 4: lambda i, t: 8 + (i >> 1)}
FWD_Y = {3: lambda i, t: 16 + t, 2: lambda i, t: 32 + 2 * t + (i >> 3), 1: lambda i, t: 64 + 4 * t + (i >> 2),
         0: lambda i, t: 128 + 8 * t + (i >> 1)}
INV_Y = {0: lambda i, t: 255 - 8 * t - (i >> 1), 1: lambda i, t: 127 - 4 * t - (i >> 2), 2: lambda i, t: 63 - 2 * t - (i >> 3),
         3: lambda i, t: 31 - t}
INV_X = {4: lambda i, t: 15 - (i >> 1), 5: lambda i, t: 7 - (i >> 2), 6: lambda i, t: 3 - (i >> 3), 7: lambda i, t: 1}


def fwd16(poly):
    X = [[poly[16 * i + t] for i in range(16)] for t in range(16)]  # layout X: lane t, register i = coefficient 16 i + t
    for t in range(16):
        v = X[t]
        for s in (7, 6, 5, 4):
            stride = 1 << (s - 4)
            for i in range(16):
                if not i & stride:
                    w = Z[FWD_X[s](i, t)] * v[i + stride] % Q
                    v[i + stride], v[i] = (v[i] - w) % Q, (v[i] + w) % Q
    Y = [[X[i][t] for i in range(16)] for t in range(16)]           # layout Y: lane t, register i = coefficient 16 t + i
    for t in range(16):
        v = Y[t]
        for s in (3, 2, 1, 0):
            stride = 1 << s
            for i in range(16):
                if not i & stride:
                    w = Z[FWD_Y[s](i, t)] * v[i + stride] % Q
                    v[i + stride], v[i] = (v[i] - w) % Q, (v[i] + w) % Q
    return [Y[t][i] for t in range(16) for i in range(16)]


def inv16(poly):
    Y = [[poly[16 * t + i] for i in range(16)] for t in range(16)]
    for t in range(16):
        v = Y[t]
        for s in (0, 1, 2, 3):
            stride = 1 << s
            for i in range(16):
                if not i & stride:
                    a = v[i]
                    v[i], v[i + stride] = (a + v[i + stride]) % Q, (a - v[i + stride]) * (-Z[INV_Y[s](i, t)]) % Q
    X = [[Y[i][t] for i in range(16)] for t in range(16)]
    for t in range(16):
        v = X[t]
        for s in (4, 5, 6, 7):
            stride = 1 << (s - 4)
            for i in range(16):
                if not i & stride:
                    a = v[i]
                    v[i], v[i + stride] = (a + v[i + stride]) % Q, (a - v[i + stride]) * (-Z[INV_X[s](i, t)]) % Q
    out = [0] * 256
    for t in range(16):
        for i in range(16):
            out[16 * i + t] = X[t][i]
    return out


def test_index_maps_and_twiddle_indices_against_a_direct_transform():
    rng = np.random.default_rng(204)
    for _ in range(3):
        p = [int(x) for x in rng.integers(0, Q, 256)]
        assert fwd16(p) == ntt_ref(p)
        assert inv16(p) == intt_ref_unscaled(p)


def _table(text, name):
    m = re.search(r"static __device__ const Tw16 %s = \{\{(.*?)\}\};" % name, text, re.S)
    nums = [int(x) for x in re.findall(r"-?\d+", m.group(1))]
    assert len(nums) == 8 * 16 * 4
    return np.array(nums, dtype=np.int64).reshape(8, 16, 4)  # [vector][lane][z, z', z, z']


def _centred(v):
    v %= Q
    return v - Q if v > Q // 2 else v


def test_packed_tables_hold_the_lanes_twiddles_in_order_of_use():
    text = (CSRC / "zetas.inc").read_text()
    for name, fwd in (("ZS16_FWD", True), ("ZS16_INV", False), ("ZM16_FWD", True)):
        tab = _table(text, name)
        for t in range(16):
            if fwd:
                idx = [16 + t, 32 + 2 * t, 33 + 2 * t] + [64 + 4 * t + g for g in range(4)] + [128 + 8 * t + k for k in range(8)]
            else:
                idx = [255 - 8 * t - k for k in range(8)] + [127 - 4 * t - g for g in range(4)] + [63 - 2 * t - g for g in range(2)] + [31 - t]
            pairs = tab[:, t, :].reshape(16, 2)  # 16 (z, z') pairs of lane t, the last one padding
            for k, m in enumerate(idx):
                z = int(pairs[k][0])
                if name == "ZM16_FWD":    # Montgomery form: zeta * 2^32 mod q and its product with q^-1 mod 2^32
                    assert z == Z[m] * (1 << 32) % Q
                    assert (int(pairs[k][1]) - z * 58728449) % (1 << 32) == 0
                else:                      # plain twiddle, centred, with z' = round(z 2^32 / q)
                    want = _centred(Z[m] if fwd else -Z[m])
                    assert z == want
                    assert abs(int(pairs[k][1]) - want * (1 << 32) / Q) <= 1
            assert tuple(pairs[15]) == (0, 0)


def test_zetas_inc_is_what_the_generator_prints():
    out = subprocess.run([sys.executable, str(CSRC / "gen_zetas.py")], capture_output=True, text=True, check=True).stdout
    assert out == (CSRC / "zetas.inc").read_text()
