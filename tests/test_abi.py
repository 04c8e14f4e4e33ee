"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/mldsa_b200.h declares, and the host mirror fails loudly (no CPU fallback) without a GPU."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def built():
    from fips204_b200 import build
    return build.build()


def test_header_symbols_exported(built):
    import ctypes
    hdr = (ROOT / "include" / "mldsa_b200.h").read_text()
    names = sorted(set(re.findall(r"\b(mldsa_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    lib = ctypes.CDLL(str(built))
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    from fips204_b200 import _lib
    assert sorted(_lib.SIGNATURES) == names


def test_sizes_and_errors(built):
    import ctypes as C
    from fips204_b200 import _lib
    l = _lib.lib()
    for ps, want in ((44, (1312, 2560, 2420)), (65, (1952, 4032, 3309)), (87, (2592, 4896, 4627))):
        a, b, c = C.c_size_t(), C.c_size_t(), C.c_size_t()
        assert l.mldsa_sizes(ps, C.byref(a), C.byref(b), C.byref(c)) == 0
        assert (a.value, b.value, c.value) == want  # src/lib.rs:649-653, 691-695, 733-737
    assert l.mldsa_sizes(99, None, None, None) == -1
    assert l.mldsa_strerror(-2) == b"ML-DSA.Sign: ctx too long"  # src/lib.rs:274


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import fips204_b200 as F
    with pytest.raises(F.MldsaError):
        F.ml_dsa_44.keygen_batch(bytes(32))
    with pytest.raises(F.MldsaError):
        F.micro.ntt([[0] * 256])


def test_host_mirror_argument_rules(built):
    import fips204_b200 as F
    m = F.ml_dsa_44
    sk = m.PrivateKey.try_from_bytes(bytes(m.SK_LEN))
    with pytest.raises(ValueError, match="ctx too long"):  # src/lib.rs:274
        sk.try_sign_with_seed(bytes(32), b"m", b"\0" * 256)
    with pytest.raises(ValueError):
        m.PrivateKey.try_from_bytes(bytes(10))
    pk = m.PublicKey.try_from_bytes(bytes(m.PK_LEN))
    assert pk.verify(b"m", bytes(m.SIG_LEN), b"\0" * 256) is False  # src/lib.rs:368-370
    assert pk.into_bytes() == bytes(m.PK_LEN)

    class BadRng:
        def try_fill_bytes(self, n):
            raise OSError("entropy source failed")

    with pytest.raises(ValueError, match="KeyGen: Random number generator failed"):  # src/ml_dsa.rs:41
        m.try_keygen_with_rng(BadRng())
    with pytest.raises(ValueError, match="random number generator failed"):  # src/lib.rs:283
        sk.try_sign_with_rng(BadRng(), b"m", b"")


def _build_cpp_example():
    import subprocess
    host = ROOT / "fips204_b200" / "host"
    exe = host / "example_smoke"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", str(host / "example_smoke.cpp"),
                           "-L" + str(ROOT / "fips204_b200"), "-lmldsa_b200", "-Wl,-rpath," + str(ROOT / "fips204_b200"),
                           "-o", str(exe)])
    return exe


def test_cpp_host_mirror_compiles_and_links(built):
    """fips204_b200/host/fips204.hpp (the C++ mirror of the reference surface, incl. PrivateKeySet) compiles without
    warnings against include/mldsa_b200.h and links against the library; without a GPU the program reports the
    library's error instead of computing anything on the CPU."""
    import subprocess
    exe = _build_cpp_example()
    import fips204_b200 as F
    if F.device_count() == 0:
        r = subprocess.run([str(exe)], capture_output=True, text=True)
        assert r.returncode == 2 and "error -3" in r.stdout, (r.returncode, r.stdout)


@pytest.mark.gpu
def test_cpp_host_mirror_smoke_on_gpu(built):
    """the reference's in-crate smoke test (src/lib.rs:503-551) through the C++ mirror, on the device"""
    import subprocess
    r = subprocess.run([str(_build_cpp_example())], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.count(": ok") == 3, (r.returncode, r.stdout, r.stderr)


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) needs no GPU: exactly one JSON line on
    stdout with the contract's keys, the oracle as a `port` baseline, and e2e equal to the line's own value."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--ref-sample", "32"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [x for x in r.stdout.splitlines() if x.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert k in d, k
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["value"] > 0
    assert "workload" in d["config"] and d["unit"] == "signs/s"


def test_host_pipeline_chunk_plan(built):
    """Chunk sizes of the host entry points (mldsa_debug_chunk_plan, no device needed): units of 32 768 lanes ramping
    1-2-4-6...6-4-2-1, every lane covered exactly once, no chunk above the workspace capacity, small batches in one
    chunk, and a fallback (0) when not even two units fit."""
    import ctypes as C
    from fips204_b200 import _lib
    L = _lib.lib()

    def plan(n, cap):
        out = (C.c_size_t * 256)()
        k = L.mldsa_debug_chunk_plan(n, cap, out, 256)
        return list(out[:k])

    U = 32768
    assert plan(1 << 20, 1 << 20) == [U * x for x in (1, 2, 4, 6, 6, 6, 4, 2, 1)]
    assert plan(1 << 17, 1 << 20) == [U, 2 * U, U]
    assert plan(1 << 16, 1 << 20) == [U, U]
    assert plan(50000, 1 << 20) == [] and plan(1 << 20, 40000) == []          # one chunk / plain chunks of cap
    for n, cap in ((3000001, 455000), (1 << 20, 100000), (200000, 70000), (65537, 1 << 20), (10 ** 7, 1 << 19)):
        p = plan(n, cap)
        assert p and sum(p) == n and max(p) <= cap and min(p) >= U, (n, cap, p)
        assert max(p) <= 6 * U + U - 1 and (len(p) < 3 or (p[0] == U and p[-1] == U)), (n, cap, p)
