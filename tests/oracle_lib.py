"""ctypes binding of the CPU oracle (oracle/liboracle.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
_SO = ROOT / "oracle" / "liboracle.so"


class Params(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "ps", "k", "l", "eta", "tau", "lambda_div4", "gamma1", "gamma2", "omega", "beta",
        "pk_len", "sk_len", "sig_len", "w1_len", "z_bits", "w1_bits", "eta_bits")]


def build(force=False):
    src = ROOT / "oracle" / "mldsa_oracle.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(ROOT / "oracle"), "-B", "liboracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def use_native_build():
    """Switch this module to oracle/liboracle_native.so, compiled here and now with -march=native (bench.py's CPU
    baseline: the oracle should run as fast as this host allows).  Returns the flags description; falls back to the
    portable build if the compile fails."""
    global _lib
    native = ROOT / "oracle" / "liboracle_native.so"
    try:
        subprocess.check_call(["make", "-C", str(ROOT / "oracle"), "-B", "liboracle_native.so"], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
        new = C.CDLL(str(native))
    except (subprocess.CalledProcessError, OSError):
        lib()
        return "gcc -O3 -march=x86-64-v3 (portable build; native rebuild failed)"
    new.oracle_mont_reduce.restype = C.c_int32
    new.oracle_mont_reduce.argtypes = [C.c_int64]
    new.oracle_zeta_mont.restype = C.c_int32
    new.oracle_use_hint.restype = C.c_int32
    _lib = new
    return "gcc -O3 -march=native, built on this host"


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_SO))
        _lib.oracle_mont_reduce.restype = C.c_int32
        _lib.oracle_mont_reduce.argtypes = [C.c_int64]
        _lib.oracle_zeta_mont.restype = C.c_int32
        _lib.oracle_use_hint.restype = C.c_int32
    return _lib


def _buf(b):
    if b is None:
        return None
    return (C.c_uint8 * len(b)).from_buffer_copy(bytes(b)) if len(b) else (C.c_uint8 * 1)()


def params(ps):
    p = Params()
    assert lib().oracle_get_params(ps, C.byref(p)) == 0
    return p


def shake128(data, n):
    out = (C.c_uint8 * n)()
    lib().oracle_shake128(_buf(data), C.c_size_t(len(data)), out, C.c_size_t(n))
    return bytes(out)


def shake256(data, n):
    out = (C.c_uint8 * n)()
    lib().oracle_shake256(_buf(data), C.c_size_t(len(data)), out, C.c_size_t(n))
    return bytes(out)


def sha256(data):
    out = (C.c_uint8 * 32)()
    lib().oracle_sha256(_buf(data), C.c_size_t(len(data)), out)
    return bytes(out)


def sha512(data):
    out = (C.c_uint8 * 64)()
    lib().oracle_sha512(_buf(data), C.c_size_t(len(data)), out)
    return bytes(out)


def keygen(ps, xi):
    p = params(ps)
    pk = (C.c_uint8 * p.pk_len)()
    sk = (C.c_uint8 * p.sk_len)()
    assert lib().oracle_keygen(ps, _buf(xi), pk, sk) == 0
    return bytes(pk), bytes(sk)


def sign(ps, sk, msg, rnd, ctx=b"", framing=0, ph=0, want_iters=False):
    p = params(ps)
    assert len(sk) == p.sk_len and len(rnd) == 32
    sig = (C.c_uint8 * p.sig_len)()
    it = C.c_uint32(0)
    rc = lib().oracle_sign(ps, framing, ph, _buf(sk), _buf(msg), C.c_size_t(len(msg)), _buf(ctx),
                           C.c_size_t(len(ctx)), _buf(rnd), sig, C.byref(it))
    if rc != 0:
        raise ValueError(f"oracle_sign rc={rc}")
    return (bytes(sig), it.value) if want_iters else bytes(sig)


def verify(ps, pk, msg, sig, ctx=b"", framing=0, ph=0):
    p = params(ps)
    assert len(pk) == p.pk_len and len(sig) == p.sig_len
    rc = lib().oracle_verify(ps, framing, ph, _buf(pk), _buf(msg), C.c_size_t(len(msg)), _buf(ctx),
                             C.c_size_t(len(ctx)), _buf(sig))
    assert rc in (0, 1), rc
    return bool(rc)


def sk_to_pk(ps, sk):
    p = params(ps)
    pk = (C.c_uint8 * p.pk_len)()
    assert lib().oracle_sk_to_pk(ps, _buf(sk), pk) == 0
    return bytes(pk)


def _np_u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint8))


def ntt_batch(polys):
    a = np.ascontiguousarray(polys, dtype=np.int32).copy()
    lib().oracle_ntt_batch(a.ctypes.data_as(C.POINTER(C.c_int32)), C.c_size_t(a.size // 256))
    return a


def inv_ntt_batch(polys):
    a = np.ascontiguousarray(polys, dtype=np.int32).copy()
    lib().oracle_inv_ntt_batch(a.ctypes.data_as(C.POINTER(C.c_int32)), C.c_size_t(a.size // 256))
    return a


def expand_a_batch(seeds34):
    s, sp = _np_u8(seeds34)
    n = s.size // 34
    out = np.empty((n, 256), dtype=np.int32)
    lib().oracle_expand_a_batch(sp, C.c_size_t(n), out.ctypes.data_as(C.POINTER(C.c_int32)))
    return out


def ncores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def _offsets(items, dtype):
    off = np.zeros(len(items) + 1, dtype=dtype)
    if len(items):
        off[1:] = np.cumsum([len(x) for x in items])
    return off


def keygen_batch(ps, xis, nthreads=None):
    p = params(ps)
    xi, xip = _np_u8(xis)
    n = xi.size // 32
    pk = np.empty((n, p.pk_len), dtype=np.uint8)
    sk = np.empty((n, p.sk_len), dtype=np.uint8)
    rc = lib().oracle_keygen_batch(ps, nthreads or ncores(), C.c_size_t(n), xip,
                                   pk.ctypes.data_as(C.POINTER(C.c_uint8)), sk.ctypes.data_as(C.POINTER(C.c_uint8)))
    assert rc == 0
    return pk, sk


def sign_batch(ps, sks, msgs, rnds, ctxs=None, framing=0, nthreads=None, shared_key=False):
    """sks: (n, SK_LEN) array (or (SK_LEN,) with shared_key); msgs: list of bytes; rnds: (n, 32)."""
    p = params(ps)
    n = len(msgs)
    sk, skp = _np_u8(sks)
    rnd, rndp = _np_u8(rnds)
    mo = _offsets(msgs, np.uint64)
    mb, mbp = _np_u8(np.frombuffer(b"".join(msgs) or b"\0", dtype=np.uint8))
    if ctxs is not None:
        co = _offsets(ctxs, np.uint32)
        cb, cbp = _np_u8(np.frombuffer(b"".join(ctxs) or b"\0", dtype=np.uint8))
        cop = co.ctypes.data_as(C.POINTER(C.c_uint32))
    else:
        cbp, cop = None, None
    sig = np.empty((n, p.sig_len), dtype=np.uint8)
    iters = np.zeros(n, dtype=np.uint32)
    rc = lib().oracle_sign_batch(ps, framing, nthreads or ncores(), C.c_size_t(n), skp,
                                 C.c_size_t(0 if shared_key else p.sk_len), mbp,
                                 mo.ctypes.data_as(C.POINTER(C.c_uint64)), cbp, cop, rndp,
                                 sig.ctypes.data_as(C.POINTER(C.c_uint8)),
                                 iters.ctypes.data_as(C.POINTER(C.c_uint32)))
    if rc != 0:
        raise ValueError(f"oracle_sign_batch rc={rc}")
    return sig, iters


def verify_batch(ps, pks, msgs, sigs, ctxs=None, framing=0, nthreads=None, shared_key=False):
    p = params(ps)
    n = len(msgs)
    pk, pkp = _np_u8(pks)
    sg, sgp = _np_u8(sigs)
    mo = _offsets(msgs, np.uint64)
    mb, mbp = _np_u8(np.frombuffer(b"".join(msgs) or b"\0", dtype=np.uint8))
    if ctxs is not None:
        co = _offsets(ctxs, np.uint32)
        cb, cbp = _np_u8(np.frombuffer(b"".join(ctxs) or b"\0", dtype=np.uint8))
        cop = co.ctypes.data_as(C.POINTER(C.c_uint32))
    else:
        cbp, cop = None, None
    ok = np.zeros(n, dtype=np.uint8)
    rc = lib().oracle_verify_batch(ps, framing, nthreads or ncores(), C.c_size_t(n), pkp,
                                   C.c_size_t(0 if shared_key else p.pk_len), mbp,
                                   mo.ctypes.data_as(C.POINTER(C.c_uint64)), cbp, cop, sgp,
                                   ok.ctypes.data_as(C.POINTER(C.c_uint8)))
    assert rc == 0, rc
    return ok
