"""fips204_b200 -- B200-native batched ML-DSA (FIPS 204), a drop-in for the hot path of the Rust crate
integritychain/fips204.

Host-side mirror of the reference's public surface (src/lib.rs:116-614, src/traits.rs):

    from fips204_b200 import ml_dsa_44, ml_dsa_65, ml_dsa_87
    pk, sk = ml_dsa_65.try_keygen()                      # lib.rs:195      KeyGen::try_keygen
    pk, sk = ml_dsa_65.try_keygen_with_rng(rng)          # lib.rs:230
    pk, sk = ml_dsa_65.KG.keygen_from_seed(xi)           # lib.rs:247
    sig = sk.try_sign(msg, ctx)                          # traits.rs:156   Signer::try_sign
    sig = sk.try_sign_with_rng(rng, msg, ctx)            # lib.rs:268
    sig = sk.try_sign_with_seed(seed, msg, ctx)          # traits.rs:228
    ok  = pk.verify(msg, sig, ctx)                       # lib.rs:364      Verifier::verify
    sk.into_bytes(); ml_dsa_65.PrivateKey.try_from_bytes(b)   # lib.rs:417-494  SerDes

plus the batch entry points the engine exists for (keygen_batch / sign_batch / verify_batch), all of
which call hand-written sm_100a CUDA through the C ABI of include/mldsa_b200.h.  Single operations
are batches of one.  There is no CPU fallback: without the CUDA library every call raises.

Error behaviour follows the reference: signing with |ctx| > 255 raises ValueError("ML-DSA.Sign: ctx
too long") (lib.rs:274), an RNG failure raises ValueError("... random number generator failed")
(lib.rs:283, ml_dsa.rs:41), verify never raises for malformed input and returns False (lib.rs:368-370).
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import MldsaError

__all__ = ["ml_dsa_44", "ml_dsa_65", "ml_dsa_87", "Ph", "MldsaError", "micro", "device_count", "kernel_launches", "bind_host_to_device"]


class Ph:
    """Pre-hash function selector of HashML-DSA (`pub enum Ph`, src/types.rs:5-12)."""
    SHA256, SHA512, SHAKE128 = 0, 1, 2


def device_count():
    return _lib.lib().mldsa_device_count()


def bind_host_to_device(dev=0):
    """Pin the calling process to the CPU cores that are local to GPU `dev` (NVML's CPU affinity of the device), so
    that the pinned host buffers it allocates afterwards live on that NUMA node and its copies do not cross the
    socket interconnect.  One process per GPU (torchrun) on a two-socket box otherwise lands most ranks' buffers on
    the wrong socket.  Returns the list of cores, or [] when NVML / the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        phys = int(vis.split(",")[dev]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else dev
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        allowed = os.sched_getaffinity(0)
        cpus = [i for i in range(ncpu) if (mask[i // 64] >> (i % 64)) & 1 and i in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return []


def kernel_launches():
    return int(_lib.lib().mldsa_kernel_launches())


class OsRng:
    """Default RNG (the reference's `default-rng` feature binds OsRng, traits.rs:44-47)."""

    def try_fill_bytes(self, n):
        return os.urandom(n)


def _draw(rng, n, what):
    # exactly ONE 32-byte draw per keygen (ml_dsa.rs:40-41) / per sign (lib.rs:282-283), in lane order
    try:
        b = rng.try_fill_bytes(n) if hasattr(rng, "try_fill_bytes") else rng(n)
    except Exception as e:  # noqa: BLE001 - mirrors map_err(|_| "...") of the reference
        raise ValueError(what) from e
    if b is None or len(b) != n:
        raise ValueError(what)
    return bytes(b)


def _u8(a, shape=None):
    arr = np.frombuffer(a, dtype=np.uint8) if isinstance(a, (bytes, bytearray, memoryview)) else np.asarray(a, dtype=np.uint8)
    arr = np.ascontiguousarray(arr)
    if shape is not None:
        arr = arr.reshape(shape)
    return arr


def _ptr(a):
    return a.ctypes.data if a is not None else None


def _ragged(items, off_dtype):
    off = np.zeros(len(items) + 1, dtype=off_dtype)
    if len(items):
        off[1:] = np.cumsum([len(x) for x in items])
    blob = np.frombuffer(b"".join(bytes(x) for x in items) or b"\0", dtype=np.uint8)
    return blob, off


class _ParamSet:
    """One of the reference's per-parameter-set modules (src/lib.rs:639-653, 681-695, 723-737)."""

    def __init__(self, ps, k, l, pk_len, sk_len, sig_len):
        self.PS, self.K, self.L = ps, k, l
        self.PK_LEN, self.SK_LEN, self.SIG_LEN = pk_len, sk_len, sig_len
        mod = self

        class PublicKey:
            __slots__ = ("_b",)

            def __init__(self, b):
                self._b = bytes(b)

            @staticmethod
            def try_from_bytes(b):  # lib.rs:469-474; never fails on a correctly sized array (SURVEY 8a note T)
                if len(b) != mod.PK_LEN:
                    raise ValueError("incorrect public key length")
                return PublicKey(b)

            def into_bytes(self):  # lib.rs:477-493
                return self._b

            def verify(self, message, sig, ctx=b""):  # lib.rs:364-380
                if len(ctx) > 255 or len(sig) != mod.SIG_LEN:
                    return False
                return bool(mod.verify_batch(self._b, [message], [sig], ctxs=[ctx], shared_key=True)[0])

            def hash_verify(self, message, sig, ctx, ph):  # lib.rs:391-411
                if len(ctx) > 255 or len(sig) != mod.SIG_LEN:
                    return False
                return bool(mod.hash_verify_batch(self._b, [message], [sig], ph, ctxs=[ctx], shared_key=True)[0])

            def __eq__(self, o):
                return isinstance(o, PublicKey) and o._b == self._b

        class PrivateKey:
            __slots__ = ("_b",)

            def __init__(self, b):
                self._b = bytes(b)

            @staticmethod
            def try_from_bytes(b):  # lib.rs:420-424
                if len(b) != mod.SK_LEN:
                    raise ValueError("incorrect private key length")
                return PrivateKey(b)

            def into_bytes(self):  # lib.rs:427-463
                return self._b

            def try_sign(self, message, ctx=b""):  # traits.rs:156
                return self.try_sign_with_rng(OsRng(), message, ctx)

            def try_sign_with_rng(self, rng, message, ctx=b""):  # lib.rs:268-296
                if len(ctx) > 255:
                    raise ValueError("ML-DSA.Sign: ctx too long")
                rnd = _draw(rng, 32, "ML-DSA.Sign: random number generator failed")
                return self.try_sign_with_seed(rnd, message, ctx)

            def try_sign_with_seed(self, seed, message, ctx=b""):  # traits.rs:228-232 (DummyRng)
                if len(ctx) > 255:
                    raise ValueError("ML-DSA.Sign: ctx too long")
                return mod.sign_batch(self._b, [message], seed, ctxs=[ctx], shared_key=True)[0].tobytes()

            def try_hash_sign(self, message, ctx, ph):  # traits.rs:247
                return self.try_hash_sign_with_rng(OsRng(), message, ctx, ph)

            def try_hash_sign_with_rng(self, rng, message, ctx, ph):  # lib.rs:310-342
                if len(ctx) > 255:
                    raise ValueError("HashML-DSA.Sign: ctx too long")
                rnd = _draw(rng, 32, "HashML-DSA.Sign: random number generator failed")
                return self.try_hash_sign_with_seed(rnd, message, ctx, ph)

            def try_hash_sign_with_seed(self, seed, message, ctx, ph):  # traits.rs:280-284
                if len(ctx) > 255:
                    raise ValueError("HashML-DSA.Sign: ctx too long")
                return mod.hash_sign_batch(self._b, [message], seed, ph, ctxs=[ctx], shared_key=True)[0].tobytes()

            def get_public_key(self):  # lib.rs:344-346 -> ml_dsa.rs:502-563
                return PublicKey(mod.sk_to_pk_batch(self._b)[0].tobytes())

        class KG:
            @staticmethod
            def try_keygen():  # traits.rs:44
                return mod.try_keygen()

            @staticmethod
            def try_keygen_with_rng(rng):  # lib.rs:241-244
                return mod.try_keygen_with_rng(rng)

            @staticmethod
            def keygen_from_seed(xi):  # lib.rs:247-250
                if len(xi) != 32:
                    raise ValueError("xi must be 32 bytes")
                pk, sk = mod.keygen_batch(xi)
                return PublicKey(pk[0].tobytes()), PrivateKey(sk[0].tobytes())

        self.PublicKey, self.PrivateKey, self.KG = PublicKey, PrivateKey, KG

    # ----- single-op module functions (lib.rs:195, :230) -----
    def try_keygen(self):
        return self.try_keygen_with_rng(OsRng())

    def try_keygen_with_rng(self, rng):
        xi = _draw(rng, 32, "KeyGen: Random number generator failed")  # ml_dsa.rs:41
        return self.KG.keygen_from_seed(xi)

    def _internal_sign(self, sk, message, ctx, rnd):  # lib.rs:586-594 (ACVP interface)
        if len(ctx) > 255:
            raise ValueError("_internal_sign: ctx too long")
        return self.sign_batch(sk.into_bytes(), [message], rnd, shared_key=True, framing=1)[0].tobytes()

    def _internal_verify(self, pk, message, sig, ctx=b""):  # lib.rs:605-612
        if len(ctx) > 255:
            return False
        return bool(self.verify_batch(pk.into_bytes(), [message], [sig], shared_key=True, framing=1)[0])

    # ----- batch entry points (the engine) -----
    def keygen_batch(self, xis, dev=0):
        """xis: n x 32 seed bytes -> (pk[n, PK_LEN], sk[n, SK_LEN]) uint8 arrays."""
        xi = _u8(xis).reshape(-1, 32)
        n = xi.shape[0]
        pk = np.empty((n, self.PK_LEN), dtype=np.uint8)
        sk = np.empty((n, self.SK_LEN), dtype=np.uint8)
        _lib.check(_lib.lib().mldsa_keygen_batch(self.PS, dev, n, _ptr(xi), _ptr(pk), _ptr(sk)))
        return pk, sk

    def keygen_batch_with_rng(self, rng, n, dev=0):
        xi = b"".join(_draw(rng, 32, "KeyGen: Random number generator failed") for _ in range(n))
        return self.keygen_batch(xi, dev)

    def sign_batch(self, sks, msgs, rnds, ctxs=None, shared_key=False, framing=0, dev=0, want_iters=False, ph=None):
        """sks: n x SK_LEN (or one key with shared_key), msgs: list of bytes, rnds: n x 32.
        ph = Ph.* selects HashML-DSA (pre-hash on the device, lib.rs:310-342)."""
        n = len(msgs)
        sk = _u8(sks)
        if sk.size != (1 if shared_key else n) * self.SK_LEN:
            raise ValueError("incorrect private key length")
        rnd = _u8(rnds)
        if rnd.size != 32 * n:
            raise ValueError("rnd must be n x 32 bytes")
        mb, mo = _ragged(msgs, np.uint64)
        cb = co = None
        if ctxs is not None:
            if len(ctxs) != n:
                raise ValueError("ctxs must have n entries")
            if any(len(c) > 255 for c in ctxs):
                raise ValueError(("ML-DSA" if ph is None else "HashML-DSA") + ".Sign: ctx too long")
            cb, co = _ragged(ctxs, np.uint32)
        sig = np.empty((n, self.SIG_LEN), dtype=np.uint8)
        iters = np.zeros(n, dtype=np.uint32)
        fn = _lib.lib().mldsa_sign_batch if ph is None else _lib.lib().mldsa_hash_sign_batch
        rc = fn(self.PS, dev, framing if ph is None else int(ph), n, _ptr(sk), 0 if shared_key else self.SK_LEN,
                _ptr(mb), _ptr(mo), _ptr(cb), _ptr(co), _ptr(rnd), _ptr(sig), _ptr(iters))
        if rc == -2:
            raise ValueError(("ML-DSA" if ph is None else "HashML-DSA") + ".Sign: ctx too long")
        _lib.check(rc)
        return (sig, iters) if want_iters else sig

    def sign_batch_with_rng(self, rng, sks, msgs, ctxs=None, **kw):
        rnd = b"".join(_draw(rng, 32, "ML-DSA.Sign: random number generator failed") for _ in range(len(msgs)))
        return self.sign_batch(sks, msgs, rnd, ctxs, **kw)

    def hash_sign_batch(self, sks, msgs, rnds, ph, **kw):
        return self.sign_batch(sks, msgs, rnds, ph=ph, **kw)

    def hash_verify_batch(self, pks, msgs, sigs, ph, **kw):
        return self.verify_batch(pks, msgs, sigs, ph=ph, **kw)

    def load_private_keys(self, sks, dev=0):
        """n x SK_LEN packed keys -> PrivateKeySet: the keys expanded once (PrivateKey::try_from_bytes ->
        expand_private, lib.rs:462-468, ml_dsa.rs:445-469) and resident in HBM for any number of sign calls."""
        sk = _u8(sks).reshape(-1, self.SK_LEN)
        h = C.c_void_p()
        _lib.check(_lib.lib().mldsa_sk_load(self.PS, dev, sk.shape[0], _ptr(sk), C.byref(h)))
        return PrivateKeySet(self, h)

    def load_public_keys(self, pks, dev=0):
        """n x PK_LEN packed keys -> PublicKeySet: expand_public (ml_dsa.rs:477-498) done once, resident in HBM."""
        pk = _u8(pks).reshape(-1, self.PK_LEN)
        h = C.c_void_p()
        _lib.check(_lib.lib().mldsa_pk_load(self.PS, dev, pk.shape[0], _ptr(pk), C.byref(h)))
        return PublicKeySet(self, h)

    def keygen_to_keysets(self, xis, dev=0, want_pk=True, want_private=True, want_public=True):
        """n x 32 seeds -> (pk bytes or None, PrivateKeySet or None, PublicKeySet or None): KG::keygen_from_seed whose
        expanded keys stay on the device (mldsa_keygen_to_keyset); no packed secret key crosses PCIe."""
        xi = _u8(xis).reshape(-1, 32)
        n = xi.shape[0]
        pk = np.empty((n, self.PK_LEN), dtype=np.uint8) if want_pk else None
        hs, hp = C.c_void_p(), C.c_void_p()
        _lib.check(_lib.lib().mldsa_keygen_to_keyset(self.PS, dev, n, _ptr(xi), _ptr(pk), C.byref(hs) if want_private else None,
                                                     C.byref(hp) if want_public else None))
        return pk, (PrivateKeySet(self, hs) if want_private else None), (PublicKeySet(self, hp) if want_public else None)

    def pipeline_batch(self, xis, msgs, rnds, ctxs=None, framing=0, devs=(0,), want_sk=False, want_iters=False,
                       want_stage_ms=False):
        """keygen -> sign -> verify chained on the device(s) (the flow of lib.rs:503-551 as one batch call): lane i
        derives its key pair from xis[i], signs msgs[i] and verifies the result.  -> dict(pk, sig, ok[, sk, iters,
        stage_ms]).  Same bytes as keygen_batch / sign_batch / verify_batch called one after the other."""
        xi = _u8(xis).reshape(-1, 32)
        n = xi.shape[0]
        if len(msgs) != n:
            raise ValueError("msgs must have n entries")
        rnd = _u8(rnds)
        if rnd.size != 32 * n:
            raise ValueError("rnd must be n x 32 bytes")
        mb, mo = _ragged(msgs, np.uint64)
        cb = co = None
        if ctxs is not None:
            if len(ctxs) != n:
                raise ValueError("ctxs must have n entries")
            if any(len(c) > 255 for c in ctxs):
                raise ValueError("ML-DSA.Sign: ctx too long")
            cb, co = _ragged(ctxs, np.uint32)
        out = {"pk": np.empty((n, self.PK_LEN), dtype=np.uint8), "sig": np.empty((n, self.SIG_LEN), dtype=np.uint8),
               "ok": np.zeros(n, dtype=np.uint8)}
        sk = np.empty((n, self.SK_LEN), dtype=np.uint8) if want_sk else None
        iters = np.zeros(n, dtype=np.uint32) if want_iters else None
        ms = (C.c_double * 3)() if want_stage_ms else None
        dv = np.ascontiguousarray(list(devs), dtype=np.int32)
        rc = _lib.lib().mldsa_pipeline_batch_multi(self.PS, _ptr(dv), len(dv), framing, n, _ptr(xi), _ptr(mb), _ptr(mo), _ptr(cb),
                                                   _ptr(co), _ptr(rnd), _ptr(out["pk"]), _ptr(sk), _ptr(out["sig"]),
                                                   _ptr(out["ok"]), _ptr(iters), C.cast(ms, C.c_void_p) if ms else None)
        if rc == -2:
            raise ValueError("ML-DSA.Sign: ctx too long")
        _lib.check(rc)
        if want_sk:
            out["sk"] = sk
        if want_iters:
            out["iters"] = iters
        if want_stage_ms:
            out["stage_ms"] = [float(x) for x in ms]
        return out

    def sk_to_pk_batch(self, sks, dev=0):
        """Signer::get_public_key over a batch: n x SK_LEN -> pk[n, PK_LEN] (ml_dsa.rs:502-563)."""
        sk = _u8(sks).reshape(-1, self.SK_LEN)
        pk = np.empty((sk.shape[0], self.PK_LEN), dtype=np.uint8)
        _lib.check(_lib.lib().mldsa_sk_to_pk_batch(self.PS, dev, sk.shape[0], _ptr(sk), _ptr(pk)))
        return pk

    def verify_batch(self, pks, msgs, sigs, ctxs=None, shared_key=False, framing=0, dev=0, ph=None):
        """-> uint8[n], 1 = valid.  Never raises for malformed signatures (lib.rs:368-370)."""
        n = len(msgs)
        pk = _u8(pks)
        if pk.size != (1 if shared_key else n) * self.PK_LEN:
            raise ValueError("incorrect public key length")
        if isinstance(sigs, (list, tuple)):
            sg = _u8(b"".join(bytes(s) for s in sigs))
        else:
            sg = _u8(sigs)
        if sg.size != n * self.SIG_LEN:
            raise ValueError("incorrect signature length")
        mb, mo = _ragged(msgs, np.uint64)
        cb = co = None
        if ctxs is not None:
            cb, co = _ragged(ctxs, np.uint32)
        ok = np.zeros(n, dtype=np.uint8)
        fn = _lib.lib().mldsa_verify_batch if ph is None else _lib.lib().mldsa_hash_verify_batch
        _lib.check(fn(self.PS, dev, framing if ph is None else int(ph), n, _ptr(pk), 0 if shared_key else self.PK_LEN,
                      _ptr(mb), _ptr(mo), _ptr(cb), _ptr(co), _ptr(sg), _ptr(ok)))
        return ok


class PrivateKeySet:
    """Expanded secret keys resident on the device (mldsa_sk_load); zeroized by free() / on garbage collection
    (ZeroizeOnDrop, types.rs:19)."""

    def __init__(self, mod, handle):
        self._mod, self._h = mod, handle

    def __len__(self):
        return int(_lib.lib().mldsa_keyset_size(self._h)) if self._h else 0

    @property
    def device_bytes(self):
        return int(_lib.lib().mldsa_keyset_device_bytes(self._h)) if self._h else 0

    def sign_batch(self, msgs, rnds, ctxs=None, key_idx=None, framing=0, want_iters=False, ph=None):
        """Lane i signs msgs[i] with key key_idx[i] (default: key i).  Same bytes as _ParamSet.sign_batch; ph = Ph.*
        selects HashML-DSA."""
        if not self._h:
            raise ValueError("key set was freed")
        m, n = self._mod, len(msgs)
        rnd = _u8(rnds)
        if rnd.size != 32 * n:
            raise ValueError("rnd must be n x 32 bytes")
        mb, mo = _ragged(msgs, np.uint64)
        cb = co = None
        if ctxs is not None:
            if any(len(c) > 255 for c in ctxs):
                raise ValueError("ML-DSA.Sign: ctx too long")
            cb, co = _ragged(ctxs, np.uint32)
        ki = None if key_idx is None else np.ascontiguousarray(key_idx, dtype=np.uint32)
        if ki is not None and ki.size != n:
            raise ValueError("key_idx must have n entries")
        sig = np.empty((n, m.SIG_LEN), dtype=np.uint8)
        iters = np.zeros(n, dtype=np.uint32)
        fn = _lib.lib().mldsa_sign_batch_keys if ph is None else _lib.lib().mldsa_hash_sign_batch_keys
        rc = fn(self._h, framing if ph is None else int(ph), n, _ptr(ki), _ptr(mb), _ptr(mo), _ptr(cb), _ptr(co), _ptr(rnd),
                _ptr(sig), _ptr(iters))
        if rc == -2:
            raise ValueError("ML-DSA.Sign: ctx too long")
        _lib.check(rc)
        return (sig, iters) if want_iters else sig

    def free(self):
        if self._h:
            _lib.lib().mldsa_keyset_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PublicKeySet(PrivateKeySet):
    """Expanded public keys resident on the device (mldsa_pk_load): A-hat, tr, NTT(t1 * 2^d) per key."""

    def sign_batch(self, *a, **kw):
        raise ValueError("a public-key set cannot sign")

    def verify_batch(self, msgs, sigs, ctxs=None, key_idx=None, framing=0, ph=None):
        """-> uint8[n]; lane i is checked under key key_idx[i] (default: key i).  Same verdicts as
        _ParamSet.verify_batch; ph = Ph.* selects HashML-DSA."""
        if not self._h:
            raise ValueError("key set was freed")
        m, n = self._mod, len(msgs)
        sg = _u8(b"".join(bytes(s) for s in sigs)) if isinstance(sigs, (list, tuple)) else _u8(sigs)
        if sg.size != n * m.SIG_LEN:
            raise ValueError("incorrect signature length")
        mb, mo = _ragged(msgs, np.uint64)
        cb = co = None
        if ctxs is not None:
            cb, co = _ragged(ctxs, np.uint32)
        ki = None if key_idx is None else np.ascontiguousarray(key_idx, dtype=np.uint32)
        if ki is not None and ki.size != n:
            raise ValueError("key_idx must have n entries")
        ok = np.zeros(n, dtype=np.uint8)
        fn = _lib.lib().mldsa_verify_batch_keys if ph is None else _lib.lib().mldsa_hash_verify_batch_keys
        _lib.check(fn(self._h, framing if ph is None else int(ph), n, _ptr(ki), _ptr(mb), _ptr(mo), _ptr(cb), _ptr(co),
                      _ptr(sg), _ptr(ok)))
        return ok


ml_dsa_44 = _ParamSet(44, 4, 4, 1312, 2560, 2420)
ml_dsa_65 = _ParamSet(65, 6, 5, 1952, 4032, 3309)
ml_dsa_87 = _ParamSet(87, 8, 7, 2592, 4896, 4627)
PARAM_SETS = {44: ml_dsa_44, 65: ml_dsa_65, 87: ml_dsa_87}


class micro:
    """Leaf entry points of BASELINE config 2 (src/ntt.rs, src/hashing.rs:111-146) and the SHAKE hook."""

    @staticmethod
    def ntt(polys, dev=0):
        a = np.ascontiguousarray(polys, dtype=np.int32).copy()
        _lib.check(_lib.lib().mldsa_ntt_batch(dev, a.size // 256, _ptr(a)))
        return a

    @staticmethod
    def inv_ntt(polys, dev=0):
        a = np.ascontiguousarray(polys, dtype=np.int32).copy()
        _lib.check(_lib.lib().mldsa_intt_batch(dev, a.size // 256, _ptr(a)))
        return a

    @staticmethod
    def expand_a(seeds34, dev=0):
        s = _u8(seeds34).reshape(-1, 34)
        out = np.empty((s.shape[0], 256), dtype=np.int32)
        _lib.check(_lib.lib().mldsa_expand_a_batch(dev, s.shape[0], _ptr(s), _ptr(out)))
        return out

    @staticmethod
    def shake(bits, rows, out_len, dev=0):
        r = np.ascontiguousarray(rows, dtype=np.uint8)
        n, in_len = r.shape
        out = np.empty((n, out_len), dtype=np.uint8)
        _lib.check(_lib.lib().mldsa_shake_batch(dev, bits, n, _ptr(r), in_len, _ptr(out), out_len))
        return out
