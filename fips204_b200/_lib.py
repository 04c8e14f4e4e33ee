"""ctypes binding of the C-ABI library (include/mldsa_b200.h).  There is NO CPU fallback: if the CUDA
library is missing or a call fails, an exception is raised."""
import ctypes as C
import os
from pathlib import Path

HERE = Path(__file__).resolve().parent
# MLDSA_B200_LIB: an A/B build of the same library (fips204_b200/build.py build_variant) for tools/ probes
SO_PATH = Path(os.environ["MLDSA_B200_LIB"]) if os.environ.get("MLDSA_B200_LIB") else HERE / "libmldsa_b200.so"

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
i32p = C.POINTER(C.c_int32)
sz = C.c_size_t
vp = C.c_void_p

# every symbol include/mldsa_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "mldsa_device_count": (C.c_int, []),
    "mldsa_strerror": (C.c_char_p, [C.c_int]),
    "mldsa_last_cuda_error": (C.c_char_p, []),
    "mldsa_sizes": (C.c_int, [C.c_int, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz)]),
    "mldsa_host_alloc": (vp, [sz]),
    "mldsa_host_free": (None, [vp]),
    "mldsa_keygen_batch": (C.c_int, [C.c_int, C.c_int, sz, vp, vp, vp]),
    "mldsa_keygen_batch_dev": (C.c_int, [C.c_int, vp, sz, vp, vp, vp]),
    "mldsa_sign_batch": (C.c_int, [C.c_int, C.c_int, C.c_int, sz, vp, sz, vp, vp, vp, vp, vp, vp, vp]),
    "mldsa_sign_batch_dev": (C.c_int, [C.c_int, vp, C.c_int, sz, vp, sz, vp, vp, vp, vp, vp, vp, vp]),
    "mldsa_verify_batch": (C.c_int, [C.c_int, C.c_int, C.c_int, sz, vp, sz, vp, vp, vp, vp, vp, vp]),
    "mldsa_verify_batch_dev": (C.c_int, [C.c_int, vp, C.c_int, sz, vp, sz, vp, vp, vp, vp, vp, vp]),
    "mldsa_hash_sign_batch": (C.c_int, [C.c_int, C.c_int, C.c_int, sz, vp, sz, vp, vp, vp, vp, vp, vp, vp]),
    "mldsa_hash_verify_batch": (C.c_int, [C.c_int, C.c_int, C.c_int, sz, vp, sz, vp, vp, vp, vp, vp, vp]),
    "mldsa_sk_to_pk_batch": (C.c_int, [C.c_int, C.c_int, sz, vp, vp]),
    "mldsa_sk_to_pk_batch_dev": (C.c_int, [C.c_int, vp, sz, vp, vp]),
    "mldsa_keygen_batch_multi": (C.c_int, [C.c_int, vp, C.c_int, sz, vp, vp, vp]),
    "mldsa_sign_batch_multi": (C.c_int, [C.c_int, vp, C.c_int, C.c_int, sz, vp, sz, vp, vp, vp, vp, vp, vp, vp]),
    "mldsa_verify_batch_multi": (C.c_int, [C.c_int, vp, C.c_int, C.c_int, sz, vp, sz, vp, vp, vp, vp, vp, vp]),
    "mldsa_ntt_batch": (C.c_int, [C.c_int, sz, vp]),
    "mldsa_intt_batch": (C.c_int, [C.c_int, sz, vp]),
    "mldsa_expand_a_batch": (C.c_int, [C.c_int, sz, vp, vp]),
    "mldsa_ntt_batch_dev": (C.c_int, [vp, sz, vp]),
    "mldsa_intt_batch_dev": (C.c_int, [vp, sz, vp]),
    "mldsa_expand_a_batch_dev": (C.c_int, [vp, sz, vp, vp]),
    "mldsa_shake_batch": (C.c_int, [C.c_int, C.c_int, sz, vp, sz, vp, sz]),
    "mldsa_set_workspace_limit": (None, [sz]),
    "mldsa_set_tuning": (None, [C.c_int, C.c_int, C.c_int]),
    "mldsa_set_depth": (None, [C.c_int]),
    "mldsa_set_tail": (None, [C.c_int, C.c_int]),
    "mldsa_debug_set_max_rounds": (None, [C.c_int]),
    "mldsa_debug_chunk_plan": (C.c_size_t, [C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t]),
    "mldsa_sk_load": (C.c_int, [C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.POINTER(C.c_void_p)]),
    "mldsa_pk_load": (C.c_int, [C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.POINTER(C.c_void_p)]),
    "mldsa_verify_batch_keys": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p]),
    "mldsa_hash_sign_batch_keys": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mldsa_hash_verify_batch_keys": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mldsa_keygen_to_keyset": (C.c_int, [C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p),
                                          C.POINTER(C.c_void_p)]),
    "mldsa_pipeline_batch": (C.c_int, [C.c_int, C.c_int, C.c_int, sz] + [vp] * 12),
    "mldsa_pipeline_batch_multi": (C.c_int, [C.c_int, vp, C.c_int, C.c_int, sz] + [vp] * 12),
    "mldsa_keyset_size": (C.c_size_t, [C.c_void_p]),
    "mldsa_keyset_device_bytes": (C.c_size_t, [C.c_void_p]),
    "mldsa_keyset_free": (None, [C.c_void_p]),
    "mldsa_sign_batch_keys": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mldsa_debug_secret_residue": (C.c_int, [C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "mldsa_debug_sampler_stream": (C.c_int, [C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mldsa_kernel_launches": (C.c_uint64, []),
    "mldsa_last_sign_rounds": (C.c_int, []),
    "mldsa_profile_enable": (None, [C.c_int]),
    "mldsa_profile_reset": (None, []),
    "mldsa_profile_count": (C.c_int, []),
    "mldsa_profile_name": (C.c_char_p, [C.c_int]),
    "mldsa_profile_get": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "mldsa_int32_peak_probe": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "mldsa_keccak_probe": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint32)]),
}

_lib = None


class MldsaError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        super().__init__(f"mldsa error {code}: {detail}")


def lib():
    """Load the CUDA library.  Raises if it has not been built (python -m fips204_b200.build)."""
    global _lib
    if _lib is None:
        if not SO_PATH.exists():
            raise ImportError(f"{SO_PATH} is missing: build it with `python -m fips204_b200.build` "
                              "(there is no CPU fallback)")
        l = C.CDLL(str(SO_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        l = lib()
        msg = l.mldsa_strerror(rc).decode()
        if rc == -3:
            msg += ": " + l.mldsa_last_cuda_error().decode()
        raise MldsaError(rc, msg)
