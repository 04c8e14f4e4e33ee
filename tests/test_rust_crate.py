"""The Rust host crate (rust/fips204-b200) is shipped source-only -- there is no Rust toolchain in the build image -- so
what can be checked mechanically is checked here: its `extern "C"` block against include/mldsa_b200.h (every function:
name, arity, return type, and per parameter the base width, constness and pointer depth), and the presence of the
reference's public surface (src/lib.rs:116-614, src/traits.rs:8-424) by name.  Both files are parsed independently of
tools/gen_rust_ffi.py, which generated src/ffi.rs."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CRATE = ROOT / "rust" / "fips204-b200"

# (width class, signedness) of the scalar types on both sides; LP64 (the only ABI the library is built for)
C_SCALAR = {"int": "i32", "size_t": "u64", "uint64_t": "u64", "uint32_t": "u32", "int32_t": "i32", "uint8_t": "u8",
            "char": "i8", "double": "f64", "void": "void", "mldsa_keyset": "keyset"}
R_SCALAR = {"c_int": "i32", "usize": "u64", "u64": "u64", "u32": "u32", "i32": "i32", "u8": "u8", "c_char": "i8",
            "f64": "f64", "c_void": "void", "()": "void", "mldsa_keyset": "keyset"}


def c_shape(t):
    """'const uint8_t*' -> ('u8', ['const'])  (pointer levels outermost-last, each 'const' or 'mut')"""
    t = t.strip()
    const = t.startswith("const ")
    if const:
        t = t[6:]
    depth = t.count("*")
    base = C_SCALAR[t.replace("*", "").strip()]
    # in this header `const` always qualifies the pointee of the innermost pointer
    levels = [("const" if (const and i == 0) else "mut") for i in range(depth)]
    return base, levels


def r_shape(t):
    t = t.strip()
    levels = []
    while t.startswith("*"):
        m = re.match(r"\*(const|mut)\s+(.*)", t)
        levels.append(m.group(1))
        t = m.group(2).strip()
    return R_SCALAR[t], list(reversed(levels))  # innermost first, like c_shape


def header_functions():
    src = re.sub(r"/\*.*?\*/", " ", (ROOT / "include" / "mldsa_b200.h").read_text(), flags=re.S)
    src = src[src.index('extern "C" {'):]
    out = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(mldsa_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        params = []
        if args and args != "void":
            for a in args.split(","):
                params.append(c_shape(re.match(r"(.*?)\w+$", a.strip()).group(1)))
        out[name] = (c_shape(ret), params)
    return out


def rust_functions():
    src = (CRATE / "src" / "ffi.rs").read_text()
    block = src[src.index('extern "C" {'):]
    out = {}
    for m in re.finditer(r"pub fn (mldsa_\w+)\(([^)]*)\)\s*(?:->\s*([^;]+))?;", block):
        name, args, ret = m.group(1), m.group(2).strip(), (m.group(3) or "()").strip()
        params = [r_shape(a.split(":", 1)[1]) for a in args.split(",") if a.strip()]
        out[name] = (r_shape(ret), params)
    return out


def test_extern_block_matches_header_symbol_by_symbol():
    h, r = header_functions(), rust_functions()
    assert len(h) >= 50, "header parse lost functions"
    assert set(h) == set(r), (sorted(set(h) - set(r)), sorted(set(r) - set(h)))
    for name in h:
        (hret, hpar), (rret, rpar) = h[name], r[name]
        assert hret == rret, (name, "return", hret, rret)
        assert len(hpar) == len(rpar), (name, "arity", len(hpar), len(rpar))
        for i, (a, b) in enumerate(zip(hpar, rpar)):
            assert a == b, (name, f"parameter {i}", a, b)


def test_library_exports_what_the_crate_binds():
    import ctypes
    so = ROOT / "fips204_b200" / "libmldsa_b200.so"
    if not so.exists():
        from fips204_b200 import build
        build.build()
    lib = ctypes.CDLL(str(so))
    for name in rust_functions():
        assert hasattr(lib, name), name


def test_reference_surface_is_present_by_name():
    lib_rs = (CRATE / "src" / "lib.rs").read_text()
    traits_rs = (CRATE / "src" / "traits.rs").read_text()
    batch_rs = (CRATE / "src" / "batch.rs").read_text()
    # modules and constants (reference src/lib.rs:639-737)
    for mod, sk, pk, sig in (("ml_dsa_44", 2560, 1312, 2420), ("ml_dsa_65", 4032, 1952, 3309), ("ml_dsa_87", 4896, 2592, 4627)):
        body = lib_rs[lib_rs.index(f"pub mod {mod}"):]
        body = body[:body.index("functionality!")]
        assert f"SK_LEN: usize = {sk}" in body and f"PK_LEN: usize = {pk}" in body and f"SIG_LEN: usize = {sig}" in body
    # items of the functionality! macro (reference src/lib.rs:116-614)
    for item in ("pub struct KG()", "pub struct PrivateKey", "pub struct PublicKey", "pub fn try_keygen()",
                 "pub fn try_keygen_with_rng(", "impl KeyGen for KG", "impl Signer for PrivateKey", "impl Verifier for PublicKey",
                 "impl SerDes for PrivateKey", "impl SerDes for PublicKey", "pub fn _internal_sign(", "pub fn _internal_verify(",
                 "ZeroizeOnDrop"):
        assert item in lib_rs, item
    # traits and their methods (reference src/traits.rs:8-424)
    for trait, methods in (("KeyGen", ("try_keygen", "try_keygen_with_rng", "keygen_from_seed")),
                           ("Signer", ("try_sign", "try_sign_with_rng", "try_sign_with_seed", "try_hash_sign",
                                       "try_hash_sign_with_rng", "try_hash_sign_with_seed", "get_public_key")),
                           ("Verifier", ("verify", "hash_verify")),
                           ("SerDes", ("into_bytes", "try_from_bytes"))):
        body = traits_rs[traits_rs.index(f"pub trait {trait}"):]
        for m in methods:
            assert re.search(rf"fn {m}\b", body), (trait, m)
    # the reference's error strings (src/lib.rs:274, :283, :316; src/ml_dsa.rs:41)
    for text in ("ML-DSA.Sign: ctx too long", "ML-DSA.Sign: random number generator failed", "HashML-DSA.Sign: ctx too long",
                 "KeyGen: Random number generator failed"):
        assert text in lib_rs, text
    # batch layer
    for item in ("pub trait BatchKeyGen", "pub trait BatchSigner", "pub trait BatchVerifier", "pub struct PrivateKeySet",
                 "pub struct PublicKeySet", "impl Drop for PrivateKeySet", "impl Drop for PublicKeySet", "pub fn keygen_to_keysets",
                 "pub fn pipeline"):
        assert item in batch_rs, item
    cargo = (CRATE / "Cargo.toml").read_text()
    assert 'rand_core = { version = "0.6.4"' in cargo and "zeroize" in cargo
    assert (CRATE / "build.rs").exists() and "not compiled" in (CRATE / "README.md").read_text().lower()
