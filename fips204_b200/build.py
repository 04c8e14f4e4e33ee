"""Builds fips204_b200/libmldsa_b200.so (the C-ABI library, include/mldsa_b200.h) for sm_100a.

nvcc cross-compiles without a GPU.  The .so is built IN-TREE so that it travels with the repo
snapshot to the GPU box; it is git-ignored.  `python -m fips204_b200.build [--force]`.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "libmldsa_b200.so"
OBJ = HERE / "build"
UNITS = ["k_common", "k_keygen", "k_verify", "k_sign", "k_fused", "k_probe", "engine"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC,-O3", "-Xptxas", "-v"]


def _stamp():
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*")) + [HERE.parent / "include" / "mldsa_b200.h", Path(__file__)]):
        if f.is_file():
            h.update(f.name.encode())
            h.update(f.read_bytes())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    stamp = _stamp()
    stamp_file = OBJ / "stamp"
    if not force and OUT.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return OUT
    OBJ.mkdir(exist_ok=True)

    def cc(unit):
        cmd = [NVCC, *FLAGS, "-c", str(CSRC / f"{unit}.cu"), "-o", str(OBJ / f"{unit}.o")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (OBJ / f"{unit}.log").write_text(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {unit}:\n{r.stdout}\n{r.stderr}")
        return unit

    with ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        for u in ex.map(cc, UNITS):
            if verbose:
                print("compiled", u)
    cmd = [NVCC, "-shared", "-o", str(OUT), *[str(OBJ / f"{u}.o") for u in UNITS],
           "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp_file.write_text(stamp)
    return OUT


def build_variant(name, defs, units=None):
    """A/B builds for tools/ (never loaded by the package itself unless MLDSA_B200_LIB names one):
    fips204_b200/variants/libmldsa_b200_<name>.so compiled with extra -D flags.  Only `units` are recompiled with
    the flags; the other objects come from the regular build."""
    build()
    vdir = HERE / "variants"
    odir = OBJ / f"v_{name}"
    vdir.mkdir(exist_ok=True)
    odir.mkdir(exist_ok=True)
    units = units or UNITS
    objs = []
    for u in UNITS:
        if u in units:
            o = odir / f"{u}.o"
            r = subprocess.run([NVCC, *FLAGS, *defs, "-c", str(CSRC / f"{u}.cu"), "-o", str(o)], capture_output=True, text=True)
            (odir / f"{u}.log").write_text(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {u} ({name}):\n{r.stderr}")
            objs.append(str(o))
        else:
            objs.append(str(OBJ / f"{u}.o"))
    out = vdir / f"libmldsa_b200_{name}.so"
    r = subprocess.run([NVCC, "-shared", "-o", str(out), *objs, "-gencode", "arch=compute_100a,code=sm_100a",
                        "-Xcompiler", "-fPIC", "-cudart", "static"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:  # python -m fips204_b200.build --variant NAME unit[,unit] -DX=1 ...
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], [a for a in sys.argv[i + 3:]], sys.argv[i + 2].split(",")))
    else:
        print(build(force="--force" in sys.argv, verbose=True))
