#!/usr/bin/env python3
"""Regenerate the golden fixtures under tests/golden/ from the reference checkout.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py [/root/reference]

Sources (all are the reference's OWN test pins for the hot path):
  * tests/nist_vectors/ML-DSA-{keyGen,sigGen,sigVer}-FIPS204/internalProjection.json
    (NIST ACVP-Server @65370b8 public vectors; harness tests/nist_vectors/mod.rs:55-203)
  * tests/messages.rs:9-21           ChaCha8(123) end-to-end KAT, external framing
  * tests/integration.rs:60-74       bad_sig malleability regression, internal framing

Only the fields the tests consume are kept (hex strings), grouped per parameter set, and the
result is gzip-compressed so the fixtures stay small.  No reference *source* is copied.
"""
import gzip
import json
import re
import sys
from pathlib import Path

ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
out = Path(__file__).resolve().parent
nv = ref / "tests" / "nist_vectors"


def dump(name, obj):
    raw = json.dumps(obj, separators=(",", ":"), sort_keys=True).encode()
    with gzip.GzipFile(out / name, "wb", mtime=0) as f:
        f.write(raw)
    print(name, len(raw), "->", (out / name).stat().st_size)


kg = json.load(open(nv / "ML-DSA-keyGen-FIPS204" / "internalProjection.json"))
dump("acvp_keygen.json.gz", [
    {"ps": g["parameterSet"], "tc": t["tcId"], "seed": t["seed"], "pk": t["pk"], "sk": t["sk"]}
    for g in kg["testGroups"] for t in g["tests"]])

sg = json.load(open(nv / "ML-DSA-sigGen-FIPS204" / "internalProjection.json"))
dump("acvp_siggen.json.gz", [
    {"ps": g["parameterSet"], "tc": t["tcId"], "det": g["deterministic"], "sk": t["sk"],
     "message": t["message"], "rnd": t.get("rnd"), "signature": t["signature"]}
    for g in sg["testGroups"] for t in g["tests"]])

sv = json.load(open(nv / "ML-DSA-sigVer-FIPS204" / "internalProjection.json"))
dump("acvp_sigver.json.gz", [
    {"ps": g["parameterSet"], "tc": t["tcId"], "pk": g["pk"], "message": t["message"],
     "signature": t["signature"], "passed": t["testPassed"], "reason": t["reason"]}
    for g in sv["testGroups"] for t in g["tests"]])

# tests/messages.rs: three hex literals, in order sk, sig, pk.  xi / rnd are the first two 32-byte
# draws of ChaCha8Rng::seed_from_u64(123).
hexes = re.findall(r'hex::decode\("([0-9a-f]+)"\)', (ref / "tests" / "messages.rs").read_text())
assert [len(h) // 2 for h in hexes] == [2560, 2420, 1312]
msgs = {"ps": "ML-DSA-44", "msg": b"asdf".hex(), "ctx": "",
        "xi": "932b30e756257dda01d47dd4a1b0e62abb8fa01f222ad8554ed821a89f82cbbb",
        "rnd": "3f537f509949e758624a70946a77698673597bd30ea2d1bc2b1207afa75145c2",
        "sk": hexes[0], "sig": hexes[1], "pk": hexes[2]}

# tests/integration.rs bad_sig: msg, sk, pk, good_sig, bad_sig (internal framing)
txt = (ref / "tests" / "integration.rs").read_text()
body = txt[txt.index("fn bad_sig()"):txt.index("fn test_44_no_verif()")]
hx = re.findall(r'hex::decode\("([0-9a-f]+)"\)', body)
assert [len(h) // 2 for h in hx] == [32, 2560, 1312, 2420, 2420]
bad = {"ps": "ML-DSA-44", "msg": hx[0], "sk": hx[1], "pk": hx[2], "good_sig": hx[3], "bad_sig": hx[4]}
dump("reference_kats.json.gz", {"messages_rs": msgs, "bad_sig": bad})
