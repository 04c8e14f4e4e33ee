"""Pins the CPU oracle against every golden vector the reference's own tests hold for the hot path
(SURVEY.md section 8c): 75 + 60 + 45 ACVP vectors (tests/nist_vectors/mod.rs:55-203), the
tests/messages.rs KAT, tests/integration.rs::bad_sig and the src/lib.rs:536-545 pk[0] pins."""
import oracle_lib as O
from conftest import PS_OF


def test_acvp_keygen(acvp_keygen):
    assert len(acvp_keygen) == 75
    for t in acvp_keygen:
        pk, sk = O.keygen(PS_OF[t["ps"]], bytes.fromhex(t["seed"]))
        assert pk.hex().upper() == t["pk"].upper(), t["tc"]
        assert sk.hex().upper() == t["sk"].upper(), t["tc"]


def test_acvp_siggen(acvp_siggen):
    assert len(acvp_siggen) == 60
    for t in acvp_siggen:
        rnd = bytes.fromhex(t["rnd"]) if t["rnd"] else bytes(32)  # mod.rs:107-112
        sig = O.sign(PS_OF[t["ps"]], bytes.fromhex(t["sk"]), bytes.fromhex(t["message"]), rnd, framing=1)
        assert sig.hex().upper() == t["signature"].upper(), t["tc"]


def test_acvp_sigver(acvp_sigver):
    assert len(acvp_sigver) == 45
    for t in acvp_sigver:
        ok = O.verify(PS_OF[t["ps"]], bytes.fromhex(t["pk"]), bytes.fromhex(t["message"]),
                      bytes.fromhex(t["signature"]), framing=1)
        assert ok == t["passed"], (t["tc"], t["reason"])


def test_messages_rs_kat(reference_kats):
    m = reference_kats["messages_rs"]
    pk, sk = O.keygen(44, bytes.fromhex(m["xi"]))
    assert pk.hex() == m["pk"] and sk.hex() == m["sk"]
    sig = O.sign(44, sk, bytes.fromhex(m["msg"]), bytes.fromhex(m["rnd"]), ctx=b"", framing=0)
    assert sig.hex() == m["sig"]
    assert O.verify(44, pk, bytes.fromhex(m["msg"]), sig, ctx=b"", framing=0)
    assert O.sk_to_pk(44, sk) == pk


def test_bad_sig(reference_kats):
    b = reference_kats["bad_sig"]
    pk, msg = bytes.fromhex(b["pk"]), bytes.fromhex(b["msg"])
    assert O.verify(44, pk, msg, bytes.fromhex(b["good_sig"]), framing=1)
    assert not O.verify(44, pk, msg, bytes.fromhex(b["bad_sig"]), framing=1)


def test_pk0_pins():
    # src/lib.rs:536-545: keygen_from_seed([0x11; 32]) -> pk[0] = 197 / 177 / 16
    for ps, want in ((44, 197), (65, 177), (87, 16)):
        pk, sk = O.keygen(ps, b"\x11" * 32)
        assert pk[0] == want
        assert O.sk_to_pk(ps, sk) == pk
