"""The C oracle (oracle/*.c, the checker of every GPU parity test) against an INDEPENDENT second restatement of the reference:
oracle/pyref.py, pure-Python big integers written from the Rust source and the public specifications, sharing no code with the C
oracle or the engine (VERDICT r1 "what's missing" 1 / "next round" 9).

  * tests/golden/pyref_vectors.json (made by tests/golden/make_pyref_vectors.py, ~2 minutes of pure Python, committed) holds the
    document oracle/rust_vectors would print for its six shapes -- commitments, seed nonces, proof bytes, recovered masks, the verdict
    of a flipped-r1 proof -- as computed by pyref; the C oracle, replaying the same ChaCha12 stream in the reference's drawing order,
    must reproduce every byte (same checker as tests/test_rust_vectors.py uses for the real crate's document);
  * one small shape is recomputed live by pyref, so the committed file cannot drift from the module;
  * pyref's own primitives are pinned to external known answers (RFC 9496, Merlin's KAT, FIPS 202 via hashlib, rand_chacha's stream).
This does not replace the run against the Rust crate (still impossible here: no cargo), it removes the single-author reading.
"""
import hashlib
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
sys.path.insert(0, os.path.join(HERE, "golden"))

import pyref as R  # noqa: E402
import orc  # noqa: E402
from test_rust_vectors import check_document  # noqa: E402

PATH = os.path.join(HERE, "golden", "pyref_vectors.json")


def test_pyref_primitives_against_external_known_answers():
    # RFC 9496 A.1: multiples of the generator
    assert R.encode(R.IDENTITY) == bytes(32)
    assert R.encode(R.pt_mul(2, R.BASEPOINT)).hex() == "6a493210f7499cd17fecb510ae0cea23a110e8d5b901f8acadd3095c73a3b919"
    assert R.encode(R.pt_mul(15, R.BASEPOINT)).hex() == "e0c418f7c8d9c4cdd7395b93ea124f3ad99021bb681dfc3302a9d99a2e53e64e"
    # RFC 9496 A.3: one-way map
    u = hashlib.sha512(b"Ristretto is traditionally a short shot of espresso coffee").digest()
    assert R.encode(R.from_uniform_bytes(u)).hex() == "3066f82a1a747d45120d1740f14358531a8f04bbffe6a819f86dfe50f44a0a46"
    # RFC 9496 A.2: invalid encodings (non-canonical field element, negative, non-square, negative xy, y = 0)
    for bad in ("00ffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff", "0100000000000000000000000000000000000000000000000000000000000000",
                "26948d35ca62e643e26a83177332e6b6afeb9d08e4268b650f1f5bbd8d81d371", "3eb858e78f5a7254d8c9731174a94f76755fd3941c0ac93735c07ba14579630e",
                "edffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff7f"):
        assert R.decode(bytes.fromhex(bad)) is None
    # Keccak-f[1600] through a SHA3-256 sponge vs hashlib
    def sha3_256(msg):
        st, rate = bytearray(200), 136
        m = bytearray(msg) + b"\x06"
        m += bytes(-len(m) % rate)
        m[-1] |= 0x80
        for i in range(0, len(m), rate):
            for j in range(rate):
                st[j] ^= m[i + j]
            R.keccak_f1600(st)
        return bytes(st[:32])
    for msg in (b"", b"abc", b"q" * 137, b"z" * 1000):
        assert sha3_256(msg) == hashlib.sha3_256(msg).digest()
    # merlin's own test vector
    t = R.Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"
    # ChaCha: the 20-round core of RFC 8439 2.3.2 is not reachable (12 rounds here); pin the 12-round stream to the C oracle's
    # independent implementation and the word / u64 consumption rules to each other
    a, b = R.ChaCha12Rng.seed_from_u64(8675309), orc.Rng("chacha", 8675309)
    for _ in range(70):
        assert a.next_u64() == b.next_u64()
    assert a.fill_bytes(61) == b.fill(61)
    assert a.next_u64() == b.next_u64()


def test_pyref_wire_state_matches_c_oracle_transcripts():
    t = R.Transcript(b"BatchedRangeProofTest")
    assert t.s.to_wire() == orc.transcript_new(b"BatchedRangeProofTest")


def test_committed_pyref_vectors_are_reproduced_by_the_c_oracle():
    doc = json.load(open(PATH))
    assert doc["crate"].startswith("pyref")
    assert len(doc["cases"]) == 6
    assert check_document(doc) == 1 + 4 + 3 + 2 + 1 + 3


def test_live_pyref_case_matches_committed_file_and_c_oracle():
    import make_pyref_vectors as M

    case = json.loads(json.dumps(M.run_case(8, [1, 2, 4], 2, "third")))
    committed = json.load(open(PATH))["cases"][2]
    assert case == committed
    assert check_document({"cases": [case]}) == 3


def test_pyref_rejects_what_the_reference_rejects():
    """error variants of pyref.verify_batch agree with the C oracle on the reference's error-path checks (a small shape)"""
    import make_pyref_vectors as M
    import workload

    rng = R.ChaCha12Rng.seed_from_u64(11)
    params = R.Params(8, 1, 1)
    v = rng.next_u64() % 128
    bl = [R.random_not_zero(rng)]
    st = R.Statement(params, [params.commit(v, bl)], [v // 3], R.random_not_zero(rng))
    pr = R.prove_with_rng(R.Transcript(M.LABEL), st, [v], [bl], rng)
    assert R.verify_batch([R.Transcript(M.LABEL)], [st], [pr], R.RECOVER_AND_VERIFY) == [bl]
    # tampered promise -> VerificationFailed (tests/ristretto.rs:320-356)
    st_bad = R.Statement(params, st.commitments, [v // 3 + 1], st.seed_nonce)
    with pytest.raises(R.ProofError) as e:
        R.verify_batch([R.Transcript(M.LABEL)], [st_bad], [pr], R.VERIFY_ONLY)
    assert e.value.variant == "VerificationFailed"
    # wrong seed nonce: verifies, other mask (tests/ristretto.rs:291-318)
    st_seed = R.Statement(params, st.commitments, st.mins, (st.seed_nonce + 1) % R.L)
    assert R.verify_batch([R.Transcript(M.LABEL)], [st_seed], [pr], R.RECOVER_AND_VERIFY) != [bl]
    # identity / non-canonical encodings in the proof
    for field, variant in (("a", "VerificationFailed"), ("b", "VerificationFailed")):
        bad = R.Proof.from_bytes(pr.to_bytes())
        setattr(bad, field, bytes(32))
        with pytest.raises(R.ProofError) as e:
            R.verify_batch([R.Transcript(M.LABEL)], [st], [bad], R.VERIFY_ONLY)
        assert e.value.variant == variant
    bad = R.Proof.from_bytes(pr.to_bytes())
    bad.li[0] = bytes([1]) + bytes(31)
    with pytest.raises(R.ProofError) as e:
        R.verify_batch([R.Transcript(M.LABEL)], [st], [bad], R.VERIFY_ONLY)
    assert e.value.variant == "InvalidArgument"
    # the same proof bytes through the C oracle
    rc, parsed = orc.proof_from_bytes(pr.to_bytes())
    assert rc == 0
    op = orc.Params(8, 1, 1)
    ost = orc.St(op, st.commitments_c, st.mins, st.seed_nonce)
    rc, masks = orc.verify_batch([orc.transcript_new(M.LABEL)], [ost], [parsed], orc.RECOVER_AND_VERIFY)
    assert rc == 0 and masks == [bl]
    assert workload.LABEL == M.LABEL
