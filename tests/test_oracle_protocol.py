"""Oracle protocol self-consistency, mirroring the reference's own test matrix
(/root/reference/tests/ristretto.rs:25-373 and /root/reference/src/range_proof.rs:1328-1855).
The reference pins no bytes, so these are round trips and error-variant checks, as upstream."""
import pytest

import orc
import workload
from orc import (INVALID_ARGUMENT, INVALID_LENGTH, OK, RECOVER_AND_VERIFY, RECOVER_ONLY, VERIFICATION_FAILED,
                 VERIFY_ONLY)

L = orc.L


def test_derived_constants():
    """masking bases (ristretto.rs:92-95) and the head of the party-0 chains
    (bulletproof_gens.rs:91-96) as derived independently with hashlib + libsodium."""
    p = orc.Params(64, 1, 6)
    exp = [
        "044fad914b346d1623f0a123c90bec712c6bac717f2acbc48e12db5f6dcaef79",
        "429127c12411a4580d2606d2437a410a52254198b614e2d5c52ab8bb06576d55",
        "3255d8182cba353e52515411dcf0e2c28e926c5ca32689e41b722233e2bcfe70",
        "34f782cde81b8f949c34384912c4f978e612bd4a0609537682440e39a1864d15",
        "12a1a8238bf87962e8b102935ea88e14b2eee8d3d1719c63c44727b855205307",
        "aef270d5ac749567f6ffd256a692e1d4fb718565929693ca0aab94295b937b5e",
    ]
    for i in range(6):
        assert p.point(1, i).hex() == exp[i]
    assert p.point(0).hex() == "e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76"
    assert p.point(2, 0).hex() == "fc3b25801422672a6a8d3adb5d8457d4301fe92324b4fc56ae934c8713ddfe2d"
    assert p.point(2, 1).hex() == "ae817fdef62f713dd169dc8a26406f68be0bd3cd53652614636b0801567c4264"
    assert p.point(3, 0).hex() == "ba698f6dd08c501e32b55d2ee7259f6019d629fa2ba4d7039c5de157cba4df73"
    assert p.point(3, 1).hex() == "acf2d2b95428fac99b12da3bab92edf8ea3788c2fd16769e586397eede7b5052"


def test_generators_independent_derivation():
    """Gi/Hi chain vs hashlib.shake_256 + libsodium from_hash (independent of the oracle's own hash/map)."""
    import ctypes as C
    import hashlib

    s = orc.sodium()
    if s is None:
        pytest.skip("no libsodium")
    p = orc.Params(8, 4, 1)
    for party in range(4):
        for tag, which in ((b"G", 2), (b"H", 3)):
            stream = hashlib.shake_256(b"GeneratorsChain" + tag + party.to_bytes(4, "little")).digest(64 * 8)
            for i in range(8):
                o = C.create_string_buffer(32)
                s.crypto_core_ristretto255_from_hash(o, stream[64 * i : 64 * i + 64])
                assert p.point(which, party * 8 + i) == o.raw


CASES = [
    # (bit_length, aggregation sizes, ext, promise)   tests/ristretto.rs:25-142
    (8, [1], 1, "none"), (64, [1], 1, "none"),
    (4, [4], 2, "third"), (32, [4], 2, "third"),
    (64, [1, 1], 3, "equal"),
    (64, [1, 2], 1, "third"),
]


@pytest.mark.parametrize("n,ms,ext,promise", CASES)
def test_prove_and_verify(n, ms, ext, promise):
    c = workload.make_case(n, ms, ext, promise=promise, same_blinding=True)
    for action in (RECOVER_ONLY, RECOVER_AND_VERIFY, VERIFY_ONLY):
        rc, masks = orc.verify_batch(c.transcripts, c.statements, c.proofs, action)
        assert rc == OK
        for st, w, mk in zip(c.statements, c.witnesses, masks):
            if action != VERIFY_ONLY and st.seed_nonce is not None:
                assert mk == w.blindings[0]                       # ristretto.rs:254-289
            else:
                assert mk is None
    # wrong seed nonce: still verifies, masks differ (ristretto.rs:291-318)
    wrong = [orc.St(c.params, s.commitments, s.min_values, (s.seed_nonce + 1) % L if s.seed_nonce is not None else None)
             for s in c.statements]
    rc, masks = orc.verify_batch(c.transcripts, wrong, c.proofs, RECOVER_AND_VERIFY)
    assert rc == OK
    for st, w, mk in zip(wrong, c.witnesses, masks):
        if st.seed_nonce is not None:
            assert mk != w.blindings[0]
    # tampered promises (ristretto.rs:320-356)
    if promise != "none":
        bad = [orc.St(c.params, s.commitments, [v + 1 for v in s.min_values], s.seed_nonce) for s in c.statements]
        assert orc.verify_batch(c.transcripts, bad, c.proofs, VERIFY_ONLY)[0] == VERIFICATION_FAILED
    # serialisation round trip (ristretto.rs:359-371)
    for pr in c.proofs:
        b = orc.proof_to_bytes(pr)
        assert len(b) == 1 + 32 * (ext + 5 + 2 * pr.n_li)
        rc, q = orc.proof_from_bytes(b)
        assert rc == OK and orc.proof_to_bytes(q) == b


def test_promise_larger_than_value_fails_prover():
    """tests/ristretto.rs:231-241"""
    params = orc.Params(64, 1, 1)
    rng = orc.Rng("chacha", 1)
    v, b = 1000, [rng.random_not_zero()]
    st = orc.St(params, [params.commit(v, b)], [v + 1], None)
    rc, _, _ = orc.prove(orc.transcript_new(b"x"), st, orc.Wit([v], [b]), rng)
    assert rc == INVALID_ARGUMENT


def test_prover_consistency_errors():
    """src/range_proof.rs test_prover_consistency_errors"""
    params = orc.Params(4, 2, 1)
    rng = orc.Rng("chacha", 2)
    b = [rng.random_not_zero()]
    t = orc.transcript_new(b"x")
    # value exceeds bit length
    st = orc.St(params, [params.commit(16, b)], [None], None)
    assert orc.prove(t, st, orc.Wit([16], [b]), rng)[0] == INVALID_LENGTH
    # openings != commitments
    st = orc.St(params, [params.commit(1, b)], [None], None)
    assert orc.prove(t, st, orc.Wit([1, 2], [b, b]), rng)[0] == INVALID_LENGTH
    # witness extension degree mismatch
    assert orc.prove(t, st, orc.Wit([1], [b + b]), rng)[0] == INVALID_LENGTH
    # wrong opening
    assert orc.prove(t, st, orc.Wit([2], [b]), rng)[0] == INVALID_ARGUMENT
    # statement errors (range_statement.rs:36-73)
    with pytest.raises(orc.OracleError):
        orc.St(params, [params.commit(1, b)] * 3, [None] * 3, None)
    with pytest.raises(orc.OracleError):
        orc.St(params, [params.commit(1, b)] * 4, [None] * 4, None)
    with pytest.raises(orc.OracleError):
        orc.St(params, [params.commit(1, b)] * 2, [None] * 2, 5)
    with pytest.raises(orc.OracleError):
        orc.St(params, [params.commit(1, b)], [None, None], None)
    for bad in [(3, 1, 1), (128, 1, 1), (8, 3, 1), (8, 1, 0), (8, 1, 7)]:
        with pytest.raises(orc.OracleError):
            orc.Params(*bad)


def test_from_bytes_errors():
    """src/range_proof.rs test_from_bytes + fuzz target: ok => canonical"""
    c = workload.make_case(4, [1], 2)
    b = orc.proof_to_bytes(c.proofs[0])
    for cut in range(len(b)):
        # a truncation parses only when it still holds >= 1 whole (L, R) pair and no stray bytes
        whole = cut >= 1 + 32 * (2 + 5 + 2) and (cut - 1 - 32 * (2 + 5)) % 64 == 0
        assert (orc.proof_from_bytes(b[:cut])[0] == OK) == whole
    assert orc.proof_from_bytes(b + b"\x00" * 32)[0] == INVALID_LENGTH
    assert orc.proof_from_bytes(b + b"\x00" * 64)[0] == OK
    assert orc.proof_from_bytes(b + b"\x00")[0] == INVALID_LENGTH
    assert orc.proof_from_bytes(bytes([0]) + b[1:])[0] == INVALID_ARGUMENT
    assert orc.proof_from_bytes(bytes([7]) + b[1:])[0] == INVALID_ARGUMENT
    noncanon = b[:1] + (L).to_bytes(32, "little") + b[33:]
    assert orc.proof_from_bytes(noncanon)[0] == INVALID_ARGUMENT


def test_verify_errors():
    """src/range_proof.rs test_verify_errors / test_getters / test_consistency_errors"""
    c = workload.make_case(4, [1, 1], 1)
    T, S, Pf = c.transcripts, c.statements, c.proofs
    assert orc.verify_batch([], [], [], VERIFY_ONLY)[0] == INVALID_ARGUMENT
    assert orc.verify_batch(T, S, Pf[:1], VERIFY_ONLY)[0] == INVALID_ARGUMENT
    assert orc.verify_batch(T[:1], S, Pf, VERIFY_ONLY)[0] == INVALID_ARGUMENT
    # popped L / R
    p = Pf[0].copy(); p.n_li -= 1
    assert orc.verify_batch(T, S, [p, Pf[1]], VERIFY_ONLY)[0] == INVALID_LENGTH
    p = Pf[0].copy(); p.n_ri -= 1
    assert orc.verify_batch(T, S, [p, Pf[1]], VERIFY_ONLY)[0] == INVALID_LENGTH
    # non-canonical point encodings => InvalidArgument, identity => VerificationFailed (transcript validation)
    for fld in ("a", "a1", "b"):
        p = Pf[0].copy()
        for i in range(32):
            getattr(p, fld)[i] = 0
        getattr(p, fld)[0] = 1
        assert orc.verify_batch(T, S, [p, Pf[1]], VERIFY_ONLY)[0] == INVALID_ARGUMENT
        for i in range(32):
            getattr(p, fld)[i] = 0
        assert orc.verify_batch(T, S, [p, Pf[1]], VERIFY_ONLY)[0] == VERIFICATION_FAILED
    p = Pf[1].copy()
    p.li[0][0] ^= 1
    assert orc.verify_batch(T, S, [Pf[0], p], VERIFY_ONLY)[0] in (INVALID_ARGUMENT, VERIFICATION_FAILED)
    # inconsistent generators: different ext / bit length
    other = workload.make_case(8, [1], 1)
    assert orc.verify_batch(T, [S[0], other.statements[0]], [Pf[0], other.proofs[0]], VERIFY_ONLY)[0] == INVALID_ARGUMENT
    other = workload.make_case(4, [1], 2)
    assert orc.verify_batch(T, [S[0], other.statements[0]], [Pf[0], other.proofs[0]], VERIFY_ONLY)[0] == INVALID_ARGUMENT
    # oversize promise
    bad = orc.St(c.params, S[0].commitments, [16], S[0].seed_nonce)
    assert orc.verify_batch(T, [bad, S[1]], Pf, VERIFY_ONLY)[0] == INVALID_LENGTH


def test_aggregation_lower_than_generators():
    """src/range_proof.rs test_aggregation_lower_than_generators: m=1 under M=2 parameters (zero padding)"""
    c = workload.make_case(8, [1], 1, max_aggregation=2)
    assert orc.verify_batch(c.transcripts, c.statements, c.proofs, VERIFY_ONLY)[0] == OK


def test_verify_batch_only_first_256():
    """src/range_proof.rs:739-751: entries beyond index 255 are never looked at."""
    c = workload.make_case(2, [1] * 3, 1)
    bad = c.proofs[2].copy()
    bad.r1[0] ^= 1
    T = c.transcripts[:2] * 128 + [c.transcripts[2]]
    S = c.statements[:2] * 128 + [c.statements[2]]
    Pf = c.proofs[:2] * 128 + [bad]
    rc, masks = orc.verify_batch(T, S, Pf, VERIFY_ONLY)
    assert rc == OK and len(masks) == 256
    rc, _ = orc.verify_batch(T[1:], S[1:], Pf[1:], VERIFY_ONLY)
    assert rc == VERIFICATION_FAILED
