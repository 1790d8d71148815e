"""Committed golden vectors (tests/golden/bpp_golden.json, made by tests/golden/make_golden.py from the CPU oracle; the Rust
reference cannot run here and holds no vectors of its own).
  CPU: the oracle still reproduces every frozen byte (proofs, transcripts after proving, verdicts, recovered masks, the error code
       of every corruption) and its generators equal the constants SURVEY.md §8c derived independently with hashlib + libsodium.
  GPU: the device prover reproduces the proof bytes and advanced transcripts from the frozen inputs, the device verifier the
       verdicts, masks and error codes, and the device-derived generators equal the same constants."""
import json
import os

import pytest

import orc

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "bpp_golden.json")))


def _ints(xs):
    return [int(x, 16) for x in xs]


def _oracle_objects(case):
    params = orc.Params(case["bit_length"], case["max_aggregation"], case["extension_degree"])
    sts, wits = [], []
    for p in case["proofs"]:
        seed = int(p["seed_nonce"], 16) if p["seed_nonce"] is not None else None
        sts.append(orc.St(params, [bytes.fromhex(c) for c in p["commitments"]], p["minimum_value_promises"], seed))
        wits.append(orc.Wit(p["values"], [_ints(b) for b in p["blindings"]]))
    return params, sts, wits


@pytest.mark.parametrize("ci", range(len(GOLD["cases"])))
def test_oracle_reproduces_golden(ci):
    case = GOLD["cases"][ci]
    label = case["label"].encode()
    params, sts, wits = _oracle_objects(case)
    prs = []
    for p, st, w in zip(case["proofs"], sts, wits):
        assert [params.commit(v, b) for v, b in zip(w.values, w.blindings)] == st.commitments
        rc, pr, t_after = orc.prove(orc.transcript_new(label), st, w, orc.Rng("buffer", data=bytes.fromhex(p["rng_bytes"])))
        assert rc == 0 and orc.proof_to_bytes(pr).hex() == p["proof"] and t_after.hex() == p["prover_transcript_after"]
        prs.append(pr)
    rc, masks = orc.verify_batch([orc.transcript_new(label)] * len(prs), sts, prs, orc.RECOVER_AND_VERIFY)
    assert rc == case["verify"]["status"] == 0
    assert [None if m is None else [hex(x) for x in m] for m in masks] == case["verify"]["masks"]
    for c in case["corruptions_of_last_proof"]:
        rcb, bad = orc.proof_from_bytes(bytes.fromhex(c["proof"]))
        assert rcb == 0
        rc2, _ = orc.verify_batch([orc.transcript_new(label)] * len(prs), sts, prs[:-1] + [bad], orc.VERIFY_ONLY)
        assert rc2 == c["status"] != 0, c


def test_oracle_generators_equal_survey_constants():
    g = GOLD["generators_survey_8c"]
    p = orc.Params(64, 1, 6)
    assert p.point(0).hex() == g["H"] and [p.point(1, k).hex() for k in range(6)] == g["G"]
    assert (p.point(2, 0).hex(), p.point(2, 1).hex(), p.point(3, 0).hex(), p.point(3, 1).hex()) == (g["Gi_0_0"], g["Gi_0_1"], g["Hi_0_0"], g["Hi_0_1"])


@pytest.mark.gpu
@pytest.mark.parametrize("ci", range(len(GOLD["cases"])))
def test_device_reproduces_golden(ci):
    import bpp

    api = bpp.pkg.api
    case = GOLD["cases"][ci]
    label = case["label"].encode()
    gp = api.RangeParameters.init(bpp.engine(), case["bit_length"], case["max_aggregation"], case["extension_degree"])
    sts, wits, streams = [], [], []
    for p in case["proofs"]:
        seed = int(p["seed_nonce"], 16) if p["seed_nonce"] is not None else None
        sts.append(api.RangeStatement.init(gp, [bytes.fromhex(c) for c in p["commitments"]], p["minimum_value_promises"], seed))
        wits.append(api.RangeWitness.init([api.CommitmentOpening(v, _ints(b)) for v, b in zip(p["values"], p["blindings"])]))
        streams.append(bytes.fromhex(p["rng_bytes"]))
        assert gp.gens.commit_batch(p["values"], [_ints(b) for b in p["blindings"]]) == [bytes.fromhex(c) for c in p["commitments"]]
    # prover: statements of one shape per call
    proofs = []
    for i, p in enumerate(case["proofs"]):
        t = api.Transcript(label)
        (res,) = api.RangeProof.prove_batch([t], [sts[i]], [wits[i]], [streams[i]])
        assert not isinstance(res, Exception), res
        assert res.to_bytes().hex() == p["proof"] and t.state.hex() == p["prover_transcript_after"]
        proofs.append(res)
    # verifier: verdict + masks, then every frozen corruption with the frozen error code
    masks = api.RangeProof.verify_batch([api.Transcript(label) for _ in proofs], sts, proofs, api.VerifyAction.RecoverAndVerify)
    assert [None if m is None else [hex(x) for x in m.blindings()] for m in masks] == case["verify"]["masks"]
    for c in case["corruptions_of_last_proof"]:
        try:
            bad = api.RangeProof.from_bytes(bytes.fromhex(c["proof"]))
            api.RangeProof.verify_batch([api.Transcript(label) for _ in proofs], sts, proofs[:-1] + [bad], api.VerifyAction.VerifyOnly)
            raise AssertionError("corrupted proof accepted: %r" % (c["field"],))
        except bpp.pkg.EngineError as e:
            assert e.code == c["status"], (c["field"], e.code, c["status"])


@pytest.mark.gpu
def test_device_generators_equal_survey_constants():
    import bpp

    g = GOLD["generators_survey_8c"]
    gens = bpp.pkg.Gens(bpp.engine(), 64, 1, 6)
    assert gens.point(0).hex() == g["H"] and [gens.point(1, k).hex() for k in range(6)] == g["G"]
    assert (gens.point(2, 0).hex(), gens.point(2, 1).hex(), gens.point(3, 0).hex(), gens.point(3, 1).hex()) == (g["Gi_0_0"], g["Gi_0_1"], g["Hi_0_0"], g["Hi_0_1"])
    gens.close()
