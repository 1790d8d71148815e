"""BASELINE.json configs at their full sizes, through size-independent properties and cross-checks between the two device
paths and the CPU oracle:
  configs[1]  1024 non-aggregated 64-bit proofs = 4 reference calls of 256
  configs[2]  aggregated 64-bit proof, aggregation 32 (N = 2048, 11 rounds): prove + verify
  configs[3]  256 proofs, extension degree 3, minimum-value promises
Proofs are made by the device prover (byte-identical to the oracle prover on a sample), verified by the device verifier
(all) and by the oracle verifier (a sample chunk); corruptions must fail exactly their own chunk."""
import hashlib

import pytest

import bpp
import orc

pytestmark = pytest.mark.gpu
api = bpp.pkg.api
LABEL = b"BatchedRangeProofTest"


def _make(n, m, ext, count, seed, with_seed):
    eng = bpp.engine()
    gp = api.RangeParameters.init(eng, n, m, ext)
    rng = orc.Rng("chacha", seed)
    vals = [[rng.next_u64() % (1 << (n - 1)) for _ in range(m)] for _ in range(count)]
    blinds = [[[rng.random_not_zero() for _ in range(ext)] for _ in range(m)] for _ in range(count)]
    flat_c = gp.gens.commit_batch([v for vs in vals for v in vs], [b for bs in blinds for b in bs])
    commits = [flat_c[i * m:(i + 1) * m] for i in range(count)]
    mins = [[v // 3 for v in vs] for vs in vals]
    seeds = [rng.random_not_zero() if (with_seed and m == 1) else None for _ in range(count)]
    sts = [api.RangeStatement.init(gp, commits[i], mins[i], seeds[i]) for i in range(count)]
    wits = [api.RangeWitness.init([api.CommitmentOpening(v, b) for v, b in zip(vals[i], blinds[i])]) for i in range(count)]
    need = api.RangeProof.rng_bytes_needed(gp, m)
    streams = [hashlib.shake_256(b"cfg-%d-%d-%d-%d" % (n, m, ext, i)).digest(need) for i in range(count)]
    trs = [api.Transcript(LABEL) for _ in range(count)]
    proofs = api.RangeProof.prove_batch(trs, sts, wits, streams)
    assert not any(isinstance(p, Exception) for p in proofs)
    return gp, vals, blinds, commits, mins, seeds, sts, proofs, streams


def _oracle_statements(op, commits, mins, seeds, idx):
    return [orc.St(op, commits[i], mins[i], seeds[i]) for i in idx]


def test_config1_1024_proofs_as_four_calls():
    n, m, ext, count = 64, 1, 1, 1024
    gp, vals, blinds, commits, mins, seeds, sts, proofs, streams = _make(n, m, ext, count, 11, True)
    op = orc.Params(n, m, ext)
    # device-proved == oracle-proved on a sample
    for i in (0, 511, 1023):
        rc, pr, _ = orc.prove(orc.transcript_new(LABEL), orc.St(op, commits[i], mins[i], seeds[i]), orc.Wit(vals[i], blinds[i]),
                              orc.Rng("buffer", data=streams[i]))
        assert rc == 0 and orc.proof_to_bytes(pr) == proofs[i].to_bytes()
    calls = [([api.Transcript(LABEL) for _ in range(256)], sts[c * 256:(c + 1) * 256], proofs[c * 256:(c + 1) * 256]) for c in range(4)]
    status, masks = api.verify_chunks(gp, calls, api.VerifyAction.RecoverAndVerify)
    assert status == [0, 0, 0, 0]
    for c in range(4):
        for k, mk in enumerate(masks[c]):
            assert mk.blindings() == blinds[c * 256 + k][0]
    # the oracle verifier accepts one whole chunk of device-made proofs and recovers the same masks
    idx = list(range(256, 512))
    o_proofs = [orc.proof_from_bytes(proofs[i].to_bytes())[1] for i in idx]
    rc, o_masks = orc.verify_batch([orc.transcript_new(LABEL)] * 256, _oracle_statements(op, commits, mins, seeds, idx), o_proofs, orc.RECOVER_AND_VERIFY)
    assert rc == 0 and [mk for mk in o_masks] == [blinds[i][0] for i in idx]
    # one corrupted proof fails its own call only; an undecodable point gives InvalidArgument for its call
    bad = bytearray(proofs[700].to_bytes()); bad[1 + 32 * 4 + 5] ^= 0x10           # r1
    bad2 = bytearray(proofs[10].to_bytes()); bad2[1 + 32:1 + 64] = b"\x01" + bytes(31)  # A = [1, 0, ...]: not a valid encoding
    proofs2 = list(proofs)
    proofs2[700] = api.RangeProof.from_bytes(bytes(bad))
    proofs2[10] = api.RangeProof.from_bytes(bytes(bad2))
    calls = [([api.Transcript(LABEL) for _ in range(256)], sts[c * 256:(c + 1) * 256], proofs2[c * 256:(c + 1) * 256]) for c in range(4)]
    status, _ = api.verify_chunks(gp, calls, api.VerifyAction.VerifyOnly)
    assert status == [orc.INVALID_ARGUMENT, 0, orc.VERIFICATION_FAILED, 0]


def test_config2_aggregation_32():
    n, m, ext, count = 64, 32, 1, 2
    gp, vals, blinds, commits, mins, seeds, sts, proofs, streams = _make(n, m, ext, count, 12, False)
    assert len(proofs[0].to_bytes()) == 897 and len(proofs[0].li()) == 11
    op = orc.Params(n, m, ext)
    rc, pr, t_after = orc.prove(orc.transcript_new(LABEL), orc.St(op, commits[0], mins[0], None), orc.Wit(vals[0], blinds[0]),
                                orc.Rng("buffer", data=streams[0]))
    assert rc == 0 and orc.proof_to_bytes(pr) == proofs[0].to_bytes()
    got = api.RangeProof.verify_batch([api.Transcript(LABEL) for _ in range(count)], sts, proofs, api.VerifyAction.VerifyOnly)
    assert got == [None, None]
    rc, _ = orc.verify_batch([orc.transcript_new(LABEL)] * count, _oracle_statements(op, commits, mins, seeds, range(count)),
                             [orc.proof_from_bytes(p.to_bytes())[1] for p in proofs], orc.VERIFY_ONLY)
    assert rc == 0
    # a promise raised by one makes the batch fail on both sides (tests/ristretto.rs:320-356)
    st_bad = api.RangeStatement.init(gp, commits[1], [mins[1][0] + 1] + mins[1][1:], None)
    with pytest.raises(bpp.pkg.EngineError) as ei:
        api.RangeProof.verify_batch([api.Transcript(LABEL) for _ in range(count)], [sts[0], st_bad], proofs, api.VerifyAction.VerifyOnly)
    assert ei.value.code == orc.VERIFICATION_FAILED


def test_config3_256_proofs_extension_degree_3():
    n, m, ext, count = 64, 1, 3, 256
    gp, vals, blinds, commits, mins, seeds, sts, proofs, streams = _make(n, m, ext, count, 13, True)
    assert len(proofs[0].to_bytes()) == 641
    op = orc.Params(n, m, ext)
    for i in (0, 255):
        rc, pr, _ = orc.prove(orc.transcript_new(LABEL), orc.St(op, commits[i], mins[i], seeds[i]), orc.Wit(vals[i], blinds[i]),
                              orc.Rng("buffer", data=streams[i]))
        assert rc == 0 and orc.proof_to_bytes(pr) == proofs[i].to_bytes()
    masks = api.RangeProof.verify_batch([api.Transcript(LABEL) for _ in range(count)], sts, proofs, api.VerifyAction.RecoverAndVerify)
    assert [mk.blindings() for mk in masks] == [b[0] for b in blinds]
    rc, o_masks = orc.verify_batch([orc.transcript_new(LABEL)] * count, _oracle_statements(op, commits, mins, seeds, range(count)),
                                   [orc.proof_from_bytes(p.to_bytes())[1] for p in proofs], orc.RECOVER_AND_VERIFY)
    assert rc == 0 and o_masks == [b[0] for b in blinds]


def test_config4_msm_large_properties():
    """raw MSM at 2^18: linearity and shard-sum == whole (sizes the oracle cannot follow)"""
    import ctypes as C

    e = bpp.engine()
    n = 1 << 18
    base = e.from_uniform(hashlib.shake_256(b"big-msm").digest(64 * 4096))
    pts = (base * (n // 4096))
    s1 = bytearray(hashlib.shake_256(b"s1").digest(32 * n)); s2 = bytearray(hashlib.shake_256(b"s2").digest(32 * n))
    for buf in (s1, s2):
        for i in range(31, len(buf), 32):
            buf[i] &= 0x07
    plan = bpp.pkg.MsmPlan(e, pts)
    plan.set_scalars(bytes(s1)); r1 = plan.run()
    plan.set_scalars(bytes(s2)); r2 = plan.run()
    ssum = bytearray(32 * n)
    for i in range(n):
        v = int.from_bytes(s1[32 * i:32 * i + 32], "little") + int.from_bytes(s2[32 * i:32 * i + 32], "little")
        ssum[32 * i:32 * i + 32] = v.to_bytes(32, "little")                        # < 2^252: canonical without reduction
    plan.set_scalars(bytes(ssum)); r3 = plan.run()
    o = C.create_string_buffer(32)
    assert orc.lib().orc_ristretto_add(r1, r2, o) == 1 and o.raw == r3
    # shard-sum == whole, as bpp_msm_segmented over 8 shards then a unit-scalar MSM of the 8 partials
    offs = [i * (n // 8) for i in range(9)]
    parts = e.msm_segmented(offs, bytes(s1), pts)
    assert e.msm((1).to_bytes(32, "little") * 8, b"".join(parts)) == r1
    plan.close()
