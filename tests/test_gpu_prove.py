"""GPU parity for the batched prover: with the same transcript, statement, witness and external-RNG byte stream the proof
bytes produced through bpp_prove_batch must be IDENTICAL to the CPU oracle's prove_with_rng, the transcripts must advance
identically, and the proofs must verify (reference round trip: /root/reference/tests/ristretto.rs:152-373,
prover error paths: /root/reference/src/range_proof.rs test_prover_consistency_errors)."""
import hashlib

import pytest

import bpp
import orc

pytestmark = pytest.mark.gpu
api = bpp.pkg.api
L = orc.L
LABEL = b"BatchedRangeProofTest"

_params = {}


def params_pair(n, M, ext):
    key = (n, M, ext)
    if key not in _params:
        _params[key] = (api.RangeParameters.init(bpp.engine(), n, M, ext), orc.Params(n, M, ext))
    return _params[key]


def make_inputs(n, m, ext, count, seed, with_seed_nonce, promise):
    """deterministic statements / witnesses for `count` proofs"""
    gp, op = params_pair(n, m, ext)
    rng = orc.Rng("chacha", seed)
    items = []
    for _ in range(count):
        vals, blinds, mins = [], [], []
        for _ in range(m):
            v = rng.next_u64() % (1 << (n - 1))
            vals.append(v)
            blinds.append([rng.random_not_zero() for _ in range(ext)])
            mins.append({"none": None, "third": v // 3, "equal": v}[promise])
        commits = [op.commit(v, b) for v, b in zip(vals, blinds)]
        sn = rng.random_not_zero() if (with_seed_nonce and m == 1) else None
        items.append((vals, blinds, mins, commits, sn))
    return gp, op, items


def rng_stream(tag, i, nbytes):
    return hashlib.shake_256(b"ext-rng-%s-%d" % (tag, i)).digest(nbytes)


def oracle_prove(op, item, stream, label=LABEL):
    vals, blinds, mins, commits, sn = item
    st = orc.St(op, commits, mins, sn)
    wit = orc.Wit(vals, blinds)
    rc, pr, t_after = orc.prove(orc.transcript_new(label), st, wit, orc.Rng("buffer", data=stream))
    return rc, (orc.proof_to_bytes(pr) if rc == 0 else None), t_after


@pytest.mark.parametrize("n,m,ext,seeded,promise", [
    (64, 1, 1, True, "third"), (64, 1, 1, False, "none"), (8, 1, 2, True, "equal"), (8, 4, 2, False, "third"),
    (64, 2, 3, False, "none"), (32, 4, 1, False, "equal"), (4, 1, 6, True, "none"), (64, 1, 3, False, "third"),
])
def test_proof_bytes_identical_to_oracle(n, m, ext, seeded, promise):
    count = 19          # two groups of eight through the lock-step host sponge (engine_prove.cu, Lock8) + three one at a time
    gp, op, items = make_inputs(n, m, ext, count, 1000 + n + 10 * m + ext, seeded, promise)
    need = api.RangeProof.rng_bytes_needed(gp, m)
    streams = [rng_stream(b"%d-%d-%d" % (n, m, ext), i, need) for i in range(count)]
    sts = [api.RangeStatement.init(gp, it[3], it[2], it[4]) for it in items]
    wits = [api.RangeWitness.init([api.CommitmentOpening(v, b) for v, b in zip(it[0], it[1])]) for it in items]
    trs = [api.Transcript(LABEL) for _ in items]
    got = api.RangeProof.prove_batch(trs, sts, wits, streams)
    for i, it in enumerate(items):
        rc, want, t_after = oracle_prove(op, it, streams[i])
        assert rc == 0
        assert not isinstance(got[i], Exception), got[i]
        assert got[i].to_bytes() == want, (i, got[i].to_bytes().hex()[:80], want.hex()[:80])
        assert trs[i].state == t_after
    # and they verify (with mask recovery where a seed nonce was used)
    vtrs = [api.Transcript(LABEL) for _ in items]
    masks = api.RangeProof.verify_batch(vtrs, sts, got, api.VerifyAction.RecoverAndVerify)
    for it, mk in zip(items, masks):
        if it[4] is not None:
            assert mk.blindings() == it[1][0]
        else:
            assert mk is None


def test_lockstep_groups_fall_back_lane_by_lane():
    """The host advances eight proofs per vectorised sponge only while the eight agree on everything that steers the transcript
    (engine_prove.cu, Lock8::uniform).  Groups that do not -- mixed nonce sources, a transcript label of another length, a proof that
    fails its checks -- take the one-at-a-time path; every proof still has the oracle's bytes, status and transcript."""
    n, m, ext = 16, 1, 2
    gp, op, seeded = make_inputs(n, m, ext, 32, 9001, True, "third")
    _, _, plain = make_inputs(n, m, ext, 32, 9002, False, "none")
    need = api.RangeProof.rng_bytes_needed(gp, m)
    items, labels = [], []
    for i in range(32):
        g = i // 8
        it = seeded[i] if (g == 0 or (g == 2 and i % 2)) else plain[i]         # group 0 seeded, 1 and 3 plain, 2 mixed
        items.append(it)
        labels.append(b"another label, longer" if i == 13 else LABEL)           # group 1: one transcript at another sponge position
    vals, blinds, mins, commits, sn = items[27]                                 # group 3: one value below its promise
    items[27] = (vals, blinds, [vals[0] + 1], commits, sn)
    streams = [rng_stream(b"lk", i, need) for i in range(32)]
    sts = [api.RangeStatement.init(gp, it[3], it[2], it[4]) for it in items]
    wits = [api.RangeWitness.init([api.CommitmentOpening(v, b) for v, b in zip(it[0], it[1])]) for it in items]
    trs = [api.Transcript(lb) for lb in labels]
    got = api.RangeProof.prove_batch(trs, sts, wits, streams)
    for i, it in enumerate(items):
        rc, want, t_after = oracle_prove(op, it, streams[i], labels[i])
        if rc:
            assert i == 27 and isinstance(got[i], bpp.pkg.EngineError) and got[i].code == rc
        else:
            assert not isinstance(got[i], Exception), (i, got[i])
            assert got[i].to_bytes() == want, i
            assert trs[i].state == t_after, i


def test_lockstep_off_gives_the_same_bytes():
    """BPP_PROVE_LOCKSTEP=0: every proof through the one-at-a-time host sponge (the round-1 path)"""
    import os
    import subprocess
    import sys

    if os.environ.get("BPP_PROVE_LOCKSTEP") or os.environ.get("BPP_PROVE_FOLD"):
        pytest.skip("already inside a variant run")
    env = dict(os.environ, BPP_PROVE_LOCKSTEP="0")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_prove.py"), "-x", "-q", "-m", "gpu",
                        "-k", "identical_to_oracle or lane_by_lane"], env=env, capture_output=True, text=True, timeout=1200, cwd=root)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_single_prove_with_rng_object():
    gp, op, items = make_inputs(64, 1, 1, 1, 77, True, "third")
    it = items[0]
    st = api.RangeStatement.init(gp, it[3], it[2], it[4])
    wit = api.RangeWitness.init([api.CommitmentOpening(it[0][0], it[1][0])])
    rng_a, rng_b = orc.Rng("chacha", 4242), orc.Rng("chacha", 4242)
    t = api.Transcript(LABEL)
    proof = api.RangeProof.prove_with_rng(t, st, wit, rng_a)
    rc, pr, t_after = orc.prove(orc.transcript_new(LABEL), orc.St(op, it[3], it[2], it[4]), orc.Wit(it[0], it[1]), rng_b)
    assert rc == 0 and proof.to_bytes() == orc.proof_to_bytes(pr) and t.state == t_after
    assert len(proof.to_bytes()) == 577          # 1 + 32 * (ext + 5 + 2 * log2(n*m)), SURVEY §3.4


def test_prover_error_paths_match_oracle():
    """range_proof.rs test_prover_consistency_errors: wrong opening, value below its promise, value beyond the bit length"""
    n, m, ext = 8, 1, 1
    gp, op, items = make_inputs(n, m, ext, 4, 5, False, "none")
    need = api.RangeProof.rng_bytes_needed(gp, m)
    streams = [rng_stream(b"err", i, need) for i in range(4)]
    # 0: good; 1: blinding off by one; 2: promise above the value; 3: good
    vals = [it[0] for it in items]
    blinds = [it[1] for it in items]
    mins = [it[2] for it in items]
    blinds[1] = [[(blinds[1][0][0] + 1) % L]]
    mins[2] = [vals[2][0] + 1]
    sts = [api.RangeStatement.init(gp, it[3], mn, it[4]) for it, mn in zip(items, mins)]
    wits = [api.RangeWitness.init([api.CommitmentOpening(v[0], b[0])]) for v, b in zip(vals, blinds)]
    trs = [api.Transcript(LABEL) for _ in items]
    got = api.RangeProof.prove_batch(trs, sts, wits, streams)
    for i in range(4):
        rc, want, t_after = oracle_prove(op, (vals[i], blinds[i], mins[i], items[i][3], items[i][4]), streams[i])
        if rc:
            assert isinstance(got[i], bpp.pkg.EngineError) and got[i].code == rc, (i, got[i], rc)
        else:
            assert got[i].to_bytes() == want
    assert isinstance(got[1], bpp.pkg.EngineError) and isinstance(got[2], bpp.pkg.EngineError)
    # value does not fit the bit length -> InvalidLength (:264-271); the commitment is consistent with the oversized value
    big = 1 << n
    c = op.commit(big, blinds[0][0])
    st = api.RangeStatement.init(gp, [c], [None], None)
    res = api.RangeProof.prove_batch([api.Transcript(LABEL)], [st], [api.RangeWitness.init([api.CommitmentOpening(big, blinds[0][0])])], [streams[0]])
    assert isinstance(res[0], bpp.pkg.EngineError) and res[0].code == orc.INVALID_LENGTH
    # witness / statement shape mismatches (:248-260)
    w2 = api.RangeWitness.init([api.CommitmentOpening(1, [1, 2])])
    res = api.RangeProof.prove_batch([api.Transcript(LABEL)], [sts[0]], [w2], [streams[0]])
    assert isinstance(res[0], bpp.pkg.EngineError) and res[0].code == orc.INVALID_LENGTH


def test_folding_path_still_byte_identical():
    """the prover's default path evaluates every commitment from fixed-base window tables (k_fb.cu); the generator-folding path it
    replaces stays for generator sets whose tables exceed the memory budget -- run this file's byte-identity tests through it"""
    import os
    import subprocess
    import sys

    if os.environ.get("BPP_PROVE_FOLD"):
        pytest.skip("already inside the folding-path run")
    env = dict(os.environ, BPP_PROVE_FOLD="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_prove.py"), "-x", "-q", "-m", "gpu"],
                       env=env, capture_output=True, text=True, timeout=1200, cwd=root)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
