"""Verification lanes (api.VerifierPool: one bpp_ctx + host thread per lane) and the alternative kernel paths.
  * batches verified concurrently on S lanes give exactly the oracle's statuses, masks and advanced transcripts, valid and
    corrupted, whatever S is;
  * the kernel variants that are picked by size (thread-per-bucket / quad-per-bucket / split-thread MSM, long-vector scalar prep) are forced
    through their environment switches in a child process and must reproduce the oracle on the same cases."""
import os
import subprocess
import sys

import pytest

import bpp
import orc
import workload

pytestmark = pytest.mark.gpu
api = bpp.pkg.api
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _calls(params, case, lo, hi):
    sts = [api.RangeStatement.init(params, s.commitments, s.min_values, s.seed_nonce) for s in case.statements[lo:hi]]
    prs = [api.RangeProof.from_bytes(orc.proof_to_bytes(p)) for p in case.proofs[lo:hi]]
    trs = [api.Transcript(state=t) for t in case.transcripts[lo:hi]]
    return (trs, sts, prs)


@pytest.mark.parametrize("lanes", [1, 3])
def test_pool_matches_oracle(lanes):
    case = workload.make_case(64, [1] * 12 + [2, 4], 1, max_aggregation=4, promise="third", rng_seed=77)
    # 7 batches of 2 proofs; batch 3 carries a corrupted proof, batch 5 a proof with a flipped challenge-relevant point
    bad = {3: 0, 5: 1}
    pool = api.VerifierPool(0, 64, 4, 1, lanes=lanes)
    try:
        batches, expect = [], []
        for bi in range(7):
            lo, hi = 2 * bi, 2 * bi + 2
            proofs = [p.copy() for p in case.proofs[lo:hi]]
            if bi in bad:
                proofs[bad[bi]].r1[3] ^= 0x10
            rc, masks = orc.verify_batch(list(case.transcripts[lo:hi]), case.statements[lo:hi], proofs, orc.RECOVER_AND_VERIFY)
            expect.append((rc, masks))
            _, params = pool.lanes[bi % lanes]
            trs, sts, _ = _calls(params, case, lo, hi)
            prs = [api.RangeProof.from_bytes(orc.proof_to_bytes(p)) for p in proofs]
            batches.append([(trs, sts, prs)])
        got = pool.verify_many(batches, api.VerifyAction.RecoverAndVerify)
        assert pool.launch_count() > 0
        for bi, ((status, masks), (rc, want)) in enumerate(zip(got, expect)):
            assert status == [rc], (bi, status, rc)
            if rc == 0:
                for g, w in zip(masks[0], want):
                    assert (g is None) == (w is None) and (g is None or g.blindings() == w)
            else:
                assert rc == orc.VERIFICATION_FAILED and bi in bad
    finally:
        pool.close()


_CHILD = r"""
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import bpp, orc, workload
api = bpp.pkg.api
eng = bpp.pkg.Engine(0)
# 40 single proofs + aggregated ones: 2 calls; enough MSM entries / buckets for every kernel variant to do real work
case = workload.make_case(64, [1] * 20 + [2, 4] + [1] * 18, 1, max_aggregation=4, promise="third", rng_seed=5)
params = api.RangeParameters.init(eng, 64, 4, 1)
def calls(lo, hi, proofs):
    sts = [api.RangeStatement.init(params, s.commitments, s.min_values, s.seed_nonce) for s in case.statements[lo:hi]]
    prs = [api.RangeProof.from_bytes(orc.proof_to_bytes(p)) for p in proofs]
    trs = [api.Transcript(state=t) for t in case.transcripts[lo:hi]]
    return (trs, sts, prs)
good = [p.copy() for p in case.proofs]
bad = [p.copy() for p in case.proofs]
bad[25].s1[0] ^= 1
for proofs, want in ((good, [0, 0]), (bad, [0, orc.VERIFICATION_FAILED])):
    status, masks = api.verify_chunks(params, [calls(0, 20, proofs[:20]), calls(20, 40, proofs[20:])], api.VerifyAction.RecoverAndVerify)
    assert status == want, (status, want)
    rc, om = orc.verify_batch(list(case.transcripts[:20]), case.statements[:20], proofs[:20], orc.RECOVER_AND_VERIFY)
    assert rc == 0
    for g, w in zip(masks[0], om):
        assert (g is None) == (w is None) and (g is None or g.blindings() == w)
# raw MSM against the oracle
import hashlib
n = 3000
pts = eng.from_uniform(hashlib.shake_256(b"lanes-msm").digest(64 * n))
sc = bytearray(hashlib.shake_256(b"lanes-sc").digest(32 * n))
for i in range(31, len(sc), 32):
    sc[i] &= 0x0f
import ctypes as C
o = C.create_string_buffer(32)
assert orc.lib().orc_msm(bytes(sc), pts, n, 0, o) == 1
assert eng.msm(bytes(sc), pts) == o.raw
print("child ok")
"""


@pytest.mark.parametrize("env", [{"BPP_MSM_BUCKET": "1"}, {"BPP_MSM_BUCKET": "2"}, {"BPP_MSM_BUCKET": "3", "BPP_MSM_SPLIT": "2"},
                                 {"BPP_MSM_BUCKET": "3", "BPP_MSM_SPLIT": "4"}, {"BPP_MSM_BUCKET": "3", "BPP_MSM_SPLIT": "8"}, {"BPP_MSM_REDUCE": "1"}, {"BPP_MSM_REDUCE": "1", "BPP_MSM_REDUCE_PARTS": "2"}, {"BPP_MSM_REDUCE": "1", "BPP_MSM_REDUCE_PARTS": "4"},
                                 {"BPP_MSM_REDUCE": "2"}, {"BPP_MSM_SCAN_SORT": "1"}, {"BPP_MSM_WINDOW_SORT": "1"}, {"BPP_MSM_LOCAL_RANK": "1"}, {"BPP_MSM_LOCAL_RANK": "1", "BPP_MSM_BUCKET": "2"}, {"BPP_MSM_SCAN_SORT": "1", "BPP_MSM_BUCKET": "1"}, {"BPP_VPREP_WARP": "100000"},
                                 {"BPP_VPREP_DIRECT": "1"}, {"BPP_NO_GRAPHS": "1"}, {"BPP_SCALAR_WEIGHTS": "1"},
                                 {"BPP_THROUGHPUT_MODE": "2"}, {"BPP_THROUGHPUT_MODE": "2", "BPP_WEIGHTS_WARP": "1"}, {"BPP_NO_DIRECT_DMA": "1"},
                                 {"BPP_MERGED_CHECK": "1"}, {"BPP_MERGED_CHECK": "1", "BPP_THROUGHPUT_MODE": "2"}, {"BPP_MERGED_CHECK": "1", "BPP_MSM_SCAN_SORT": "1"}])
def test_kernel_variants_match_oracle(env):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", _CHILD % (ROOT, os.path.join(ROOT, "tests"))], env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "child ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_pool_prove_many_byte_identical_to_oracle():
    """prove_many: jobs proved concurrently on several lanes give exactly the oracle prover's bytes and advanced transcripts"""
    import hashlib

    n, ext = 64, 1
    op = orc.Params(n, 2, ext)
    pool = api.VerifierPool(0, n, 2, ext, lanes=3)
    try:
        rng = orc.Rng("chacha", 909)
        jobs, want = [], []
        for j, (m, count) in enumerate([(1, 5), (2, 3), (1, 2), (1, 4)]):
            _, params = pool.lanes[j % 3]
            trs, sts, wits, rbs, exp = [], [], [], [], []
            for k in range(count):
                vals = [rng.next_u64() % (1 << 63) for _ in range(m)]
                blinds = [[rng.random_not_zero()] for _ in range(m)]
                commits = [op.commit(v, b) for v, b in zip(vals, blinds)]
                mins = [v // 3 for v in vals]
                seed = rng.random_not_zero() if m == 1 else None
                stream = hashlib.shake_256(b"pm-%d-%d" % (j, k)).digest(api.RangeProof.rng_bytes_needed(params, m))
                rc, pr, t_after = orc.prove(orc.transcript_new(workload.LABEL), orc.St(op, commits, mins, seed), orc.Wit(vals, blinds),
                                            orc.Rng("buffer", data=stream))
                assert rc == 0
                exp.append((orc.proof_to_bytes(pr), t_after))
                trs.append(api.Transcript(workload.LABEL))
                sts.append(api.RangeStatement.init(params, commits, mins, seed))
                wits.append(api.RangeWitness.init([api.CommitmentOpening(v, b) for v, b in zip(vals, blinds)]))
                rbs.append(stream)
            jobs.append((trs, sts, wits, rbs))
            want.append(exp)
        got = pool.prove_many(jobs)
        for j, (res, exp) in enumerate(zip(got, want)):
            for k, (r, (pb, ta)) in enumerate(zip(res, exp)):
                assert not isinstance(r, Exception), (j, k, r)
                assert r.to_bytes() == pb and jobs[j][0][k].state == ta, (j, k)
    finally:
        pool.close()
