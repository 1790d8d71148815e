"""GPU parity for RangeProof::verify_batch through the C ABI: verdicts, error variants, recovered masks and the advanced
transcripts must equal the CPU oracle's on the same inputs.  Shapes and strategies follow the reference's own tests
(/root/reference/tests/ristretto.rs:25-373, /root/reference/src/range_proof.rs:1328-1855)."""
import random

import pytest

import bpp
import orc
import workload

pytestmark = pytest.mark.gpu
api = bpp.pkg.api
VA = api.VerifyAction

_params_cache = {}


@pytest.fixture(autouse=True, params=["device_replay_sm", "device_replay_warp", "device_replay_thread", "host_replay", "device_replay_no_graphs", "blocking_waits",
                                      "device_weights", "zero_weight_fallback", "merged_check", "merged_check_forced_fallback", "merged_check_device_weights",
                                      "merged_check_zero_weight_fallback"])
def replay_mode(request):
    """every test runs with loop 1 (the Merlin transcript replay) on the device (both kernels) and on host threads, and with the
    pass issued as captured CUDA graphs (default), kernel by kernel, with blocking waits, and as the single graph with the
    verifier-weight transcripts on the device: same results"""
    bpp.engine().set_replay_mode({"device_replay_warp": 3, "device_replay_thread": 2, "host_replay": 0}.get(request.param, 1))
    bpp.engine().set_graphs(request.param != "device_replay_no_graphs")
    # blocking waits; one graph per pass with the verifier weights drawn on the device
    bpp.engine().set_throughput_mode({"blocking_waits": 1, "device_weights": 2, "merged_check_device_weights": 2}.get(request.param, 0))
    # the path a zero batch weight takes (Scalar::random_not_zero redraws): forced, results unchanged
    # merged check (one multiscalar check per call of several chunks, chunk by chunk only when it fails): alone, with the chunk-by-chunk
    # pass forced, on top of device-side weights, and combined with the zero-weight path
    bpp.engine().set_merged_check(request.param.startswith("merged_check"))
    bpp.engine().set_test_hooks({"zero_weight_fallback": 1, "merged_check_forced_fallback": 2, "merged_check_zero_weight_fallback": 1}.get(request.param, 0))
    yield request.param
    bpp.engine().set_test_hooks(0)
    bpp.engine().set_merged_check(False)
    bpp.engine().set_replay_mode(True)
    bpp.engine().set_graphs(True)
    bpp.engine().set_throughput_mode(False)


def gpu_params(n, M, ext):
    key = (n, M, ext)
    if key not in _params_cache:
        _params_cache[key] = api.RangeParameters.init(bpp.engine(), n, M, ext)
    return _params_cache[key]


def to_api(case, params, seed_override=None, promise_delta=0):
    sts, prs, trs = [], [], []
    for i, (st, pr, tr) in enumerate(zip(case.statements, case.proofs, case.transcripts)):
        mins = [None if v is None else v + promise_delta for v in st.min_values]
        seed = st.seed_nonce if seed_override is None else (seed_override if st.seed_nonce is not None else None)
        sts.append(api.RangeStatement.init(params, st.commitments, mins, seed))
        prs.append(api.RangeProof.from_bytes(orc.proof_to_bytes(pr)))
        trs.append(api.Transcript(state=tr))
    return trs, sts, prs


def oracle_verify(case, action, seed_override=None, promise_delta=0, proofs=None):
    sts = []
    for st in case.statements:
        mins = [None if v is None else v + promise_delta for v in st.min_values]
        seed = st.seed_nonce if seed_override is None else (seed_override if st.seed_nonce is not None else None)
        sts.append(orc.St(case.params, st.commitments, mins, seed))
    rc, masks = orc.verify_batch(list(case.transcripts), sts, proofs or case.proofs, action)
    return rc, masks


def check_against_oracle(case, params, action, **kw):
    trs, sts, prs = to_api(case, params, **kw)
    rc_o, masks_o = oracle_verify(case, action, **kw)
    if rc_o:
        with pytest.raises(bpp.pkg.EngineError) as ei:
            api.RangeProof.verify_batch(trs, sts, prs, action)
        assert ei.value.code == rc_o
        return rc_o, None
    got = api.RangeProof.verify_batch(trs, sts, prs, action)
    assert len(got) == len(masks_o)
    for g, o in zip(got, masks_o):
        assert (g is None) == (o is None)
        if g is not None:
            assert g.blindings() == o
    return 0, got


CONFIGS = [  # (bit_length, aggregation sizes, ext, promise strategy) -- tests/ristretto.rs:25-142
    (8, [1], 1, "none"), (64, [1], 1, "none"), (64, [1], 2, "third"), (64, [1], 3, "equal"),
    (4, [4], 1, "none"), (32, [4], 2, "third"), (32, [4], 3, "equal"),
    (64, [1, 1], 1, "third"), (64, [1, 2], 1, "none"), (64, [1, 2], 2, "third"), (64, [2, 1, 4], 3, "equal"),
]


@pytest.mark.parametrize("n,aggs,ext,promise", CONFIGS)
def test_prove_oracle_verify_gpu_matrix(n, aggs, ext, promise):
    case = workload.make_case(n, aggs, ext, promise=promise, same_blinding=True, rng_seed=8675309 + n + ext)
    params = gpu_params(n, max(aggs), ext)
    for action in (VA.VerifyOnly, VA.RecoverOnly, VA.RecoverAndVerify):
        rc, masks = check_against_oracle(case, params, action)
        assert rc == 0
        for st, w, mk in zip(case.statements, case.witnesses, masks):
            if action == VA.VerifyOnly or st.seed_nonce is None:
                assert mk is None
            else:
                assert mk.blindings() == w.blindings[0]        # the embedded mask is recovered (ristretto.rs:254-289)
    # wrong seed nonce: still verifies, masks differ (ristretto.rs:291-318)
    rc, masks = check_against_oracle(case, params, VA.RecoverAndVerify, seed_override=12345)
    assert rc == 0
    for st, w, mk in zip(case.statements, case.witnesses, masks):
        if st.seed_nonce is not None:
            assert mk.blindings() != w.blindings[0]
    # promises + 1 must fail (ristretto.rs:320-356)
    if promise != "none":
        rc, _ = check_against_oracle(case, params, VA.VerifyOnly, promise_delta=1)
        assert rc == orc.VERIFICATION_FAILED


def test_transcripts_advance_like_the_oracle():
    case = workload.make_case(64, [1, 2], 1, promise="third")
    params = gpu_params(64, 2, 1)
    trs, sts, prs = to_api(case, params)
    api.RangeProof.verify_batch(trs, sts, prs, VA.VerifyOnly)
    import ctypes as C

    tb = C.create_string_buffer(b"".join(case.transcripts), 203 * 2)
    sa = (orc.Statement * 2)(*[s.c for s in case.statements])
    pa = (orc.Proof * 2)(*case.proofs)
    nres = C.c_size_t()
    rc = orc.lib().orc_verify_batch(tb, 2, sa, 2, pa, 2, 2, C.create_string_buffer(256 * 6 * 32), C.create_string_buffer(256), C.byref(nres))
    assert rc == 0
    assert trs[0].state == tb.raw[:203] and trs[1].state == tb.raw[203:406]
    assert trs[0].state != case.transcripts[0]


def _mutations(pr, rnd):
    """systematic corruption of every proof field"""
    out = []
    for name in ("a", "a1", "b", "r1", "s1"):
        q = pr.copy()
        f = getattr(q, name)
        f[rnd.randrange(31)] ^= 1 << rnd.randrange(8)
        out.append((name + " bitflip", q))
    for arr, cnt in (("li", pr.n_li), ("ri", pr.n_ri), ("d1", pr.n_d1)):
        for j in range(cnt):
            q = pr.copy()
            getattr(q, arr)[j][rnd.randrange(31)] ^= 1 << rnd.randrange(8)
            out.append(("%s[%d] bitflip" % (arr, j), q))
    for name in ("a", "a1", "b"):
        q = pr.copy()
        for k in range(32):
            getattr(q, name)[k] = 0
        out.append((name + " identity encoding", q))
        q = pr.copy()
        getattr(q, name)[0] = 1
        for k in range(1, 32):
            getattr(q, name)[k] = 0
        out.append((name + " = [1,0..] non-decodable", q))         # range_proof.rs test_getters
    q = pr.copy()
    for k in range(32):
        q.li[0][k] = 0
    out.append(("li[0] identity", q))
    q = pr.copy()
    q.ri[pr.n_ri - 1][0] ^= 1   # odd -> negative s
    out.append(("ri[last] parity", q))
    # swap two rounds
    if pr.n_li >= 2:
        q = pr.copy()
        for k in range(32):
            q.li[0][k], q.li[1][k] = q.li[1][k], q.li[0][k]
        out.append(("swap li[0], li[1]", q))
    # r1 <-> s1
    q = pr.copy()
    for k in range(32):
        q.r1[k], q.s1[k] = q.s1[k], q.r1[k]
    out.append(("swap r1 s1", q))
    # drop / add a round (wrong length): range_proof.rs test_verify_errors
    q = pr.copy()
    q.n_li -= 1; q.n_ri -= 1
    out.append(("pop L and R", q))
    return out


@pytest.mark.parametrize("n,aggs,ext", [(64, [1, 1, 1], 1), (8, [2, 1], 2)])
def test_corrupted_proofs_same_verdict_as_oracle(n, aggs, ext):
    rnd = random.Random(n + ext)
    case = workload.make_case(n, aggs, ext, promise="third")
    params = gpu_params(n, max(aggs), ext)
    seen = set()
    for which in range(len(aggs)):
        for label, bad in _mutations(case.proofs[which], rnd):
            proofs = list(case.proofs)
            proofs[which] = bad
            data = orc.proof_to_bytes(bad)
            rc_o, _ = orc.verify_batch(list(case.transcripts), case.statements, proofs, orc.VERIFY_ONLY)
            trs, sts, prs = to_api(case, params)
            try:
                prs[which] = api.RangeProof.from_bytes(data)
                api.RangeProof.verify_batch(trs, sts, prs, VA.VerifyOnly)
                rc_g = 0
            except bpp.pkg.EngineError as e:
                rc_g = e.code
            if rc_o == 0 and label.startswith(("r1", "s1", "d1")) and orc.proof_from_bytes(data)[0] != 0:
                continue  # non-canonical scalar: rejected at parse time on both sides, not comparable through orc_proof structs
            assert rc_g == rc_o, (label, which, rc_g, rc_o)
            assert rc_o != 0, label
            seen.add(rc_o)
    assert {orc.VERIFICATION_FAILED, orc.INVALID_ARGUMENT, orc.INVALID_LENGTH} <= seen


def test_many_chunks_one_call_independent_verdicts():
    """K reference calls in one device pass: one bad proof only fails its own chunk; results equal per-call results"""
    case = workload.make_case(64, [1] * 12, 1, promise="third")
    params = gpu_params(64, 1, 1)
    trs, sts, prs = to_api(case, params)
    bad = case.proofs[7].copy()
    bad.r1[3] ^= 4
    prs[7] = api.RangeProof.from_bytes(orc.proof_to_bytes(bad))
    calls = [(trs[0:5], sts[0:5], prs[0:5]), (trs[5:9], sts[5:9], prs[5:9]), (trs[9:12], sts[9:12], prs[9:12])]
    status, masks = api.verify_chunks(params, calls, VA.RecoverAndVerify)
    assert status == [0, orc.VERIFICATION_FAILED, 0]
    for c, (lo, hi) in enumerate([(0, 5), (5, 9), (9, 12)]):
        for i, mk in zip(range(lo, hi), masks[c]):
            if status[c] == 0:
                assert mk.blindings() == case.witnesses[i].blindings[0]
            else:
                assert mk is None
    # split form gives the same answer and can be re-run
    trs2, sts2, prs2 = to_api(case, params)
    vb = api.VerifyBatch(params, [(trs2, sts2, prs2)], VA.VerifyOnly)
    assert vb.run()[0] == [0] and vb.run()[0] == [0]
    vb.close()


def test_argument_errors():
    case = workload.make_case(8, [1, 1], 1)
    params = gpu_params(8, 1, 1)
    trs, sts, prs = to_api(case, params)
    for args in [([], sts, prs), (trs, [], prs), (trs, sts, []), (trs[:1], sts, prs), (trs, sts[:1], prs), (trs, sts, prs[:1])]:
        with pytest.raises(bpp.pkg.EngineError) as ei:
            api.RangeProof.verify_batch(*args, VA.VerifyOnly)
        assert ei.value.code == orc.INVALID_ARGUMENT
    # extension degree of the proof differs from the parameters (range_proof.rs test_consistency_errors)
    case2 = workload.make_case(8, [1], 2)
    _, _, prs2 = to_api(case2, gpu_params(8, 1, 2))
    with pytest.raises(bpp.pkg.EngineError) as ei:
        api.RangeProof.verify_batch(trs[:1], sts[:1], prs2, VA.VerifyOnly)
    assert ei.value.code == orc.INVALID_ARGUMENT
    # promise exceeding the bit length -> InvalidLength (:675-682)
    st_bad = api.RangeStatement.init(params, sts[0].commitments, [1 << 8], None)
    with pytest.raises(bpp.pkg.EngineError) as ei:
        api.RangeProof.verify_batch(trs[:1], [st_bad], prs[:1], VA.VerifyOnly)
    assert ei.value.code == orc.INVALID_LENGTH
    # statement construction errors (range_statement.rs:42-61)
    c = sts[0].commitments[0]
    for commits, mins, seed in [([c] * 3, [None] * 3, None), ([c], [None, None], None), ([c] * 2, [None] * 2, None), ([c], [None], None)]:
        if len(commits) == 1 and len(mins) == 1:
            continue
        with pytest.raises(bpp.pkg.EngineError) as ei:
            api.RangeStatement.init(params, commits, mins, seed)
        assert ei.value.code == orc.INVALID_ARGUMENT
    p2 = gpu_params(8, 2, 1)
    with pytest.raises(bpp.pkg.EngineError):
        api.RangeStatement.init(p2, [c, c], [None, None], 5)


def test_aggregation_lower_than_generators():
    """m = 1 proof under M = 2 parameters: exercises the zero padding (range_proof.rs test_aggregation_lower_than_generators)"""
    case = workload.make_case(64, [1], 1, max_aggregation=2, promise="third")
    params = gpu_params(64, 2, 1)
    rc, masks = check_against_oracle(case, params, VA.RecoverAndVerify)
    assert rc == 0 and masks[0] is not None


def test_batch_truncates_to_256():
    """verify_batch only looks at the first 256 proofs (range_proof.rs:739-751)"""
    base = workload.make_case(8, [1] * 4, 1, promise="none")
    params = gpu_params(8, 1, 1)
    trs, sts, prs = to_api(base, params)
    n = 258
    T = [trs[i % 4].clone() for i in range(n)]
    S = [sts[i % 4] for i in range(n)]
    P = [prs[i % 4] for i in range(n)]
    bad = base.proofs[1].copy()
    bad.s1[0] ^= 2
    P[257] = api.RangeProof.from_bytes(orc.proof_to_bytes(bad))     # beyond the first 256: never looked at
    got = api.RangeProof.verify_batch(T, S, P, VA.VerifyOnly)
    assert len(got) == 256
    assert T[257].state == trs[1].state and T[0].state != trs[0].state
    P[255] = P[257]
    with pytest.raises(bpp.pkg.EngineError) as ei:
        api.RangeProof.verify_batch([t.clone() for t in T], S, P, VA.VerifyOnly)
    assert ei.value.code == orc.VERIFICATION_FAILED
