"""Round-2 boundary and front end, on the GPU, against the oracle and the independent pure-Python restatement (oracle/pyref.py):
  * bpp_vqueue: calls submitted from one thread, merged by the lanes into multi-call passes, give exactly the statuses, masks and
    advanced transcripts of every call alone (and of the oracle), valid and corrupted, whatever the lane count / pass size;
  * bpp_vbatch_create_multi: the merged pass directly;
  * bpp_verify_chunks_ch: the challenge-input form a stock-merlin host binds -- loop 1 and the weights made by pyref's Merlin
    (what src/transcripts.rs does on the Rust host) -- equals the state-passing form;
  * bpp_gens_create_with_bases: caller-made Pedersen bases, commitments / proofs / verdicts against pyref under the same bases;
  * statements built on different generator sets in one batch -> InvalidArgument (range_proof.rs:637-705);
  * proofs with 32..63 and >= 64 (L, R) pairs: InvalidLength / SizeOverflow where the reference raises them (:875-888)."""
import ctypes as C
import os
import sys

import pytest

import bpp
import orc
import workload

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import pyref as R  # noqa: E402

pytestmark = pytest.mark.gpu
api = bpp.pkg.api
ffi = bpp.ffi


def _calls(params, case, lo, hi, proofs=None):
    sts = [api.RangeStatement.init(params, s.commitments, s.min_values, s.seed_nonce) for s in case.statements[lo:hi]]
    prs = [api.RangeProof.from_bytes(orc.proof_to_bytes(p)) for p in (proofs or case.proofs[lo:hi])]
    trs = [api.Transcript(state=t) for t in case.transcripts[lo:hi]]
    return (trs, sts, prs)


def _orc_advanced(transcripts, statements, proofs, action):
    """(rc, masks, advanced transcripts) of the oracle: `&mut [Transcript]` after RangeProof::verify_batch"""
    n = len(statements)
    tb = C.create_string_buffer(b"".join(transcripts), 203 * n)
    sa = (orc.Statement * n)(*[s.c for s in statements])
    pa = (orc.Proof * n)(*proofs)
    nres = C.c_size_t()
    orc.lib().orc_verify_batch(tb, n, sa, n, pa, n, action, C.create_string_buffer(256 * 6 * 32), C.create_string_buffer(256), C.byref(nres))
    rc, masks = orc.verify_batch(list(transcripts), statements, proofs, action)
    return rc, masks, [tb.raw[203 * i:203 * (i + 1)] for i in range(n)]


def _make_batches(case, n_batches, per, bad):
    """n_batches calls of `per` proofs each; bad: {batch: index of the proof to corrupt}"""
    out = []
    for bi in range(n_batches):
        lo, hi = per * bi, per * bi + per
        proofs = [p.copy() for p in case.proofs[lo:hi]]
        if bi in bad:
            proofs[bad[bi]].r1[3] ^= 0x10
        rc, masks, ts = _orc_advanced(case.transcripts[lo:hi], case.statements[lo:hi], proofs, orc.RECOVER_AND_VERIFY)
        out.append((lo, hi, proofs, rc, masks, ts))
    return out


@pytest.mark.parametrize("lanes,max_calls", [(1, 1), (1, 4), (2, 3), (3, 16)])
def test_queue_matches_single_calls_and_oracle(lanes, max_calls):
    case = workload.make_case(64, [1] * 16 + [2, 4] + [1] * 6, 1, max_aggregation=4, promise="third", rng_seed=4711)
    specs = _make_batches(case, 12, 2, {3: 0, 7: 1, 8: 0})
    q = api.VerifyQueue(0, 64, 4, 1, lanes=lanes, max_calls_per_pass=max_calls)
    eng = bpp.pkg.Engine(0)
    params = api.RangeParameters.init(eng, 64, 4, 1)
    try:
        batches = [[_calls(params, case, lo, hi, proofs)] for lo, hi, proofs, _, _, _ in specs]
        got = q.verify_many(batches, api.VerifyAction.RecoverAndVerify)
        st = q.stats()
        assert st["calls"] == 12 and st["kernels"] > 0 and st["passes"] <= 12
        for bi, ((status, masks), (lo, hi, proofs, rc, want, ts)) in enumerate(zip(got, specs)):
            assert status == [rc], (bi, status, rc)
            alone = api.verify_chunks(params, [_calls(params, case, lo, hi, proofs)], api.VerifyAction.RecoverAndVerify)
            assert alone[0] == status
            if rc == 0:
                for g, a, w in zip(masks[0], alone[1][0], want):
                    assert (g is None) == (w is None) and (g is None or (g.blindings() == w and g == a))
            # `&mut Transcript`: advanced exactly as the oracle leaves them
            for t, o in zip(batches[bi][0][0], ts):
                assert t.state == o
    finally:
        q.close()
        eng.close()


def test_page_locked_proof_bytes_are_read_in_place():
    """proof bytes in page-locked memory (bpp_host_alloc) are uploaded by the copy engine from where they are, pageable ones are staged;
    one pass mixes both kinds (staged runs and direct copies interleave) and must give what separate staged calls give"""
    case = workload.make_case(64, [1] * 16 + [2, 4] + [1] * 6, 1, max_aggregation=4, promise="third", rng_seed=4711)
    specs = _make_batches(case, 12, 2, {3: 0, 7: 1, 8: 0})
    q = api.VerifyQueue(0, 64, 4, 1, lanes=1, max_calls_per_pass=16)
    try:
        packed = [q.pack([_calls(q.shape, case, lo, hi, proofs)], api.VerifyAction.RecoverAndVerify, pinned=(i % 3 != 1))
                  for i, (lo, hi, proofs, _, _, _) in enumerate(specs)]
        tickets = [q.submit(pk) for pk in packed]
        for t in tickets:
            q.wait(t)
        for pk, (lo, hi, proofs, rc, want, ts) in zip(packed, specs):
            status, masks = pk.results()
            assert status == [rc]
            if rc == 0:
                for g, w in zip(masks[0], want):
                    assert (g is None) == (w is None) and (g is None or g.blindings() == w)
            for t, o in zip(pk.transcripts, ts):
                assert t.state == o
        # and through the plain entry point
        pk = q.pack([_calls(q.shape, case, 0, 6)], api.VerifyAction.VerifyOnly, pinned=True)
        eng = bpp.pkg.Engine(0)
        params = api.RangeParameters.init(eng, 64, 4, 1)
        assert ffi.lib().bpp_verify_chunks(params.gens.h, C.byref(pk.args), pk.status, pk.masks, pk.mask_present) == 0
        assert list(pk.status)[:1] == [0]
        eng.close()
    finally:
        q.close()


@pytest.mark.parametrize("device_weights", [False, True])
def test_queue_merged_check_gives_the_per_call_results(device_weights):
    """merged check: one multiscalar check per device pass; a pass that holds an invalid proof is settled call by call -- statuses,
    masks and transcripts are those of separate calls either way"""
    case = workload.make_case(64, [1] * 16 + [2, 4] + [1] * 6, 1, max_aggregation=4, promise="third", rng_seed=4711)
    for bad in ({}, {3: 0, 7: 1, 8: 0}):
        specs = _make_batches(case, 12, 2, bad)
        q = api.VerifyQueue(0, 64, 4, 1, lanes=1, max_calls_per_pass=16, merged_check=True, device_weights=device_weights)
        try:
            packed = [q.pack([_calls(q.shape, case, lo, hi, proofs)], api.VerifyAction.RecoverAndVerify) for lo, hi, proofs, _, _, _ in specs]
            tickets = [q.submit(pk) for pk in packed]
            for t in tickets:
                q.wait(t)
            for pk, (lo, hi, proofs, rc, want, ts) in zip(packed, specs):
                status, masks = pk.results()
                assert status == [rc]
                if rc == 0:
                    for g, w in zip(masks[0], want):
                        assert (g is None) == (w is None) and (g is None or g.blindings() == w)
                for t, o in zip(pk.transcripts, ts):
                    assert t.state == o
        finally:
            q.close()


def test_queue_mixed_actions_and_many_threads():
    import threading

    case = workload.make_case(64, [1] * 24, 1, promise="third", rng_seed=99)
    q = api.VerifyQueue(0, 64, 1, 1, lanes=2, max_calls_per_pass=8)
    results = {}
    try:
        def worker(i):
            lo, hi = 2 * i, 2 * i + 2
            action = api.VerifyAction.VerifyOnly if i % 3 else api.VerifyAction.RecoverAndVerify
            pk = q.pack([_calls(q.shape, case, lo, hi)], action)
            rc = ffi.lib().bpp_vqueue_verify(q.h, C.byref(pk.args), pk.status, pk.masks, pk.mask_present)
            results[i] = (rc, pk.results(), action)
        ths = [threading.Thread(target=worker, args=(i,)) for i in range(12)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        for i in range(12):
            rc, (status, masks), action = results[i]
            assert rc == 0 and status == [0]
            orc_rc, want = orc.verify_batch(list(case.transcripts[2 * i:2 * i + 2]), case.statements[2 * i:2 * i + 2], case.proofs[2 * i:2 * i + 2],
                                            orc.RECOVER_AND_VERIFY if action == api.VerifyAction.RecoverAndVerify else orc.VERIFY_ONLY)
            assert orc_rc == 0
            for g, w in zip(masks[0], want):
                assert (g is None) == (w is None) and (g is None or g.blindings() == w)
    finally:
        q.close()


def test_transcripts_in_different_states_within_one_call():
    """calls whose transcripts all hold the same state upload ONE state (the usual case: one label); a call with per-proof labels uploads
    them all -- both in one pass, every proof replayed from its own state"""
    labels = [b"label-%d" % (i % 3) for i in range(8)]
    mixed = workload.make_case(16, [1] * 8, 1, promise="third", rng_seed=12, labels=labels)
    same = workload.make_case(16, [1] * 8, 1, promise="third", rng_seed=13)
    eng = bpp.pkg.Engine(0)
    params = api.RangeParameters.init(eng, 16, 1, 1)
    q = api.VerifyQueue(0, 16, 1, 1, lanes=1, max_calls_per_pass=4)
    try:
        assert len(set(mixed.transcripts)) == 3 and len(set(same.transcripts)) == 1
        batches = [[_calls(params, mixed, 0, 8)], [_calls(params, same, 0, 8)], [_calls(params, mixed, 2, 6)]]
        got = q.verify_many(batches, api.VerifyAction.RecoverAndVerify)
        for (status, masks), (case, lo, hi), b in zip(got, ((mixed, 0, 8), (same, 0, 8), (mixed, 2, 6)), batches):
            rc, want, ts = _orc_advanced(case.transcripts[lo:hi], case.statements[lo:hi], case.proofs[lo:hi], orc.RECOVER_AND_VERIFY)
            assert rc == 0 and status == [0]
            assert [m.blindings() for m in masks[0]] == want
            assert [t.state for t in b[0][0]] == ts
        # a proof replayed from the wrong state does not verify
        trs, sts, prs = _calls(params, mixed, 0, 8)
        trs[1], trs[2] = trs[2], trs[1]
        status, _ = api.verify_chunks(params, [(trs, sts, prs)], api.VerifyAction.VerifyOnly)
        assert status == [orc.VERIFICATION_FAILED]
    finally:
        q.close()
        eng.close()


def test_create_multi_equals_separate_passes():
    case = workload.make_case(32, [1, 2, 1, 4, 1, 1, 2, 1], 2, max_aggregation=4, promise="third", rng_seed=31)
    eng = bpp.pkg.Engine(0)
    params = api.RangeParameters.init(eng, 32, 4, 2)
    try:
        bad = [p.copy() for p in case.proofs]
        bad[4].d1[1][0] ^= 2
        groups = [(0, 3), (3, 5), (5, 8)]
        pks = [api._Packed(params, [_calls(params, case, lo, hi, bad[lo:hi])], api.VerifyAction.RecoverAndVerify) for lo, hi in groups]
        ptrs = (C.c_void_p * 3)(*[C.addressof(pk.args) for pk in pks])
        vb = C.c_void_p()
        assert ffi.lib().bpp_vbatch_create_multi(params.gens.h, 3, ptrs, C.byref(vb)) == 0
        assert ffi.lib().bpp_vbatch_call_count(vb) == 3
        st = (C.c_void_p * 3)(*[C.addressof(pk.status) for pk in pks])
        mk = (C.c_void_p * 3)(*[C.addressof(pk.masks) for pk in pks])
        mp = (C.c_void_p * 3)(*[C.addressof(pk.mask_present) for pk in pks])
        assert ffi.lib().bpp_vbatch_run_multi(vb, st, mk, mp) == 0
        for i, pk in enumerate(pks):
            assert ffi.lib().bpp_vbatch_transcripts_call(vb, i, C.addressof(pk.tbuf)) == 0
        ffi.lib().bpp_vbatch_destroy(vb)
        for (lo, hi), pk in zip(groups, pks):
            status, masks = pk.results()
            rc, want, ts = _orc_advanced(case.transcripts[lo:hi], case.statements[lo:hi], bad[lo:hi], orc.RECOVER_AND_VERIFY)
            assert status == [rc]
            assert rc == (orc.VERIFICATION_FAILED if lo <= 4 < hi else 0)
            if rc == 0:
                for g, w in zip(masks[0], want):
                    assert (g is None) == (w is None) and (g is None or g.blindings() == w)
            assert [t.state for t in pk.transcripts] == ts
    finally:
        eng.close()


def _pyref_loop1(params_r, statements_r, proofs_r, label):
    """what a stock-merlin Rust host does before bpp_verify_chunks_ch: src/transcripts.rs over merlin (pyref's pure-Python STROBE),
    the verifier-weight transcript and the weight draws (range_proof.rs:811-853, :894)"""
    weight_t = R.Transcript(b"Bulletproofs+ verifier weights")
    chal = []
    for st, pr in zip(statements_r, proofs_r):
        rpt = R.RangeProofTranscript(R.Transcript(label), params_r, st, None, R.NullRng())
        y, z = rpt.challenges_y_z(pr.a)
        ej = [rpt.challenge_round_e(l, r) for l, r in zip(pr.li, pr.ri)]
        e = rpt.challenge_final_e(pr.a1, pr.b)
        chal.append([y, z, e] + ej)
        weight_t.append_message(b"proof", rpt.to_verifier_rng(pr.r1, pr.s1, pr.d1).fill_bytes(32))
    wrng = weight_t.build_rng().finalize(R.NullRng())
    return chal, [R.random_not_zero(wrng) for _ in proofs_r]


def test_challenge_input_form_equals_state_passing_form():
    case = workload.make_case(16, [1, 2, 1, 1], 2, max_aggregation=2, promise="third", rng_seed=2024)
    eng = bpp.pkg.Engine(0)
    params = api.RangeParameters.init(eng, 16, 2, 2)
    try:
        prm_r = R.Params(16, 2, 2)
        sts_r = [R.Statement(prm_r, [R.decode(c) for c in s.commitments], s.min_values, s.seed_nonce) for s in case.statements]
        for corrupt in (None, 2):
            proofs = [p.copy() for p in case.proofs]
            if corrupt is not None:
                proofs[corrupt].s1[7] ^= 4
            prs_r = [R.Proof.from_bytes(orc.proof_to_bytes(p)) for p in proofs]
            chal, weights = _pyref_loop1(prm_r, sts_r, prs_r, workload.LABEL)
            want_status, want_masks = api.verify_chunks(params, [_calls(params, case, 0, 4, proofs)], api.VerifyAction.RecoverAndVerify)
            got_status, got_masks = api.verify_chunks_ch(params, [_calls(params, case, 0, 4, proofs)], chal, weights, api.VerifyAction.RecoverAndVerify)
            assert got_status == want_status == [orc.VERIFICATION_FAILED if corrupt is not None else 0]
            assert got_masks == want_masks
            # a wrong weight is the caller's problem, a wrong challenge must break the check
            if corrupt is None:
                chal[1][0] = (chal[1][0] + 1) % R.L
                st2, _ = api.verify_chunks_ch(params, [_calls(params, case, 0, 4, proofs)], chal, weights, api.VerifyAction.VerifyOnly)
                assert st2 == [orc.VERIFICATION_FAILED]
    finally:
        eng.close()


def test_custom_pedersen_bases_against_pyref():
    """RangeParameters::init with caller-made PedersenGens (range_parameters.rs:32-58): commitments, device-made proofs and
    verdicts under bases that are NOT the reference's constants, against pyref with the same bases"""
    import hashlib

    n, ext = 8, 2
    prm = R.Params(n, 1, ext)
    prm.h = R.pt_mul(5, R.BASEPOINT)
    prm.g = [R.from_uniform_bytes(hashlib.sha3_512(b"custom masking base %d" % k).digest()) for k in range(ext)]
    prm.h_c, prm.g_c = R.encode(prm.h), [R.encode(g) for g in prm.g]
    eng = bpp.pkg.Engine(0)
    gens = bpp.pkg.Gens(eng, n, 1, ext, h_base=prm.h_c, g_bases=prm.g_c)
    params = api.RangeParameters(gens)
    try:
        assert params.h_base() == prm.h_c and params.g_bases() == prm.g_c
        rng = R.ChaCha12Rng.seed_from_u64(5)
        proofs, sts_r, sts = [], [], []
        for i in range(3):
            v = rng.next_u64() % 128
            bl = [R.random_not_zero(rng) for _ in range(ext)]
            c = prm.commit(v, bl)
            assert gens.commit_batch([v], [bl])[0] == R.encode(c)
            seed = R.random_not_zero(rng)
            st_r = R.Statement(prm, [c], [v // 3], seed)
            stream = hashlib.shake_256(b"custom-%d" % i).digest(32 * 8)
            pr = R.prove_with_rng(R.Transcript(b"custom bases"), st_r, [v], [bl], R.BufferRng(stream))
            st = api.RangeStatement.init(params, [R.encode(c)], [v // 3], seed)
            dev = api.RangeProof.prove_batch([api.Transcript(b"custom bases")], [st], [api.RangeWitness.init([api.CommitmentOpening(v, bl)])], [stream])[0]
            assert dev.to_bytes() == pr.to_bytes()          # byte-identical proof under the custom bases
            proofs.append((pr, bl))
            sts_r.append(st_r)
            sts.append(st)
        want = R.verify_batch([R.Transcript(b"custom bases") for _ in proofs], sts_r, [p for p, _ in proofs], R.RECOVER_AND_VERIFY)
        got = api.RangeProof.verify_batch([api.Transcript(b"custom bases") for _ in proofs], sts,
                                          [api.RangeProof.from_bytes(p.to_bytes()) for p, _ in proofs], api.VerifyAction.RecoverAndVerify)
        assert [g.blindings() for g in got] == want == [bl for _, bl in proofs]
        # the same proofs under the STANDARD bases do not verify
        std = api.RangeParameters.init(eng, n, 1, ext)
        sts2 = [api.RangeStatement.init(std, s.commitments, s.minimum_value_promises, s.seed_nonce) for s in sts]
        with pytest.raises(bpp.pkg.EngineError) as e:
            api.RangeProof.verify_batch([api.Transcript(b"custom bases") for _ in proofs], sts2,
                                        [api.RangeProof.from_bytes(p.to_bytes()) for p, _ in proofs], api.VerifyAction.VerifyOnly)
        assert e.value.code == orc.VERIFICATION_FAILED
        # statements of one batch on different generator sets: InvalidArgument (range_proof.rs:637-705, test :1438-1620)
        with pytest.raises(bpp.pkg.EngineError) as e:
            api.RangeProof.verify_batch([api.Transcript(b"custom bases") for _ in proofs], [sts[0], sts2[1], sts[2]],
                                        [api.RangeProof.from_bytes(p.to_bytes()) for p, _ in proofs], api.VerifyAction.VerifyOnly)
        assert e.value.code == orc.INVALID_ARGUMENT
        # an encoding that does not decode is rejected at construction
        with pytest.raises(bpp.pkg.EngineError) as e:
            bpp.pkg.Gens(eng, n, 1, ext, h_base=bytes([1]) + bytes(31))
        assert e.value.code == orc.INVALID_ARGUMENT
    finally:
        eng.close()


def test_oversized_round_counts_fail_where_the_reference_fails():
    """from_bytes accepts any number of (L, R) pairs; verify raises InvalidLength for 2^rounds != n*m and SizeOverflow from 64 rounds
    on (range_proof.rs:875-888), after the transcript replay and the decompression of every point.  pyref is the checker (the C
    oracle's proof struct stops at 32 pairs)."""
    case = workload.make_case(8, [1], 1, promise="third", rng_seed=8)
    eng = bpp.pkg.Engine(0)
    params = api.RangeParameters.init(eng, 8, 1, 1)
    prm_r = R.Params(8, 1, 1)
    st_r = R.Statement(prm_r, [R.decode(c) for c in case.statements[0].commitments], case.statements[0].min_values, case.statements[0].seed_nonce)
    try:
        good = orc.proof_to_bytes(case.proofs[0])
        pair = good[-64:]
        for pairs, variant, code in ((33, "InvalidLength", orc.INVALID_LENGTH), (64, "SizeOverflow", orc.SIZE_OVERFLOW), (2, "InvalidLength", orc.INVALID_LENGTH)):
            raw = good[: 1 + 32 * 6] + pair * pairs
            with pytest.raises(R.ProofError) as e:
                R.verify_batch([R.Transcript(workload.LABEL)], [st_r], [R.Proof.from_bytes(raw)], R.VERIFY_ONLY)
            assert e.value.variant == variant
            trs, sts, _ = _calls(params, case, 0, 1)
            status, _ = api.verify_chunks(params, [(trs, sts, [api.RangeProof.from_bytes(raw)])], api.VerifyAction.VerifyOnly)
            assert status == [code], (pairs, status)
        # an undecodable L inside an oversized proof wins over the length error (decompression comes first, :859-866)
        raw = bytearray(good[: 1 + 32 * 6] + pair * 33)
        raw[1 + 32 * 6 + 64 * 20] = 1
        raw[1 + 32 * 6 + 64 * 20 + 1: 1 + 32 * 6 + 64 * 20 + 32] = bytes(31)
        trs, sts, _ = _calls(params, case, 0, 1)
        status, _ = api.verify_chunks(params, [(trs, sts, [api.RangeProof.from_bytes(bytes(raw))])], api.VerifyAction.VerifyOnly)
        assert status == [orc.INVALID_ARGUMENT]
    finally:
        eng.close()
