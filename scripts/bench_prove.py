"""proving throughput: P non-aggregated 64-bit proofs per bpp_prove_batch call (wall clock, host Fiat-Shamir included)"""
import hashlib, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bpp, orc
api = bpp.pkg.api
eng = bpp.pkg.Engine(0)
n, m, ext = 64, int(sys.argv[2]) if len(sys.argv) > 2 else 1, 1
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
gp = api.RangeParameters.init(eng, n, m, ext)
op = orc.Params(n, m, ext)
rng = orc.Rng("chacha", 99)
vals = [[rng.next_u64() % (1 << 63) for _ in range(m)] for _ in range(P)]
blinds = [[[rng.random_not_zero()] for _ in range(m)] for _ in range(P)]
t0 = time.time()
commits = gp.gens.commit_batch([v for vs in vals for v in vs], [b for bs in blinds for b in bs])
print("commit %d openings on the device: %.1f ms" % (P * m, (time.time() - t0) * 1e3))
commits = [commits[i * m:(i + 1) * m] for i in range(P)]
sts = [api.RangeStatement.init(gp, commits[i], [v // 3 for v in vals[i]], (rng.random_not_zero() if m == 1 else None)) for i in range(P)]
wits = [api.RangeWitness.init([api.CommitmentOpening(v, b) for v, b in zip(vals[i], blinds[i])]) for i in range(P)]
need = api.RangeProof.rng_bytes_needed(gp, m)
streams = [hashlib.shake_256(b"s%d" % i).digest(need) for i in range(P)]
best = 1e9
for it in range(4):
    trs = [api.Transcript(b"BatchedRangeProofTest") for _ in range(P)]
    t0 = time.perf_counter()
    proofs = api.RangeProof.prove_batch(trs, sts, wits, streams)
    dt = time.perf_counter() - t0
    best = min(best, dt)
    print("prove_batch P=%d m=%d: %.2f ms  (%.0f proofs/s); C-ABI call alone %.2f ms (%.0f proofs/s)" % (
        P, m, dt * 1e3, P / dt, api.RangeProof.last_prove_call_ms, P / api.RangeProof.last_prove_call_ms * 1e3))
assert not any(isinstance(p, Exception) for p in proofs)
# check one against the oracle and verify all
st0 = orc.St(op, commits[0], [v // 3 for v in vals[0]], sts[0].seed_nonce)
rc, pr, _ = orc.prove(orc.transcript_new(b"BatchedRangeProofTest"), st0, orc.Wit(vals[0], blinds[0]), orc.Rng("buffer", data=streams[0]))
assert rc == 0 and orc.proof_to_bytes(pr) == proofs[0].to_bytes()
t0 = time.perf_counter()
for lo in range(0, P, 256):
    api.RangeProof.verify_batch([api.Transcript(b"BatchedRangeProofTest") for _ in range(lo, min(P, lo + 256))], sts[lo:lo + 256], proofs[lo:lo + 256], api.VerifyAction.VerifyOnly)
print("verified all %d proofs (python API, %d calls): %.1f ms; best prove %.0f proofs/s" % (P, (P + 255) // 256, (time.perf_counter() - t0) * 1e3, P / best))
