"""host-side timing breakdown of bpp_vbatch_create / run / destroy (wall clock), 1024 proofs"""
import ctypes as C, sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bpp, orc, bench
api = bpp.pkg.api
eng = bpp.pkg.Engine(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
params_o, cases = bench.make_workload(n)
params = api.RangeParameters.init(eng, 64, 1, 1)
calls = []
for c in cases:
    sts = [api.RangeStatement.init(params, s.commitments, s.min_values, s.seed_nonce) for s in c.statements]
    prs = [api.RangeProof.from_bytes(orc.proof_to_bytes(p)) for p in c.proofs]
    trs = [api.Transcript(state=t) for t in c.transcripts]
    calls.append((trs, sts, prs))
pk = api._Packed(params, calls, api.VerifyAction.VerifyOnly)
t_init = bytes(pk.tbuf.raw)
lib = bpp.ffi.lib()
for threads in (16, 8, 4, 1):
    eng.set_host_threads(threads)
    tc = tr = td = 0.0
    for it in range(8):
        C.memmove(pk.tbuf, t_init, len(t_init))
        h = C.c_void_p()
        t0 = time.perf_counter()
        rc = lib.bpp_vbatch_create(params.gens.h, C.byref(pk.args), C.byref(h)); assert rc == 0
        t1 = time.perf_counter()
        rc = lib.bpp_vbatch_run(h, pk.status, pk.masks, pk.mask_present); assert rc == 0
        t2 = time.perf_counter()
        lib.bpp_vbatch_destroy(h)
        t3 = time.perf_counter()
        if it >= 3:
            tc += t1 - t0; tr += t2 - t1; td += t3 - t2
    print("   host phases (last call):", {k: round(v, 3) for k, v in eng.host_ms().items()})
    print("threads=%2d  create %.3f ms  run %.3f ms  destroy %.3f ms" % (threads, tc / 5 * 1e3, tr / 5 * 1e3, td / 5 * 1e3))
