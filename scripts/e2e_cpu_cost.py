#!/usr/bin/env python
"""host CPU time (user + system, all threads of the process) per 1024-proof verification call in blocking-wait mode, one lane, one
host thread: what a pass costs the host when the calling thread sleeps while the device works"""
import ctypes as C, os, resource, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, bpp, orc
api, lib = bpp.pkg.api, bpp.ffi.lib()
_, cases = bench.make_workload(1024)
eng = bpp.pkg.Engine(0)
eng.set_host_threads(1)
eng.set_throughput_mode(1)
params = api.RangeParameters.init(eng, 64, 1, 1)
def calls():
    out = []
    for c in cases:
        sts = [api.RangeStatement.init(params, s.commitments, s.min_values, s.seed_nonce) for s in c.statements]
        prs = [api.RangeProof.from_bytes(orc.proof_to_bytes(p)) for p in c.proofs]
        out.append(([api.Transcript(state=t) for t in c.transcripts], sts, prs))
    return out
vb = api.VerifyBatch(params, calls(), api.VerifyAction.VerifyOnly)
pk = api._Packed(params, calls(), api.VerifyAction.VerifyOnly)
t_init = bytes(pk.tbuf.raw)
def cpu():
    r = resource.getrusage(resource.RUSAGE_SELF)
    return r.ru_utime + r.ru_stime
N = 300
for name, fn in (("device-resident (bpp_vbatch_run)", lambda: lib.bpp_vbatch_run(vb.h, vb.pk.status, vb.pk.masks, vb.pk.mask_present)),
                 ("end to end (bpp_verify_chunks)", lambda: (C.memmove(pk.tbuf, t_init, len(t_init)), lib.bpp_verify_chunks(params.gens.h, C.byref(pk.args), pk.status, pk.masks, pk.mask_present))[1])):
    for _ in range(20): assert fn() == 0
    c0, t0 = cpu(), time.perf_counter()
    for _ in range(N): assert fn() == 0
    c1, t1 = cpu(), time.perf_counter()
    print("%-36s wall %.3f ms/call   host CPU %.3f ms/call" % (name, (t1 - t0) / N * 1e3, (c1 - c0) / N * 1e3))
    print("   host phases (last call):", {k: round(v, 3) for k, v in eng.host_ms().items()})
