#!/usr/bin/env python
"""How much of the GPU does one 1024-proof verification step leave idle?  Runs S independent bpp_ctx (one CUDA stream pair each,
one host thread each) that verify the same 1024-proof workload concurrently and prints the aggregate proofs/s for
S = 1, 2, 4, 8 ..., device-resident (bpp_vbatch_run) and end to end (bpp_verify_chunks with host buffers), for both device
replay kernels, with the host CPU time (user + system, all threads) the process spent per step.  usage: pipeline_probe.py
[steps_per_stream] [S ...]; PROBE_ONE_MODE=1: only the default replay kernel; run under `taskset -c 0-3` to see what 4 host cores
per GPU (the 8-GPU box) can feed."""
import ctypes as C
import os
import resource
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import bpp  # noqa: E402
import orc  # noqa: E402

api = bpp.pkg.api
lib = bpp.ffi.lib()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
S_list = [int(x) for x in sys.argv[2:]] or [1, 2, 4, 8]
N = int(os.environ.get("PROBE_PROOFS", "1024"))
_, cases = bench.make_workload(N)
cores = os.cpu_count() or 1


class Lane:
    def __init__(self, host_threads):
        self.eng = bpp.pkg.Engine(0)
        self.eng.set_throughput_mode(int(os.environ.get("PROBE_TPUT", "1")))
        lib.bpp_ctx_set_host_threads(self.eng.h, host_threads)
        self.params = api.RangeParameters.init(self.eng, 64, 1, 1)
        self.vb = api.VerifyBatch(self.params, self.calls(), api.VerifyAction.VerifyOnly)
        self.pk = api._Packed(self.params, self.calls(), api.VerifyAction.VerifyOnly)
        self.t_init = bytes(self.pk.tbuf.raw)

    def calls(self):
        out = []
        for c in cases:
            sts = [api.RangeStatement.init(self.params, s.commitments, s.min_values, s.seed_nonce) for s in c.statements]
            prs = [api.RangeProof.from_bytes(orc.proof_to_bytes(p)) for p in c.proofs]
            trs = [api.Transcript(state=t) for t in c.transcripts]
            out.append((trs, sts, prs))
        return out

    def dev_steps(self, k):
        for _ in range(k):
            rc = lib.bpp_vbatch_run(self.vb.h, self.vb.pk.status, self.vb.pk.masks, self.vb.pk.mask_present)
            assert rc == 0 and all(self.vb.pk.status[c] == 0 for c in range(len(cases)))

    def e2e_steps(self, k):
        pk = self.pk
        for _ in range(k):
            C.memmove(pk.tbuf, self.t_init, len(self.t_init))
            rc = lib.bpp_verify_chunks(self.params.gens.h, C.byref(pk.args), pk.status, pk.masks, pk.mask_present)
            assert rc == 0 and all(pk.status[c] == 0 for c in range(pk.k))


def timed(lanes, fn_name, k):
    ths = [threading.Thread(target=getattr(ln, fn_name), args=(k,)) for ln in lanes]
    r0 = resource.getrusage(resource.RUSAGE_SELF)
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    r1 = resource.getrusage(resource.RUSAGE_SELF)
    timed.cpu_s = (r1.ru_utime + r1.ru_stime) - (r0.ru_utime + r0.ru_stime)
    return dt


for S in S_list:
    lanes = [Lane(max(1, cores // S)) for _ in range(S)]
    for mode, name in (((1, "default"),) if os.environ.get("PROBE_ONE_MODE") else ((2, "thread/proof"), (3, "warp/proof"))):
        for ln in lanes:
            ln.eng.set_replay_mode(mode)
        timed(lanes, "dev_steps", 3)
        td = timed(lanes, "dev_steps", steps)
        cd = timed.cpu_s
        timed(lanes, "e2e_steps", 3)
        te = timed(lanes, "e2e_steps", steps)
        ce = timed.cpu_s
        print("S=%d replay=%-12s device-resident %9.0f proofs/s (%.3f ms/step/stream, host CPU %.3f ms/step)   e2e %9.0f proofs/s (%.3f ms/step/stream, host CPU %.3f ms/step)" % (
            S, name, S * steps * N / td, 1e3 * td / steps, 1e3 * cd / (S * steps), S * steps * N / te, 1e3 * te / steps, 1e3 * ce / (S * steps)), flush=True)
    for ln in lanes:
        ln.eng.close()          # closes the lane's batches and generator tables first
