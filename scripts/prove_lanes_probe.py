#!/usr/bin/env python
"""Prover lanes: P 64-bit proofs proved as S concurrent bpp_prove_batch calls of P/S proofs (one bpp_ctx + host thread each).
Times the C-ABI calls only (arguments packed beforehand).  usage: prove_lanes_probe.py [P] [S ...]"""
import ctypes as C
import hashlib
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bpp  # noqa: E402
import orc  # noqa: E402

api, ffi = bpp.pkg.api, bpp.ffi
lib = ffi.lib()
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
S_list = [int(x) for x in sys.argv[2:]] or [1, 2, 4]
n, m, ext = 64, 1, 1
cores = os.cpu_count() or 1
rng = orc.Rng("chacha", 99)
vals = [rng.next_u64() % (1 << 63) for _ in range(P)]
blinds = [rng.random_not_zero() for _ in range(P)]
seeds = [rng.random_not_zero() for _ in range(P)]


class Lane:
    def __init__(self, lo, hi, host_threads):
        self.eng = bpp.pkg.Engine(0)
        self.eng.set_host_threads(host_threads)
        self.gp = api.RangeParameters.init(self.eng, n, m, ext)
        k = hi - lo
        commits = self.gp.gens.commit_batch(vals[lo:hi], [[b] for b in blinds[lo:hi]])
        need = api.RangeProof.rng_bytes_needed(self.gp, m)
        self.bufs = dict(
            commits=C.create_string_buffer(b"".join(commits), 32 * k),
            values=(C.c_uint64 * k)(*vals[lo:hi]),
            blind=C.create_string_buffer(b"".join(int(b).to_bytes(32, "little") for b in blinds[lo:hi]), 32 * k),
            minv=(C.c_uint64 * k)(*[v // 3 for v in vals[lo:hi]]),
            minp=(C.c_uint8 * k)(*([1] * k)),
            seeds=C.create_string_buffer(b"".join(int(s).to_bytes(32, "little") for s in seeds[lo:hi]), 32 * k),
            seedp=(C.c_uint8 * k)(*([1] * k)),
            rbuf=C.create_string_buffer(b"".join(hashlib.shake_256(b"s%d" % i).digest(need) for i in range(lo, hi)), need * k))
        self.t0 = api.Transcript(b"BatchedRangeProofTest").state * k
        self.tbuf = C.create_string_buffer(self.t0, len(self.t0))
        b = self.bufs
        self.args = ffi.ProveArgs(k, m, C.addressof(b["commits"]), C.addressof(b["values"]), C.addressof(b["blind"]), C.addressof(b["minv"]),
                                  C.addressof(b["minp"]), C.addressof(b["seeds"]), C.addressof(b["seedp"]), C.addressof(self.tbuf), C.addressof(b["rbuf"]), need)
        self.plen = lib.bpp_proof_size(ext, 6)
        self.out = C.create_string_buffer(self.plen * k)
        self.status = (C.c_int32 * k)()
        self.k = k

    def run(self, reps):
        for _ in range(reps):
            C.memmove(self.tbuf, self.t0, len(self.t0))
            rc = lib.bpp_prove_batch(self.gp.gens.h, C.byref(self.args), self.out, self.plen, self.status)
            assert rc == 0 and not any(self.status)


ref = None
for S in S_list:
    per = P // S
    lanes = [Lane(i * per, (i + 1) * per, max(1, cores // S)) for i in range(S)]

    def go(reps):
        ths = [threading.Thread(target=ln.run, args=(reps,)) for ln in lanes]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        return time.perf_counter() - t0

    go(2)
    dt = go(4) / 4
    proofs = b"".join(ln.out.raw for ln in lanes)
    if ref is None:
        ref = proofs
    assert proofs == ref, "proof bytes depend on the lane count"
    print("P=%d as %d concurrent calls of %d: %.2f ms  %.0f proofs/s" % (P, S, per, dt * 1e3, P / dt), flush=True)
    for ln in lanes:
        ln.eng.close()
