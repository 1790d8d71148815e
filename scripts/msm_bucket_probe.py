"""bucket-sum kernel variants over mid-size raw MSMs: python scripts/msm_bucket_probe.py [log2 sizes...]
For every size: quads (BPP_MSM_BUCKET=1), whole threads (2), split threads with 2 / 4 / 8 lanes per bucket (3 + BPP_MSM_SPLIT), each
at the automatic window width c and at c + 1, c + 2; prints total and per-phase device times.  The switches are read per launch."""
import hashlib, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import bpp
eng = bpp.engine()
sizes = [int(x) for x in sys.argv[1:]] or [13, 14, 15, 16, 17, 18]
base = eng.from_uniform(hashlib.shake_256(b"sweep").digest(64 * (1 << 13)))
variants = [("auto", {}), ("quads", {"BPP_MSM_BUCKET": "1"}), ("threads", {"BPP_MSM_BUCKET": "2"})] + [
    ("split%d" % t, {"BPP_MSM_BUCKET": "3", "BPP_MSM_SPLIT": str(t)}) for t in (2, 4, 8)]
for lg in sizes:
    n = 1 << lg
    pts = (base * ((n * 32 + len(base) - 1) // len(base)))[: 32 * n]
    sc = bytearray(hashlib.shake_256(b"sc%d" % lg).digest(32 * n))
    for i in range(31, len(sc), 32):
        sc[i] &= 0x0F
    auto = bpp.pkg.MsmPlan(eng, pts, 0)
    c0 = auto.window_bits
    auto.close()
    want = None
    for c in (c0 - 1, c0, c0 + 1, c0 + 2):
        plan = bpp.pkg.MsmPlan(eng, pts, c)
        plan.set_scalars(bytes(sc))
        for name, env in variants:
            for k in ("BPP_MSM_BUCKET", "BPP_MSM_SPLIT"):
                os.environ.pop(k, None)
            os.environ.update(env)
            got = plan.run(True)
            want = want or got
            assert got == want, (lg, c, name)
            eng.phase_timing(True)
            plan.run(True)
            ph = eng.phase_ms()
            eng.phase_timing(False)
            reps = 10
            eng.timer_start()
            for _ in range(reps):
                plan.run(False)
            t = eng.timer_stop() / reps
            print("2^%d c=%d %-8s %.3f ms  %6.1f Mpoints/s   sort %.3f bucket %.3f reduce %.3f combine %.3f" % (
                lg, c, name, t, n / t / 1e3, ph["msm_sort"], ph["msm_bucket"], ph["msm_reduce"], ph["msm_combine"]), flush=True)
        plan.close()
