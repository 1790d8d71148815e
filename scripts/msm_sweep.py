"""raw MSM sweep with per-phase device times: python scripts/msm_sweep.py [log2 sizes...]; window bits 0 = auto;
SWEEP_AUTO=1: only the automatic window width"""
import hashlib, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import bpp
eng = bpp.engine()
sizes = [int(x) for x in sys.argv[1:]] or [12, 16, 20]
base = eng.from_uniform(hashlib.shake_256(b"sweep").digest(64 * (1 << 14)))
for lg in sizes:
    n = 1 << lg
    pts = (base * ((n * 32 + len(base) - 1) // len(base)))[: 32 * n]
    sc = bytearray(hashlib.shake_256(b"sc%d" % lg).digest(32 * n))
    for i in range(31, len(sc), 32):
        sc[i] &= 0x0F
    for c in ([0] if lg < 16 or os.environ.get("SWEEP_AUTO") else [0, 12, 13, 14, 15, 16]):
        plan = bpp.pkg.MsmPlan(eng, pts, c)
        plan.set_scalars(bytes(sc))
        ref = plan.run(True)
        eng.phase_timing(True)
        plan.run(True)
        ph = eng.phase_ms()
        eng.phase_timing(False)
        reps = 5
        eng.timer_start()
        for _ in range(reps):
            plan.run(False)
        t = eng.timer_stop() / reps
        print("2^%d c=%d (W=%d): %.3f ms  %.1f Mpoints/s   sort %.3f bucket %.3f reduce %.3f combine %.3f" % (
            lg, plan.window_bits, (252 + plan.window_bits - 1) // plan.window_bits, t, n / t / 1e3, ph["msm_sort"], ph["msm_bucket"], ph["msm_reduce"], ph["msm_combine"]))
        plan.close()
