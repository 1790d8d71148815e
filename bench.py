#!/usr/bin/env python
"""bench.py — batched verification throughput of 64-bit Bulletproofs+ range proofs on B200 (BASELINE.json metric).

Unit of work ("job") = BASELINE.json configs[1]: verify_batch of 1024 non-aggregated 64-bit proofs, issued as 4 reference calls
of 256 (RangeProof::verify_batch looks at 256 proofs per call, /root/reference/src/range_proof.rs:739-751).  One STEP = `reps`
independent jobs back to back (reps = ceil(4096 / steps), so that the K timed steps span >= 0.4 s whatever K the driver picks;
`config.jobs_per_step`).
  value  : proofs/s with inputs resident in HBM.  Jobs are verified the way the engine's coalescing queue verifies them: `pass_jobs`
           jobs (16 x 1024 proofs = 64 reference calls) per device pass (bpp_vbatch_create_multi), `lanes` passes in flight.
  e2e    : proofs/s through the queue's C-ABI entry points (bpp_vqueue_submit / bpp_vqueue_wait) with HOST buffers: every job's
           proof bytes, commitments and transcripts are copied host->device and its statuses and advanced transcripts come back
           inside the timed region.
  one_batch_at_a_time / one_shot_4096 : latency of one 1024-proof job alone, and of 4096 proofs split over the GPUs of the run.
  N > 1  : one process per GPU (torchrun), every rank verifies its own jobs ("weak"), no collective on the data path;
           barrier + max-over-ranks timing.  extras.msm_sharded: the raw MSM sharded over the ranks (one 32-byte partial per GPU).
  --impl reference : the CPU restatement of the reference (oracle/, multi-threaded over independent verify_batch calls).
Rank 0 prints ONE JSON line.
"""
import argparse
import ctypes as C
import hashlib
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

# Lanes are independent CUDA streams; by default the driver multiplexes all streams of a process onto 8 hardware work queues, which
# falsely serialises kernels of different lanes.  Must be set before the CUDA context exists, i.e. before torch touches the device.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "64-bit range proofs verified/sec (batched)"
UNIT = "proofs/s"
BIT_LENGTH, EXT = 64, 1
CHUNK = 256
JOB = 1024
# per launch of the kernel over a 16-job pass, from profiles/r02_ncu_full_raw.csv (read + write)
NCU_DRAM_BYTES_PER_LAUNCH = {"decompress": 11.16e6, "msm_bucket": 49.05e6}      # (msm_bucket: the merged form; one sum per call: 106.9e6)
TARGET_JOBS = 4096                    # jobs in the timed region (>= 0.4 s at 10 M proofs/s)
# algorithmic 32x32->64 multiplies (field mul = 72, field square = 44, scalar Montgomery mul = 96 + 32)
MUL32_FE_MUL, MUL32_FE_SQ = 72, 44
MUL32_DECODE = 257 * MUL32_FE_SQ + 25 * MUL32_FE_MUL          # Ristretto decode + affine-Niels entry, per point
MUL32_MADD = 7 * MUL32_FE_MUL                                  # extended + affine-Niels mixed addition
LABEL = b"BatchedRangeProofTest"                               # benches/range_proof.rs:49


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def jobs_per_step(steps):
    return max(1, math.ceil(TARGET_JOBS / max(1, steps)))


def workload_config(args):
    reps = jobs_per_step(args.steps)
    return {"workload": "verify_batch of %d non-aggregated 64-bit proofs (aggregation 1, extension degree 1, minimum-value promises, seed nonces: "
                        "benches/range_proof.rs:206-292) = %d reference calls of <=256 per job (BASELINE.json configs[1]); one step = %d independent "
                        "jobs per GPU" % (JOB, JOB // CHUNK, reps),
            "proofs_per_job": JOB, "jobs_per_step": reps, "proofs_per_step_per_gpu": JOB * reps, "bit_length": BIT_LENGTH, "extension_degree": EXT,
            "action": "VerifyOnly",
            "l2": "flushed: every lane overwrites a %d MiB device buffer on its stream before each device pass, inside the timed region"
                  % int(os.environ.get("BPP_BENCH_FLUSH_MIB", "144"))}


# ------------------------------------------------------------------------------------------------ synthetic workload
def seeded_inputs(n_proofs, seed):
    """values, blindings, seed nonces and rng streams from SHAKE256(seed): the same bytes feed the device prover, the oracle and the
    CPU baseline (value = u64 % 2^63, promise value / 3: benches/range_proof.rs:236, :249)"""
    L = 2**252 + 27742317777372353535851937790883648493
    xof = hashlib.shake_256(b"bpp-bench-%d" % seed).digest(n_proofs * (8 + 64 + 64))
    vals, blinds, seeds = [], [], []
    for i in range(n_proofs):
        o = i * 136
        vals.append(int.from_bytes(xof[o:o + 8], "little") % (1 << 63))
        blinds.append([int.from_bytes(xof[o + 8:o + 72], "little") % (L - 1) + 1])
        seeds.append(int.from_bytes(xof[o + 72:o + 136], "little") % (L - 1) + 1)
    return vals, blinds, seeds


def make_workload_device(api, params, n_proofs, seed):
    """n_proofs non-aggregated 64-bit proofs made by the DEVICE prover (bpp_prove_batch) -> list of (commitment, min, seed, proof bytes)"""
    vals, blinds, seeds = seeded_inputs(n_proofs, seed)
    commits = params.gens.commit_batch(vals, blinds)
    need = api.RangeProof.rng_bytes_needed(params, 1)
    out = []
    for lo in range(0, n_proofs, 2048):
        hi = min(n_proofs, lo + 2048)
        sts = [api.RangeStatement.init(params, [commits[i]], [vals[i] // 3], seeds[i]) for i in range(lo, hi)]
        wits = [api.RangeWitness.init([api.CommitmentOpening(vals[i], blinds[i])]) for i in range(lo, hi)]
        streams = [hashlib.shake_256(b"bench-rng-%d-%d" % (seed, i)).digest(need) for i in range(lo, hi)]
        prs = api.RangeProof.prove_batch([api.Transcript(LABEL) for _ in range(lo, hi)], sts, wits, streams)
        for i, p in zip(range(lo, hi), prs):
            assert not isinstance(p, Exception), p
            out.append((commits[i], vals[i] // 3, seeds[i], p.to_bytes()))
    return out


def make_workload_oracle(n_proofs, seed):
    """the same proofs made by the CPU oracle prover (the reference arm runs where there may be no GPU)"""
    import orc
    from concurrent.futures import ThreadPoolExecutor

    vals, blinds, seeds = seeded_inputs(n_proofs, seed)
    op = orc.Params(BIT_LENGTH, 1, EXT)
    t0 = orc.transcript_new(LABEL)

    def one(i):
        need = 32 * 9
        st = orc.St(op, [op.commit(vals[i], blinds[i])], [vals[i] // 3], seeds[i])
        stream = hashlib.shake_256(b"bench-rng-%d-%d" % (seed, i)).digest(need)
        rc, pr, _ = orc.prove(t0, st, orc.Wit([vals[i]], [blinds[i]]), orc.Rng("buffer", data=stream))
        assert rc == 0
        return (st.commitments[0], vals[i] // 3, seeds[i], orc.proof_to_bytes(pr))

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
        return list(ex.map(one, range(n_proofs)))


def oracle_arrays(items, reps):
    """orc_verify_chunks_mt arguments for `reps` copies of the job made of `items`"""
    import orc

    op = orc.Params(BIT_LENGTH, 1, EXT)
    sts, prs = [], []
    for c, mn, sd, pb in items:
        sts.append(orc.St(op, [c], [mn], sd))
        rc, pr = orc.proof_from_bytes(pb)
        assert rc == 0
        prs.append(pr)
    t0 = orc.transcript_new(LABEL)
    n = len(items)
    offs = [0]
    for r in range(reps):
        for lo in range(0, n, CHUNK):
            offs.append(r * n + min(n, lo + CHUNK))
    sa = (orc.Statement * (n * reps))(*([s.c for s in sts] * reps))
    pa = (orc.Proof * (n * reps))(*(prs * reps))
    tb = C.create_string_buffer(t0 * (n * reps), 203 * n * reps)
    oa = (C.c_size_t * len(offs))(*offs)
    codes = (C.c_int32 * (len(offs) - 1))()
    keep = (op, sts, prs)
    return tb, sa, pa, oa, codes, keep


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons of the GPUs in use while the timed region runs"""

    def __init__(self, indices):
        self.indices, self.rows, self.proc = indices, [], None

    def start(self):
        if not self.indices:
            return
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", ",".join(str(i) for i in self.indices), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, windows=()):
        if not self.indices:
            return None
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for t, r in self.rows if any(a <= t <= b + 0.25 for a, b in windows)]
        for r in (inside or [r for _, r in self.rows]):
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "samples_inside_timed_regions": len(inside), "gpus_sampled": list(self.indices)}


# ------------------------------------------------------------------------------------------------ reference arm (CPU)
def cpu_arm(items, steps, warmup, threads):
    """the oracle's restatement of RangeProof::verify_batch, `threads` pthreads over independent 256-proof calls; one step = enough
    copies of the 1024-proof job to occupy every core (a bounded sample of the GPU arm's step)"""
    import orc

    n_calls = (len(items) + CHUNK - 1) // CHUNK
    reps = max(1, (threads + n_calls - 1) // n_calls)
    tb, sa, pa, oa, codes, keep = oracle_arrays(items, reps)
    lib = orc.lib()
    n = len(items) * reps

    def step():
        sec = lib.orc_verify_chunks_mt(tb, sa, pa, oa, len(oa) - 1, orc.VERIFY_ONLY, threads, codes)
        assert all(c == 0 for c in codes), list(codes)
        return sec

    for _ in range(warmup):
        step()
    total = sum(step() for _ in range(steps))
    sample = "%d steps x %d proofs = %d verify_batch calls of %d (the %d-proof job x%d), VerifyOnly, %d pthreads" % (
        steps, n, len(oa) - 1, CHUNK, len(items), reps, threads)
    return n * steps / total, 1e3 * total / steps, sample


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    items = make_workload_oracle(JOB, 8675309)
    value, ms, sample = cpu_arm(items, args.steps, args.warmup, threads)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 limbs (GF(2^255-19), scalars mod l)", "data": "synthetic", "impl": "reference",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "C restatement of tari_bulletproofs_plus 0.4.1 + curve25519-dalek's serial algorithms (no Rust toolchain in the "
                                 "image); not dalek itself -- its AVX2 backend would be an estimated 1.5-2x faster"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ secondary metrics
def msm_inputs(eng, nn, dist_kind, tag):
    """points: a 16 k-point set repeated (duplicates are legal MSM input); scalars by distribution"""
    L = 2**252 + 27742317777372353535851937790883648493
    uni = hashlib.shake_256(b"msm-points").digest(64 * min(nn, 1 << 14))
    pts = eng.from_uniform(uni)
    if dist_kind == "duplicated_points":
        pts = pts[:32 * 64]                                                        # 64 distinct points
    pts = (pts * ((nn * 32 + len(pts) - 1) // len(pts)))[: 32 * nn]
    sc = bytearray(hashlib.shake_256(b"msm-scalars-%s" % tag).digest(32 * nn))
    if dist_kind in ("uniform", "duplicated_points", "zeros5"):
        for i in range(31, len(sc), 32):
            sc[i] &= 0x0F                                                           # < 2^252 < l: canonical
        if dist_kind == "zeros5":
            z = bytes(32)
            for i in range(0, nn, 20):
                sc[32 * i:32 * i + 32] = z
    elif dist_kind == "small64":
        for i in range(nn):
            sc[32 * i + 8:32 * i + 32] = bytes(24)
    elif dist_kind == "prover_like":                                               # {0, 1, l - 1}
        vals = [bytes(32), (1).to_bytes(32, "little"), (L - 1).to_bytes(32, "little")]
        src = bytes(sc)
        for i in range(nn):
            sc[32 * i:32 * i + 32] = vals[src[32 * i] % 3]
    return bytes(sc), pts


def time_msm(bpp, eng, nn, sc, pts, msm_peak, hbm_gbs):
    plan = bpp.pkg.MsmPlan(eng, pts)
    plan.set_scalars(sc)
    res = plan.run(True)
    eng.phase_timing(True)                       # per-phase CUDA events of one run: sort / bucket sums / window reduction / Horner
    plan.run(True)
    ph = eng.phase_ms()
    eng.phase_timing(False)
    reps = 5 if nn <= (1 << 20) else 2
    eng.timer_start()
    for _ in range(reps):
        plan.run(False)
    t = eng.timer_stop() / reps
    c = plan.window_bits
    W = (252 + c - 1) // c
    adds_mul32 = nn * W * MUL32_MADD              # bucket additions only (the algorithmic work of SURVEY.md 8d minus the reduction)
    # sort phase, algorithmic bytes: the scalars are read by both digit passes (2 x 32 B), every non-zero digit is one 4-byte counter
    # update per pass and one 4-byte record written once
    sort_bytes = nn * (64 + 12 * W)
    out = {"mpoints_per_s": nn / t / 1e3, "ms": t, "window_bits": c, "windows": W,
           "achieved_tmul32_per_s": adds_mul32 / (t * 1e-3) / 1e12, "frac_of_int32_mul_peak": adds_mul32 / (t * 1e-3) / msm_peak,
           "bucket_kernel": {"ms": ph["msm_bucket"], "achieved_tmul32_per_s": adds_mul32 / (ph["msm_bucket"] * 1e-3) / 1e12,
                             "frac_of_int32_mul_peak": adds_mul32 / (ph["msm_bucket"] * 1e-3) / msm_peak},
           "sort_phase": {"ms": ph["msm_sort"], "algorithmic_bytes": sort_bytes, "gb_per_s": sort_bytes / (ph["msm_sort"] * 1e-3) / 1e9,
                          "frac_of_hbm_peak": (sort_bytes / (ph["msm_sort"] * 1e-3) / 1e9 / hbm_gbs) if hbm_gbs else None},
           "phase_ms": {k: ph[k] for k in ("msm_sort", "msm_bucket", "msm_reduce", "msm_combine")}}
    plan.close()
    return out, res


def run_extras(eng, api, bpp, orc, args, hbm_gbs):
    """proving throughput (batched lock-step prover, C-ABI call with host buffers) and raw MSM throughput (device-resident
    points, scalars uploaded once), both on this GPU; CPU oracle beside the prover on a bounded sample"""
    out = {}
    # ---- prover: P non-aggregated 64-bit proofs (BASELINE.json configs[0] shape, batched), as `lanes` concurrent bpp_prove_batch
    # calls of P / lanes proofs each (one bpp_ctx + host thread per call: the host Fiat-Shamir of one call overlaps the device
    # work of the others); timed: the C-ABI calls with host buffers, arguments packed beforehand
    P, PL = args.prove_batch, max(1, args.prove_lanes)
    per = P // PL
    ppool = api.VerifierPool(eng.device, BIT_LENGTH, 1, EXT, lanes=PL, blocking_waits=False)
    gp = ppool.lanes[0][1]
    vals, blinds, seeds = seeded_inputs(P, 4242)
    commits = gp.gens.commit_batch(vals, blinds)
    need = api.RangeProof.rng_bytes_needed(gp, 1)
    streams = [hashlib.shake_256(b"bench-rng-%d" % i).digest(need) for i in range(P)]
    packs = []
    for li in range(PL):
        prm = ppool.lanes[li][1]
        idx = range(li * per, (li + 1) * per)
        sts = [api.RangeStatement.init(prm, [commits[i]], [vals[i] // 3], seeds[i]) for i in idx]
        wits = [api.RangeWitness.init([api.CommitmentOpening(vals[i], blinds[i])]) for i in idx]
        packs.append(api._PackedProve(prm, [api.Transcript(LABEL) for _ in idx], sts, wits, [streams[i] for i in idx]))

    def prove_job(li, e, prm, i):
        packs[i].reset_transcripts()
        packs[i].run()

    times = []
    for _ in range(5):
        t0 = time.perf_counter()
        ppool.run(prove_job, PL)
        times.append((time.perf_counter() - t0) * 1e3)
    ms = sorted(times[1:])[len(times[1:]) // 2]
    proofs = [r for pk in packs for r in pk.results()]
    assert not any(isinstance(p, Exception) for p in proofs)
    # byte-identical to the CPU oracle on the first proofs, and timed there on a bounded sample (single thread)
    op = orc.Params(BIT_LENGTH, 1, EXT)
    t0 = time.perf_counter()
    n_cpu = 8
    for i in range(n_cpu):
        st = orc.St(op, [commits[i]], [vals[i] // 3], seeds[i])
        rc, pr, _ = orc.prove(orc.transcript_new(LABEL), st, orc.Wit([vals[i]], [blinds[i]]), orc.Rng("buffer", data=streams[i]))
        assert rc == 0 and orc.proof_to_bytes(pr) == proofs[i].to_bytes()
    cpu_s = (time.perf_counter() - t0) / n_cpu
    # algorithmic multiplies per proof on the fixed-base path: (2N + ext) + rounds * 2 * (1 + ext + N) + (2N + 1 + ext) + (1 + ext) table
    # additions of W = 28 windows each, minus the zero digits of A's 0 / +-1 scalars, 7 field multiplications each
    N_, R_ = BIT_LENGTH, 6
    terms = R_ * 2 * (1 + EXT + N_) + (2 * N_ + 1 + EXT) + (1 + EXT)
    mul32_per_proof = (terms * 28 + N_ + 28 * EXT) * MUL32_MADD
    peak_ops, _ = eng.microbench(1, 2000)              # IMAD.HI issue rate: the int32-multiply ceiling (see the roofline object)
    out["prove"] = {"metric": "64-bit range proofs proved/sec (batched lock-step, 1 GPU, through bpp_prove_batch with host buffers)",
                    "value": per * PL / (ms * 1e-3), "unit": "proofs/s", "batch": per * PL, "concurrent_calls": PL, "ms_per_batch": ms,
                    "path": "fixed-base window tables (k_fb.cu), no generator folding",
                    "mul32_per_proof": mul32_per_proof, "achieved_tmul32_per_s": mul32_per_proof * per * PL / (ms * 1e-3) / 1e12,
                    "frac_of_int32_mul_peak": mul32_per_proof * per * PL / (ms * 1e-3) / peak_ops,
                    "cpu_oracle_proofs_per_s_per_core": 1.0 / cpu_s, "byte_identical_to_oracle_checked": n_cpu}
    ppool.close()
    # ---- raw MSM (BASELINE.json configs[4]), device-resident decoded points
    msm = {}
    msm_peak, _ = eng.microbench(1, 2000)
    for lg in args.msm_log2:
        nn = 1 << lg
        sc, pts = msm_inputs(eng, nn, "uniform", b"%d" % lg)
        msm["2^%d" % lg], _ = time_msm(bpp, eng, nn, sc, pts, msm_peak, hbm_gbs)
    out["msm"] = msm
    dists = {}
    lg = args.msm_dist_log2
    for kind in ("uniform", "prover_like", "small64", "zeros5", "duplicated_points"):
        sc, pts = msm_inputs(eng, 1 << lg, kind, b"%d" % lg)
        r, _ = time_msm(bpp, eng, 1 << lg, sc, pts, msm_peak, hbm_gbs)
        dists[kind] = {k: r[k] for k in ("mpoints_per_s", "ms", "window_bits", "phase_ms")}
    out["msm_scalar_distributions"] = {"points": "2^%d" % lg, "results": dists}
    return out


def run_msm_sharded(eng, bpp, dist, torch, rank, world, sizes):
    """raw MSM sharded over the ranks: rank r keeps [r N / G, (r + 1) N / G) device-resident, reduces it to ONE 32-byte
    partial, the partials are gathered (all_gather of 32 B per rank) and summed on the host; rank 0 also runs the whole MSM alone and
    the two results must be equal.  Every rank takes part (collective)."""
    par = __import__("importlib").import_module("bulletproofs-plus_b200.parallel")
    out = {}
    for lg in sizes:
        nn = 1 << lg
        sc, pts = msm_inputs(eng, nn, "uniform", b"%d" % lg)
        lo, hi = par.shard_range(nn, world, rank)
        plan = bpp.pkg.MsmPlan(eng, pts[32 * lo:32 * hi])
        plan.set_scalars(sc[32 * lo:32 * hi])
        buf = torch.zeros(32, dtype=torch.uint8, device="cuda")
        gathered = torch.zeros(32 * world, dtype=torch.uint8, device="cuda")

        def once():
            partial = plan.run(True)
            buf.copy_(torch.frombuffer(bytearray(partial), dtype=torch.uint8))
            dist.all_gather_into_tensor(gathered, buf)
            return par.sum_partials(eng, bytes(gathered.cpu().numpy()))

        res = once()
        reps = 5 if lg <= 20 else 3
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            r2 = once()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        assert r2 == res
        eng.timer_start()
        for _ in range(reps):
            plan.run(False)
        dev_ms = eng.timer_stop() / reps
        t = torch.tensor([dt, dev_ms * 1e-3], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        plan.close()
        single_ms = None
        if rank == 0:
            whole = bpp.pkg.MsmPlan(eng, pts)
            whole.set_scalars(sc)
            assert whole.run(True) == res, "sharded MSM differs from the single-GPU result"
            eng.timer_start()
            whole.run(False)
            single_ms = eng.timer_stop()
            whole.close()
            out["2^%d" % lg] = {"gpus": world, "mpoints_per_s": nn / float(t[0]) / 1e6, "ms_end_to_end_max_over_ranks": float(t[0]) * 1e3,
                                "ms_kernels_max_over_ranks": float(t[1]) * 1e3, "single_gpu_ms": single_ms, "equals_single_gpu_result": True}
        dist.barrier()
    return out


# ------------------------------------------------------------------------------------------------ the B200 arm
def run_b200(args, rank, local_rank, world):
    import torch

    import bpp
    import orc

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        # NCCL prints its banner to stdout (fd 1) when NCCL_DEBUG is set; stdout carries the one JSON line, so fd 1 points at stderr
        # while the communicator comes up (NCCL_DEBUG itself is left alone: the driver reads the rank count from that log)
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    api = bpp.pkg.api
    lib = bpp.ffi.lib()
    try:
        all_cores = sorted(os.sched_getaffinity(0))
    except AttributeError:
        all_cores = list(range(os.cpu_count() or 1))
    cores = len(all_cores)
    # N > 1: every rank keeps to its own slice of the host cores (what `numactl` / a launcher's --cpu-bind would do): the lanes, workers and
    # submitting threads of eight ranks no longer migrate across each other's caches and their first-touch memory stays on their slice's
    # node.  BPP_BENCH_NO_PIN=1 leaves the scheduler alone.
    pinned_cores = None
    if world > 1 and cores >= world and not os.environ.get("BPP_BENCH_NO_PIN") and hasattr(os, "sched_setaffinity"):
        per = cores // world
        pinned_cores = all_cores[local_rank * per:(local_rank + 1) * per]
        try:
            os.sched_setaffinity(0, pinned_cores)
        except OSError:
            pinned_cores = None
    reps = jobs_per_step(args.steps)
    n_jobs = args.steps * reps
    K = max(1, args.pass_jobs)
    # Hosts with few cores per GPU (the 8-GPU boxes of this pool: 4) hash the verifier-weight transcripts on the device (k_weights_sm, the
    # pass is one graph launch, ~3 core-ms of Keccak per pass less on the host; 1 % slower device-resident) and run 8 lanes of one thread
    per_rank = len(pinned_cores) if pinned_cores else max(1, cores // world)
    if args.device_weights < 0:
        args.device_weights = 1 if per_rank < 8 else 0
    if args.lanes <= 0:
        args.lanes = 8 if (args.device_weights or args.merged_check) else 6      # (merged check: 10.6 M proofs/s with 6 passes in flight, 10.8 M with 8)
    S = max(1, args.lanes)
    htl = args.host_threads_per_lane or max(1, cores // (S * world))
    pool = api.VerifierPool(local_rank, BIT_LENGTH, 1, EXT, lanes=S, host_threads_per_lane=htl, blocking_waits=True, device_weights=bool(args.device_weights), merged_check=bool(args.merged_check))
    eng, params = pool.lanes[0]
    FLUSH = int(os.environ.get("BPP_BENCH_FLUSH_MIB", "144")) << 20
    action = api.VerifyAction.VerifyOnly

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- workload: K distinct jobs made by the device prover (different proofs on every rank)
    items = make_workload_device(api, params, K * JOB, seed=8675309 + 1000 * rank)
    t_label = api.Transcript(LABEL).state

    def job_calls(prm, j):
        calls = []
        for lo in range(j * JOB, (j + 1) * JOB, CHUNK):
            part = items[lo:lo + CHUNK]
            sts = [api.RangeStatement.init(prm, [c], [mn], sd) for c, mn, sd, _ in part]
            prs = [api.RangeProof(pb, EXT, 6) for _, _, _, pb in part]
            calls.append(([api.Transcript(state=t_label) for _ in part], sts, prs))
        return calls

    # ---------------- device-resident arm (value): every lane owns one resident pass of K jobs (bpp_vbatch_create_multi)
    n_pass = (n_jobs + K - 1) // K
    lane_pks = [[api._Packed(prm, job_calls(prm, j), action) for j in range(K)] for _, prm in pool.lanes]
    lane_vb = []
    for (e, prm), pks in zip(pool.lanes, lane_pks):
        ptrs = (C.c_void_p * K)(*[C.addressof(pk.args) for pk in pks])
        vb = C.c_void_p()
        rc = lib.bpp_vbatch_create_multi(prm.gens.h, K, ptrs, C.byref(vb))
        assert rc == 0, rc
        st = (C.c_int32 * (K * (JOB // CHUNK)))()
        lane_vb.append((vb, st))

    def dev_pass(li, e, prm, i):
        vb, st = lane_vb[li]
        if FLUSH:
            e.l2_flush(FLUSH)
        rc = lib.bpp_vbatch_run(vb, st, None, None)
        assert rc == 0 and not any(st), (rc, list(st))

    sampler = ClockSampler(list(range(world)) if rank == 0 else [])
    sampler.start()
    pool.run(dev_pass, S * args.warmup)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = pool.launch_count()
    t_wall0 = time.perf_counter()
    ev0.record()                       # the device is idle here (barrier above): ev0 precedes every kernel of the timed steps
    pool.run(dev_pass, n_pass)         # every C call returns after its stream has drained
    torch.cuda.synchronize()
    ev1.record()
    ev1.synchronize()
    dev_ms = ev0.elapsed_time(ev1)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = pool.launch_count() - launches0
    win_dev = (t_wall0, t_wall0 + t_wall)
    dev_jobs = n_pass * K

    # the same arm with one multiscalar check per reference call (the library's default; the merged check is a queue / ctx option):
    # a quarter of the passes, reported next to the headline as engine.per_call_check
    per_call_ms = None
    if args.merged_check:
        for e, _ in pool.lanes:
            e.set_merged_check(False)
        main_vb = lane_vb
        lane_vb = []
        for (e, prm), pks in zip(pool.lanes, lane_pks):
            ptrs = (C.c_void_p * K)(*[C.addressof(pk.args) for pk in pks])
            vb = C.c_void_p()
            assert lib.bpp_vbatch_create_multi(prm.gens.h, K, ptrs, C.byref(vb)) == 0
            lane_vb.append((vb, (C.c_int32 * (K * (JOB // CHUNK)))()))
        n_pass2 = max(4 * S, n_pass // 4)
        pool.run(dev_pass, S * args.warmup)
        barrier()
        ev0.record()
        pool.run(dev_pass, n_pass2)
        torch.cuda.synchronize()
        ev1.record()
        ev1.synchronize()
        per_call_ms = ev0.elapsed_time(ev1) / (n_pass2 * K)          # per job
        barrier()
        for vb, _ in lane_vb:
            lib.bpp_vbatch_destroy(vb)
        lane_vb = main_vb
        for e, _ in pool.lanes:
            e.set_merged_check(True)

    # ---------------- one job at a time on one lane (latency) and the per-kernel durations of a full pass / of one job
    eng.set_throughput_mode(0)                             # alone: spin-wait, all host threads of this rank's share
    eng.set_host_threads(max(1, min(64, cores // world)))
    vb1 = api.VerifyBatch(params, job_calls(params, 0), action)
    n_seq = 20

    def time_alone(run, n):
        ms = 0.0
        for _ in range(n):
            eng.l2_flush(FLUSH or (144 << 20))
            eng.sync()
            eng.timer_start()
            run()
            ms += eng.timer_stop()
        return ms / n

    def run_vb1():
        rc = lib.bpp_vbatch_run(vb1.h, vb1.pk.status, None, None)
        assert rc == 0 and not any(vb1.pk.status[c] for c in range(JOB // CHUNK))

    def run_pass0():
        dev_pass(0, eng, params, 0)

    run_vb1()
    seq_ms = time_alone(run_vb1, n_seq)
    pass_ms = time_alone(lambda: (lib.bpp_vbatch_run(lane_vb[0][0], lane_vb[0][1], None, None)), 5)

    def phases(run, n):
        acc = {}
        eng.phase_timing(True)
        for _ in range(n):
            eng.l2_flush(FLUSH or (144 << 20))
            run()
            for k, v in eng.phase_ms().items():
                acc[k] = acc.get(k, 0.0) + v / n
        eng.phase_timing(False)
        return acc

    ph_job = phases(run_vb1, 5)
    ph_pass = phases(lambda: lib.bpp_vbatch_run(lane_vb[0][0], lane_vb[0][1], None, None), 3)
    # 4096 proofs in one shot over the GPUs of this run: every rank verifies 4096 / world proofs as ONE pass (north_star's target case)
    shot_proofs = 4096 // world
    shot_pks = [api._Packed(params, job_calls(params, j)[: max(1, min(JOB, shot_proofs - j * JOB) // CHUNK)], action)
                for j in range((shot_proofs + JOB - 1) // JOB)]
    ptrs = (C.c_void_p * len(shot_pks))(*[C.addressof(pk.args) for pk in shot_pks])
    vb_shot = C.c_void_p()
    assert lib.bpp_vbatch_create_multi(params.gens.h, len(shot_pks), ptrs, C.byref(vb_shot)) == 0
    st_shot = (C.c_int32 * 64)()
    lib.bpp_vbatch_run(vb_shot, st_shot, None, None)
    barrier()
    shot_ms = time_alone(lambda: lib.bpp_vbatch_run(vb_shot, st_shot, None, None), 10)
    # the same end to end: host buffers -> verdicts (create + run + destroy of one multi-call pass)
    shot_times = []
    for it in range(12):                                 # the first two allocate the pooled workspace and capture the graphs
        t0 = time.perf_counter()
        vbx = C.c_void_p()
        assert lib.bpp_vbatch_create_multi(params.gens.h, len(shot_pks), ptrs, C.byref(vbx)) == 0
        assert lib.bpp_vbatch_run(vbx, st_shot, None, None) == 0
        lib.bpp_vbatch_destroy(vbx)
        if it >= 2:
            shot_times.append((time.perf_counter() - t0) * 1e3)
    shot_e2e_ms = statistics.median(shot_times)
    lib.bpp_vbatch_destroy(vb_shot)
    for vb, _ in lane_vb:
        lib.bpp_vbatch_destroy(vb)
    vb1.close()
    barrier()

    # ---------------- end-to-end arm (e2e): the coalescing queue with HOST buffers; jobs submitted from this thread
    # The queue hides a lane's host phases (building a pass: 12 MB of caller buffers parsed and staged, 17 % of a lane's time with 6 lanes;
    # handing results back) behind the device work of the other lanes, so it wants more lanes than the device-resident arm: measured on a
    # 16-core box 6 / 8 / 12 lanes -> 8.2 / 8.2-9.0 / 9.7 M proofs/s against 9.8 M device-resident.  Fewer on hosts with few cores per GPU.
    QS = max(1, args.queue_lanes or (8 if args.device_weights else 16 if per_rank >= 16 else 12 if per_rank >= 12 else 8))    # (merged check, 16 cores: 12 -> 10.7 M, 16 -> 11.3 M)
    qhtl = args.host_threads_per_lane or (1 if args.device_weights else max(1, min(2, (2 * per_rank) // QS)))
    q = api.VerifyQueue(local_rank, BIT_LENGTH, 1, EXT, lanes=QS, max_calls_per_pass=K, host_threads_per_lane=qhtl, device_weights=bool(args.device_weights), merged_check=bool(args.merged_check))
    n_slots = min(n_jobs, 2 * QS * K)
    pin = not args.pageable_inputs
    slots = [q.pack(job_calls(q.shape, j % K), action, pinned=pin) for j in range(n_slots)]      # proof bytes in page-locked host memory
    t_init = bytes(slots[0].tbuf.raw)
    tickets = [None] * n_slots

    n_sub = max(1, min(args.submitters, n_slots))        # submitting host threads (the queue takes calls from any number of threads)

    def e2e_part(t, count):
        mine = list(range(t, n_slots, n_sub))              # this thread's slots
        for j in range(count):
            s = mine[j % len(mine)]
            pk = slots[s]
            if tickets[s] is not None:
                q.wait(tickets[s])
                assert not any(pk.status[c] for c in range(pk.k)), list(pk.status)
                C.memmove(pk.tbuf, t_init, len(t_init))          # `&mut Transcript`s were advanced by the call
            tickets[s] = q.submit(pk)
        for s in mine:
            if tickets[s] is not None:
                q.wait(tickets[s])
                assert not any(slots[s].status[c] for c in range(slots[s].k))
                C.memmove(slots[s].tbuf, t_init, len(t_init))
                tickets[s] = None

    def e2e_run(count):
        if n_sub == 1:
            return e2e_part(0, count)
        errs = []

        def body(t, c):
            try:
                e2e_part(t, c)
            except BaseException as exc:                   # noqa: BLE001 - re-raised on the main thread
                errs.append(exc)
        ths = [threading.Thread(target=body, args=(t, count // n_sub + (1 if t < count % n_sub else 0))) for t in range(n_sub)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        if errs:
            raise errs[0]

    e2e_run(min(n_jobs, args.warmup * QS * K))
    # Three timed regions of the same n_jobs each; the line reports the MEDIAN (over regions) of the max over ranks and lists all three:
    # the arm depends on the host's scheduler and memory system, and on a shared box one region in three or four is 10-40 % slow.
    e2e_regions, e2e_windows, shares = [], [], []
    qs0 = q.stats()
    for _ in range(3):
        barrier()
        lm0 = q.lane_ms()
        t0 = time.perf_counter()
        e2e_run(n_jobs)                                  # returns after the last job's statuses and transcripts are back on the host
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        lm1 = q.lane_ms()
        e2e_regions.append(dt)
        e2e_windows.append((t0, t0 + dt))
        shares.append({k: round((lm1[k] - lm0[k]) / (1e3 * dt * QS), 3) for k in lm1})     # fraction of the region, mean over the lanes
    qs1 = q.stats()
    qs1 = {k: (qs1[k] - qs0[k]) // 3 + qs0[k] for k in qs1}      # per region
    mid = sorted(range(3), key=lambda i: e2e_regions[i])[1]
    e2e_s = e2e_regions[mid]
    lane_share = shares[mid]
    t0 = e2e_windows[0][0]
    clocks = sampler.stop([win_dev] + e2e_windows)
    barrier()
    # bytes one job moves (counted by the engine from the buffers it copies): one single-job call through the plain entry point
    pk1 = api._Packed(params, job_calls(params, 0), action)
    rc = lib.bpp_verify_chunks(params.gens.h, C.byref(pk1.args), pk1.status, pk1.masks, pk1.mask_present)
    assert rc == 0
    io_h2d, io_d2h = eng.io_bytes()
    host_one = eng.host_ms()
    # host cost of a coalesced pass (K jobs): create_multi alone, on this thread
    ptrsK = (C.c_void_p * K)(*[C.addressof(pk.args) for pk in lane_pks[0]])
    tc = []
    for _ in range(5):
        vbx = C.c_void_p()
        t1 = time.perf_counter()
        assert lib.bpp_vbatch_create_multi(params.gens.h, K, ptrsK, C.byref(vbx)) == 0
        tc.append((time.perf_counter() - t1) * 1e3)
        hm = eng.host_ms()
        lib.bpp_vbatch_destroy(vbx)
    q.close()
    barrier()

    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs")
    except Exception:
        pass

    # ---------------- secondary metrics (BASELINE.json: "proving at 1 GPU", "MSM Mpoints/s")
    extras = {}
    if args.extras:
        if world > 1:
            try:
                extras["msm_sharded"] = run_msm_sharded(eng, bpp, dist, torch, rank, world, args.msm_sharded_log2)
            except Exception as exc:
                extras["msm_sharded"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
        if rank == 0:
            if pinned_cores:                               # secondary metrics are single-GPU figures: the whole host, as at N = 1
                os.sched_setaffinity(0, all_cores)
            eng.set_host_threads(min(64, cores))
            try:
                extras.update(run_extras(eng, api, bpp, orc, args, hbm_peak))
            except Exception as exc:                       # secondary metrics must not take the headline line down with them
                extras["error"] = "%s: %s" % (type(exc).__name__, exc)
    barrier()

    # ---------------- reduce over ranks (max time)
    times = torch.tensor([dev_ms, e2e_s, seq_ms, shot_ms, shot_e2e_ms, pass_ms] + e2e_regions + [per_call_ms or 0.0], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms_max, _, seq_ms_max, shot_ms_max, shot_e2e_ms_max, pass_ms_max = (float(x) for x in times[:6])
    e2e_regions_max = [float(x) for x in times[6:9]]
    per_call_ms_max = float(times[9])
    e2e_s_max = sorted(e2e_regions_max)[1]

    # ---------------- roofline + cpu baseline (rank 0)
    if rank == 0:
        # int32-multiply ceiling, measured now on this GPU: every 32x32->64 product needs one high-half multiply (IMAD.HI, the slow half:
        # 8.6 T/s on this pool's B200s); the low halves issue on the other FMA sub-pipe (IMAD.lo alone: 18.6 T/s).
        peak_ops, _ = eng.microbench(1, 2000)
        wide_ops, _ = eng.microbench(2, 2000)
        pair_ops, _ = eng.microbench(12, 2000)
        alu_ops, _ = eng.microbench(3, 2000)
        def work_of(n_proofs):
            # the multiscalar sums as they are executed: one per reference call (c = 9 -> 28 windows of 256 buckets for 4226-entry sums), or
            # ONE per pass with the merged check (c = 14 at 270 k entries -> 19 windows of 8192 buckets)
            n_chunks = n_proofs // CHUNK
            n_pts = n_proofs * (3 + 2 * 6 + 1)
            entries = n_chunks * (2 * BIT_LENGTH + EXT + 1) + n_proofs * (3 + 2 * 6 + 1)
            n_sums = 1 if (args.merged_check and n_chunks >= 2) else n_chunks
            c_bits = lib.bpp_msm_window_bits(entries, n_sums) or 9
            W, B = (252 + c_bits - 1) // c_bits, 1 << (c_bits - 1)
            return {"decompress": n_pts * MUL32_DECODE,
                    "msm_bucket": entries * W * MUL32_MADD,
                    "msm_reduce": n_sums * W * 2 * B * 9 * MUL32_FE_MUL,
                    "msm_combine": n_sums * (W - 1) * (c_bits * (4 * MUL32_FE_MUL + 4 * MUL32_FE_SQ) + 9 * MUL32_FE_MUL),
                    "vprep_proof": n_proofs * (130 + 380) * 100,    # ~130 scalar products + one inversion (~380 at a^(l-2) cost), 100 mul32 each
                    "vprep_vector": n_proofs * (BIT_LENGTH * 4 + 3 * 14) * 100,      # 4 products per (proof, i) + three 8+8-entry tables
                    "vprep_weigh": (entries + n_proofs * 2 * BIT_LENGTH) * 100}

        def per_kernel_of(ph, n_proofs):
            work = work_of(n_proofs)
            # Keccak-f[1600] of the transcript replay: 197 ALU instructions per round (SASS of k_replay_sm), 24 rounds, ~21 permutations per proof
            alu_work = {"replay": n_proofs * 21 * 24 * 197}
            pk = {}
            for k, ms in ph.items():
                if ms <= 0:
                    continue
                if k in work:
                    ach = work[k] / (ms * 1e-3) / 1e12
                    pk[k] = {"ms": ms, "mul32": work[k], "achieved": ach, "peak": peak_ops / 1e12, "unit": "Tmul32/s", "frac": ach / (peak_ops / 1e12)}
                elif k in alu_work:
                    ach = alu_work[k] / (ms * 1e-3) / 1e12
                    pk[k] = {"ms": ms, "alu_ops": alu_work[k], "achieved": ach, "peak": alu_ops / 1e12, "unit": "Top32/s (ALU pipe)", "frac": ach / (alu_ops / 1e12)}
                else:
                    pk[k] = {"ms": ms}
            return pk, work

        pk_pass, work_pass = per_kernel_of(ph_pass, K * JOB)
        pk_job, _ = per_kernel_of(ph_job, JOB)
        dominant = max(work_pass, key=lambda k: work_pass[k] if k in pk_pass else -1)
        dom = pk_pass[dominant]
        job_mul32 = sum(work_of(K * JOB).values()) / K      # per job of a K-job pass (with the merged check a pass is one sum)
        ms_per_job = dev_ms_max / dev_jobs
        roof = {"bound": "int32-mul",
                "bound_note": "int32-multiply issue rate (IMAD.HI, one per 32x32->64 product); the path is modular big-integer arithmetic, neither HBM- nor "
                              "tensor-bound (north_star; DRAM traffic per 1024 proofs: a few MB, profiles/).  A FULL product cannot be issued at that rate: "
                              "IMAD.WIDE and a lo/hi pair both measure ~6 T/s (imad_wide_tops, imad_lo_hi_pairs_tops), so ~0.70-0.75 of this ceiling is the "
                              "hardware limit for the arithmetic (DESIGN.md 4)",
                "kernel": "k_" + dominant, "unit": dom.get("unit"), "achieved": dom.get("achieved"), "peak": dom.get("peak"), "frac": dom.get("frac"),
                "peak_source": "bpp_microbench, measured in this run: IMAD.HI issue rate (round 1's denominator, kept so that the fractions stay "
                               "comparable); LOP3+IADD3 rate for the Keccak kernel.  MEASURED_PEAKS.json has no integer figure",
                "imad_wide_tops": wide_ops / 1e12,
                "imad_lo_hi_pairs_tops": pair_ops / 1e12,
                "algorithmic_work_per_launch": work_pass[dominant],
                "kernel_ms": dom["ms"], "launch_covers_proofs": K * JOB,
                "whole_step": {"mul32_per_job": job_mul32, "ms_per_job_lanes_overlapped": ms_per_job,
                               "achieved": job_mul32 / (ms_per_job * 1e-3) / 1e12, "unit": "Tmul32/s",
                               "frac": job_mul32 / (ms_per_job * 1e-3) / peak_ops,
                               "note": "all arithmetic kernels of a job over the measured time per job with %d passes of %d jobs in flight" % (S, K)},
                "one_pass_alone": {"proofs": K * JOB, "ms": pass_ms_max, "frac": K * job_mul32 / (pass_ms_max * 1e-3) / peak_ops},
                "per_kernel": pk_pass, "per_kernel_one_job_alone": pk_job,
                # dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full` captures under profiles/
                "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get(dominant), "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel "
                "over a 16-job pass, ncu --set full with flushed caches (profiles/r02_ncu_summary.md); algorithmic bytes: 8.4 MB of encodings in + 25 MB of "
                "table entries out for the decompression (the table stays in L2 for the bucket sums), 30 MB of sorted entry lists + 25 MB of table + 20 MB of "
                "bucket results for the bucket sums",
                "hbm_peak_gbs_measured": hbm_peak}
        # CPU baseline beside it: the oracle on this box's cores, a bounded sample of the same workload
        if pinned_cores:                                   # the CPU arm gets the whole host, as at N = 1
            os.sched_setaffinity(0, all_cores)
        threads = cores
        cpu_value, _, cpu_sample = cpu_arm(items[:JOB], 3, 1, threads)
        cpu = {"value": cpu_value, "unit": UNIT, "cores": threads, "kind": "port", "sample": cpu_sample + "; C restatement, not dalek"}
        cfg = workload_config(args)
        line = {
            "metric": METRIC, "value": world * dev_jobs * JOB / (dev_ms_max * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / dev_jobs * reps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (GF(2^255-19), scalars mod l)", "data": "synthetic",
            "config": cfg, "host_cores": cores,
            "engine": {"lanes_per_gpu": S, "queue_lanes_per_gpu": QS, "queue_host_threads_per_lane": qhtl, "submitting_threads": n_sub,
                       "host_cores_of_this_rank": (pinned_cores if pinned_cores else "not pinned"),
                       "merged_check": ("on: ONE multiscalar check per device pass (sum over the pass's reference calls of rho_c x the call's own check, rho_c from "
                                        "the call's weight transcript), settled call by call when it fails; per-call statuses are the reference's "
                                        "(bpp_vqueue_set_merged_check, DESIGN.md 4.4)") if args.merged_check else "off: one multiscalar check per reference call",
                       "per_call_check": ({"value": world * JOB / (per_call_ms_max * 1e-3), "unit": UNIT,
                                           "note": "the device-resident arm with one multiscalar check per reference call (merged check off), a quarter of the passes"}
                                          if per_call_ms_max else None), "e2e_proof_bytes_in": "pageable host memory, staged by the engine" if args.pageable_inputs else "page-locked host memory, read by the copy engine in place",
                       "jobs_per_device_pass": K, "host_threads_per_lane": htl,
                       "timed_jobs_per_gpu": dev_jobs, "timed_device_passes_per_gpu": n_pass,
                       "transcript_replay": "host threads" if os.environ.get("BPP_HOST_REPLAY", "0") not in ("", "0") else "device (k_replay_sm)",
                       "verifier_weights": "device (k_weights_sm), one graph launch per pass" if args.device_weights else "host threads (8-way Keccak) between two graph launches",
                       "workload_made_by": "device prover (bpp_prove_batch); K = %d distinct jobs per rank" % K},
            "e2e": {"value": world * n_jobs * JOB / e2e_s_max, "unit": UNIT, "ms_per_step": 1e3 * e2e_s_max / args.steps,
                    "h2d_bytes_per_step": io_h2d * reps, "d2h_bytes_per_step": io_d2h * reps, "h2d_bytes_per_job": io_h2d, "d2h_bytes_per_job": io_d2h,
                    "through": "bpp_vqueue_submit / bpp_vqueue_wait, %d submitting thread(s) per GPU, %d lanes with %d host threads each, <= %d jobs per pass" % (n_sub, QS, qhtl, K),
                    "queue": {k: qs1[k] - qs0[k] for k in qs1},
                    "lane_time_share": lane_share,
                    "timed_regions_s": [round(x, 4) for x in e2e_regions_max], "timed_regions": "3 regions of the same jobs; value = jobs / median",
                    "host_ms_per_pass_of_%d_jobs" % K: {"create_multi_wall": statistics.median(tc), **{k: round(v, 4) for k, v in hm.items()}},
                    "host_ms_one_job_call": {k: round(v, 4) for k, v in host_one.items()}},
            "one_batch_at_a_time": {"value": world * JOB / (seq_ms_max * 1e-3), "ms_per_job": seq_ms_max, "jobs": n_seq,
                                    "note": "one 1024-proof job alone on one lane, device-resident, L2 flushed outside the CUDA-event bracket"},
            "one_shot_4096": {"proofs": shot_proofs * world, "gpus": world, "ms_device_resident": shot_ms_max, "ms_end_to_end": shot_e2e_ms_max,
                              "value_device_resident": shot_proofs * world / (shot_ms_max * 1e-3), "value_end_to_end": shot_proofs * world / (shot_e2e_ms_max * 1e-3),
                              "vs_cpu_baseline_end_to_end": shot_proofs * world / (shot_e2e_ms_max * 1e-3) / cpu_value,
                              "note": "4096 proofs verified ONCE, split evenly over the GPUs of this run, max over ranks (north_star's target case)"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "extras": extras,
            "wall_s_timed_region": t_wall,
        }
        print(json.dumps(line), flush=True)
    pool.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--lanes", type=int, default=0, help="device passes in flight per GPU (one bpp_ctx + host thread each); 0: 8, or 6 with host-side weights and per-call checks")
    ap.add_argument("--pass-jobs", type=int, default=16, help="1024-proof jobs merged into one device pass")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--merged-check", type=int, default=1, help="1: one multiscalar check per device pass, call by call only when it fails (bpp_vqueue_set_merged_check)")
    ap.add_argument("--pageable-inputs", type=int, default=0, help="1: the e2e arm's proof bytes in pageable memory (staged by the engine)")
    ap.add_argument("--submitters", type=int, default=2, help="host threads submitting calls to the end-to-end queue")
    ap.add_argument("--queue-lanes", type=int, default=0, help="lanes of the end-to-end queue (0 = --lanes)")
    ap.add_argument("--device-weights", type=int, default=-1, help="1: weight transcripts hashed on the device (k_weights_sm), a pass is ONE graph launch; -1: when the rank has fewer than 8 host cores")
    ap.add_argument("--host-threads-per-lane", type=int, default=0, help="0 = host cores / (lanes * ranks)")
    ap.add_argument("--extras", type=int, default=1, help="also measure proving and raw MSM throughput (secondary metrics)")
    ap.add_argument("--prove-batch", type=int, default=8192)
    ap.add_argument("--prove-lanes", type=int, default=8, help="concurrent bpp_prove_batch calls the proving batch is split into")
    ap.add_argument("--msm-log2", type=int, nargs="*", default=[12, 16, 20, 22, 24])
    ap.add_argument("--msm-dist-log2", type=int, default=20)
    ap.add_argument("--msm-sharded-log2", type=int, nargs="*", default=[20, 22, 24])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank, local_rank, world = dist_env()
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
