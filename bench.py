#!/usr/bin/env python
"""bench.py — batched verification throughput of 64-bit Bulletproofs+ range proofs on B200 (BASELINE.json metric).

One "step" = one pass of the verification hot path over one batch of synthetic proofs: BASELINE.json configs[1],
"verify_batch of 1024 non-aggregated 64-bit proofs on 1 B200", issued as 4 reference calls of 256 proofs
(RangeProof::verify_batch looks at 256 proofs per call, /root/reference/src/range_proof.rs:739-751).
  value  : proofs/s with inputs resident in HBM (bpp_vbatch_run: replay -> decompress -> scalar prep -> segmented MSM -> verdicts)
  e2e    : proofs/s through bpp_verify_chunks with HOST buffers (parsing, Fiat-Shamir, H2D, kernels, D2H inside the timing)
  Steps are independent; they are issued from --lanes lanes (api.VerifierPool: one bpp_ctx + host thread each) so that
  consecutive steps overlap on the GPU.  `one_batch_at_a_time` repeats the measurement with a single lane (latency).
  N > 1  : one process per GPU (torchrun), every rank verifies its own 1024 proofs per step ("weak"), no collective
           on the data path; barrier + max-over-ranks timing.
  --impl reference : the CPU restatement of the reference (oracle/, multi-threaded over independent verify_batch calls).
Rank 0 prints ONE JSON line.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

# Lanes are independent CUDA streams; by default the driver multiplexes all streams of a process onto 8 hardware work queues, which
# falsely serialises kernels of different lanes (measured: 32 lanes 4.9 M -> 5.8 M proofs/s with 32 queues).  Must be set before the
# CUDA context exists, i.e. before torch touches the device.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "64-bit range proofs verified/sec (batched)"
UNIT = "proofs/s"
BIT_LENGTH, EXT = 64, 1
CHUNK = 256
# algorithmic 32x32->64 multiplies (field mul = 72, field square = 44, scalar Montgomery mul = 96 + 32)
MUL32_FE_MUL, MUL32_FE_SQ = 72, 44
MUL32_DECODE = 257 * MUL32_FE_SQ + 25 * MUL32_FE_MUL          # Ristretto decode + affine-Niels entry, per point
MUL32_MADD = 7 * MUL32_FE_MUL                                  # extended + affine-Niels mixed addition


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def make_workload(n_proofs, seed=8675309):
    """n_proofs non-aggregated 64-bit proofs (value % 2^63, promise value/3, seed_nonce present: benches/range_proof.rs:206-292),
    generated with the CPU oracle prover, chunks built in parallel threads."""
    import workload
    from concurrent.futures import ThreadPoolExecutor

    n_chunks = (n_proofs + CHUNK - 1) // CHUNK
    sizes = [min(CHUNK, n_proofs - c * CHUNK) for c in range(n_chunks)]
    import orc

    params = orc.Params(BIT_LENGTH, 1, EXT)
    with ThreadPoolExecutor(max_workers=min(n_chunks, os.cpu_count() or 1)) as ex:
        cases = list(ex.map(lambda c: workload.make_case(BIT_LENGTH, [1] * sizes[c], EXT, promise="third", rng_seed=seed + c, params=params),
                            range(n_chunks)))
    return params, cases


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs"""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        if self.index == "off":
            return
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            sel = [] if self.index is None else ["-i", str(self.index)]          # None: every GPU of the node from ONE process
            self.proc = subprocess.Popen(["nvidia-smi"] + sel + ["--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, windows=()):
        """windows: (t0, t1) perf_counter intervals of the timed regions; samples inside them are preferred, else every sample taken
        while the sampler ran (it runs from before the warm-up to after the last timed step, i.e. under load throughout)"""
        if self.index == "off":
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
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_inside_timed_regions": len(inside)}


def run_reference(args, rank, world):
    """CPU arm: the oracle's restatement of RangeProof::verify_batch, T threads over independent 256-proof calls."""
    if rank != 0:
        return
    import orc

    threads = os.cpu_count() or 1
    params, cases = make_workload(args.proofs)
    reps = max(1, (threads + len(cases) - 1) // len(cases))       # enough independent calls to occupy every core
    sts, prs, trs, offs = [], [], [], [0]
    for _ in range(reps):
        for c in cases:
            sts += [s.c for s in c.statements]; prs += c.proofs; trs += c.transcripts
            offs.append(len(prs))
    n = len(prs)
    sa = (orc.Statement * n)(*sts)
    pa = (orc.Proof * n)(*prs)
    tb = C.create_string_buffer(b"".join(trs), 203 * n)
    oa = (C.c_size_t * len(offs))(*offs)
    codes = (C.c_int32 * (len(offs) - 1))()
    lib = orc.lib()

    def step():
        sec = lib.orc_verify_chunks_mt(tb, sa, pa, oa, len(offs) - 1, orc.VERIFY_ONLY, threads, codes)
        assert all(c == 0 for c in codes), list(codes)
        return sec

    for _ in range(args.warmup):
        step()
    total = sum(step() for _ in range(args.steps))
    value = n * args.steps / total
    sample = "%d proofs/step = %d verify_batch calls of %d (the %d-proof workload x%d), VerifyOnly, %d pthreads" % (
        n, len(offs) - 1, CHUNK, args.proofs, reps, threads)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 limbs (GF(2^255-19), scalars mod l)", "data": "synthetic", "impl": "reference",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "C restatement of tari_bulletproofs_plus 0.4.1 + curve25519-dalek algorithms (no Rust toolchain in the image); not dalek itself"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": "verify_batch of %d non-aggregated 64-bit proofs (aggregation 1, extension degree 1, minimum-value promises) "
                        "per GPU per step, as %d reference calls of <=256 (BASELINE.json configs[1])" % (args.proofs, (args.proofs + CHUNK - 1) // CHUNK),
            "proofs_per_step_per_gpu": args.proofs, "bit_length": BIT_LENGTH, "extension_degree": EXT, "action": "VerifyOnly",
            "transcript_replay": "host threads" if os.environ.get("BPP_HOST_REPLAY", "0") not in ("", "0") else "device (k_replay)",
            "steps_in_flight": "independent steps are issued from `lanes_per_gpu` lanes (one bpp_ctx + host thread each) and overlap on the GPU; "
                               "the K timed steps are bracketed once (barrier + synchronize + CUDA events on both sides)",
            "l2": "flushed between timed steps: every lane overwrites a %d MiB device buffer on its stream before each of its steps, "
                  "inside the timed region (inputs of a step are ~1.3 MB, far below L2)" % int(os.environ.get("BPP_BENCH_FLUSH_MIB", "144"))}


def run_extras(eng, api, bpp, orc, args):
    """proving throughput (batched lock-step prover, C-ABI call with host buffers) and raw MSM throughput (device-resident
    points, scalars uploaded once), both on this GPU; CPU oracle beside the prover on a bounded sample"""
    import hashlib

    out = {}
    # ---- prover: P non-aggregated 64-bit proofs (BASELINE.json configs[0] shape, batched), as `lanes` concurrent bpp_prove_batch
    # calls of P / lanes proofs each (one bpp_ctx + host thread per call: the host Fiat-Shamir of one call overlaps the device
    # work of the others); timed: the C-ABI calls with host buffers, arguments packed beforehand
    P, PL = args.prove_batch, max(1, args.prove_lanes)
    per = P // PL
    ppool = api.VerifierPool(eng.device, BIT_LENGTH, 1, EXT, lanes=PL, blocking_waits=False)
    gp = ppool.lanes[0][1]
    rng = orc.Rng("chacha", 4242)
    vals = [rng.next_u64() % (1 << 63) for _ in range(P)]
    blinds = [[rng.random_not_zero()] for _ in range(P)]
    commits = gp.gens.commit_batch(vals, blinds)
    seeds = [rng.random_not_zero() for _ in range(P)]
    need = api.RangeProof.rng_bytes_needed(gp, 1)
    streams = [hashlib.shake_256(b"bench-rng-%d" % i).digest(need) for i in range(P)]
    packs = []
    for li in range(PL):
        prm = ppool.lanes[li][1]
        idx = range(li * per, (li + 1) * per)
        sts = [api.RangeStatement.init(prm, [commits[i]], [vals[i] // 3], seeds[i]) for i in idx]
        wits = [api.RangeWitness.init([api.CommitmentOpening(vals[i], blinds[i])]) for i in idx]
        packs.append(api._PackedProve(prm, [api.Transcript(b"BatchedRangeProofTest") for _ in idx], sts, wits, [streams[i] for i in idx]))

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
        rc, pr, _ = orc.prove(orc.transcript_new(b"BatchedRangeProofTest"), st, orc.Wit([vals[i]], [blinds[i]]), orc.Rng("buffer", data=streams[i]))
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
    msm_peak, _ = eng.microbench(1, 2000)            # IMAD.HI issue rate (see the roofline object)
    for lg in args.msm_log2:
        nn = 1 << lg
        seed = hashlib.shake_256(b"msm-points").digest(64)
        uni = hashlib.shake_256(seed).digest(64 * min(nn, 1 << 14))
        pts = eng.from_uniform(uni)
        pts = (pts * ((nn * 32 + len(pts) - 1) // len(pts)))[: 32 * nn]          # repeat a 16 k-point set (duplicates are legal MSM input)
        plan = bpp.pkg.MsmPlan(eng, pts)
        sc = hashlib.shake_256(b"msm-scalars-%d" % lg).digest(32 * nn)
        sc = bytearray(sc)
        for i in range(31, len(sc), 32):
            sc[i] &= 0x0F                                                           # < 2^252 < l: canonical
        plan.set_scalars(bytes(sc))
        plan.run(True)
        eng.phase_timing(True)                       # per-phase CUDA events of one run: sort / bucket sums / window reduction / Horner
        plan.run(True)
        ph = eng.phase_ms()
        eng.phase_timing(False)
        reps = 5 if lg <= 20 else 2
        eng.timer_start()
        for _ in range(reps):
            plan.run(False)
        t = eng.timer_stop() / reps
        W = (252 + plan.window_bits - 1) // plan.window_bits
        adds_mul32 = nn * W * MUL32_MADD              # bucket additions only (the algorithmic work of SURVEY.md 8d minus the reduction)
        msm["2^%d" % lg] = {"mpoints_per_s": nn / t / 1e3, "ms": t, "window_bits": plan.window_bits, "windows": W,
                            "achieved_tmul32_per_s": adds_mul32 / (t * 1e-3) / 1e12, "frac_of_int32_mul_peak": adds_mul32 / (t * 1e-3) / msm_peak,
                            "bucket_kernel": {"ms": ph["msm_bucket"], "achieved_tmul32_per_s": adds_mul32 / (ph["msm_bucket"] * 1e-3) / 1e12,
                                              "frac_of_int32_mul_peak": adds_mul32 / (ph["msm_bucket"] * 1e-3) / msm_peak},
                            "phase_ms": {k: ph[k] for k in ("msm_sort", "msm_bucket", "msm_reduce", "msm_combine")}}
        plan.close()
    out["msm"] = msm
    return out


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

        # NCCL prints its version banner to stdout (fd 1) when NCCL_DEBUG is set; stdout carries the one JSON line, so fd 1 points
        # at stderr while the communicator comes up
        os.environ.pop("NCCL_DEBUG", None)
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
    S = max(1, min(args.lanes, args.steps))
    cores = os.cpu_count() or 1
    # S independent lanes (bpp_ctx + host thread each) on this GPU; the host cores are shared by the ranks of the node
    htl = args.host_threads_per_lane or max(1, cores // (S * world))
    blocking = (S > 1 and S * world >= cores) if args.blocking_waits < 0 else bool(args.blocking_waits)
    pool = api.VerifierPool(local_rank, BIT_LENGTH, 1, EXT, lanes=S, host_threads_per_lane=htl, blocking_waits=blocking)
    eng, params = pool.lanes[0]
    params_o, cases = make_workload(args.proofs, seed=8675309 + 1000 * rank)

    def build_calls(prm):
        calls = []
        for c in cases:
            sts = [api.RangeStatement.init(prm, s.commitments, s.min_values, s.seed_nonce) for s in c.statements]
            prs = [api.RangeProof.from_bytes(orc.proof_to_bytes(p)) for p in c.proofs]
            trs = [api.Transcript(state=t) for t in c.transcripts]
            calls.append((trs, sts, prs))
        return calls

    action = api.VerifyAction.VerifyOnly
    FLUSH_MIB = int(os.environ.get("BPP_BENCH_FLUSH_MIB", "144"))      # 151 MB > 126 MB of L2
    FLUSH = FLUSH_MIB << 20

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    # every lane owns its device-resident batch (bpp_vbatch: inputs in HBM) and its host-side argument block (e2e)
    vbs = [api.VerifyBatch(prm, build_calls(prm), action) for _, prm in pool.lanes]
    pks = [api._Packed(prm, build_calls(prm), action) for _, prm in pool.lanes]
    t_init = bytes(pks[0].tbuf.raw)
    for vb in vbs:
        status, _ = vb.run()
        assert status == [0] * len(cases), status

    # ---------------- device-resident arm (value): K steps over S lanes, L2 overwritten by every lane before every step
    def dev_step(li, e, prm, i):
        vb = vbs[li]
        if FLUSH:
            e.l2_flush(FLUSH)
        rc = lib.bpp_vbatch_run(vb.h, vb.pk.status, vb.pk.masks, vb.pk.mask_present)
        assert rc == 0 and all(vb.pk.status[c] == 0 for c in range(len(cases))), (rc, list(vb.pk.status))

    # one nvidia-smi process per NODE (rank 0, all GPUs): eight of them polling at once stall the driver's submission path (measured:
    # the 8-GPU device-resident arm dropped to 0.56 M proofs/s per GPU with one sampler per rank)
    sampler = ClockSampler((local_rank if world == 1 else None) if rank == 0 else "off")
    sampler.start()
    pool.run(dev_step, S * args.warmup)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = pool.launch_count()
    t_wall0 = time.perf_counter()
    ev0.record()                       # the device is idle here (barrier above): ev0 precedes every kernel of the timed steps
    pool.run(dev_step, args.steps)     # every C call returns after its stream has drained
    torch.cuda.synchronize()
    ev1.record()
    ev1.synchronize()
    dev_ms = ev0.elapsed_time(ev1)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = pool.launch_count() - launches0
    win_dev = (t_wall0, t_wall0 + t_wall)

    # ---------------- end-to-end arm (e2e): the C-ABI call with HOST buffers, K calls over S lanes
    host_acc = {}

    def e2e_step(li, e, prm, i, flush=True):
        pk = pks[li]
        C.memmove(pk.tbuf, t_init, len(t_init))          # `&mut Transcript`s are advanced by the call
        if FLUSH and flush:
            e.l2_flush(FLUSH)
        rc = lib.bpp_verify_chunks(prm.gens.h, C.byref(pk.args), pk.status, pk.masks, pk.mask_present)
        assert rc == 0 and all(pk.status[c] == 0 for c in range(pk.k)), (rc, list(pk.status))

    pool.run(e2e_step, S * args.warmup)
    barrier()
    t0 = time.perf_counter()
    pool.run(e2e_step, args.steps)                       # synchronous calls: each returns after the D2H of its verdicts
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop([win_dev, (t0, t0 + e2e_s)])
    barrier()
    # ---------------- one batch at a time on one lane (latency; the per-kernel figures of the roofline come from here)
    vb = vbs[0]
    seq_ms = 0.0
    n_seq = min(args.steps, 20)
    eng.set_throughput_mode(0)                             # alone: spin-wait, all host threads of this rank's share
    eng.set_host_threads(max(1, min(64, cores // world)))
    for _ in range(n_seq):
        eng.l2_flush(FLUSH or (144 << 20))
        eng.sync()
        eng.timer_start()
        rc = lib.bpp_vbatch_run(vb.h, vb.pk.status, vb.pk.masks, vb.pk.mask_present)
        seq_ms += eng.timer_stop()
        assert rc == 0
    # per-kernel durations: the same steps with CUDA events between the kernels (the events serialise the decompression
    # with the scalar prep, which otherwise overlap on two streams)
    phase_acc = {}
    eng.phase_timing(True)
    for _ in range(n_seq):
        eng.l2_flush(FLUSH or (144 << 20))
        rc = lib.bpp_vbatch_run(vb.h, vb.pk.status, vb.pk.masks, vb.pk.mask_present)
        assert rc == 0
        for k, v in eng.phase_ms().items():
            phase_acc[k] = phase_acc.get(k, 0.0) + v
    eng.phase_timing(False)

    e2e_seq_s = 0.0
    for _ in range(n_seq):
        eng.l2_flush(FLUSH or (144 << 20))
        eng.sync()
        t0 = time.perf_counter()
        e2e_step(0, eng, params, 0, flush=False)
        e2e_seq_s += time.perf_counter() - t0
        for k, v in eng.host_ms().items():
            host_acc[k] = host_acc.get(k, 0.0) + v / n_seq
    io_h2d, io_d2h = eng.io_bytes()
    barrier()

    # ---------------- secondary metrics (BASELINE.json: "proving at 1 GPU", "MSM Mpoints/s"), rank 0 only, not the headline
    extras = {}
    if rank == 0 and args.extras:
        eng.set_host_threads(min(64, cores))           # the lanes shared the host cores; the prover call below is alone
        try:
            extras = run_extras(eng, api, bpp, orc, args)
        except Exception as exc:                       # secondary metrics must not take the headline line down with them
            extras = {"error": "%s: %s" % (type(exc).__name__, exc)}

    # ---------------- reduce over ranks (max time)
    times = torch.tensor([dev_ms, e2e_s, seq_ms, e2e_seq_s], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_s_max, seq_ms_max, e2e_seq_s_max = (float(x) for x in times)

    # ---------------- roofline + cpu baseline (rank 0)
    if rank == 0:
        n_pts = args.proofs * (3 + 2 * 6 + 1)
        n_chunks = len(cases)
        entries = n_chunks * (2 * BIT_LENGTH + EXT + 1) + args.proofs * (3 + 2 * 6 + 1)
        per_launch = {k: v / n_seq for k, v in phase_acc.items()}
        # int32-multiply ceiling, measured now on this GPU: every 32x32->64 product needs one high-half multiply (IMAD.HI, the slow half:
        # 8.6 T/s on this pool's B200s); the low halves issue on the other FMA sub-pipe (IMAD.lo alone: 18.6 T/s).  A single IMAD.WIDE
        # per product is slower (6.1 T/s), which is why arith.cuh multiplies with mad.lo.cc / madc.hi.cc pairs.
        peak_ops, _ = eng.microbench(1, 2000)
        wide_ops, _ = eng.microbench(2, 2000)
        c_bits, W, B = 9, 28, 256                          # c = 9 -> ceil(252 / 9) = 28 windows of 256 buckets for 4226-entry segments
        work = {"decompress": n_pts * MUL32_DECODE,
                "msm_bucket": entries * W * MUL32_MADD,
                "msm_reduce": n_chunks * W * 2 * B * 9 * MUL32_FE_MUL,
                "msm_combine": n_chunks * (W - 1) * (c_bits * (4 * MUL32_FE_MUL + 4 * MUL32_FE_SQ) + 9 * MUL32_FE_MUL),
                "vprep_proof": args.proofs * (130 + 380) * 100,    # ~130 scalar products + one inversion (~380 at a^(l-2) cost), 100 mul32 each
                "vprep_vector": args.proofs * (BIT_LENGTH * 4 + 3 * 14) * 100,      # 4 products per (proof, i) + three 8+8-entry tables
                "vprep_weigh": (entries + args.proofs * 2 * BIT_LENGTH) * 100}
        # Keccak-f[1600] of the transcript replay: ~130 64-bit logic / rotate ops per round = 260 32-bit ALU ops, 24 rounds, ~21
        # permutations per 64-bit proof; its ceiling is the ALU pipe (LOP3 / IADD3 / SHF), measured by bpp_microbench(3)
        alu_ops, _ = eng.microbench(3, 2000)
        alu_work = {"replay": args.proofs * 21 * 24 * 260}
        per_kernel = {}
        for k, ms in per_launch.items():
            if ms <= 0:
                continue
            if k in work:
                ach = work[k] / (ms * 1e-3) / 1e12
                per_kernel[k] = {"ms": ms, "mul32": work[k], "achieved": ach, "peak": peak_ops / 1e12, "unit": "Tmul32/s", "frac": ach / (peak_ops / 1e12)}
            elif k in alu_work:
                ach = alu_work[k] / (ms * 1e-3) / 1e12
                per_kernel[k] = {"ms": ms, "alu_ops": alu_work[k], "achieved": ach, "peak": alu_ops / 1e12, "unit": "Top32/s (ALU pipe)", "frac": ach / (alu_ops / 1e12)}
            else:
                per_kernel[k] = {"ms": ms}
        # the dominant kernel = the one that carries the most algorithmic multiplies (and the most issued instructions in the ncu
        # launch list): the MSM bucket accumulation.  Its duration is the live CUDA-event figure of the one-batch-at-a-time pass.
        dominant = max(work, key=lambda k: work[k] if k in per_kernel else -1)
        dom = per_kernel[dominant]
        step_mul32 = sum(work[k] for k in work if k in per_kernel)
        step_ms = dev_ms_max / args.steps
        roof = {"bound": "int32-mul",
                "bound_note": "int32-multiply issue rate (IMAD.HI, one per 32x32->64 product); the path is modular big-integer arithmetic, neither HBM- nor "
                              "tensor-bound (north_star; DRAM traffic per step: a few MB, profiles/r01_ncu_summary.md)",
                "kernel": "k_" + dominant, "unit": dom.get("unit"), "achieved": dom.get("achieved"), "peak": dom.get("peak"), "frac": dom.get("frac"),
                "peak_source": "bpp_microbench, measured in this run: IMAD.HI issue rate (one per 32x32->64 product; the IMAD.lo half issues on the "
                               "other FMA sub-pipe); LOP3+IADD3 rate for the Keccak kernel.  MEASURED_PEAKS.json has no integer figure",
                "imad_wide_tops": wide_ops / 1e12,
                "algorithmic_work_per_launch": work[dominant],
                "kernel_ms": dom["ms"],
                "whole_step": {"mul32_per_step": step_mul32, "ms_per_step_lanes_overlapped": step_ms,
                               "achieved": step_mul32 / (step_ms * 1e-3) / 1e12, "unit": "Tmul32/s",
                               "frac": step_mul32 / (step_ms * 1e-3) / peak_ops,
                               "note": "all arithmetic kernels of a step over the measured time per step with %d lanes in flight" % S},
                "longest_kernel_one_batch_alone": max(per_launch, key=per_launch.get),
                "per_kernel": per_kernel,
                # dram__bytes_read.sum + dram__bytes_write.sum per launch from the round-1 `ncu --set full` captures (profiles/r01_ncu_summary.md)
                "traffic": {"replay": 961024, "msm_combine": 34304, "decompress": 588032, "vprep_proof": 519936, "msm_bucket": 3680000,
                            "msm_reduce": 3780000, "vprep_vector": 917000}.get(dominant),
                "traffic_unit": "bytes per launch (ncu, round 1)",
                "note": "one 1024-proof batch alone is a chain of latency-bound kernels (2-30 % occupancy each); the lanes overlap "
                        "independent batches, which is what `value` measures; per_kernel holds the one-batch-alone durations"}
        hbm_peak = None
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs")
        except Exception:
            pass
        roof["hbm_peak_gbs_measured"] = hbm_peak
        # CPU baseline beside it: oracle, bounded sample
        threads = os.cpu_count() or 1
        reps = max(1, (threads + n_chunks - 1) // n_chunks)
        sts, prs, trs, offs = [], [], [], [0]
        for _ in range(reps):
            for c in cases:
                sts += [s.c for s in c.statements]; prs += c.proofs; trs += c.transcripts
                offs.append(len(prs))
        n = len(prs)
        sa = (orc.Statement * n)(*sts); pa = (orc.Proof * n)(*prs)
        tb = C.create_string_buffer(b"".join(trs), 203 * n)
        oa = (C.c_size_t * len(offs))(*offs)
        codes = (C.c_int32 * (len(offs) - 1))()
        ol = orc.lib()
        ol.orc_verify_chunks_mt(tb, sa, pa, oa, len(offs) - 1, orc.VERIFY_ONLY, threads, codes)
        cpu_reps = 3
        sec = sum(ol.orc_verify_chunks_mt(tb, sa, pa, oa, len(offs) - 1, orc.VERIFY_ONLY, threads, codes) for _ in range(cpu_reps))
        cpu = {"value": n * cpu_reps / sec, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d x (%d proofs = %d verify_batch calls of %d), VerifyOnly, %d pthreads; C restatement, not dalek" % (
                   cpu_reps, n, len(offs) - 1, CHUNK, threads)}
        h2d, d2h = io_h2d, io_d2h          # counted by the engine from the buffers it copies (bpp_ctx_io_bytes)
        cfg = workload_config(args)
        cfg["lanes_per_gpu"] = S
        line = {
            "metric": METRIC, "value": world * args.proofs * args.steps / (dev_ms_max * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (GF(2^255-19), scalars mod l)", "data": "synthetic",
            "config": cfg, "host_cores": cores,
            "e2e": {"value": world * args.proofs * args.steps / e2e_s_max, "unit": UNIT, "ms_per_step": 1e3 * e2e_s_max / args.steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "lanes": S, "host_threads_per_lane": htl, "blocking_waits": blocking,
                    "one_call_at_a_time": {"value": world * args.proofs * n_seq / e2e_seq_s_max, "ms_per_call": 1e3 * e2e_seq_s_max / n_seq,
                                           "host_ms_per_call": {k: round(v, 4) for k, v in host_acc.items()}}},
            "one_batch_at_a_time": {"value": world * args.proofs * n_seq / (seq_ms_max * 1e-3), "ms_per_step": seq_ms_max / n_seq, "steps": n_seq,
                                    "note": "one lane, L2 flushed outside the per-step CUDA-event bracket (the round-1 `value`)"},
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
    ap.add_argument("--steps", type=int, default=512)
    ap.add_argument("--lanes", type=int, default=32, help="independent verification lanes (bpp_ctx + host thread) per GPU")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--host-threads-per-lane", type=int, default=0, help="0 = host cores / (lanes * ranks)")
    ap.add_argument("--blocking-waits", type=int, default=-1, help="-1 = when lanes * ranks >= host cores")
    ap.add_argument("--proofs", type=int, default=1024, help="proofs per GPU per step")
    ap.add_argument("--extras", type=int, default=1, help="also measure proving and raw MSM throughput on rank 0 (secondary metrics)")
    ap.add_argument("--prove-batch", type=int, default=8192)
    ap.add_argument("--prove-lanes", type=int, default=8, help="concurrent bpp_prove_batch calls the proving batch is split into")
    ap.add_argument("--msm-log2", type=int, nargs="*", default=[12, 16, 20, 22])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank, local_rank, world = dist_env()
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
