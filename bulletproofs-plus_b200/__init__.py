"""bulletproofs-plus_b200 — python host side over the C ABI of libbpp_b200.so (sm_100a CUDA engine).

Mirrors the reference's public surface for the accelerated path (names and error behaviour follow
/root/reference/src/range_proof.rs, range_parameters.rs, range_statement.rs, range_witness.rs,
commitment_opening.rs, extended_mask.rs, generators/pedersen_gens.rs); see api.py.  This module holds the thin
engine-level wrappers.  There is no CPU fallback: constructing an Engine without a CUDA device raises.
"""
import ctypes as C
import weakref

from . import _ffi
from ._ffi import (EngineError, OK, VERIFICATION_FAILED, INVALID_ARGUMENT, INVALID_LENGTH, INVALID_BLAKE2B, SIZE_OVERFLOW,
                   RECOVER_ONLY, RECOVER_AND_VERIFY, VERIFY_ONLY, TRANSCRIPT_BYTES, MAX_BATCH, build)

L = 2**252 + 27742317777372353535851937790883648493


def _chk(ctx, rc):
    if rc != 0:
        msg = ""
        if ctx is not None and ctx.h:
            msg = (_ffi.lib().bpp_last_error(ctx.h) or b"").decode(errors="replace")
        raise EngineError(rc, msg)


def _u64arr(vals):
    return (C.c_uint64 * max(1, len(vals)))(*vals)


class Engine:
    """One bpp_ctx: a CUDA stream on one device.  Calls on one Engine are serialised on its stream."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        rc = _ffi.lib().bpp_ctx_create(device, C.byref(self.h))
        if rc != 0:
            self.h = C.c_void_p()
            raise EngineError(rc, "bpp_ctx_create failed: no usable CUDA device %d (there is no CPU fallback)" % device)
        self.device = device
        self._children = weakref.WeakSet()       # Gens / MsmPlan / VerifyBatch handles that live inside this ctx

    def adopt(self, child):
        """handles created from this ctx are closed before the ctx is (their destructors touch its stream)"""
        self._children.add(child)

    def close(self):
        for ch in sorted(getattr(self, "_children", ()), key=lambda c: isinstance(c, Gens)):      # batches and plans before their Gens
            ch.close()
        if self.h:
            _ffi.lib().bpp_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self):
        return int(_ffi.lib().bpp_ctx_launch_count(self.h))

    @property
    def stream(self):
        return _ffi.lib().bpp_ctx_stream(self.h)

    def sync(self):
        _chk(self, _ffi.lib().bpp_ctx_sync(self.h))

    def l2_flush(self, nbytes=256 << 20):
        """enqueue a write of nbytes (> L2) of scratch on the engine's stream (benchmarks: cold L2 for the next call)"""
        _chk(self, _ffi.lib().bpp_ctx_l2_flush(self.h, nbytes))

    def timer_start(self):
        _chk(self, _ffi.lib().bpp_ctx_timer_start(self.h))

    def timer_stop(self):
        """milliseconds between timer_start and now on the engine's stream (CUDA events)"""
        ms = C.c_float()
        _chk(self, _ffi.lib().bpp_ctx_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def phase_timing(self, enable):
        _chk(self, _ffi.lib().bpp_ctx_phase_timing(self.h, 1 if enable else 0))

    PHASES = ("replay", "decompress", "vprep_proof", "vprep_vector", "host_weights_wait", "vprep_weigh", "msm_sort", "msm_bucket",
              "msm_reduce", "msm_combine", "encode")

    def phase_ms(self):
        arr = (C.c_float * 11)()
        _chk(self, _ffi.lib().bpp_ctx_phase_ms(self.h, arr))
        return dict(zip(self.PHASES, [float(x) for x in arr]))

    HOST_PHASES = ("parse", "layout", "fill_replay", "weights", "h2d", "run_wall")

    def io_bytes(self):
        """(host->device, device->host) bytes of the last verification call"""
        arr = (C.c_uint64 * 2)()
        _chk(self, _ffi.lib().bpp_ctx_io_bytes(self.h, arr))
        return int(arr[0]), int(arr[1])

    def set_replay_mode(self, on_device):
        """loop 1 (transcript replay): 0 / False = host threads, 1 / True = device (kernel picked by batch size),
        2 = device, one thread per proof, 3 = device, one warp per proof"""
        _chk(self, _ffi.lib().bpp_ctx_set_replay_mode(self.h, int(on_device)))

    def set_graphs(self, enable):
        """verification passes as captured CUDA graphs (default) or kernel by kernel"""
        _chk(self, _ffi.lib().bpp_ctx_set_graphs(self.h, 1 if enable else 0))

    def set_throughput_mode(self, enable):
        """0 / False: spin-wait (lowest latency for one call); 1 / True: blocking waits (many calls in flight from many ctxs);
        2: additionally verifier weights on the device, one graph per pass (measured slower; see include/bpp_b200.h)"""
        _chk(self, _ffi.lib().bpp_ctx_set_throughput_mode(self.h, int(enable)))

    def set_merged_check(self, enable):
        """one multiscalar check per pass (passes of four or more reference calls), chunk by chunk only when it fails (bpp_ctx_set_merged_check)"""
        _chk(self, _ffi.lib().bpp_ctx_set_merged_check(self.h, 1 if enable else 0))

    @property
    def merged_fallbacks(self):
        return int(_ffi.lib().bpp_ctx_merged_fallbacks(self.h))

    def set_test_hooks(self, flags):
        """bit 0: every verification pass is repeated through the zero-weight fallback; bit 1: a merged check is always followed by the
        chunk-by-chunk pass (results must not change)"""
        _chk(self, _ffi.lib().bpp_ctx_set_test_hooks(self.h, int(flags)))

    @property
    def graph_launch_count(self):
        return int(_ffi.lib().bpp_ctx_graph_launch_count(self.h))

    def host_ms(self):
        arr = (C.c_double * 6)()
        _chk(self, _ffi.lib().bpp_ctx_host_ms(self.h, arr))
        return dict(zip(self.HOST_PHASES, [float(x) for x in arr]))

    def set_host_threads(self, n):
        _chk(self, _ffi.lib().bpp_ctx_set_host_threads(self.h, n))

    # ---- batched point primitives
    def decompress_check(self, encodings, reencode=True):
        """encodings: bytes (n*32). Returns (ok list[int], re-encoded bytes or None)."""
        n = len(encodings) // 32
        ok = C.create_string_buffer(max(1, n))
        out = C.create_string_buffer(max(1, 32 * n)) if reencode else None
        _chk(self, _ffi.lib().bpp_decompress_check(self.h, n, encodings, ok, out))
        return list(ok.raw[:n]), (out.raw[: 32 * n] if reencode else None)

    def from_uniform(self, data):
        n = len(data) // 64
        out = C.create_string_buffer(max(1, 32 * n))
        _chk(self, _ffi.lib().bpp_from_uniform_batch(self.h, n, data, out))
        return out.raw[: 32 * n]

    # ---- MSM
    def msm(self, scalars, points):
        """scalars, points: bytes (n*32 each) -> 32-byte encoding of sum s_i * P_i"""
        n = len(scalars) // 32
        out = C.create_string_buffer(32)
        _chk(self, _ffi.lib().bpp_msm(self.h, n, scalars, points, out))
        return out.raw

    def msm_segmented(self, offsets, scalars, points):
        k = len(offsets) - 1
        out = C.create_string_buffer(32 * max(1, k))
        off = _u64arr(offsets)
        _chk(self, _ffi.lib().bpp_msm_segmented(self.h, k, C.cast(off, C.c_void_p), scalars, points, out))
        return [out.raw[32 * i: 32 * i + 32] for i in range(k)]

    def microbench(self, which, iters):
        ops, sec = C.c_double(), C.c_double()
        _chk(self, _ffi.lib().bpp_microbench(self.h, which, iters, C.byref(ops), C.byref(sec)))
        return ops.value, sec.value


def points_sum_host(points):
    """sum of <= 64 points given as concatenated 32-byte encodings, on the host (last step of a multi-GPU MSM)"""
    out = C.create_string_buffer(32)
    rc = _ffi.lib().bpp_points_sum_host(len(points) // 32, points, out)
    if rc != 0:
        raise EngineError(rc, "bpp_points_sum_host")
    return out.raw


class MsmPlan:
    """Device-resident point set for repeated MSMs (BASELINE.json configs[4])."""

    def __init__(self, engine, points, window_bits=0):
        self.engine = engine
        self.n = len(points) // 32
        self.h = C.c_void_p()
        _chk(engine, _ffi.lib().bpp_msm_plan_create(engine.h, self.n, points, window_bits, C.byref(self.h)))
        engine.adopt(self)

    @property
    def window_bits(self):
        return _ffi.lib().bpp_msm_plan_window_bits(self.h)

    def set_scalars(self, scalars):
        assert len(scalars) == 32 * self.n
        _chk(self.engine, _ffi.lib().bpp_msm_plan_set_scalars(self.h, scalars))

    def run(self, want_result=True):
        out = C.create_string_buffer(32) if want_result else None
        _chk(self.engine, _ffi.lib().bpp_msm_plan_run(self.h, out))
        return out.raw if want_result else None

    def close(self):
        # a handle whose ctx is already gone (finalisers of a garbage cycle run in arbitrary order) is dropped, not destroyed
        if self.h and self.engine.h:
            _ffi.lib().bpp_msm_plan_destroy(self.h)
        self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Gens:
    """Device-resident generator tables: RangeParameters::init(bit_length, aggregation_factor,
    create_pedersen_gens_with_extension_degree(ext)) (/root/reference/src/range_parameters.rs:32-58,
    src/ristretto.rs:67-76)."""

    def __init__(self, engine, bit_length, max_aggregation, extension_degree, h_base=None, g_bases=None):
        """h_base / g_bases: caller-made PedersenGens as 32-byte encodings (None = the reference's Ristretto constants)"""
        self.engine = engine
        self.h = C.c_void_p()
        gb = b"".join(g_bases) if g_bases else None
        _chk(engine, _ffi.lib().bpp_gens_create_with_bases(engine.h, bit_length, max_aggregation, extension_degree, h_base, gb, C.byref(self.h)))
        engine.adopt(self)
        self.bit_length, self.max_aggregation, self.extension_degree = bit_length, max_aggregation, extension_degree

    def point(self, which, index=0):
        out = C.create_string_buffer(32)
        _chk(self.engine, _ffi.lib().bpp_gens_get(self.h, which, index, out))
        return out.raw

    def fixed_base_msm(self, scalars, gidx, n_seg):
        """n_seg sums over the generator set (order Gi | Hi | G_k | H): scalars = n_seg * len(gidx) canonical 32-byte scalars,
        segment-major -> list of n_seg 32-byte encodings"""
        seg_len = len(gidx)
        out = C.create_string_buffer(32 * n_seg)
        gi = (C.c_uint32 * seg_len)(*gidx)
        _chk(self.engine, _ffi.lib().bpp_gens_fixed_base_msm(self.h, n_seg, seg_len, scalars, C.cast(gi, C.c_void_p), out))
        return [out.raw[32 * i: 32 * i + 32] for i in range(n_seg)]

    def commit_batch(self, values, blindings):
        """values: list[int]; blindings: list (per opening) of lists of ints (same length each) -> list of 32-byte encodings"""
        n = len(values)
        nb = len(blindings[0]) if n else 1
        vals = _u64arr(values)
        bl = b"".join(int(x % L).to_bytes(32, "little") for b in blindings for x in b)
        out = C.create_string_buffer(max(1, 32 * n))
        _chk(self.engine, _ffi.lib().bpp_pedersen_commit_batch(self.h, n, C.cast(vals, C.c_void_p), bl, nb, out))
        return [out.raw[32 * i: 32 * i + 32] for i in range(n)]

    def close(self):
        if self.h and self.engine.h:       # see MsmPlan.close
            _ffi.lib().bpp_gens_destroy(self.h)
        self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- Merlin transcripts on the 203-byte wire form (host)
def transcript_new(label):
    out = C.create_string_buffer(TRANSCRIPT_BYTES)
    _ffi.lib().bpp_transcript_new(label, len(label), out)
    return out.raw


def transcript_append_message(t, label, msg):
    buf = C.create_string_buffer(t, TRANSCRIPT_BYTES)
    _ffi.lib().bpp_transcript_append_message(buf, label, len(label), msg, len(msg))
    return buf.raw


def transcript_challenge_bytes(t, label, n):
    buf = C.create_string_buffer(t, TRANSCRIPT_BYTES)
    out = C.create_string_buffer(n)
    _ffi.lib().bpp_transcript_challenge_bytes(buf, label, len(label), out, n)
    return buf.raw, out.raw


def proof_check_bytes(data):
    """RangeProof::from_bytes validation: returns (status, extension_degree, rounds)."""
    ext, rounds = C.c_int32(), C.c_int32()
    rc = _ffi.lib().bpp_proof_check_bytes(data, len(data), C.byref(ext), C.byref(rounds))
    return rc, ext.value, rounds.value

from . import api  # noqa: E402  (host-side mirror of the reference API; needs the definitions above)
