"""ctypes binding of libbpp_b200.so (the C ABI declared in include/bpp_b200.h).

The library is the product: there is no CPU fallback here.  If the shared object is missing, or no CUDA device is
usable, every compute entry point fails loudly (EngineError / OSError).
"""
import ctypes as C
import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libbpp_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(PKG_DIR), "include", "bpp_b200.h")

OK, VERIFICATION_FAILED, INVALID_ARGUMENT, INVALID_LENGTH, INVALID_BLAKE2B, SIZE_OVERFLOW = range(6)
ERR_CUDA, ERR_INTERNAL = 100, 101
RECOVER_ONLY, RECOVER_AND_VERIFY, VERIFY_ONLY = 0, 1, 2
TRANSCRIPT_BYTES = 203
MAX_BATCH = 256

STATUS_NAMES = {0: "Ok", 1: "VerificationFailed", 2: "InvalidArgument", 3: "InvalidLength", 4: "InvalidBlake2b",
                5: "SizeOverflow", 100: "CudaError", 101: "InternalError"}


class EngineError(Exception):
    """A non-zero bpp_status; `.code` mirrors ProofError (/root/reference/src/errors.rs:12-28) for 1..5."""

    def __init__(self, code, msg=""):
        super().__init__("%s (%d)%s" % (STATUS_NAMES.get(code, "?"), code, (": " + msg) if msg else ""))
        self.code = code


def build(force=False, verbose=False):
    """Compile libbpp_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    if force:
        subprocess.run(["make", "-C", CSRC_DIR, "clean"], check=True, capture_output=not verbose)
    r = subprocess.run(["make", "-C", CSRC_DIR, "-j8"], capture_output=not verbose)
    if r.returncode != 0:
        raise RuntimeError("building libbpp_b200.so failed:\n" + ((r.stdout or b"") + (r.stderr or b"")).decode(errors="replace")[-4000:])
    return LIB_PATH


class VerifyArgs(C.Structure):
    _fields_ = [
        ("n_proofs", C.c_size_t),
        ("n_chunks", C.c_size_t),
        ("chunk_offsets", C.c_void_p),
        ("proof_bytes", C.c_void_p),
        ("proof_offsets", C.c_void_p),
        ("commitments32", C.c_void_p),
        ("commit_offsets", C.c_void_p),
        ("min_values", C.c_void_p),
        ("min_present", C.c_void_p),
        ("seed_nonces32", C.c_void_p),
        ("seed_present", C.c_void_p),
        ("transcripts", C.c_void_p),
        ("action", C.c_int32),
    ]


class VerifyChallenges(C.Structure):
    _fields_ = [("challenges32", C.c_void_p), ("challenge_offsets", C.c_void_p), ("weights32", C.c_void_p)]


class ProveArgs(C.Structure):
    _fields_ = [
        ("n_proofs", C.c_size_t),
        ("aggregation", C.c_int32),
        ("commitments32", C.c_void_p),
        ("values", C.c_void_p),
        ("blindings32", C.c_void_p),
        ("min_values", C.c_void_p),
        ("min_present", C.c_void_p),
        ("seed_nonces32", C.c_void_p),
        ("seed_present", C.c_void_p),
        ("transcripts", C.c_void_p),
        ("rng_bytes", C.c_void_p),
        ("rng_stride", C.c_size_t),
    ]


_lib = None


def lib():
    """Load the shared object (raises OSError if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError("libbpp_b200.so is not built (run __graft_entry__.build()); there is no CPU fallback")
    l = C.CDLL(LIB_PATH)
    vp, sz, cp, i32 = C.c_void_p, C.c_size_t, C.c_char_p, C.c_int32
    P = C.POINTER
    sig = {
        "bpp_ctx_create": (i32, [i32, P(vp)]),
        "bpp_ctx_destroy": (None, [vp]),
        "bpp_last_error": (cp, [vp]),
        "bpp_ctx_sync": (i32, [vp]),
        "bpp_ctx_launch_count": (C.c_uint64, [vp]),
        "bpp_ctx_stream": (vp, [vp]),
        "bpp_ctx_set_host_threads": (i32, [vp, i32]),
        "bpp_ctx_l2_flush": (i32, [vp, sz]),
        "bpp_ctx_timer_start": (i32, [vp]),
        "bpp_ctx_timer_stop": (i32, [vp, P(C.c_float)]),
        "bpp_ctx_phase_timing": (i32, [vp, i32]),
        "bpp_ctx_phase_ms": (i32, [vp, P(C.c_float)]),
        "bpp_ctx_host_ms": (i32, [vp, P(C.c_double)]),
        "bpp_ctx_io_bytes": (i32, [vp, P(C.c_uint64)]),
        "bpp_decompress_check": (i32, [vp, sz, cp, cp, cp]),
        "bpp_from_uniform_batch": (i32, [vp, sz, cp, cp]),
        "bpp_points_sum_host": (i32, [sz, cp, cp]),
        "bpp_host_verifier_weights": (i32, [cp, sz, sz, i32, cp]),
        "bpp_host_sc_from_wide64": (None, [cp, cp]),
        "bpp_host_sc_mul64": (None, [cp, cp, cp]),
        "bpp_host_sc_generic64": (None, [cp, cp, cp]),
        "bpp_keccak_f1600_x1": (None, [vp]),
        "bpp_keccak_f1600_x1_generic": (None, [vp]),
        "bpp_host_simd_level": (i32, []),
        "bpp_msm": (i32, [vp, sz, cp, cp, cp]),
        "bpp_msm_segmented": (i32, [vp, sz, vp, cp, cp, cp]),
        "bpp_msm_plan_create": (i32, [vp, sz, cp, i32, P(vp)]),
        "bpp_msm_plan_set_scalars": (i32, [vp, cp]),
        "bpp_msm_plan_run": (i32, [vp, cp]),
        "bpp_msm_plan_window_bits": (i32, [vp]),
        "bpp_msm_window_bits": (i32, [C.c_size_t, C.c_size_t]),
        "bpp_msm_plan_destroy": (None, [vp]),
        "bpp_gens_create": (i32, [vp, i32, i32, i32, P(vp)]),
        "bpp_gens_destroy": (None, [vp]),
        "bpp_gens_get": (i32, [vp, i32, sz, cp]),
        "bpp_gens_fixed_base_msm": (i32, [vp, sz, sz, cp, vp, cp]),
        "bpp_pedersen_commit_batch": (i32, [vp, sz, vp, cp, i32, cp]),
        "bpp_verify_chunks": (i32, [vp, P(VerifyArgs), vp, vp, vp]),
        "bpp_vbatch_create": (i32, [vp, P(VerifyArgs), P(vp)]),
        "bpp_vbatch_create_multi": (i32, [vp, sz, vp, P(vp)]),
        "bpp_vbatch_call_count": (sz, [vp]),
        "bpp_vbatch_transcripts_call": (i32, [vp, sz, vp]),
        "bpp_verify_chunks_ch": (i32, [vp, P(VerifyArgs), P(VerifyChallenges), vp, vp, vp]),
        "bpp_vbatch_run": (i32, [vp, vp, vp, vp]),
        "bpp_vbatch_run_multi": (i32, [vp, vp, vp, vp]),
        "bpp_gens_create_with_bases": (i32, [vp, i32, i32, i32, cp, cp, P(vp)]),
        "bpp_vqueue_create": (i32, [i32, i32, i32, i32, cp, cp, i32, i32, i32, P(vp)]),
        "bpp_vqueue_destroy": (None, [vp]),
        "bpp_vqueue_submit": (i32, [vp, P(VerifyArgs), vp, vp, vp, P(C.c_uint64)]),
        "bpp_vqueue_wait": (i32, [vp, C.c_uint64]),
        "bpp_vqueue_verify": (i32, [vp, P(VerifyArgs), vp, vp, vp]),
        "bpp_vqueue_stats": (i32, [vp, P(C.c_uint64)]),
        "bpp_host_alloc": (i32, [C.c_size_t, P(vp)]),
        "bpp_host_free": (None, [vp]),
        "bpp_vqueue_lanes": (i32, [vp]),
        "bpp_vqueue_lane_ms": (i32, [vp, P(C.c_double)]),
        "bpp_vqueue_set_device_weights": (i32, [vp, i32]),
        "bpp_vqueue_set_merged_check": (i32, [vp, i32]),
        "bpp_ctx_set_merged_check": (i32, [vp, i32]),
        "bpp_ctx_merged_fallbacks": (C.c_uint64, [vp]),
        "bpp_vbatch_transcripts": (i32, [vp, vp]),
        "bpp_vbatch_destroy": (None, [vp]),
        "bpp_ctx_set_replay_mode": (i32, [vp, i32]),
        "bpp_ctx_set_graphs": (i32, [vp, i32]),
        "bpp_ctx_set_throughput_mode": (i32, [vp, i32]),
        "bpp_ctx_graph_launch_count": (C.c_uint64, [vp]),
        "bpp_ctx_set_test_hooks": (i32, [vp, C.c_uint32]),
        "bpp_proof_check_bytes": (i32, [cp, sz, P(i32), P(i32)]),
        "bpp_proof_size": (sz, [i32, i32]),
        "bpp_prove_batch": (i32, [vp, P(ProveArgs), vp, sz, vp]),
        "bpp_transcript_new": (None, [cp, sz, cp]),
        "bpp_transcript_append_message": (None, [cp, cp, sz, cp, sz]),
        "bpp_transcript_challenge_bytes": (None, [cp, cp, sz, cp, sz]),
        "bpp_hash_sha3_512": (None, [cp, sz, cp]),
        "bpp_hash_shake256": (None, [cp, sz, cp, sz]),
        "bpp_hash_blake2b_nonce_bytes": (i32, [cp, sz, cp, sz, cp]),
        "bpp_scalar_from_wide": (None, [cp, cp]),
        "bpp_microbench": (i32, [vp, i32, i32, P(C.c_double), P(C.c_double)]),
    }
    for name, (res, args) in sig.items():
        if not hasattr(l, name):
            continue  # optional while the library grows; test_abi_exports checks the header against the .so
        f = getattr(l, name)
        f.restype, f.argtypes = res, args
    _lib = l
    return l
