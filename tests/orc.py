"""ctypes binding of the CPU oracle (oracle/, test infrastructure only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product package never does.
"""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "libbpp_oracle.so")

OK, VERIFICATION_FAILED, INVALID_ARGUMENT, INVALID_LENGTH, INVALID_BLAKE2B, SIZE_OVERFLOW = range(6)
RECOVER_ONLY, RECOVER_AND_VERIFY, VERIFY_ONLY = 0, 1, 2
TRANSCRIPT_BYTES = 203
MAX_ROUNDS = 32
L = 2**252 + 27742317777372353535851937790883648493
P = 2**255 - 19


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(os.path.join(ORACLE_DIR, f)) > os.path.getmtime(LIB_PATH)
        for f in os.listdir(ORACLE_DIR)
        if f.endswith((".c", ".h"))
    ):
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
    return LIB_PATH


class Proof(C.Structure):
    _fields_ = [
        ("extension_degree", C.c_int32),
        ("n_d1", C.c_int32),
        ("n_li", C.c_int32),
        ("n_ri", C.c_int32),
        ("d1", (C.c_uint8 * 32) * 6),
        ("a", C.c_uint8 * 32),
        ("a1", C.c_uint8 * 32),
        ("b", C.c_uint8 * 32),
        ("r1", C.c_uint8 * 32),
        ("s1", C.c_uint8 * 32),
        ("li", (C.c_uint8 * 32) * MAX_ROUNDS),
        ("ri", (C.c_uint8 * 32) * MAX_ROUNDS),
    ]

    def copy(self):
        q = Proof()
        C.memmove(C.byref(q), C.byref(self), C.sizeof(Proof))
        return q


class Statement(C.Structure):
    _fields_ = [
        ("params", C.c_void_p),
        ("m", C.c_int32),
        ("commitments", C.c_void_p),
        ("min_values", C.c_void_p),
        ("min_present", C.c_void_p),
        ("seed_nonce", C.c_void_p),
        ("n_min", C.c_int32),
    ]


class Witness(C.Structure):
    _fields_ = [
        ("n_openings", C.c_int32),
        ("values", C.c_void_p),
        ("blindings", C.c_void_p),
        ("r_len", C.c_int32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    l = C.CDLL(LIB_PATH)
    vp, sz, u8p = C.c_void_p, C.c_size_t, C.c_char_p
    l.orc_params_new.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    l.orc_params_free.argtypes = [vp]
    l.orc_params_point.argtypes = [vp, C.c_int, sz, u8p]
    l.orc_commit.argtypes = [vp, C.c_uint64, u8p, C.c_int, u8p]
    l.orc_statement_check.argtypes = [C.POINTER(Statement)]
    l.orc_transcript_new.argtypes = [u8p, sz, u8p]
    l.orc_transcript_append_message.argtypes = [u8p, u8p, u8p, sz]
    l.orc_transcript_challenge_bytes.argtypes = [u8p, u8p, u8p, sz]
    l.orc_rng_chacha12_seed_from_u64.argtypes = [C.c_uint64]
    l.orc_rng_chacha12_seed_from_u64.restype = vp
    l.orc_rng_null.restype = vp
    l.orc_rng_chacha12_from_seed.argtypes = [u8p]
    l.orc_rng_chacha12_from_seed.restype = vp
    l.orc_rng_buffer.argtypes = [u8p, sz]
    l.orc_rng_buffer.restype = vp
    l.orc_rng_fill.argtypes = [vp, u8p, sz]
    l.orc_rng_next_u64.argtypes = [vp]
    l.orc_rng_next_u64.restype = C.c_uint64
    l.orc_rng_free.argtypes = [vp]
    l.orc_random_not_zero.argtypes = [vp, u8p]
    l.orc_prove.argtypes = [u8p, C.POINTER(Statement), C.POINTER(Witness), vp, C.POINTER(Proof)]
    l.orc_proof_to_bytes.argtypes = [C.POINTER(Proof), u8p, sz, C.POINTER(sz)]
    l.orc_proof_from_bytes.argtypes = [u8p, sz, C.POINTER(Proof)]
    l.orc_verify_batch.argtypes = [u8p, sz, C.POINTER(Statement), sz, C.POINTER(Proof), sz, C.c_int, u8p, u8p, C.POINTER(sz)]
    l.orc_ristretto_decode_encode.argtypes = [u8p, u8p]
    l.orc_ristretto_from_uniform.argtypes = [u8p, u8p]
    l.orc_ristretto_add.argtypes = [u8p, u8p, u8p]
    l.orc_ristretto_scalarmult.argtypes = [u8p, u8p, u8p]
    l.orc_msm.argtypes = [u8p, u8p, sz, C.c_int, u8p]
    for f in ("orc_sc_mul", "orc_sc_add", "orc_sc_sub", "orc_fe_mul"):
        getattr(l, f).argtypes = [u8p, u8p, u8p]
    for f in ("orc_sc_invert", "orc_sc_from_wide", "orc_fe_invert"):
        getattr(l, f).argtypes = [u8p, u8p]
    l.orc_sc_is_canonical.argtypes = [u8p]
    l.orc_fe_sqrt_ratio_i.argtypes = [u8p, u8p, u8p]
    l.orc_sha3_512.argtypes = [u8p, sz, u8p]
    l.orc_shake256.argtypes = [u8p, sz, u8p, sz]
    l.orc_blake2b_nonce.argtypes = [u8p, u8p, C.c_int, C.c_uint32, C.c_int, C.c_uint32, u8p]
    l.orc_keccak_f1600.argtypes = [C.POINTER(C.c_uint64)]
    l.orc_verifier_weights.argtypes = [u8p, sz, u8p]
    l.orc_verifier_weights.restype = None
    l.orc_verify_chunks_mt.argtypes = [u8p, C.POINTER(Statement), C.POINTER(Proof), C.POINTER(sz), sz, C.c_int, C.c_int, C.POINTER(C.c_int32)]
    l.orc_verify_chunks_mt.restype = C.c_double
    _lib = l
    return l


# ---------------------------------------------------------------- convenience layer
def sc_bytes(x):
    return int(x % L).to_bytes(32, "little")


def sc_int(b):
    return int.from_bytes(bytes(b), "little")


class Params:
    """RangeParameters::init(bit_length, max_aggregation, create_pedersen_gens_with_extension_degree(ext))"""

    def __init__(self, bit_length, max_aggregation, ext):
        h = C.c_void_p()
        rc = lib().orc_params_new(bit_length, max_aggregation, ext, C.byref(h))
        if rc:
            raise OracleError(rc)
        self.h, self.bit_length, self.max_aggregation, self.ext = h, bit_length, max_aggregation, ext

    def point(self, which, index=0):
        out = C.create_string_buffer(32)
        rc = lib().orc_params_point(self.h, which, index, out)
        if rc:
            raise OracleError(rc)
        return out.raw

    def commit(self, value, blindings):
        out = C.create_string_buffer(32)
        rc = lib().orc_commit(self.h, value, b"".join(sc_bytes(b) for b in blindings), len(blindings), out)
        if rc:
            raise OracleError(rc)
        return out.raw

    def __del__(self):
        try:
            lib().orc_params_free(self.h)
        except Exception:
            pass


class OracleError(Exception):
    def __init__(self, code):
        super().__init__("oracle error %d" % code)
        self.code = code


class Rng:
    def __init__(self, kind="chacha", seed=0, data=b""):
        l = lib()
        self._data = data
        if kind == "chacha":
            self.h = l.orc_rng_chacha12_seed_from_u64(seed)
        elif kind == "chacha_seed":
            self.h = l.orc_rng_chacha12_from_seed(data)
        elif kind == "null":
            self.h = l.orc_rng_null()
        else:
            self.h = l.orc_rng_buffer(self._data, len(self._data))

    def fill(self, n):
        out = C.create_string_buffer(n)
        lib().orc_rng_fill(self.h, out, n)
        return out.raw

    def next_u64(self):
        return lib().orc_rng_next_u64(self.h)

    def random_not_zero(self):
        out = C.create_string_buffer(32)
        lib().orc_random_not_zero(self.h, out)
        return sc_int(out.raw)

    def __del__(self):
        try:
            lib().orc_rng_free(self.h)
        except Exception:
            pass


class St:
    """RangeStatement as plain data; keeps the ctypes buffers alive."""

    def __init__(self, params, commitments, min_values, seed_nonce=None, check=True):
        self.params = params
        self.m = len(commitments)
        self.commitments = list(commitments)
        self.min_values = list(min_values)
        self.seed_nonce = seed_nonce
        self._c = C.create_string_buffer(b"".join(commitments), 32 * max(1, self.m))
        nm = len(min_values)
        self._mv = (C.c_uint64 * max(1, nm))(*[(v or 0) for v in min_values])
        self._mp = (C.c_uint8 * max(1, nm))(*[0 if v is None else 1 for v in min_values])
        self._sn = C.create_string_buffer(sc_bytes(seed_nonce), 32) if seed_nonce is not None else None
        self.c = Statement(
            params.h.value, self.m, C.addressof(self._c), C.addressof(self._mv), C.addressof(self._mp),
            C.addressof(self._sn) if self._sn is not None else None, nm,
        )
        if check:
            rc = lib().orc_statement_check(C.byref(self.c))
            if rc:
                raise OracleError(rc)


class Wit:
    def __init__(self, values, blindings):
        """blindings: list (per opening) of lists of ints"""
        self.values = list(values)
        self.blindings = [list(b) for b in blindings]
        n = len(values)
        r_len = len(blindings[0]) if blindings else 0
        self._v = (C.c_uint64 * max(1, n))(*values)
        self._b = C.create_string_buffer(b"".join(sc_bytes(x) for b in blindings for x in b), max(1, 32 * n * r_len))
        self.c = Witness(n, C.addressof(self._v), C.addressof(self._b), r_len)


def transcript_new(label):
    out = C.create_string_buffer(TRANSCRIPT_BYTES)
    lib().orc_transcript_new(label, len(label), out)
    return out.raw


def prove(transcript, st, wit, rng):
    """returns (rc, Proof, advanced_transcript)"""
    t = C.create_string_buffer(transcript, TRANSCRIPT_BYTES)
    pr = Proof()
    rc = lib().orc_prove(t, C.byref(st.c), C.byref(wit.c), rng.h, C.byref(pr))
    return rc, pr, t.raw


def proof_to_bytes(pr):
    out = C.create_string_buffer(1 + 32 * (6 + 5 + 2 * MAX_ROUNDS))
    n = C.c_size_t()
    rc = lib().orc_proof_to_bytes(C.byref(pr), out, len(out), C.byref(n))
    assert rc == 0
    return out.raw[: n.value]


def proof_from_bytes(b):
    pr = Proof()
    rc = lib().orc_proof_from_bytes(b, len(b), C.byref(pr))
    return rc, pr


def verify_batch(transcripts, statements, proofs, action):
    """returns (rc, masks) with masks = list of None | list[int]"""
    n = len(statements)
    nt = len(transcripts)
    tb = C.create_string_buffer(b"".join(transcripts), max(1, TRANSCRIPT_BYTES * nt))
    sa = (Statement * max(1, n))(*[s.c for s in statements])
    npf = len(proofs)
    pa = (Proof * max(1, npf))(*proofs)
    ext = statements[0].params.ext if n else 1
    masks = C.create_string_buffer(max(1, 256 * 6 * 32))
    present = C.create_string_buffer(max(1, 256))
    nres = C.c_size_t()
    rc = lib().orc_verify_batch(tb, nt, sa, n, pa, npf, action, masks, present, C.byref(nres))
    out = []
    if rc == 0:
        for i in range(nres.value):
            if present.raw[i]:
                out.append([sc_int(masks.raw[(i * ext + k) * 32 : (i * ext + k + 1) * 32]) for k in range(ext)])
            else:
                out.append(None)
    return rc, out


# ---------------------------------------------------------------- libsodium (independent group oracle)
_sodium = None


def sodium():
    global _sodium
    if _sodium is None:
        import glob
        import sys

        cands = []
        for sp in sys.path:
            cands += glob.glob(os.path.join(sp, "pyzmq.libs", "libsodium-*.so*"))
        cands += glob.glob("/usr/lib/x86_64-linux-gnu/libsodium.so*")
        for c in cands:
            try:
                l = C.CDLL(c)
                l.crypto_core_ristretto255_from_hash
                _sodium = l
                break
            except (OSError, AttributeError):
                continue
        if _sodium is None:
            _sodium = False
    return _sodium or None
