"""Host-side mirror of the reference's public API for the accelerated path, over the C ABI.

Same names, argument meaning and error behaviour as tari_bulletproofs_plus 0.4.1:
  RangeParameters::init            /root/reference/src/range_parameters.rs:32-58
  RangeStatement::init             /root/reference/src/range_statement.rs:36-73
  RangeWitness / CommitmentOpening /root/reference/src/range_witness.rs:15-41, src/commitment_opening.rs:15-38
  ExtendedMask                     /root/reference/src/extended_mask.rs:15-41
  PedersenGens::commit             /root/reference/src/generators/pedersen_gens.rs:112-122
  RangeProof::{verify_batch, from_bytes, to_bytes, prove_with_rng}  /root/reference/src/range_proof.rs:232-608,712-752,1120-1257
  merlin::Transcript               (re-exported by the reference, src/lib.rs)
ProofError variants map to EngineError.code 1..5 (src/errors.rs:12-28).
Points cross the boundary as 32-byte Ristretto encodings, scalars as python ints (mod l).
"""
import ctypes as C
import enum

from . import _ffi
from . import Engine, Gens, EngineError, L, _chk, _u64arr

ProofError = EngineError
MAX_RANGE_PROOF_BIT_LENGTH = 64      # range_proof.rs:71
MAX_RANGE_PROOF_BATCH_SIZE = 256     # range_proof.rs:76
SERIALIZED_ELEMENT_SIZE = 32         # range_proof.rs:85


class VerifyAction(enum.IntEnum):
    RecoverOnly = 0
    RecoverAndVerify = 1
    VerifyOnly = 2


class ExtensionDegree(enum.IntEnum):
    DefaultPedersen = 1
    AddOneBasePoint = 2
    AddTwoBasePoints = 3
    AddThreeBasePoints = 4
    AddFourBasePoints = 5
    AddFiveBasePoints = 6

    @classmethod
    def try_from(cls, v):
        try:
            return cls(v)
        except ValueError:
            raise EngineError(_ffi.INVALID_ARGUMENT, "Extension degree not valid")


def _sc(x):
    return int(x % L).to_bytes(32, "little")


class Transcript:
    """merlin::Transcript on the 203-byte STROBE state the C ABI exchanges."""

    def __init__(self, label=None, state=None):
        if state is not None:
            self.state = bytes(state)
        else:
            out = C.create_string_buffer(_ffi.TRANSCRIPT_BYTES)
            _ffi.lib().bpp_transcript_new(label, len(label), out)
            self.state = out.raw

    def clone(self):
        return Transcript(state=self.state)

    def append_message(self, label, msg):
        buf = C.create_string_buffer(self.state, _ffi.TRANSCRIPT_BYTES)
        _ffi.lib().bpp_transcript_append_message(buf, label, len(label), msg, len(msg))
        self.state = buf.raw

    def challenge_bytes(self, label, n):
        buf = C.create_string_buffer(self.state, _ffi.TRANSCRIPT_BYTES)
        out = C.create_string_buffer(n)
        _ffi.lib().bpp_transcript_challenge_bytes(buf, label, len(label), out, n)
        self.state = buf.raw
        return out.raw


class ExtendedMask:
    def __init__(self, blindings):
        self._b = list(blindings)

    @classmethod
    def assign(cls, extension_degree, blindings):
        if not blindings or len(blindings) != int(extension_degree):
            raise EngineError(_ffi.INVALID_LENGTH, "Extended mask length must correspond to the extension degree")
        return cls(blindings)

    def blindings(self):
        if not self._b:
            raise EngineError(_ffi.INVALID_LENGTH, "Extended mask length cannot be 0")
        return list(self._b)

    def __eq__(self, o):
        return isinstance(o, ExtendedMask) and self._b == o._b

    def __repr__(self):
        return "ExtendedMask(%r)" % (self._b,)


class CommitmentOpening:
    def __init__(self, v, r):
        if not r:
            raise EngineError(_ffi.INVALID_LENGTH, "Extended mask length cannot be 0")
        self.v, self.r = int(v), [int(x) % L for x in r]


class RangeWitness:
    def __init__(self, openings):
        self.openings = list(openings)
        self.extension_degree = ExtensionDegree.try_from(len(self.openings[0].r)) if self.openings else None

    @classmethod
    def init(cls, openings):
        if not openings:
            raise EngineError(_ffi.INVALID_LENGTH, "Vector openings_vec length cannot be 0")
        n = len(openings[0].r)
        if any(len(o.r) != n for o in openings):
            raise EngineError(_ffi.INVALID_LENGTH, "Extended mask length must be consistent")
        return cls(openings)


class RangeParameters:
    """bp_gens + pc_gens, resident on the device (bpp_gens handle)."""

    def __init__(self, gens):
        self.gens = gens

    @classmethod
    def init(cls, engine, bit_length, aggregation_factor, extension_degree):
        return cls(Gens(engine, bit_length, aggregation_factor, int(extension_degree)))

    def bit_length(self):
        return self.gens.bit_length

    def max_aggregation_factor(self):
        return self.gens.max_aggregation

    def extension_degree(self):
        return ExtensionDegree(self.gens.extension_degree)

    def h_base(self):
        return self.gens.point(0)

    def g_bases(self):
        return [self.gens.point(1, k) for k in range(self.gens.extension_degree)]

    def gi_base(self, i):
        return self.gens.point(2, i)

    def hi_base(self, i):
        return self.gens.point(3, i)

    def commit(self, value, blindings):
        """PedersenGens::commit"""
        return self.gens.commit_batch([value], [list(blindings)])[0]


class RangeStatement:
    def __init__(self, generators, commitments, minimum_value_promises, seed_nonce):
        self.generators = generators
        self.commitments = list(commitments)
        self.minimum_value_promises = list(minimum_value_promises)
        self.seed_nonce = seed_nonce

    @classmethod
    def init(cls, generators, commitments, minimum_value_promises, seed_nonce=None):
        n = len(commitments)
        if n == 0 or n & (n - 1):
            raise EngineError(_ffi.INVALID_ARGUMENT, "Number of commitments must be a power of two")
        if len(minimum_value_promises) != n:
            raise EngineError(_ffi.INVALID_ARGUMENT, "Incorrect number of minimum value promises")
        if generators.max_aggregation_factor() < n:
            raise EngineError(_ffi.INVALID_ARGUMENT, "Not enough generators for this statement")
        if seed_nonce is not None and n > 1:
            raise EngineError(_ffi.INVALID_ARGUMENT, "Mask recovery is not supported with an aggregated statement")
        return cls(generators, commitments, minimum_value_promises, seed_nonce)


class RangeProof:
    """A serialised proof (range_proof.rs:58-68 in its to_bytes layout, :1120-1150)."""

    def __init__(self, data, extension_degree, rounds):
        self._bytes = bytes(data)
        self._ext, self._rounds = extension_degree, rounds

    @classmethod
    def from_bytes(cls, data):
        ext, rounds = C.c_int32(), C.c_int32()
        rc = _ffi.lib().bpp_proof_check_bytes(bytes(data), len(data), C.byref(ext), C.byref(rounds))
        if rc:
            raise EngineError(rc, "Invalid serialized proof")
        return cls(data, ext.value, rounds.value)

    def to_bytes(self):
        return self._bytes

    def extension_degree(self):
        return ExtensionDegree(self._ext)

    @staticmethod
    def extension_degree_from_proof_bytes(data):
        if not data:
            raise EngineError(_ffi.INVALID_LENGTH, "Serialized proof bytes cannot be empty")
        return ExtensionDegree.try_from(data[0])

    # ---- field views (the reference's getters a(), a1(), b(), li(), ri(), r1(), s1(), d1())
    def _el(self, idx):
        o = 1 + 32 * idx
        return self._bytes[o:o + 32]

    def d1(self):
        return [int.from_bytes(self._el(k), "little") for k in range(self._ext)]

    def a(self):
        return self._el(self._ext)

    def a1(self):
        return self._el(self._ext + 1)

    def b(self):
        return self._el(self._ext + 2)

    def r1(self):
        return int.from_bytes(self._el(self._ext + 3), "little")

    def s1(self):
        return int.from_bytes(self._el(self._ext + 4), "little")

    def li(self):
        return [self._el(self._ext + 5 + 2 * j) for j in range(self._rounds)]

    def ri(self):
        return [self._el(self._ext + 6 + 2 * j) for j in range(self._rounds)]

    # ---- proving
    @staticmethod
    def rng_bytes_needed(generators, aggregation):
        """bytes the external RNG contributes to one proof: 32 per TranscriptRng rebuild (transcripts.rs:185-194)"""
        n = generators.bit_length() * aggregation
        return 32 * ((n - 1).bit_length() + 3)

    @staticmethod
    def prove_batch(transcripts, statements, witnesses, rng_bytes):
        """P x RangeProof::prove_with_rng for statements of one shape, in lock-step on the device.
        rng_bytes: per proof, the bytes its external RNG delivers (rng_bytes_needed each).
        Returns a list with a RangeProof or a ProofError per proof; transcripts are advanced in place."""
        P = len(statements)
        if not (len(transcripts) == len(witnesses) == len(rng_bytes) == P) or P == 0:
            raise EngineError(_ffi.INVALID_ARGUMENT, "prove_batch: length mismatch")
        params = statements[0].generators
        ext, m = params.gens.extension_degree, len(statements[0].commitments)
        for s, w in zip(statements, witnesses):
            if len(s.commitments) != m or s.generators.gens is not params.gens:
                raise EngineError(_ffi.INVALID_ARGUMENT, "prove_batch: statements must share one shape and one parameter set")
        results = [None] * P
        live = []
        for i, (s, w) in enumerate(zip(statements, witnesses)):
            # range_proof.rs:248-260: witness / statement shape checks
            if len(w.openings) != len(s.commitments):
                results[i] = EngineError(_ffi.INVALID_LENGTH, "Witness openings and statement commitments do not match!")
            elif w.extension_degree != params.extension_degree():
                results[i] = EngineError(_ffi.INVALID_LENGTH, "Witness and statement extension degrees do not match!")
            else:
                live.append(i)
        if live:
            pk = _PackedProve(params, [transcripts[i] for i in live], [statements[i] for i in live], [witnesses[i] for i in live],
                              [rng_bytes[i] for i in live])
            import time as _time
            _t0 = _time.perf_counter()
            pk.run()
            RangeProof.last_prove_call_ms = (_time.perf_counter() - _t0) * 1e3      # the C-ABI call alone (bench)
            for i, res in zip(live, pk.results()):
                results[i] = res
        return results

    @staticmethod
    def prove_with_rng(transcript, statement, witness, rng):
        """RangeProof::prove_with_rng; `rng` is any object with fill(n) -> bytes (the external CryptoRng)."""
        need = RangeProof.rng_bytes_needed(statement.generators, len(statement.commitments))
        res = RangeProof.prove_batch([transcript], [statement], [witness], [rng.fill(need)])[0]
        if isinstance(res, EngineError):
            raise res
        return res

    # ---- verification
    @staticmethod
    def verify_batch(transcripts, statements, proofs, action):
        """RangeProof::verify_batch: returns Vec<Option<ExtendedMask>> (min(len, 256) entries) or raises ProofError.
        `transcripts` (list of Transcript) are advanced in place like `&mut [Transcript]`."""
        if not statements or not proofs or not transcripts:
            raise EngineError(_ffi.INVALID_ARGUMENT, "Range statements or proofs length empty")
        if len(statements) != len(proofs):
            raise EngineError(_ffi.INVALID_ARGUMENT, "Range statements and proofs length mismatch")
        if len(transcripts) != len(statements):
            raise EngineError(_ffi.INVALID_ARGUMENT, "Range statements and transcripts length mismatch")
        status, masks = verify_chunks(statements[0].generators, [(transcripts, statements, proofs)], action)
        if status[0]:
            raise EngineError(status[0], "verify_batch")
        return masks[0][:MAX_RANGE_PROOF_BATCH_SIZE]


class _PackedProve:
    """Flat host buffers for bpp_prove_args: P statements of one shape (kept alive for the duration of the call)."""

    def __init__(self, params, transcripts, statements, witnesses, rng_bytes):
        ext, m = params.gens.extension_degree, len(statements[0].commitments)
        n = len(statements)
        self.params, self.transcripts, self.n, self.ext = params, transcripts, n, ext
        need = RangeProof.rng_bytes_needed(params, m)
        self.rounds = (params.bit_length() * m - 1).bit_length()
        self.plen = _ffi.lib().bpp_proof_size(ext, self.rounds)
        for r in rng_bytes:
            if len(r) < need:
                raise EngineError(_ffi.INVALID_LENGTH, "rng_bytes: %d bytes needed per proof" % need)
        self.commits = C.create_string_buffer(b"".join(c for s in statements for c in s.commitments), 32 * n * m)
        self.values = _u64arr([o.v for w in witnesses for o in w.openings])
        self.blind = C.create_string_buffer(b"".join(_sc(r) for w in witnesses for o in w.openings for r in o.r), 32 * n * m * ext)
        self.minv = _u64arr([(v or 0) for s in statements for v in s.minimum_value_promises])
        self.minp = (C.c_uint8 * (n * m))(*[0 if v is None else 1 for s in statements for v in s.minimum_value_promises])
        self.seeds = C.create_string_buffer(b"".join(_sc(s.seed_nonce) if s.seed_nonce is not None else bytes(32) for s in statements), 32 * n)
        self.seedp = (C.c_uint8 * n)(*[0 if s.seed_nonce is None else 1 for s in statements])
        self.t_init = b"".join(t.state for t in transcripts)
        self.tbuf = C.create_string_buffer(self.t_init, _ffi.TRANSCRIPT_BYTES * n)
        self.rbuf = C.create_string_buffer(b"".join(bytes(r[:need]) for r in rng_bytes), need * n)
        self.args = _ffi.ProveArgs(n, m, C.addressof(self.commits), C.addressof(self.values), C.addressof(self.blind), C.addressof(self.minv),
                                   C.addressof(self.minp), C.addressof(self.seeds), C.addressof(self.seedp), C.addressof(self.tbuf),
                                   C.addressof(self.rbuf), need)
        self.out = C.create_string_buffer(self.plen * n)
        self.status = (C.c_int32 * n)()

    def reset_transcripts(self):
        C.memmove(self.tbuf, self.t_init, len(self.t_init))

    def run(self):
        _chk(self.params.gens.engine, _ffi.lib().bpp_prove_batch(self.params.gens.h, C.byref(self.args), self.out, self.plen, self.status))

    def results(self):
        """RangeProof or ProofError per statement; the transcripts are advanced in place"""
        res = []
        for k, t in enumerate(self.transcripts):
            t.state = self.tbuf.raw[_ffi.TRANSCRIPT_BYTES * k: _ffi.TRANSCRIPT_BYTES * (k + 1)]
            if self.status[k]:
                res.append(EngineError(self.status[k], "prove_with_rng"))
            else:
                res.append(RangeProof(self.out.raw[self.plen * k: self.plen * (k + 1)], self.ext, self.rounds))
        return res


class _Packed:
    """Flat host buffers for bpp_verify_args (kept alive for the duration of a call)."""

    def __init__(self, params, calls, action, pinned=False):
        """pinned: the serialised proofs go into page-locked memory (bpp_host_alloc): the engine then uploads them without a staging copy"""
        self.params = params
        self._pinned = None
        ext = params.gens.extension_degree
        chunk_offsets, proof_offsets, commit_offsets = [0], [0], [0]
        pbytes, commits, minv, minp, seeds, seedp, tstates = [], [], [], [], [], [], []
        self.transcripts = []
        for transcripts, statements, proofs in calls:
            if not (len(transcripts) == len(statements) == len(proofs)):
                raise EngineError(_ffi.INVALID_ARGUMENT, "Range statements and proofs length mismatch")
            for t, s, p in zip(transcripts, statements, proofs):
                g = s.generators.gens
                # verify_statements_and_generators_consistency (range_proof.rs:637-705): every statement must carry the same G, H,
                # bit length, extension degree and Gi / Hi vectors.  Device-resident tables are compared by handle first; statements
                # built on ANOTHER handle are compared by parameters and by the encodings of their Pedersen bases (Gi / Hi are a
                # function of (bit length, aggregation factor) alone: generators/bulletproof_gens.rs:83-112)
                if g is not params.gens:
                    if (g.bit_length, g.extension_degree) != (params.gens.bit_length, ext):
                        raise EngineError(_ffi.INVALID_ARGUMENT, "Inconsistent generators in batch statement")
                    if hasattr(g, "point") and hasattr(params.gens, "point"):
                        if g.point(0) != params.gens.point(0) or any(g.point(1, k) != params.gens.point(1, k) for k in range(ext)):
                            raise EngineError(_ffi.INVALID_ARGUMENT, "Inconsistent generator point in batch statement")
                b = p.to_bytes()
                pbytes.append(b)
                proof_offsets.append(proof_offsets[-1] + len(b))
                commits.extend(s.commitments)
                commit_offsets.append(commit_offsets[-1] + len(s.commitments))
                minv.extend((v or 0) for v in s.minimum_value_promises)
                minp.extend(0 if v is None else 1 for v in s.minimum_value_promises)
                seeds.append(_sc(s.seed_nonce) if s.seed_nonce is not None else bytes(32))
                seedp.append(0 if s.seed_nonce is None else 1)
                tstates.append(t.state)
                self.transcripts.append(t)
            chunk_offsets.append(len(pbytes))
        n = len(pbytes)
        self.n, self.k, self.ext = n, len(calls), ext
        self.chunk_offsets = _u64arr(chunk_offsets)
        self.proof_offsets = _u64arr(proof_offsets)
        self.commit_offsets = _u64arr(commit_offsets)
        raw = b"".join(pbytes)
        if pinned and raw:
            ptr = C.c_void_p()
            rc = _ffi.lib().bpp_host_alloc(len(raw), C.byref(ptr))
            if rc:
                raise EngineError(rc, "bpp_host_alloc")
            self._pinned = ptr
            C.memmove(ptr, raw, len(raw))
            self.proof_bytes = (C.c_char * len(raw)).from_address(ptr.value)
        else:
            self.proof_bytes = C.create_string_buffer(raw, max(1, proof_offsets[-1]))
        self.commitments = C.create_string_buffer(b"".join(commits), max(1, 32 * len(commits)))
        self.min_values = _u64arr(minv)
        self.min_present = (C.c_uint8 * max(1, len(minp)))(*minp)
        self.seeds = C.create_string_buffer(b"".join(seeds), max(1, 32 * n))
        self.seed_present = (C.c_uint8 * max(1, n))(*seedp)
        self.tbuf = C.create_string_buffer(b"".join(tstates), max(1, _ffi.TRANSCRIPT_BYTES * n))
        a = _ffi.VerifyArgs()
        a.n_proofs, a.n_chunks = n, self.k
        a.chunk_offsets = C.addressof(self.chunk_offsets)
        a.proof_bytes = C.addressof(self.proof_bytes)
        a.proof_offsets = C.addressof(self.proof_offsets)
        a.commitments32 = C.addressof(self.commitments)
        a.commit_offsets = C.addressof(self.commit_offsets)
        a.min_values = C.addressof(self.min_values)
        a.min_present = C.addressof(self.min_present)
        a.seed_nonces32 = C.addressof(self.seeds)
        a.seed_present = C.addressof(self.seed_present)
        a.transcripts = C.addressof(self.tbuf)
        a.action = int(action)
        self.args = a
        self.status = (C.c_int32 * max(1, self.k))()
        self.masks = C.create_string_buffer(max(1, 32 * n * ext))
        self.mask_present = C.create_string_buffer(max(1, n))

    def __del__(self):
        try:
            if self._pinned is not None:
                _ffi.lib().bpp_host_free(self._pinned)
                self._pinned = None
        except Exception:
            pass

    def results(self, update_transcripts=True):
        if update_transcripts:
            raw = self.tbuf.raw
            for i, t in enumerate(self.transcripts):
                t.state = raw[_ffi.TRANSCRIPT_BYTES * i: _ffi.TRANSCRIPT_BYTES * (i + 1)]
        status = [self.status[c] for c in range(self.k)]
        out = []
        for c in range(self.k):
            lo, hi = self.chunk_offsets[c], self.chunk_offsets[c + 1]
            row = []
            for i in range(lo, hi):
                if self.mask_present.raw[i]:
                    row.append(ExtendedMask([int.from_bytes(self.masks.raw[32 * (i * self.ext + k): 32 * (i * self.ext + k + 1)], "little")
                                             for k in range(self.ext)]))
                else:
                    row.append(None)
            out.append(row)
        return status, out


def verify_chunks(params, calls, action):
    """K independent RangeProof::verify_batch calls in one device pass.
    calls: list of (transcripts, statements, proofs).  Returns (status per call, masks per call)."""
    pk = _Packed(params, calls, action)
    rc = _ffi.lib().bpp_verify_chunks(params.gens.h, C.byref(pk.args), pk.status, pk.masks, pk.mask_present)
    _chk(params.gens.engine, rc)
    return pk.results()


class VerifyBatch:
    """Split form (bpp_vbatch_*): create = host Fiat-Shamir + upload, run = device work + verdict readback."""

    def __init__(self, params, calls, action):
        self.pk = _Packed(params, calls, action)
        self.params = params
        self.h = C.c_void_p()
        _chk(params.gens.engine, _ffi.lib().bpp_vbatch_create(params.gens.h, C.byref(self.pk.args), C.byref(self.h)))
        params.gens.engine.adopt(self)

    def run(self):
        pk = self.pk
        _chk(self.params.gens.engine, _ffi.lib().bpp_vbatch_run(self.h, pk.status, pk.masks, pk.mask_present))
        _chk(self.params.gens.engine, _ffi.lib().bpp_vbatch_transcripts(self.h, C.addressof(pk.tbuf)))
        return pk.results()

    def close(self):
        if self.h and self.params.gens.h and self.params.gens.engine.h:       # never touch a destroyed ctx / generator set
            _ffi.lib().bpp_vbatch_destroy(self.h)
        self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def verify_chunks_ch(params, calls, challenges, weights, action):
    """bpp_verify_chunks_ch: the caller keeps the Merlin transcripts and hands over, per proof, the challenges [y, z, e, e_0..]
    (list of ints per proof) and the batch weight (int).  calls as in verify_chunks (transcripts are ignored)."""
    pk = _Packed(params, calls, action)
    flat = [c for ch in challenges for c in ch]
    offs = [0]
    for ch in challenges:
        offs.append(offs[-1] + len(ch))
    cbuf = C.create_string_buffer(b"".join(_sc(c) for c in flat), max(1, 32 * len(flat)))
    wbuf = C.create_string_buffer(b"".join(_sc(w) for w in weights), max(1, 32 * len(weights)))
    obuf = _u64arr(offs)
    vc = _ffi.VerifyChallenges(C.addressof(cbuf), C.addressof(obuf), C.addressof(wbuf))
    rc = _ffi.lib().bpp_verify_chunks_ch(params.gens.h, C.byref(pk.args), C.byref(vc), pk.status, pk.masks, pk.mask_present)
    _chk(params.gens.engine, rc)
    return pk.results(update_transcripts=False)


class VerifyQueue:
    """bpp_vqueue: callers submit verify_batch calls, a few lanes verify whatever is waiting as ONE device pass each
    (include/bpp_b200.h, csrc/engine_queue.cpp).  Results are those of verify_chunks on each call alone."""

    def __init__(self, device, bit_length, max_aggregation, extension_degree, lanes=3, max_calls_per_pass=16, host_threads_per_lane=0,
                 h_base=None, g_bases=None, device_weights=False, merged_check=False):
        self.h = C.c_void_p()
        gb = b"".join(g_bases) if g_bases else None
        rc = _ffi.lib().bpp_vqueue_create(device, bit_length, max_aggregation, int(extension_degree), h_base, gb, lanes, max_calls_per_pass,
                                          host_threads_per_lane, C.byref(self.h))
        if rc:
            self.h = C.c_void_p()
            raise EngineError(rc, "bpp_vqueue_create")
        self.shape = _QueueShape(bit_length, max_aggregation, int(extension_degree))
        if device_weights:
            _ffi.lib().bpp_vqueue_set_device_weights(self.h, 1)
        if merged_check:
            _ffi.lib().bpp_vqueue_set_merged_check(self.h, 1)

    def submit(self, packed):
        """packed: a _Packed kept alive by the caller until wait() returns"""
        t = C.c_uint64()
        rc = _ffi.lib().bpp_vqueue_submit(self.h, C.byref(packed.args), packed.status, packed.masks, packed.mask_present, C.byref(t))
        if rc:
            raise EngineError(rc, "bpp_vqueue_submit")
        return t.value

    def wait(self, ticket):
        rc = _ffi.lib().bpp_vqueue_wait(self.h, ticket)
        if rc:
            raise EngineError(rc, "bpp_vqueue_wait")

    def pack(self, calls, action, pinned=False):
        return _Packed(self.shape, calls, action, pinned=pinned)

    def verify_many(self, batches, action=VerifyAction.VerifyOnly):
        """batches: list of `calls` (each a list of (transcripts, statements, proofs)); all submitted at once, then waited for.
        Returns [(status per call, masks per call)] in input order; transcripts are advanced in place."""
        pks = [self.pack(calls, action) for calls in batches]
        tickets = [self.submit(pk) for pk in pks]
        for t in tickets:
            self.wait(t)
        return [pk.results() for pk in pks]

    def lane_ms(self):
        arr = (C.c_double * 4)()
        _ffi.lib().bpp_vqueue_lane_ms(self.h, arr)
        return dict(zip(("waiting_for_calls", "building_passes", "running_passes", "handing_back"), [float(x) for x in arr]))

    def stats(self):
        arr = (C.c_uint64 * 5)()
        _ffi.lib().bpp_vqueue_stats(self.h, arr)
        return dict(zip(("passes", "calls", "proofs", "kernels", "graph_launches"), [int(x) for x in arr]))

    def close(self):
        if self.h:
            _ffi.lib().bpp_vqueue_destroy(self.h)
        self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _QueueShape:
    """what _Packed needs to know about a parameter set when there is no single Gens handle (the queue's lanes hold one each)"""

    class _G:
        pass

    def __init__(self, bit_length, max_aggregation, extension_degree):
        self.gens = self._G()
        self.gens.bit_length, self.gens.max_aggregation, self.gens.extension_degree = bit_length, max_aggregation, extension_degree

    def bit_length(self):
        return self.gens.bit_length

    def max_aggregation_factor(self):
        return self.gens.max_aggregation

    def extension_degree(self):
        return ExtensionDegree(self.gens.extension_degree)


class VerifierPool:
    """S independent verification lanes on ONE GPU: one bpp_ctx (stream pair, pooled workspace, generator tables) and one
    host thread per lane.  A single verify_batch call of a few hundred proofs is a chain of short dependent kernels that
    leaves most of the 148 SMs idle; independent calls issued from different lanes overlap on the device, and
    the host side of call i+1 (parsing, Fiat-Shamir weights, H2D) overlaps the device side of call i.  This is the C ABI's
    threading model ("one bpp_ctx per (thread, device); different ctxs are independent", include/bpp_b200.h) wrapped for
    Python callers; a Rust caller does the same with one ctx per worker thread.

    verify_many(batches) runs batches[i] on lane i % S and returns the per-batch (status, masks) in input order; results do
    not depend on S (every batch is verified by the same code path as RangeProof.verify_batch)."""

    def __init__(self, device, bit_length, max_aggregation, extension_degree, lanes=8, host_threads_per_lane=None, blocking_waits=None,
                 device_weights=False, merged_check=False):
        import os

        from . import Engine

        self.lanes = []
        per = host_threads_per_lane or max(1, (os.cpu_count() or 1) // max(1, lanes))
        if blocking_waits is None:          # spinning lane threads only pay while every one of them has a core to itself
            blocking_waits = lanes > 1 and lanes >= (os.cpu_count() or 1)
        for _ in range(lanes):
            eng = Engine(device)
            eng.set_host_threads(per)
            eng.set_merged_check(merged_check)
            eng.set_throughput_mode(2 if device_weights else 1 if blocking_waits else 0)   # lane threads sleep while their pass runs (bpp_ctx_set_throughput_mode)
            self.lanes.append((eng, RangeParameters.init(eng, bit_length, max_aggregation, extension_degree)))

    def __len__(self):
        return len(self.lanes)

    def _start_threads(self):
        """one persistent host thread per lane (created on first use): a job is (fn, n_items); lane li takes items li, li + S, ..."""
        import queue
        import threading

        self._jobs = [queue.SimpleQueue() for _ in self.lanes]
        self._done = queue.SimpleQueue()

        def loop(li):
            eng, params = self.lanes[li]
            S = len(self.lanes)
            while True:
                job = self._jobs[li].get()
                if job is None:
                    return
                fn, n_items, out = job
                err = None
                try:
                    for i in range(li, n_items, S):
                        out[i] = fn(li, eng, params, i)
                except BaseException as e:  # noqa: BLE001 - re-raised in the caller's thread
                    err = e
                self._done.put(err)

        self._threads = [threading.Thread(target=loop, args=(li,), daemon=True) for li in range(len(self.lanes))]
        for t in self._threads:
            t.start()

    def run(self, fn, n_items):
        """fn(lane_index, engine, params, item_index) for item_index in range(n_items); item i runs on lane i % S, every lane in
        its own (persistent) thread -- the C calls release the GIL; returns the results in item order, re-raises the first exception"""
        if not getattr(self, "_threads", None):
            self._start_threads()
        out = [None] * n_items
        active = min(len(self.lanes), n_items)
        for li in range(active):
            self._jobs[li].put((fn, n_items, out))
        errs = [e for e in (self._done.get() for _ in range(active)) if e is not None]
        if errs:
            raise errs[0]
        return out

    def verify_many(self, batches, action=VerifyAction.VerifyOnly):
        """batches: list of `calls` (each a list of (transcripts, statements, proofs), one entry per reference verify_batch
        call).  Returns [(status per call, masks per call)] in input order; transcripts are advanced in place."""
        return self.run(lambda li, eng, params, i: verify_chunks(params, batches[i], action), len(batches))

    def prove_many(self, jobs):
        """jobs: list of (transcripts, statements, witnesses, rng_bytes), each one RangeProof.prove_batch call (statements of one
        shape); job i runs on lane i % S.  Returns the per-job result lists in input order.  Proof bytes do not depend on S."""
        def one(li, eng, params, i):
            trs, sts, wits, rbs = jobs[i]
            pk = _PackedProve(params, trs, sts, wits, rbs)
            pk.run()
            return pk.results()
        return self.run(one, len(jobs))

    def launch_count(self):
        return sum(eng.launch_count for eng, _ in self.lanes)

    def close(self):
        for q in getattr(self, "_jobs", []):
            q.put(None)
        for t in getattr(self, "_threads", []):
            t.join(timeout=5)
        self._threads = []
        for eng, _ in self.lanes:
            eng.close()
        self.lanes = []
