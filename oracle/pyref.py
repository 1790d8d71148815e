"""pyref — an INDEPENDENT pure-Python (big-integer) restatement of tari_bulletproofs_plus 0.4.1's prover and batch verifier.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): nothing in the product imports it.  Its purpose is to remove the single
point of failure VERDICT r1 named: the C oracle (oracle/*.c) and the CUDA engine were written by the same hand from the same reading
of the reference.  This module was written separately, straight from the reference's Rust source and from the public specifications
of the crates it depends on, sharing NO code with oracle/*.c or the engine:
  group          RFC 9496 (ristretto255 decode / encode / one-way map), RFC 8032 (field, l)
  transcripts    STROBE v1.0.2 + Merlin v1.0 (merlin 3.0.0: Transcript, TranscriptRngBuilder, TranscriptRng), FIPS 202 (Keccak-f[1600])
  hashes         python hashlib for SHA3-512 / SHAKE256 / BLAKE2b (keyed + personalised)
  external RNG   rand_chacha 0.3.1 ChaCha12Rng over rand_core 0.6 (seed_from_u64 = PCG32 expansion, BlockRng word semantics)
  protocol       /root/reference/src/range_proof.rs:232-608 (prove_with_rng), :610-1065 (verify), :1120-1257 (bytes),
                 src/transcripts.rs:59-194, src/protocols/transcript_protocol.rs:39-79, src/protocols/scalar_protocol.rs:23-37,
                 src/utils/generic.rs:30-82, src/generators/{bulletproof_gens,generators_chain,pedersen_gens}.rs, src/ristretto.rs:67-112
Parity status: still "unpinned" against bytes produced by the Rust crate (no cargo in the image); what this adds is a second,
independent reading that must agree byte for byte with the C oracle (tests/test_pyref_vectors.py, tests/golden/pyref_vectors.json
made by tests/golden/make_pyref_vectors.py on the six shapes of oracle/rust_vectors).
Everything is slow and simple on purpose: points are affine-free extended coordinates over python ints, scalar multiplication is
plain double-and-add, an MSM is a shared-doubling Straus loop.
"""
import hashlib

P = 2**255 - 19
L = 2**252 + 27742317777372353535851937790883648493
D = (-121665 * pow(121666, P - 2, P)) % P
SQRT_M1 = pow(2, (P - 1) // 4, P)


def _is_neg(x):
    return (x % P) & 1


def _abs(x):
    x %= P
    return P - x if x & 1 else x


def sqrt_ratio_m1(u, v):
    """RFC 9496 section 4.2"""
    u %= P
    v %= P
    v3 = v * v % P * v % P
    v7 = v3 * v3 % P * v % P
    r = u * v3 % P * pow(u * v7 % P, (P - 5) // 8, P) % P
    check = v * r % P * r % P
    correct = check == u
    flipped = check == (-u) % P
    flipped_i = check == (-u) * SQRT_M1 % P
    if flipped or flipped_i:
        r = r * SQRT_M1 % P
    return (correct or flipped), _abs(r)


# constants of RFC 9496 section 4.1, derived rather than typed in (their signs are fixed by the RFC's decimal values, checked below)
ONE_MINUS_D_SQ = (1 - D * D) % P
D_MINUS_ONE_SQ = (D - 1) * (D - 1) % P
_ok, _r = sqrt_ratio_m1(-D - 1, 1)
SQRT_AD_MINUS_ONE = _r if _r == 25063068953384623474111414158702152701244531502492656460079210482610430750235 else P - _r
_ok2, INVSQRT_A_MINUS_D = sqrt_ratio_m1(1, -1 - D)
assert _ok and _ok2
assert SQRT_AD_MINUS_ONE == 25063068953384623474111414158702152701244531502492656460079210482610430750235
assert INVSQRT_A_MINUS_D == 54469307008909316920995813868745141605393597292927456921205312896311721017578
assert SQRT_M1 == 19681161376707505956807079304988542015446066515923890162744021073123829784752

IDENTITY = (0, 1, 1, 0)


def pt_add(p, q):
    """extended twisted Edwards, a = -1 (unified)"""
    x1, y1, z1, t1 = p
    x2, y2, z2, t2 = q
    a = (y1 - x1) * (y2 - x2) % P
    b = (y1 + x1) * (y2 + x2) % P
    c = 2 * D * t1 % P * t2 % P
    d = 2 * z1 * z2 % P
    e, f, g, h = b - a, d - c, d + c, b + a
    return (e * f % P, g * h % P, f * g % P, e * h % P)


def pt_neg(p):
    return ((-p[0]) % P, p[1], p[2], (-p[3]) % P)


def pt_mul(k, p):
    k %= L
    acc = IDENTITY
    for bit in range(k.bit_length() - 1, -1, -1):
        acc = pt_add(acc, acc)
        if (k >> bit) & 1:
            acc = pt_add(acc, p)
    return acc


def msm(scalars, points):
    """sum k_i * P_i, shared doublings"""
    ks = [k % L for k in scalars]
    top = max([k.bit_length() for k in ks] + [0])
    acc = IDENTITY
    for bit in range(top - 1, -1, -1):
        acc = pt_add(acc, acc)
        for k, p in zip(ks, points):
            if (k >> bit) & 1:
                acc = pt_add(acc, p)
    return acc


def pt_eq(p, q):
    """ristretto equality"""
    return (p[0] * q[1] - p[1] * q[0]) % P == 0 or (p[1] * q[1] - p[0] * q[0]) % P == 0


def decode(b):
    """RFC 9496 4.3.1; None on failure"""
    s = int.from_bytes(b, "little")
    if s >= P or s & 1:
        return None
    ss = s * s % P
    u1, u2 = (1 - ss) % P, (1 + ss) % P
    u2s = u2 * u2 % P
    v = (-(D * u1 % P * u1) - u2s) % P
    ok, inv = sqrt_ratio_m1(1, v * u2s % P)
    den_x = inv * u2 % P
    den_y = inv * den_x % P * v % P
    x = _abs(2 * s * den_x % P)
    y = u1 * den_y % P
    t = x * y % P
    if not ok or _is_neg(t) or y == 0:
        return None
    return (x, y, 1, t)


def encode(p):
    """RFC 9496 4.3.2"""
    x0, y0, z0, t0 = p
    u1 = (z0 + y0) * (z0 - y0) % P
    u2 = x0 * y0 % P
    _, inv = sqrt_ratio_m1(1, u1 * u2 % P * u2 % P)
    den1, den2 = inv * u1 % P, inv * u2 % P
    z_inv = den1 * den2 % P * t0 % P
    if _is_neg(t0 * z_inv % P):
        x, y, den_inv = y0 * SQRT_M1 % P, x0 * SQRT_M1 % P, den1 * INVSQRT_A_MINUS_D % P
    else:
        x, y, den_inv = x0, y0, den2
    if _is_neg(x * z_inv % P):
        y = (-y) % P
    return _abs(den_inv * (z0 - y) % P).to_bytes(32, "little")


def _elligator(t):
    r = SQRT_M1 * t % P * t % P
    u = (r + 1) * ONE_MINUS_D_SQ % P
    v = (-1 - r * D) % P * ((r + D) % P) % P
    sq, s = sqrt_ratio_m1(u, v)
    if not sq:
        s = (-_abs(s * t % P)) % P
    c = P - 1 if sq else r
    n = (c * (r - 1) % P * D_MINUS_ONE_SQ - v) % P
    w0 = 2 * s * v % P
    w1 = n * SQRT_AD_MINUS_ONE % P
    w2 = (1 - s * s) % P
    w3 = (1 + s * s) % P
    return (w0 * w3 % P, w2 * w1 % P, w1 * w3 % P, w0 * w2 % P)


def from_uniform_bytes(b64):
    """RistrettoPoint::from_uniform_bytes (RFC 9496 4.3.4): each half is read as a 255-bit little-endian integer (top bit masked)"""
    t1 = int.from_bytes(b64[:32], "little") & ((1 << 255) - 1)
    t2 = int.from_bytes(b64[32:], "little") & ((1 << 255) - 1)
    return pt_add(_elligator(t1 % P), _elligator(t2 % P))


BASEPOINT = decode(bytes.fromhex("e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76"))


# ------------------------------------------------------------------------------------------------ Keccak / STROBE / Merlin
_RC = []
_r = 1
for _i in range(24):                      # FIPS 202 algorithm 5 (rc), the Keccak team's compact LFSR form
    _c = 0
    for _j in range(7):
        _r = ((_r << 1) ^ ((_r >> 7) * 0x71)) % 256
        if _r & 2:
            _c ^= 1 << ((1 << _j) - 1)
    _RC.append(_c)
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]      # [x][y]
_M64 = (1 << 64) - 1


def _rol(v, n):
    n %= 64
    return ((v << n) | (v >> (64 - n))) & _M64 if n else v


def keccak_f1600(state):
    """state: bytearray(200), in place"""
    a = [[int.from_bytes(state[8 * (x + 5 * y): 8 * (x + 5 * y) + 8], "little") for y in range(5)] for x in range(5)]
    for rnd in range(24):
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                b[y][(2 * x + 3 * y) % 5] = _rol(a[x][y], _ROT[x][y])
        a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y] & _M64) for y in range(5)] for x in range(5)]
        a[0][0] ^= _RC[rnd]
    for x in range(5):
        for y in range(5):
            state[8 * (x + 5 * y): 8 * (x + 5 * y) + 8] = a[x][y].to_bytes(8, "little")


FLAG_I, FLAG_A, FLAG_C, FLAG_T, FLAG_M, FLAG_K = 1, 2, 4, 8, 16, 32
STROBE_R = 166


class Strobe128:
    def __init__(self, protocol_label=None):
        self.st = bytearray(200)
        self.pos = self.pos_begin = self.cur_flags = 0
        if protocol_label is not None:
            self.st[0:6] = bytes([1, STROBE_R + 2, 1, 0, 1, 96])
            self.st[6:18] = b"STROBEv1.0.2"
            keccak_f1600(self.st)
            self.meta_ad(protocol_label, False)

    def clone(self):
        s = Strobe128()
        s.st, s.pos, s.pos_begin, s.cur_flags = bytearray(self.st), self.pos, self.pos_begin, self.cur_flags
        return s

    def _run_f(self):
        self.st[self.pos] ^= self.pos_begin
        self.st[self.pos + 1] ^= 0x04
        self.st[STROBE_R + 1] ^= 0x80
        keccak_f1600(self.st)
        self.pos = self.pos_begin = 0

    def _absorb(self, data):
        for byte in data:
            self.st[self.pos] ^= byte
            self.pos += 1
            if self.pos == STROBE_R:
                self._run_f()

    def _overwrite(self, data):
        for byte in data:
            self.st[self.pos] = byte
            self.pos += 1
            if self.pos == STROBE_R:
                self._run_f()

    def _squeeze(self, n):
        out = bytearray()
        for _ in range(n):
            out.append(self.st[self.pos])
            self.st[self.pos] = 0
            self.pos += 1
            if self.pos == STROBE_R:
                self._run_f()
        return bytes(out)

    def _begin_op(self, flags, more):
        if more:
            assert self.cur_flags == flags
            return
        assert not flags & FLAG_T
        old_begin = self.pos_begin
        self.pos_begin = self.pos + 1
        self.cur_flags = flags
        self._absorb(bytes([old_begin, flags]))
        if flags & (FLAG_C | FLAG_K) and self.pos != 0:
            self._run_f()

    def meta_ad(self, data, more):
        self._begin_op(FLAG_M | FLAG_A, more)
        self._absorb(data)

    def ad(self, data, more):
        self._begin_op(FLAG_A, more)
        self._absorb(data)

    def prf(self, n, more):
        self._begin_op(FLAG_I | FLAG_A | FLAG_C, more)
        return self._squeeze(n)

    def key(self, data, more):
        self._begin_op(FLAG_A | FLAG_C, more)
        self._overwrite(data)

    def to_wire(self):
        """the 203-byte form the C ABI and the C oracle exchange: 200 state bytes, pos, pos_begin, cur_flags"""
        return bytes(self.st) + bytes([self.pos, self.pos_begin, self.cur_flags])


def _le32(n):
    return int(n).to_bytes(4, "little")


class Transcript:
    """merlin::Transcript"""

    def __init__(self, label=None, strobe=None):
        if strobe is not None:
            self.s = strobe
        else:
            self.s = Strobe128(b"Merlin v1.0")
            self.append_message(b"dom-sep", label)

    def clone(self):
        return Transcript(strobe=self.s.clone())

    def append_message(self, label, msg):
        self.s.meta_ad(label, False)
        self.s.meta_ad(_le32(len(msg)), True)
        self.s.ad(msg, False)

    def append_u64(self, label, v):
        self.append_message(label, int(v).to_bytes(8, "little"))

    def challenge_bytes(self, label, n):
        self.s.meta_ad(label, False)
        self.s.meta_ad(_le32(n), True)
        return self.s.prf(n, False)

    def build_rng(self):
        return TranscriptRngBuilder(self.s.clone())


class TranscriptRngBuilder:
    def __init__(self, strobe):
        self.s = strobe

    def rekey_with_witness_bytes(self, label, witness):
        self.s.meta_ad(label, False)
        self.s.meta_ad(_le32(len(witness)), True)
        self.s.key(witness, False)
        return self

    def finalize(self, rng):
        random_bytes = rng.fill_bytes(32)
        self.s.meta_ad(b"rng", False)
        self.s.key(random_bytes, False)
        return TranscriptRng(self.s)


class TranscriptRng:
    def __init__(self, strobe):
        self.s = strobe

    def fill_bytes(self, n):
        self.s.meta_ad(_le32(n), False)
        return self.s.prf(n, False)


class NullRng:
    """src/utils/nullrng.rs:16-40"""

    def fill_bytes(self, n):
        return bytes(n)


class BufferRng:
    """an external RNG that replays a given byte stream (what the C ABI's bpp_prove_args.rng_bytes is)"""

    def __init__(self, data):
        self.data, self.off = bytes(data), 0

    def fill_bytes(self, n):
        out = self.data[self.off: self.off + n]
        assert len(out) == n, "rng stream exhausted"
        self.off += n
        return out


class ChaCha12Rng:
    """rand_chacha 0.3.1 ChaCha12Rng over rand_core 0.6 BlockRng: 64-word (4-block) buffer, 64-bit block counter in words 12-13,
    stream id 0; next_u64 / fill_bytes consume whole 32-bit words"""

    def __init__(self, seed32):
        self.key = [int.from_bytes(seed32[4 * i: 4 * i + 4], "little") for i in range(8)]
        self.counter = 0
        self.buf, self.index = [], 64

    @classmethod
    def seed_from_u64(cls, state):
        """rand_core 0.6 SeedableRng::seed_from_u64: PCG32 output per 4-byte chunk of the seed"""
        mul, inc = 6364136223846793005, 11634580027462260723
        seed = bytearray()
        for _ in range(8):
            state = (state * mul + inc) & _M64
            xorshifted = (((state >> 18) ^ state) >> 27) & 0xFFFFFFFF
            rot = state >> 59
            x = ((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & 0xFFFFFFFF
            seed += x.to_bytes(4, "little")
        return cls(bytes(seed))

    def _block(self, counter):
        c = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574]
        init = c + self.key + [counter & 0xFFFFFFFF, counter >> 32, 0, 0]
        x = list(init)

        def qr(a, b, cc, d):
            x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] ^= x[a]; x[d] = ((x[d] << 16) | (x[d] >> 16)) & 0xFFFFFFFF
            x[cc] = (x[cc] + x[d]) & 0xFFFFFFFF; x[b] ^= x[cc]; x[b] = ((x[b] << 12) | (x[b] >> 20)) & 0xFFFFFFFF
            x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] ^= x[a]; x[d] = ((x[d] << 8) | (x[d] >> 24)) & 0xFFFFFFFF
            x[cc] = (x[cc] + x[d]) & 0xFFFFFFFF; x[b] ^= x[cc]; x[b] = ((x[b] << 7) | (x[b] >> 25)) & 0xFFFFFFFF

        for _ in range(6):
            qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
            qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
        return [(a + b) & 0xFFFFFFFF for a, b in zip(x, init)]

    def _generate(self):
        self.buf = []
        for i in range(4):
            self.buf += self._block(self.counter + i)
        self.counter += 4

    def next_u64(self):
        if self.index < 63:
            lo, hi = self.buf[self.index], self.buf[self.index + 1]
            self.index += 2
        elif self.index >= 64:
            self._generate()
            lo, hi = self.buf[0], self.buf[1]
            self.index = 2
        else:
            lo = self.buf[63]
            self._generate()
            hi = self.buf[0]
            self.index = 1
        return (hi << 32) | lo

    def fill_bytes(self, n):
        out = bytearray()
        while len(out) < n:
            if self.index >= 64:
                self._generate()
                self.index = 0
            words = min(64 - self.index, (n - len(out) + 3) // 4)
            chunk = b"".join(w.to_bytes(4, "little") for w in self.buf[self.index: self.index + words])
            out += chunk[: n - len(out)]
            self.index += words
        return bytes(out)


# ------------------------------------------------------------------------------------------------ scalars, nonces, generators
def scalar_from_wide(b64):
    return int.from_bytes(b64, "little") % L


def random_not_zero(rng):
    """src/protocols/scalar_protocol.rs:23-30 over Scalar::random (64 rng bytes, wide reduction)"""
    v = 0
    while v == 0:
        v = scalar_from_wide(rng.fill_bytes(64))
    return v


def sc_bytes(x):
    return int(x % L).to_bytes(32, "little")


def inv(x):
    return pow(x % L, L - 2, L)


def nonce(seed, label, j, k):
    """src/utils/generic.rs:30-60"""
    key = b"\x00" + sc_bytes(seed)
    if j is not None:
        key += b"j" + _le32(j)
    if k is not None:
        key += b"k" + _le32(k)
    h = hashlib.blake2b(key=key, person=label.encode(), digest_size=64)
    return scalar_from_wide(h.digest())


class Params:
    """RangeParameters::init(bit_length, aggregation_factor, create_pedersen_gens_with_extension_degree(ext))"""

    def __init__(self, bit_length, max_aggregation, ext):
        self.n, self.M, self.ext = bit_length, max_aggregation, ext
        self.h = BASEPOINT                                                                             # src/ristretto.rs:70
        self.g = [from_uniform_bytes(hashlib.sha3_512(b"RISTRETTO_MASKING_BASEPOINT_%d" % (i + 1)).digest()) for i in range(ext)]
        self.h_c, self.g_c = encode(self.h), [encode(g) for g in self.g]
        self.gi, self.hi = [], []                                                                      # flat, party-major
        for party in range(max_aggregation):                                                           # bulletproof_gens.rs:88-97
            for tag, dst in ((b"G", self.gi), (b"H", self.hi)):
                xof = hashlib.shake_256(b"GeneratorsChain" + tag + _le32(party)).digest(64 * bit_length)
                dst.extend(from_uniform_bytes(xof[64 * i: 64 * i + 64]) for i in range(bit_length))

    def commit(self, value, blindings):
        assert 1 <= len(blindings) <= self.ext
        return msm([value] + list(blindings), [self.h] + self.g[: len(blindings)])


class ProofError(Exception):
    def __init__(self, variant, msg=""):
        super().__init__("%s: %s" % (variant, msg))
        self.variant = variant


class Statement:
    def __init__(self, params, commitments, minimum_value_promises, seed_nonce):
        self.params, self.commitments, self.mins, self.seed_nonce = params, list(commitments), list(minimum_value_promises), seed_nonce
        self.commitments_c = [encode(c) for c in self.commitments]


class Proof:
    def __init__(self, a, a1, b, r1, s1, d1, li, ri):
        self.a, self.a1, self.b, self.r1, self.s1, self.d1, self.li, self.ri = a, a1, b, r1, s1, list(d1), list(li), list(ri)

    def to_bytes(self):
        """range_proof.rs:1120-1150"""
        out = bytes([len(self.d1)]) + b"".join(sc_bytes(x) for x in self.d1) + self.a + self.a1 + self.b + sc_bytes(self.r1) + sc_bytes(self.s1)
        for l, r in zip(self.li, self.ri):
            out += l + r
        return out

    @classmethod
    def from_bytes(cls, data):
        """range_proof.rs:1155-1257"""
        if len(data) < 1:
            raise ProofError("InvalidLength", "too short")
        ext = data[0]
        if not 1 <= ext <= 6:
            raise ProofError("InvalidArgument", "extension degree")
        body = data[1:]
        chunks = [body[32 * i: 32 * i + 32] for i in range(len(body) // 32)]
        rem = len(body) % 32
        pos = [0]

        def nxt():
            if pos[0] >= len(chunks):
                raise ProofError("InvalidLength", "too short")
            pos[0] += 1
            return chunks[pos[0] - 1]

        def scalar():
            c = nxt()
            v = int.from_bytes(c, "little")
            if v >= L:
                raise ProofError("InvalidArgument", "non-canonical scalar")
            return v

        d1 = [scalar() for _ in range(ext)]
        a, a1, b = nxt(), nxt(), nxt()
        r1, s1 = scalar(), scalar()
        rest = chunks[pos[0]:]
        li, ri = rest[0:len(rest) - len(rest) % 2:2], rest[1::2]
        if not li or not ri:
            raise ProofError("InvalidLength", "too short")
        if len(rest) % 2 or rem:
            raise ProofError("InvalidLength", "unused data")
        return cls(a, a1, b, r1, s1, d1, li, ri)


# ------------------------------------------------------------------------------------------------ RangeProofTranscript
class RangeProofTranscript:
    """src/transcripts.rs:59-194"""

    def __init__(self, transcript, params, statement, witness_bytes, external_rng):
        t = transcript
        t.append_message(b"dom-sep", b"Bulletproofs+ Range Proof")
        self._point(t, b"H", params.h_c)
        for g in params.g_c:
            self._point(t, b"G", g)
        t.append_u64(b"N", params.n)
        t.append_u64(b"T", params.ext)
        t.append_u64(b"M", len(statement.commitments))
        for c in statement.commitments_c:
            t.append_message(b"Ci", c)
        for mv in statement.mins:
            t.append_u64(b"vi - minimum_value", 0 if mv is None else mv)
        self.t, self.bytes, self.ext_rng = t, witness_bytes, external_rng
        self.rng = self._build_rng()

    @staticmethod
    def _point(t, label, enc):
        if enc == bytes(32):
            raise ProofError("VerificationFailed", "Identity element cannot be added to the transcript")
        t.append_message(label, enc)

    def _build_rng(self):
        b = self.t.build_rng()
        if self.bytes is not None:
            b.rekey_with_witness_bytes(b"witness", self.bytes)
        return b.finalize(self.ext_rng)

    def _challenge(self, label):
        v = scalar_from_wide(self.t.challenge_bytes(label, 64))
        if v == 0:
            raise ProofError("VerificationFailed", "Transcript challenge cannot be zero")
        return v

    def challenges_y_z(self, a):
        self._point(self.t, b"A", a)
        self.rng = self._build_rng()
        return self._challenge(b"y"), self._challenge(b"z")

    def challenge_round_e(self, l, r):
        self._point(self.t, b"L", l)
        self._point(self.t, b"R", r)
        self.rng = self._build_rng()
        return self._challenge(b"e")

    def challenge_final_e(self, a1, b):
        self._point(self.t, b"A1", a1)
        self._point(self.t, b"B", b)
        self.rng = self._build_rng()
        return self._challenge(b"e")

    def to_verifier_rng(self, r1, s1, d1):
        self.t.append_message(b"r1", sc_bytes(r1))
        self.t.append_message(b"s1", sc_bytes(s1))
        for x in d1:
            self.t.append_message(b"d1", sc_bytes(x))
        self.rng = self._build_rng()
        return self.rng


# ------------------------------------------------------------------------------------------------ prover
def prove_with_rng(transcript, statement, values, blindings, rng):
    """range_proof.rs:232-608.  values[j], blindings[j][k]: the witness openings."""
    prm = statement.params
    n, m, ext = prm.n, len(statement.commitments), prm.ext
    N = n * m
    if len(values) != m:
        raise ProofError("InvalidLength", "Witness openings and statement commitments do not match!")
    if any(len(b) != ext for b in blindings):
        raise ProofError("InvalidLength", "Witness and statement extension degrees do not match!")
    for v in values:
        if n < 64 and v >> n:
            raise ProofError("InvalidLength", "Value exceeds bit vector capacity!")
    for v, b, c in zip(values, blindings, statement.commitments):
        if not pt_eq(prm.commit(v, b), c):
            raise ProofError("InvalidArgument", "Witness opening is invalid!")
    wbytes = b"".join(int(v).to_bytes(8, "little") + b"".join(sc_bytes(r) for r in b) for v, b in zip(values, blindings))
    rpt = RangeProofTranscript(transcript, prm, statement, wbytes, rng)
    a_li, a_ri = [], []
    for mv, v in zip(statement.mins, values):
        if mv is not None and mv > v:
            raise ProofError("InvalidArgument", "Minimum value is larger than value")
        off = v - (mv or 0)
        for i in range(n):
            a_li.append((off >> i) & 1)
            a_ri.append(((off >> i) & 1) - 1)
    seed = statement.seed_nonce
    alpha = [nonce(seed, "alpha", None, k) if seed is not None else random_not_zero(rpt.rng) for k in range(ext)]
    # A: gi/hi scalars interleaved (a_li, a_ri) over the interleaved precomputation, zero padding, then (alpha, G)
    A = msm(a_li + a_ri + alpha, prm.gi[:N] + prm.hi[:N] + prm.g)
    a_c = encode(A)
    y, z = rpt.challenges_y_z(a_c)
    z2 = z * z % L
    ypow = [pow(y, i, L) for i in range(N + 2)]
    d = []
    for j in range(m):
        for i in range(n):
            d.append(pow(z2, j + 1, L) * pow(2, i, L) % L)
    a_li = [(x - z) % L for x in a_li]
    a_ri = [(x + d[i] * ypow[N - i] + z) % L for i, x in enumerate(a_ri)]
    zp = 1
    for bl in blindings:
        zp = zp * z2 % L
        for k in range(ext):
            alpha[k] = (alpha[k] + zp * bl[k] % L * ypow[N + 1]) % L
    gi, hi = list(prm.gi[:N]), list(prm.hi[:N])
    li, ri = [], []
    nn, rnd = N, 0
    while nn > 1:
        nn //= 2
        a_lo, a_hi, b_lo, b_hi = a_li[:nn], a_li[nn:], a_ri[:nn], a_ri[nn:]
        gi_lo, gi_hi, hi_lo, hi_hi = gi[:nn], gi[nn:], hi[:nn], hi[nn:]
        if ypow[nn] == 0:
            raise ProofError("InvalidArgument", "Cannot invert a zero valued Scalar")
        y_n_inv = inv(ypow[nn])
        a_lo_off = [x * y_n_inv % L for x in a_lo]
        a_hi_off = [x * ypow[nn] % L for x in a_hi]
        d_l = [nonce(seed, "dL", rnd, k) if seed is not None else random_not_zero(rpt.rng) for k in range(ext)]
        d_r = [nonce(seed, "dR", rnd, k) if seed is not None else random_not_zero(rpt.rng) for k in range(ext)]
        rnd += 1
        c_l = sum(a_lo[i] * ypow[i + 1] % L * b_hi[i] for i in range(nn)) % L
        c_r = sum(a_hi[i] * ypow[nn + 1 + i] % L * b_lo[i] for i in range(nn)) % L
        Lp = msm([c_l] + d_l + a_lo_off + b_hi, [prm.h] + prm.g + gi_hi + hi_lo)
        Rp = msm([c_r] + d_r + a_hi_off + b_lo, [prm.h] + prm.g + gi_lo + hi_hi)
        l_c, r_c = encode(Lp), encode(Rp)
        li.append(l_c)
        ri.append(r_c)
        e = rpt.challenge_round_e(l_c, r_c)
        e2, e_inv = e * e % L, inv(e)
        e_inv2 = e_inv * e_inv % L
        e_y_n_inv = e * y_n_inv % L
        gi = [msm([e_inv, e_y_n_inv], [lo, hi_]) for lo, hi_ in zip(gi_lo, gi_hi)]
        hi = [msm([e, e_inv], [lo, hi_]) for lo, hi_ in zip(hi_lo, hi_hi)]
        a_li = [(lo * e + hi_ * e_inv) % L for lo, hi_ in zip(a_lo, a_hi_off)]
        a_ri = [(lo * e_inv + hi_ * e) % L for lo, hi_ in zip(b_lo, b_hi)]
        for k in range(ext):
            alpha[k] = (alpha[k] + d_l[k] * e2 + d_r[k] * e_inv2) % L
    r = random_not_zero(rpt.rng)
    s = random_not_zero(rpt.rng)
    dd = [nonce(seed, "d", None, k) if seed is not None else random_not_zero(rpt.rng) for k in range(ext)]
    eta = [nonce(seed, "eta", None, k) if seed is not None else random_not_zero(rpt.rng) for k in range(ext)]
    A1 = msm([r, s, (r * ypow[1] % L * a_ri[0] + s * ypow[1] % L * a_li[0]) % L] + dd, [gi[0], hi[0], prm.h] + prm.g)
    B = msm([r * ypow[1] % L * s % L] + eta, [prm.h] + prm.g)
    a1_c, b_c = encode(A1), encode(B)
    e = rpt.challenge_final_e(a1_c, b_c)
    e2 = e * e % L
    r1 = (r + a_li[0] * e) % L
    s1 = (s + a_ri[0] * e) % L
    d1 = [(eta[k] + dd[k] * e + alpha[k] * e2) % L for k in range(ext)]
    return Proof(a_c, a1_c, b_c, r1, s1, d1, li, ri)


# ------------------------------------------------------------------------------------------------ verifier
RECOVER_ONLY, RECOVER_AND_VERIFY, VERIFY_ONLY = 0, 1, 2


def verify_batch(transcripts, statements, proofs, action):
    """range_proof.rs:712-752 + verify :756-1065.  transcripts: list of Transcript, advanced in place.
    Returns the list of masks (None or list of ints); raises ProofError."""
    if not statements or not proofs or not transcripts:
        raise ProofError("InvalidArgument", "Range statements or proofs length empty")
    if len(statements) != len(proofs):
        raise ProofError("InvalidArgument", "Range statements and proofs length mismatch")
    if len(transcripts) != len(statements):
        raise ProofError("InvalidArgument", "Range statements and transcripts length mismatch")
    statements, proofs = statements[:256], proofs[:256]
    # consistency (:610-709); generator sets are compared by value
    first = statements[0].params
    n, ext = first.n, first.ext
    if ext != len(proofs[0].d1):
        raise ProofError("InvalidArgument", "Inconsistent extension degree")
    max_mn, max_index = len(statements[0].commitments) * n, 0
    for i, (st, pr) in enumerate(zip(statements, proofs)):
        if i == 0:
            continue
        if st.params.g_c != first.g_c or st.params.h_c != first.h_c or st.params.n != n:
            raise ProofError("InvalidArgument", "Inconsistent generators in batch statement")
        if st.params.ext != ext or len(pr.d1) != ext:
            raise ProofError("InvalidArgument", "Inconsistent extension degree")
        if len(st.commitments) * n > max_mn:
            max_mn, max_index = len(st.commitments) * n, i
    max_prm = statements[max_index].params
    for i, st in enumerate(statements):
        for mv in st.mins:
            if mv is not None and n < 64 and mv >> n:
                raise ProofError("InvalidLength", "Minimum value promise exceeds bit vector capacity")
        if i != max_index:
            if any(not pt_eq(a, b) for a, b in zip(st.params.gi, max_prm.gi)) or any(not pt_eq(a, b) for a, b in zip(st.params.hi, max_prm.hi)):
                raise ProofError("InvalidArgument", "Inconsistent generator point vector in batch statement")
    two_n_minus_one = (pow(2, n, L) - 1) % L
    g_sc, h_sc = [0] * ext, 0
    gi_sc, hi_sc = [0] * max_mn, [0] * max_mn
    dyn_s, dyn_p = [], []
    masks = []
    weight_t = Transcript(b"Bulletproofs+ verifier weights")
    challenges = []
    for pr, st, t in zip(proofs, statements, transcripts):                      # izip! truncates `transcripts` to the batch
        rpt = RangeProofTranscript(t, first, st, None, NullRng())
        y, z = rpt.challenges_y_z(pr.a)
        round_e = [rpt.challenge_round_e(l, r) for l, r in zip(pr.li, pr.ri)]
        e = rpt.challenge_final_e(pr.a1, pr.b)
        challenges.append((y, z, round_e, e))
        weight_t.append_message(b"proof", rpt.to_verifier_rng(pr.r1, pr.s1, pr.d1).fill_bytes(32))
    weight_rng = weight_t.build_rng().finalize(NullRng())
    for pr, st, (y, z, ch, e) in zip(proofs, statements, challenges):
        pts = []
        for name, enc in [("a", pr.a), ("a1", pr.a1), ("b", pr.b)] + [("L", x) for x in pr.li] + [("R", x) for x in pr.ri]:
            p = decode(enc)
            if p is None:
                raise ProofError("InvalidArgument", "Member '%s' was not the canonical encoding of a point" % name)
            pts.append(p)
        A, A1, B = pts[0], pts[1], pts[2]
        rounds = len(pr.li)
        Lp, Rp = pts[3:3 + rounds], pts[3 + rounds:]
        m = len(st.commitments)
        N = m * n
        if len(pr.li) != len(pr.ri):
            raise ProofError("InvalidLength", "Vector L length not equal to vector R length")
        if rounds >= 64:                 # 1usize.checked_shl(rounds) overflows from 64 on (range_proof.rs:880-885)
            raise ProofError("SizeOverflow")
        if (1 << rounds) != N:
            raise ProofError("InvalidLength", "Vector L/R length not adequate")
        weight = random_not_zero(weight_rng)
        ch_inv = [inv(c) for c in ch]
        y_inv, y_1_inv = inv(y), inv(y - 1)
        ch_inv_prod = 1
        for c in ch_inv:
            ch_inv_prod = ch_inv_prod * c % L
        z2, e2 = z * z % L, e * e % L
        ch_sq = [c * c % L for c in ch]
        ch_sq_inv = [c * c % L for c in ch_inv]
        y_nm = pow(y, N, L)
        y_nm_1 = y_nm * y % L
        y_sum = y * (y_nm - 1) % L * y_1_inv % L
        d = [pow(z2, j + 1, L) * pow(2, i, L) % L for j in range(m) for i in range(n)]
        d_sum = sum(pow(z2, j, L) for j in range(1, m + 1)) % L * two_n_minus_one % L
        if action == VERIFY_ONLY:
            masks.append(None)
        else:
            if st.seed_nonce is not None:
                sd = st.seed_nonce
                tm = []
                for k in range(ext):
                    mk = (pr.d1[k] - nonce(sd, "eta", None, k) - e * nonce(sd, "d", None, k)) % L * inv(e2) % L
                    mk = (mk - nonce(sd, "alpha", None, k)) % L
                    for j in range(rounds):
                        mk = (mk - ch_sq[j] * nonce(sd, "dL", j, k) - ch_sq_inv[j] * nonce(sd, "dR", j, k)) % L
                    tm.append(mk * inv(z2 * y_nm_1) % L)
                masks.append(tm)
            else:
                masks.append(None)
            if action == RECOVER_ONLY:
                continue
        s = [ch_inv_prod]
        for i in range(1, N):
            lg = i.bit_length() - 1
            s.append(s[i - (1 << lg)] * ch_sq[rounds - lg - 1] % L)
        r1_e, s1_e, e2_z = pr.r1 * e % L, pr.s1 * e % L, e2 * z % L
        y_inv_i, y_nm_i = 1, y_nm
        for i in range(N):
            gi_sc[i] = (gi_sc[i] + weight * (r1_e * y_inv_i % L * s[i] + e2_z)) % L
            hi_sc[i] = (hi_sc[i] + weight * (s1_e * s[N - 1 - i] - e2 * (d[i] * y_nm_i + z))) % L
            y_inv_i = y_inv_i * y_inv % L
            y_nm_i = y_nm_i * y_inv % L
        zp = 1
        for mv in st.mins:
            zp = zp * z2 % L
            weighted = weight * ((-e2) * zp % L * y_nm_1 % L) % L
            dyn_s.append(weighted)
            if mv is not None:
                h_sc = (h_sc - weighted * mv) % L
        dyn_p.extend(st.commitments)
        h_sc = (h_sc + weight * (pr.r1 * y % L * pr.s1 + e2 * (y_nm_1 * z % L * d_sum + (z2 - z) * y_sum))) % L
        for k in range(ext):
            g_sc[k] = (g_sc[k] + weight * pr.d1[k]) % L
        dyn_s += [weight * (-e) % L, (-weight) % L, weight * (-e2) % L]
        dyn_p += [A1, B, A]
        dyn_s += [weight * (-e2) % L * c % L for c in ch_sq]
        dyn_p += Lp
        dyn_s += [weight * (-e2) % L * c % L for c in ch_sq_inv]
        dyn_p += Rp
    if action == RECOVER_ONLY:
        return masks
    dyn_s += g_sc + [h_sc]
    dyn_p += first.g + [first.h]
    res = msm(gi_sc + hi_sc + dyn_s, max_prm.gi[:max_mn] + max_prm.hi[:max_mn] + dyn_p)
    if not pt_eq(res, IDENTITY):
        raise ProofError("VerificationFailed", "Range proof batch not valid")
    return masks
