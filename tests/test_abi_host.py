"""CPU-side checks of the product library: it loads, exports every symbol include/bpp_b200.h declares, its host hash
layer matches python hashlib / the Merlin KAT / the oracle, and the compute entry points fail loudly without a GPU."""
import ctypes as C
import hashlib
import os
import random
import re

import pytest

import bpp
import orc


@pytest.fixture(scope="module")
def lib():
    bpp.ffi.build()
    return bpp.ffi.lib()


def test_exports_match_header(lib):
    hdr = open(bpp.ffi.HEADER_PATH).read()
    names = sorted(set(re.findall(r"\b(bpp_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert lib.bpp_ctx_create(0, C.byref(h)) == bpp.ffi.ERR_CUDA
    with pytest.raises(bpp.pkg.EngineError):
        bpp.pkg.Engine(0)


def test_host_hashes_vs_hashlib(lib):
    rnd = random.Random(5)
    for n in [0, 1, 71, 72, 73, 135, 136, 137, 500]:
        data = bytes(rnd.randrange(256) for _ in range(n))
        out = C.create_string_buffer(64)
        lib.bpp_hash_sha3_512(data, n, out)
        assert out.raw == hashlib.sha3_512(data).digest()
        out2 = C.create_string_buffer(300)
        lib.bpp_hash_shake256(data, n, out2, 300)
        assert out2.raw == hashlib.shake_256(data).digest(300)
    for keylen in [1, 33, 38, 43, 64]:
        for person in [b"alpha", b"dL", b"0123456789abcdef", b""]:
            key = bytes(rnd.randrange(256) for _ in range(keylen))
            out = C.create_string_buffer(64)
            assert lib.bpp_hash_blake2b_nonce_bytes(key, keylen, person, len(person), out) == 0
            assert out.raw == hashlib.blake2b(b"", key=key, person=person, digest_size=64).digest()
    assert lib.bpp_hash_blake2b_nonce_bytes(b"x" * 65, 65, b"", 0, C.create_string_buffer(64)) == bpp.ffi.INVALID_BLAKE2B
    assert lib.bpp_hash_blake2b_nonce_bytes(b"x", 1, b"y" * 17, 17, C.create_string_buffer(64)) == bpp.ffi.INVALID_BLAKE2B
    for _ in range(50):
        w = rnd.randrange(2**512)
        out = C.create_string_buffer(32)
        lib.bpp_scalar_from_wide(w.to_bytes(64, "little"), out)
        assert int.from_bytes(out.raw, "little") == w % orc.L


def test_merlin_kat_and_oracle(lib):
    # merlin's own test vector (merlin 3.0.0 src/transcript.rs tests)
    t = bpp.pkg.transcript_new(b"test protocol")
    t = bpp.pkg.transcript_append_message(t, b"some label", b"some data")
    t, ch = bpp.pkg.transcript_challenge_bytes(t, b"challenge", 32)
    assert ch.hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"
    # long messages crossing the 166-byte rate, against the oracle's independent STROBE
    rnd = random.Random(6)
    t1 = bpp.pkg.transcript_new(b"BatchedRangeProofTest")
    t2 = orc.transcript_new(b"BatchedRangeProofTest")
    assert t1 == t2
    ol = orc.lib()
    for n in [0, 1, 32, 165, 166, 167, 400]:
        msg = bytes(rnd.randrange(256) for _ in range(n))
        t1 = bpp.pkg.transcript_append_message(t1, b"lbl", msg)
        b2 = C.create_string_buffer(t2, 203)
        ol.orc_transcript_append_message(b2, b"lbl", msg, n)
        t2 = b2.raw
        assert t1 == t2
        t1, c1 = bpp.pkg.transcript_challenge_bytes(t1, b"ch", 64)
        c2 = C.create_string_buffer(64)
        b2 = C.create_string_buffer(t2, 203)
        ol.orc_transcript_challenge_bytes(b2, b"ch", c2, 64)
        t2 = b2.raw
        assert c1 == c2.raw and t1 == t2


def test_proof_check_bytes_matches_oracle(lib):
    import workload

    case = workload.make_case(8, [1], 2, seed_nonce=True)
    good = case.proof_bytes()[0]
    rc, ext, rounds = bpp.pkg.proof_check_bytes(good)
    assert (rc, ext, rounds) == (0, 2, 3)
    rnd = random.Random(7)
    variants = [good[:k] for k in range(0, len(good))] + [good + bytes(k) for k in (1, 32, 63, 64, 65, 128)]
    for k in (0, 1, 33, 65 + 96, 65 + 96 + 32):
        b = bytearray(good)
        b[k:k + 32] = (orc.L + 5).to_bytes(32, "little") if k else b"\x07" + bytes(b[1:32])
        variants.append(bytes(b))
    variants += [bytes([e]) + good[1:] for e in (0, 1, 3, 7, 255)]
    for v in variants:
        rc_o, _ = orc.proof_from_bytes(v)
        rc_p, _, _ = bpp.pkg.proof_check_bytes(v)
        assert rc_p == rc_o, (len(v), rc_p, rc_o)


def test_points_sum_host_matches_oracle():
    """bpp_points_sum_host (host-side sum of the per-GPU partial results of a sharded MSM) against the oracle's unit-scalar MSM;
    a non-decodable encoding is refused"""
    import ctypes as C
    import hashlib

    import orc

    o = C.create_string_buffer(32)
    pts = []
    for i in range(8):
        orc.lib().orc_ristretto_from_uniform(hashlib.shake_256(b"partial-%d" % i).digest(64), o)
        pts.append(o.raw)
    for k in (0, 1, 2, 8):
        got = bpp.pkg.points_sum_host(b"".join(pts[:k]))
        if k == 0:
            assert got == bytes(32)
        else:
            assert orc.lib().orc_msm((1).to_bytes(32, "little") * k, b"".join(pts[:k]), k, 0, o) == 1
            assert got == o.raw
    with pytest.raises(bpp.pkg.EngineError):
        bpp.pkg.points_sum_host(pts[0] + b"\xff" * 32)


@pytest.mark.parametrize("shape", [(1, 1), (1, 5), (2, 7), (4, 37), (3, 100), (2, 255), (4, 256), (5, 9), (8, 256), (7, 33)])
def test_host_verifier_weights_match_oracle(shape):
    """the product's host-side verifier-weight transcripts (range_proof.rs:811-853, :894) -- one at a time and four in lock-step
    through the vectorised four-way Keccak-f (host_keccak4.cpp) -- against the oracle's restatement of the same lines"""
    n_chunks, length = shape
    lib = bpp.ffi.lib()
    wb = hashlib.shake_256(b"weights-%d-%d" % shape).digest(32 * length * n_chunks)
    want = C.create_string_buffer(32 * length * n_chunks)
    for c in range(n_chunks):
        part = C.create_string_buffer(32 * length)
        orc.lib().orc_verifier_weights(wb[32 * length * c: 32 * length * (c + 1)], length, part)
        C.memmove(C.addressof(want) + 32 * length * c, part, 32 * length)
    for lockstep in (0, 1, 2):          # one transcript at a time, four in lock-step, eight in lock-step
        if lockstep == 1 and n_chunks > 4:
            continue
        got = C.create_string_buffer(32 * length * n_chunks)
        assert lib.bpp_host_verifier_weights(wb, length, n_chunks, lockstep, got) == 0
        assert got.raw == want.raw, lockstep


def _fold_edge_cases():
    """values that sit on the edges of the three folds of the wide reduction (multiples of l, of 2^252, all-ones limbs)"""
    L = orc.L
    vals = [0, 1, L - 1, L, L + 1, 2 * L, 4 * L - 1, L * L, L * L - 1, 2**252 - 1, 2**252, 2**252 + 1, 2**256 - 1, 2**256, 2**504, 2**511, 2**512 - 1,
            2**512 - L, (2**512 // L) * L, (2**512 // L) * L - 1, 2**385 - 1, 2**385, 2**258 - 1, 2**258, (2**260 - 1) << 252, (2**133 - 1) << 252]
    vals += [k * L + d for k in (1, 2**100, 2**200, 2**259) for d in (-1, 0, 1) if 0 <= k * L + d < 2**512]
    vals += [(0xFFFFFFFFFFFFFFFF << (64 * i)) for i in range(8)] + [(1 << (64 * i)) for i in range(8)]
    return vals


def test_host_wide_reduction_64bit_limbs():
    """bpp_host_sc_from_wide64 (the folded 64-bit-limb reduction used for the weights: MULX body and portable body) against python
    integers and the shared 32-bit-limb path"""
    lib = bpp.ffi.lib()
    L = orc.L
    cases = [v.to_bytes(64, "little") for v in _fold_edge_cases()]
    cases += [hashlib.shake_256(b"wide-%d" % i).digest(64) for i in range(2000)]
    a, b, g = C.create_string_buffer(32), C.create_string_buffer(32), C.create_string_buffer(32)
    for c in cases:
        lib.bpp_host_sc_from_wide64(c, a)
        lib.bpp_scalar_from_wide(c, b)
        lib.bpp_host_sc_generic64(c, None, g)
        assert a.raw == b.raw == g.raw == (int.from_bytes(c, "little") % L).to_bytes(32, "little"), c.hex()


def test_host_scalar_mul_64bit_limbs():
    """bpp_host_sc_mul64 (the prover's host-side scalar products; both bodies): any two 256-bit values, against python integers"""
    import itertools

    lib = bpp.ffi.lib()
    L = orc.L
    o, g = C.create_string_buffer(32), C.create_string_buffer(32)
    a_vals = [0, 1, L - 1, L, L + 5, 2**255, 2**256 - 1] + [int.from_bytes(hashlib.shake_256(b"ma%d" % i).digest(32), "little") for i in range(60)]
    b_vals = [0, 1, 2, L - 1, L - 2, 2**252, 2**256 - 1] + [int.from_bytes(hashlib.shake_256(b"mb%d" % i).digest(32), "little") % L for i in range(60)]
    for a, b in itertools.product(a_vals, b_vals):
        lib.bpp_host_sc_mul64(a.to_bytes(32, "little"), b.to_bytes(32, "little"), o)
        lib.bpp_host_sc_generic64(a.to_bytes(32, "little"), b.to_bytes(32, "little"), g)
        assert int.from_bytes(o.raw, "little") == int.from_bytes(g.raw, "little") == (a * b) % L, (hex(a), hex(b))


def test_host_keccak_permutation_bodies_agree():
    """bpp_keccak_f1600_x1 (vector-register body when the CPU has one) and the plain 64-bit body: the all-zero state gives the
    published first lane of Keccak-f[1600](0), random states agree, and SHA3-512 through the library equals hashlib's"""
    lib = bpp.ffi.lib()
    z1, z2 = (C.c_uint64 * 25)(), (C.c_uint64 * 25)()
    lib.bpp_keccak_f1600_x1(z1)
    lib.bpp_keccak_f1600_x1_generic(z2)
    assert z1[0] == z2[0] == 0xF1258F7940E1DDE7 and list(z1) == list(z2)
    for i in range(50):
        raw = hashlib.shake_256(b"kf-%d" % i).digest(200)
        a, b = (C.c_uint64 * 25).from_buffer_copy(raw), (C.c_uint64 * 25).from_buffer_copy(raw)
        for _ in range(1 + i % 3):
            lib.bpp_keccak_f1600_x1(a)
            lib.bpp_keccak_f1600_x1_generic(b)
        assert list(a) == list(b), i
    out = C.create_string_buffer(64)
    for n in (0, 1, 71, 72, 73, 200, 1000):
        msg = hashlib.shake_256(b"m%d" % n).digest(n)
        lib.bpp_hash_sha3_512(msg, n, out)
        assert out.raw == hashlib.sha3_512(msg).digest()


_SIMD_CHILD = r"""
import ctypes as C, hashlib, sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import bpp, orc
lib = bpp.ffi.lib()
want_level = int(sys.argv[1])
assert lib.bpp_host_simd_level() <= want_level
# permutation: this level's body against the plain 64-bit body
for i in range(20):
    raw = hashlib.shake_256(b"lvl-%%d" %% i).digest(200)
    a, b = (C.c_uint64 * 25).from_buffer_copy(raw), (C.c_uint64 * 25).from_buffer_copy(raw)
    lib.bpp_keccak_f1600_x1(a); lib.bpp_keccak_f1600_x1_generic(b)
    assert list(a) == list(b)
# SHA3-512 / SHAKE256 through the host sponge
out = C.create_string_buffer(64)
for n in (0, 71, 72, 73, 500):
    msg = hashlib.shake_256(b"m%%d" %% n).digest(n)
    lib.bpp_hash_sha3_512(msg, n, out)
    assert out.raw == hashlib.sha3_512(msg).digest()
# verifier weights: lock-step four-way sponge and one at a time, against the oracle
length, n_chunks = 37, 4
wb = hashlib.shake_256(b"lvl-weights").digest(32 * length * n_chunks)
want = b""
for c in range(n_chunks):
    part = C.create_string_buffer(32 * length)
    orc.lib().orc_verifier_weights(wb[32 * length * c: 32 * length * (c + 1)], length, part)
    want += part.raw
for lockstep in (0, 1, 2):
    got = C.create_string_buffer(32 * length * n_chunks)
    assert lib.bpp_host_verifier_weights(wb, length, n_chunks, lockstep, got) == 0
    assert got.raw == want
# wide reduction and product mod l
L = orc.L
o = C.create_string_buffer(32)
for i in range(200):
    w = hashlib.shake_256(b"lvl-w%%d" %% i).digest(64)
    lib.bpp_host_sc_from_wide64(w, o)
    assert int.from_bytes(o.raw, "little") == int.from_bytes(w, "little") %% L
    lib.bpp_host_sc_mul64(w[:32], w[32:], o)
    assert int.from_bytes(o.raw, "little") == int.from_bytes(w[:32], "little") * int.from_bytes(w[32:], "little") %% L
print("level", lib.bpp_host_simd_level(), "ok")
"""


@pytest.mark.parametrize("level", [0, 1, 2])
def test_host_simd_bodies(level):
    """every host body (baseline ISA without MULX, AVX2, AVX-512VL) gives the same bytes: BPP_HOST_SIMD caps the dispatch in a child
    process; a level the CPU lacks simply runs the best one below it"""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env["BPP_HOST_SIMD"] = str(level)
    r = subprocess.run([sys.executable, "-c", _SIMD_CHILD % (root, os.path.join(root, "tests")), str(level)], env=env, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-1000:] + r.stderr[-3000:]


def test_host_scalar_arithmetic_property():
    """hypothesis: the folded wide reduction and the 256 x 256-bit product mod l (both bodies) against python integers over the whole
    input range, with the boundary-heavy distributions hypothesis likes to draw"""
    from hypothesis import given, settings, strategies as st

    lib = bpp.ffi.lib()
    L = orc.L
    a_buf, g_buf = C.create_string_buffer(32), C.create_string_buffer(32)

    @settings(max_examples=3000, deadline=None)
    @given(st.integers(min_value=0, max_value=2**512 - 1))
    def wide(x):
        w = x.to_bytes(64, "little")
        lib.bpp_host_sc_from_wide64(w, a_buf)
        lib.bpp_host_sc_generic64(w, None, g_buf)
        assert a_buf.raw == g_buf.raw == (x % L).to_bytes(32, "little")

    @settings(max_examples=3000, deadline=None)
    @given(st.integers(min_value=0, max_value=2**256 - 1), st.integers(min_value=0, max_value=2**256 - 1))
    def mul(a, b):
        ab, bb = a.to_bytes(32, "little"), b.to_bytes(32, "little")
        lib.bpp_host_sc_mul64(ab, bb, a_buf)
        lib.bpp_host_sc_generic64(ab, bb, g_buf)
        assert a_buf.raw == g_buf.raw == (a * b % L).to_bytes(32, "little")

    wide()
    mul()
