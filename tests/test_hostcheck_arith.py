"""Host-compiled check of the PRODUCT's shared arithmetic header (csrc/arith.cuh, portable path) against python
big integers and the oracle.  The GPU path (PTX cores) is checked by tests/test_gpu_*.py with the same vectors."""
import ctypes as C
import os
import random
import subprocess

import pytest

import orc

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
SO = os.path.join(HERE, "hostcheck", "_hostcheck.so")
HDR = os.path.join(HERE, "..", "bulletproofs-plus_b200", "csrc", "arith.cuh")
P, L = orc.P, orc.L


@pytest.fixture(scope="module")
def hc():
    if not os.path.exists(SO) or max(os.path.getmtime(SRC), os.path.getmtime(HDR)) > os.path.getmtime(SO):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-x", "c++", "-o", SO, SRC],
                       check=True)
    return C.CDLL(SO)


def b32(x):
    return int(x).to_bytes(32, "little")


def buf(n=32):
    return C.create_string_buffer(n)


def i32(b):
    return int.from_bytes(b.raw[:32], "little")


EDGE = [0, 1, 2, 18, 19, 20, 37, 38, 39, P - 2, P - 1, P, P + 1, P + 18, 2**255 - 1, 2**254, 2**255 - 38,
        2**32 - 1, 2**32, 2**224, (2**255 - 1) ^ (2**31 - 1), 2**255 - 2**32]


def test_constants(hc):
    d = (-121665 * pow(121666, -1, P)) % P
    sqrtm1 = pow(2, (P - 1) // 4, P)
    exp = [d, 2 * d % P, sqrtm1,
           54469307008909316920995813868745141605393597292927456921205312896311721017578,
           25063068953384623474111414158702152701244531502492656460079210482610430750235,
           (1 - d * d) % P, (d - 1) ** 2 % P]
    o = buf()
    for i, e in enumerate(exp):
        hc.hc_const(i, o)
        assert i32(o) == e, i
    hc.hc_sc_const(0, o); assert i32(o) == 2**256 % L
    hc.hc_sc_const(1, o); assert i32(o) == 2**512 % L
    hc.hc_sc_const(2, o); assert i32(o) == L
    assert (-pow(L, -1, 2**32)) % 2**32 == 0x12547E1B


def test_field_ops(hc):
    rnd = random.Random(11)
    vals = EDGE + [rnd.randrange(2**255) for _ in range(400)]
    o = buf()
    for a in vals:
        for b in rnd.sample(vals, 6) + EDGE[:8]:
            hc.hc_fe_mul(b32(a), b32(b), o); assert i32(o) == a * b % P
            hc.hc_fe_add(b32(a), b32(b), o); assert i32(o) == (a + b) % P
            hc.hc_fe_sub(b32(a), b32(b), o); assert i32(o) == (a - b) % P
            for f in (hc.hc_fe_mul_raw, hc.hc_fe_sub_raw, hc.hc_fe_add_raw):
                f(b32(a), b32(b), o); assert i32(o) < 2**255
        hc.hc_fe_sq(b32(a), o); assert i32(o) == a * a % P
    for a in [1, 2, P - 1, 5, rnd.randrange(P)]:
        hc.hc_fe_invert(b32(a), o); assert i32(o) == pow(a, -1, P)


def test_lazy_add_sub_feed_multiplications(hc):
    """fe_add_l / fe_sub_l produce loose (full 256-bit) values; fe_mul / fe_sq must reduce them correctly"""
    rnd, o = random.Random(15), buf()
    loose_edge = [0, 1, 37, 38, 39, 2**255 - 19, 2**255 - 1, 2**255, 2**256 - 39, 2**256 - 38, 2**256 - 1, 2**256 - 2**32]
    tight = EDGE + [rnd.randrange(2**255) for _ in range(40)]
    loose = loose_edge + [rnd.randrange(2**256) for _ in range(40)]
    for a in loose:
        for b in rnd.sample(loose, 8) + loose_edge:
            hc.hc_fe_mul_loose_raw(b32(a), b32(b), o); assert i32(o) < 2**255 and i32(o) % P == a * b % P
            hc.hc_fe_lazy_sq(b32(a), b32(b), o); assert i32(o) == (a - b) ** 2 % P
            hc.hc_fe_lazy_sq_tight(b32(a), b32(b % 2**255), o); assert i32(o) == (a - b % 2**255) ** 2 % P
    for a in tight[:30]:
        for b in rnd.sample(tight, 4):
            for c in rnd.sample(loose, 4) + loose_edge[:6]:
                d = rnd.choice(loose)
                hc.hc_fe_lazy_mul(b32(a), b32(b), b32(c), b32(d), o); assert i32(o) == (a + b) * (c - d) % P


def test_sqrt_ratio_vs_oracle(hc):
    rnd, l = random.Random(12), orc.lib()
    o, o2 = buf(), buf()
    for _ in range(60):
        u, v = rnd.randrange(P), rnd.randrange(1, P)
        assert hc.hc_fe_sqrt_ratio_i(b32(u), b32(v), o) == l.orc_fe_sqrt_ratio_i(b32(u), b32(v), o2)
        assert o.raw == o2.raw
        assert hc.hc_fe_invsqrt(b32(v), o) == l.orc_fe_sqrt_ratio_i(b32(1), b32(v), o2)
        assert o.raw == o2.raw
    # u = 0 / v = 0 corner cases
    for u, v in [(0, 5), (5, 0), (0, 0), (1, 0)]:
        assert hc.hc_fe_sqrt_ratio_i(b32(u), b32(v), o) == l.orc_fe_sqrt_ratio_i(b32(u), b32(v), o2)
        assert o.raw == o2.raw


def test_scalar_ops(hc):
    rnd, o = random.Random(13), buf()
    edge = [0, 1, 2, L - 1, L - 2, 2**252, 2**252 - 1]
    vals = edge + [rnd.randrange(L) for _ in range(300)]
    for a in vals:
        for b in rnd.sample(vals, 5) + edge:
            hc.hc_sc_mul(b32(a), b32(b), o); assert i32(o) == a * b % L
            hc.hc_sc_add(b32(a), b32(b), o); assert i32(o) == (a + b) % L
            hc.hc_sc_sub(b32(a), b32(b), o); assert i32(o) == (a - b) % L
    for a in [2**256 - 1, L, L + 1, 2 * L, rnd.randrange(2**256)]:
        hc.hc_sc_reduce256(b32(a), o); assert i32(o) == a % L
    for _ in range(100):
        w = rnd.randrange(2**512)
        hc.hc_sc_from_wide(w.to_bytes(64, "little"), o); assert i32(o) == w % L
    hc.hc_sc_from_wide((2**512 - 1).to_bytes(64, "little"), o); assert i32(o) == (2**512 - 1) % L
    for a in [1, 2, L - 1, rnd.randrange(1, L)]:
        hc.hc_sc_invert(b32(a), o); assert i32(o) == pow(a, -1, L)
    for a in [1, 2, 3, L - 1, L - 2, 2**252, 2**252 - 1, (L + 1) // 2] + [rnd.randrange(1, L) for _ in range(300)] + [2**k for k in range(0, 252, 17)]:
        hc.hc_sc_invert_gcd(b32(a), o); assert i32(o) == pow(a, -1, L), a
        hc.hc_scm_invert_gcd(b32(a), o); assert i32(o) == pow(a, -1, L), a
    hc.hc_sc_invert_gcd(b32(0), o); assert i32(o) == 0
    # safegcd (30-bit division-step batches): what the verifier's scalar prep inverts with
    for a in [1, 2, 3, L - 1, L - 2, 2**252, 2**252 - 1, (L + 1) // 2, (L - 1) // 2, 2**30 - 1, 2**30, 2**60 + 1] + [rnd.randrange(1, L) for _ in range(3000)] + \
            [2**k for k in range(0, 252)] + [L - 2**k for k in range(0, 252, 5)] + [rnd.randrange(1, 2**rnd.randrange(1, 252)) for _ in range(500)]:
        hc.hc_sc_invert_sg(b32(a), o); assert i32(o) == pow(a, -1, L), a
    hc.hc_sc_invert_sg(b32(0), o); assert i32(o) == 0
    assert hc.hc_sc_is_canonical(b32(L - 1)) == 1 and hc.hc_sc_is_canonical(b32(L)) == 0
    assert hc.hc_sc_is_canonical(b32(2**256 - 1)) == 0 and hc.hc_sc_is_canonical(b32(0)) == 1


def test_ristretto_vs_oracle(hc):
    from test_oracle_primitives import RFC9496_BAD, RFC9496_MULTIPLES

    rnd, l = random.Random(14), orc.lib()
    o, o2 = buf(), buf()
    for hx in RFC9496_MULTIPLES:
        enc = bytes.fromhex(hx)
        assert hc.hc_decode_encode(enc, o) == 1 and o.raw == enc
    for hx in RFC9496_BAD:
        assert hc.hc_decode_encode(bytes.fromhex(hx), o) == 0
    pts = []
    for _ in range(60):
        h = bytes(rnd.randrange(256) for _ in range(64))
        hc.hc_from_uniform(h, o); l.orc_ristretto_from_uniform(h, o2)
        assert o.raw == o2.raw
        pts.append(o.raw)
        assert hc.hc_decode_encode(pts[-1], o) == 1 and o.raw == pts[-1]
        rb = bytes(rnd.randrange(256) for _ in range(32))
        assert hc.hc_decode_encode(rb, o) == l.orc_ristretto_decode_encode(rb, o2)
    oa, od, om, os_ = buf(), buf(), buf(), buf()
    ident = bytes(32)
    for p, q in zip(pts, pts[1:] + [ident]):
        for pp, qq in ((p, q), (p, p), (ident, q), (p, ident)):
            assert hc.hc_point_ops(pp, qq, oa, od, om, os_) == 1
            l.orc_ristretto_add(pp, qq, o); assert oa.raw == o.raw
            l.orc_ristretto_add(o.raw, o.raw, o2); assert od.raw == o2.raw
            dbl = o2.raw
            l.orc_ristretto_add(dbl, pp, o); assert om.raw == o.raw
            neg = C.create_string_buffer(32)
            l.orc_ristretto_scalarmult(b32(L - 1), pp, neg) if pp != ident else None
            l.orc_ristretto_add(dbl, neg.raw if pp != ident else ident, o); assert os_.raw == o.raw
    assert hc.hc_is_identity(pts[0], pts[0]) == 3
    assert hc.hc_is_identity(pts[0], pts[1]) == 0
    assert hc.hc_is_identity(ident, ident) == 3
