"""Host only: the prover's eight-way lock-step sponge (csrc/strobe_n.hpp: StrobeN / MerlinN, one vectorised Keccak-f for eight
transcripts) against the one-at-a-time Merlin of the library AND against the independent python restatement (oracle/pyref.py):
append_message -> challenge_bytes -> witness-keyed TranscriptRng -> fill_bytes twice, at sponge positions that make the absorbed and
squeezed bytes straddle lanes and the rate boundary."""
import ctypes as C
import hashlib
import os
import sys

import pytest

import bpp

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import pyref  # noqa: E402

REC = 203 + 64 + 203 + 128


class _Buf:
    def __init__(self, data):
        self.data = data

    def fill_bytes(self, n):
        assert n == len(self.data)
        return self.data


@pytest.mark.parametrize("prefix_len,msg_len,wlen", [(0, 32, 40), (3, 32, 40), (97, 32, 72), (140, 64, 40), (163, 1, 8), (165, 200, 232), (20, 340, 136)])
def test_lockstep_sponge_equals_scalar_merlin_and_pyref(prefix_len, msg_len, wlen):
    lib = bpp.ffi.lib()
    lib.bpp_host_lockstep_selftest.restype = C.c_int32
    lib.bpp_host_lockstep_selftest.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_char_p, C.c_char_p, C.c_char_p]
    xof = hashlib.shake_256(b"lockstep-%d-%d-%d" % (prefix_len, msg_len, wlen)).digest(8 * (prefix_len + msg_len + wlen + 32))
    cut = [0]

    def take(n):
        cut[0] += n
        return xof[cut[0] - n:cut[0]]

    # eight transcripts with different contents at the SAME sponge position: one label, `prefix_len` different bytes each
    refs, states = [], b""
    for _ in range(8):
        t = pyref.Transcript(b"lock-step selftest")
        t.append_message(b"prefix", take(prefix_len))
        refs.append(t)
        states += t.s.to_wire()
    msgs = [take(msg_len) for _ in range(8)]
    wits = [take(wlen) for _ in range(8)]
    exts = [take(32) for _ in range(8)]
    a, b = C.create_string_buffer(8 * REC), C.create_string_buffer(8 * REC)
    rc = lib.bpp_host_lockstep_selftest(states, b"".join(msgs), msg_len, b"".join(wits), wlen, b"".join(exts), a, b)
    assert rc == 0
    assert a.raw == b.raw                                   # lock-step == one at a time
    for j, t in enumerate(refs):                            # == the independent restatement
        t.append_message(b"L", msgs[j])
        ch = t.challenge_bytes(b"e", 64)
        rng = t.build_rng().rekey_with_witness_bytes(b"witness", wits[j]).finalize(_Buf(exts[j]))
        f0, f1 = rng.fill_bytes(64), rng.fill_bytes(64)
        want = t.s.to_wire() + ch + rng.s.to_wire() + f0 + f1
        assert a.raw[REC * j:REC * (j + 1)] == want, j


def test_lockstep_sponge_random_shapes():
    """the same comparison over random prefix / message / witness lengths (hypothesis), lock-step against one-at-a-time only"""
    from hypothesis import given, settings, strategies as st

    lib = bpp.ffi.lib()
    lib.bpp_host_lockstep_selftest.restype = C.c_int32
    lib.bpp_host_lockstep_selftest.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_char_p, C.c_char_p, C.c_char_p]
    base = pyref.Transcript(b"lock-step selftest")

    @settings(max_examples=60, deadline=None)
    @given(st.integers(0, 400), st.integers(1, 400), st.integers(1, 300), st.integers(0, 2**32 - 1))
    def run(prefix_len, msg_len, wlen, seed):
        xof = hashlib.shake_256(b"rnd-%d-%d-%d-%d" % (prefix_len, msg_len, wlen, seed)).digest(8 * (prefix_len + msg_len + wlen + 32))
        off = 0
        states = b""
        for _ in range(8):
            t = base.clone()
            t.append_message(b"prefix", xof[off:off + prefix_len])
            off += prefix_len
            states += t.s.to_wire()
        msgs = xof[off:off + 8 * msg_len]
        off += 8 * msg_len
        wits = xof[off:off + 8 * wlen]
        off += 8 * wlen
        exts = xof[off:off + 256]
        a, b = C.create_string_buffer(8 * REC), C.create_string_buffer(8 * REC)
        assert lib.bpp_host_lockstep_selftest(states, msgs, msg_len, wits, wlen, exts, a, b) == 0
        assert a.raw == b.raw

    run()


def test_window_choice_is_what_the_bench_counts_with():
    lib = bpp.ffi.lib()
    lib.bpp_msm_window_bits.restype = C.c_int32
    lib.bpp_msm_window_bits.argtypes = [C.c_size_t, C.c_size_t]
    assert lib.bpp_msm_window_bits(64 * 4226, 64) == 9          # the verifier's per-call sums of a 16-job pass
    assert lib.bpp_msm_window_bits(64 * 4226, 1) == 14          # the same entries as one merged sum
    assert lib.bpp_msm_window_bits(0, 1) == 0
    for lg in range(8, 25):                                     # (not monotonic: the cost model follows the kernel variants)
        assert 4 <= lib.bpp_msm_window_bits(1 << lg, 1) <= 16
