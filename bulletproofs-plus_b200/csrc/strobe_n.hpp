// StrobeN: LANES (4 or 8) STROBE-128 sponges advancing in lock-step on the host (host_keccak4.cpp permutes the states with ONE
// vectorised Keccak-f: AVX-512 for eight, AVX2 / AVX-512VL for four).  Sponges that perform the same operations at the same positions
// -- the verifier-weight transcripts of equally long chunks (engine_verify.cu), the Fiat-Shamir transcripts and TranscriptRngs of the
// proofs of one bpp_prove_batch call (engine_prove.cu) -- differ only in the bytes they absorb.  Mirrors Strobe128 / Merlin / MerlinRng
// of hash.cuh operation by operation (merlin 3.0.0 src/strobe.rs, src/transcript.rs); tests/test_abi_host.py compares the lock-step
// weights with the one-at-a-time sponge, tests/test_gpu_prove.py the lock-step prover with the scalar one byte for byte.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include "hash.cuh"

extern "C" void bpp_keccak_f1600_x4(uint64_t *st);
extern "C" void bpp_keccak_f1600_x8(uint64_t *st);

namespace bpp {

template <int LANES> struct StrobeN {
    alignas(64) uint64_t st[25 * LANES];          // lane k of state j at st[LANES * k + j]
    uint8_t pos = 0, pos_begin = 0, cur_flags = 0;
    static constexpr int RATE = Strobe128::RATE;
    void permute() { if (LANES == 8) bpp_keccak_f1600_x8(st); else bpp_keccak_f1600_x4(st); }
    void load_all(const uint8_t *b) {      // the same 203-byte state into every lane
        for (int k = 0; k < 25; k++) {
            uint64_t x = 0;
            for (int j = 7; j >= 0; j--) x = (x << 8) | b[8 * k + j];
            for (int j = 0; j < LANES; j++) st[LANES * k + j] = x;
        }
        pos = b[200]; pos_begin = b[201]; cur_flags = b[202];
    }
    void wipe() {                           // a sponge keyed with witness bytes is a secret (the prover's TranscriptRngs)
        volatile uint64_t *p = st;
        for (int i = 0; i < 25 * LANES; i++) p[i] = 0;
    }
    // lane j <- / -> one scalar sponge; the position bytes are shared, so only sponges that agree on them may share a StrobeN
    void load_lane(int j, const Strobe128 &s) {
        for (int k = 0; k < 25; k++) st[LANES * k + j] = s.st[k];
        pos = s.pos; pos_begin = s.pos_begin; cur_flags = s.cur_flags;
    }
    void store_lane(int j, Strobe128 &s) const {
        for (int k = 0; k < 25; k++) s.st[k] = st[LANES * k + j];
        s.pos = pos; s.pos_begin = pos_begin; s.cur_flags = cur_flags;
    }
    void xor_all(int p, uint8_t v) {
        const uint64_t x = (uint64_t)v << (8 * (p & 7));
        uint64_t *l = st + LANES * (p >> 3);
        for (int j = 0; j < LANES; j++) l[j] ^= x;
    }
    void run_f() {
        xor_all(pos, pos_begin); xor_all(pos + 1, 0x04); xor_all(RATE + 1, 0x80);
        permute();
        pos = 0; pos_begin = 0;
    }
    // Spans move up to eight bytes at a time as one 64-bit word per state (a word may straddle two sponge lanes); the byte loops
    // they replace cost about as much as the permutations they fed.
    static uint64_t load_le(const uint8_t *d, size_t n) {           // n <= 8
        uint64_t x = 0;
        if (n == 8) memcpy(&x, d, 8);
        else for (size_t i = 0; i < n; i++) x |= (uint64_t)d[i] << (8 * i);
        return x;
    }
    void absorb_same(const uint8_t *d, size_t len) {
        while (len) {
            size_t n = len < 8 ? len : 8;
            if (n > (size_t)(RATE - pos)) n = (size_t)(RATE - pos);
            const uint64_t x = load_le(d, n);
            const int off = pos & 7, sh = 8 * off;
            uint64_t *l = st + LANES * (pos >> 3);
            const uint64_t lo = x << sh;
            for (int j = 0; j < LANES; j++) l[j] ^= lo;
            if (off + (int)n > 8) { const uint64_t hi = x >> (64 - sh); for (int j = 0; j < LANES; j++) l[LANES + j] ^= hi; }
            d += n; len -= n;
            pos = (uint8_t)(pos + n);
            if (pos == RATE) run_f();
        }
    }
    void absorb_each(const uint8_t *const d[LANES], size_t len) {
        size_t i = 0;
        while (i < len) {
            size_t n = len - i < 8 ? len - i : 8;
            if (n > (size_t)(RATE - pos)) n = (size_t)(RATE - pos);
            const int off = pos & 7, sh = 8 * off;
            uint64_t *l = st + LANES * (pos >> 3);
            const bool straddle = off + (int)n > 8;
            for (int j = 0; j < LANES; j++) {
                const uint64_t x = load_le(d[j] + i, n);
                l[j] ^= x << sh;
                if (straddle) l[LANES + j] ^= x >> (64 - sh);
            }
            i += n;
            pos = (uint8_t)(pos + n);
            if (pos == RATE) run_f();
        }
    }
    void overwrite_same(const uint8_t *d, size_t len) {
        for (size_t i = 0; i < len; i++) {
            const int sh = 8 * (pos & 7);
            uint64_t *l = st + LANES * (pos >> 3);
            for (int j = 0; j < LANES; j++) l[j] = (l[j] & ~(0xffULL << sh)) | ((uint64_t)d[i] << sh);
            if (++pos == RATE) run_f();
        }
    }
    void overwrite_each(const uint8_t *const d[LANES], size_t len) {
        for (size_t i = 0; i < len; i++) {
            const int sh = 8 * (pos & 7);
            uint64_t *l = st + LANES * (pos >> 3);
            for (int j = 0; j < LANES; j++) l[j] = (l[j] & ~(0xffULL << sh)) | ((uint64_t)d[j][i] << sh);
            if (++pos == RATE) run_f();
        }
    }
    void squeeze_each(uint8_t *const d[LANES], size_t len) {
        size_t i = 0;
        while (i < len) {
            if ((pos & 7) == 0 && len - i >= 8 && pos + 8 <= RATE) {         // a whole lane: read it and clear it
                uint64_t *l = st + LANES * (pos >> 3);
                for (int j = 0; j < LANES; j++) { memcpy(d[j] + i, &l[j], 8); l[j] = 0; }
                i += 8;
                pos = (uint8_t)(pos + 8);
            } else {
                const int sh = 8 * (pos & 7);
                uint64_t *l = st + LANES * (pos >> 3);
                for (int j = 0; j < LANES; j++) { d[j][i] = (uint8_t)(l[j] >> sh); l[j] &= ~(0xffULL << sh); }
                i++;
                pos++;
            }
            if (pos == RATE) run_f();
        }
    }
    void begin_op(uint8_t flags, bool more) {
        if (more) return;
        const uint8_t hdr[2] = {pos_begin, flags};
        pos_begin = (uint8_t)(pos + 1);
        cur_flags = flags;
        absorb_same(hdr, 2);
        if ((flags & (Strobe128::FC | Strobe128::FK)) && pos != 0) run_f();
    }
    void meta_ad_same(const uint8_t *d, size_t len, bool more) { begin_op(Strobe128::FM | Strobe128::FA, more); absorb_same(d, len); }
    void ad_each(const uint8_t *const d[LANES], size_t len) { begin_op(Strobe128::FA, false); absorb_each(d, len); }
    void ad_same(const uint8_t *d, size_t len) { begin_op(Strobe128::FA, false); absorb_same(d, len); }
    void key_each(const uint8_t *const d[LANES], size_t len) { begin_op(Strobe128::FA | Strobe128::FC, false); overwrite_each(d, len); }
    void key_same(const uint8_t *d, size_t len) { begin_op(Strobe128::FA | Strobe128::FC, false); overwrite_same(d, len); }
    void prf_each(uint8_t *const d[LANES], size_t len) { begin_op(Strobe128::FI | Strobe128::FA | Strobe128::FC, false); squeeze_each(d, len); }
};

// Merlin / TranscriptRng over StrobeN (hash.cuh: Merlin::append_message, ::challenge_bytes, MerlinRng::build, ::fill)
template <int LANES> struct MerlinN {
    StrobeN<LANES> s;
    void label_len(const uint8_t *label, size_t ll, size_t len) {
        uint8_t l4[4];
        le32_bytes(l4, (uint32_t)len);
        s.meta_ad_same(label, ll, false);
        s.meta_ad_same(l4, 4, true);
    }
    void append_same(const uint8_t *label, size_t ll, const uint8_t *msg, size_t len) { label_len(label, ll, len); s.ad_same(msg, len); }
    void append_each(const uint8_t *label, size_t ll, const uint8_t *const msg[LANES], size_t len) { label_len(label, ll, len); s.ad_each(msg, len); }
    void append_u64_same(const uint8_t *label, size_t ll, uint64_t x) {
        uint8_t b[8];
        le64_bytes(b, x);
        append_same(label, ll, b, 8);
    }
    void challenge_each(const uint8_t *label, size_t ll, uint8_t *const out[LANES], size_t len) { label_len(label, ll, len); s.prf_each(out, len); }
    // transcript.build_rng().rekey_with_witness_bytes("witness", w_j).finalize(rng_j) into `out`; *this is untouched
    // (built inside `out`, so that no keyed copy is left behind on the stack)
    void build_rng(StrobeN<LANES> &out, const uint8_t *const witness[LANES], size_t wlen, const uint8_t *const ext32[LANES]) const {
        const uint8_t wl[7] = {'w', 'i', 't', 'n', 'e', 's', 's'}, rl[3] = {'r', 'n', 'g'};
        uint8_t l4[4];
        le32_bytes(l4, (uint32_t)wlen);
        out = s;
        out.meta_ad_same(wl, 7, false);
        out.meta_ad_same(l4, 4, true);
        out.key_each(witness, wlen);
        out.meta_ad_same(rl, 3, false);
        out.key_each(ext32, 32);
    }
};
template <int LANES> inline void rng_fill_each(StrobeN<LANES> &r, uint8_t *const out[LANES], size_t len) {
    uint8_t l4[4];
    le32_bytes(l4, (uint32_t)len);
    r.meta_ad_same(l4, 4, false);
    r.prf_each(out, len);
}

} // namespace bpp
