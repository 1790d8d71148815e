// Warp-cooperative STROBE-128 / Merlin for the device-side transcript replay (k_replay.cu).
//
// One warp owns one sponge: lane L < 25 holds Keccak lane L (x = L % 5, y = L / 5) in a 64-bit register, pos / pos_begin /
// cur_flags are warp-uniform.  A Keccak-f round is 9 64-bit shuffles + ~25 ALU instructions per lane instead of ~190
// instructions in one thread, and the round's shuffles are independent within each of its four steps, so a permutation is
// ~4x shorter as a latency chain -- which is all that matters for a 1024-proof batch (32 warps would otherwise each walk ~21
// dependent permutations alone on their SM sub-partition).  Same interface as the thread-serial Merlin / MerlinRng of
// hash.cuh (the host path and the unit tests use those), so replay.cuh compiles one restatement of loop 1 against either.
// All 32 lanes of the warp must call every method together with identical (warp-uniform) arguments; buffers passed in are
// read by every lane, buffers passed out are written identically by every lane.
#pragma once
#include "hash.cuh"

namespace bpp {

struct WStrobe128 {
    uint64_t v;                       // this lane's Keccak lane (lanes 25..31 carry zeros)
    uint8_t pos, pos_begin, cur_flags;
    static constexpr int RATE = 166;
    static constexpr unsigned FULL = 0xffffffffu;

    __device__ __forceinline__ static int lane() { return (int)(threadIdx.x & 31); }
    __device__ __forceinline__ static uint64_t shfl64(uint64_t x, int src) { return (uint64_t)__shfl_sync(FULL, (unsigned long long)x, src); }

    __device__ __noinline__ void permute() {
        const int L = lane(), Lc = L < 25 ? L : 24, x = Lc % 5, y = Lc / 5;
        const uint8_t ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
        const int rot = ROT[Lc];
        const int col1 = (Lc + 5) % 25, col2 = (Lc + 10) % 25, col3 = (Lc + 15) % 25, col4 = (Lc + 20) % 25;
        const int xm1 = y * 5 + (x + 4) % 5, xp1 = y * 5 + (x + 1) % 5, xp2 = y * 5 + (x + 2) % 5;
        const int pi_src = 5 * x + (x + 3 * y) % 5;      // lane (X, Y) receives from (x, y) = ((X + 3Y) % 5, X)
        uint64_t a = v;
        BPP_HASH_ROLLED
        for (int round = 0; round < 24; round++) {
            uint64_t c = a ^ shfl64(a, col1) ^ shfl64(a, col2) ^ shfl64(a, col3) ^ shfl64(a, col4);     // theta: column parity
            uint64_t cm = shfl64(c, xm1), cp = shfl64(c, xp1);
            a ^= cm ^ ((cp << 1) | (cp >> 63));
            uint64_t r = rot ? ((a << rot) | (a >> (64 - rot))) : a;                                     // rho
            uint64_t b = shfl64(r, pi_src);                                                              // pi
            uint64_t b1 = shfl64(b, xp1), b2 = shfl64(b, xp2);
            a = b ^ (~b1 & b2);                                                                          // chi
            if (L == 0) a ^= keccak_rc(round);                                                           // iota
        }
        v = L < 25 ? a : 0;
    }
    __device__ __forceinline__ void xor_state_byte(int p, uint8_t bt) {      // warp-uniform p
        if (lane() == (p >> 3)) v ^= (uint64_t)bt << (8 * (p & 7));
    }
    __device__ __noinline__ void run_f() {
        xor_state_byte(pos, pos_begin);
        xor_state_byte(pos + 1, 0x04);
        xor_state_byte(RATE + 1, 0x80);
        permute();
        pos = 0;
        pos_begin = 0;
    }
    // state bytes [pos, pos + n) (op)= d[0..n), n <= RATE - pos: every lane handles the bytes that fall into its own 8
    template <int OP> __device__ __forceinline__ void apply_run(const uint8_t *d, int n) {
        const int lo = 8 * lane();
        uint64_t m = 0, keep = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            int idx = lo + j - (int)pos;
            if (idx >= 0 && idx < n) { m |= (uint64_t)d[idx] << (8 * j); keep |= 0xffull << (8 * j); }
        }
        if (OP == 0) v ^= m;                       // absorb
        else v = (v & ~keep) | m;                  // overwrite
    }
    template <int OP> __device__ __noinline__ void absorb_like(const uint8_t *d, size_t len) {
        while (len) {
            int n = (int)((size_t)(RATE - pos) < len ? (size_t)(RATE - pos) : len);
            apply_run<OP>(d, n);
            pos = (uint8_t)(pos + n); d += n; len -= (size_t)n;
            if (pos == RATE) run_f();
        }
    }
    __device__ __forceinline__ void absorb(const uint8_t *d, size_t len) { absorb_like<0>(d, len); }
    __device__ __forceinline__ void overwrite(const uint8_t *d, size_t len) { absorb_like<1>(d, len); }
    // every lane receives all squeezed bytes (d is a per-lane buffer with identical contents afterwards)
    __device__ __noinline__ void squeeze(uint8_t *d, size_t len) {
        while (len) {
            int n = (int)((size_t)(RATE - pos) < len ? (size_t)(RATE - pos) : len);
            for (int i = 0; i < n; i++) {
                int p = pos + i;
                uint64_t w = shfl64(v, p >> 3);
                d[i] = (uint8_t)(w >> (8 * (p & 7)));
            }
            const int lo = 8 * lane();
            uint64_t keep = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                int idx = lo + j - (int)pos;
                if (idx >= 0 && idx < n) keep |= 0xffull << (8 * j);
            }
            v &= ~keep;
            pos = (uint8_t)(pos + n); d += n; len -= (size_t)n;
            if (pos == RATE) run_f();
        }
    }
    __device__ __noinline__ void begin_op(uint8_t flags, bool more) {
        if (more) return;
        uint8_t hdr[2] = {pos_begin, flags};
        pos_begin = (uint8_t)(pos + 1);
        cur_flags = flags;
        absorb(hdr, 2);
        if ((flags & (Strobe128::FC | Strobe128::FK)) && pos != 0) run_f();
    }
    __device__ void meta_ad(const uint8_t *d, size_t len, bool more) { begin_op(Strobe128::FM | Strobe128::FA, more); absorb(d, len); }
    __device__ void ad(const uint8_t *d, size_t len, bool more) { begin_op(Strobe128::FA, more); absorb(d, len); }
    __device__ void prf(uint8_t *d, size_t len, bool more) { begin_op(Strobe128::FI | Strobe128::FA | Strobe128::FC, more); squeeze(d, len); }
    __device__ void key(const uint8_t *d, size_t len, bool more) { begin_op(Strobe128::FA | Strobe128::FC, more); overwrite(d, len); }

    // 203-byte wire form: 200 state bytes, pos, pos_begin, cur_flags
    __device__ void load(const uint8_t *b) {
        const int L = lane();
        uint64_t x = 0;
        if (L < 25)
            for (int j = 7; j >= 0; j--) x = (x << 8) | b[8 * L + j];
        v = x;
        pos = b[200]; pos_begin = b[201]; cur_flags = b[202];
    }
    __device__ void store(uint8_t *b) const {
        const int L = lane();
        if (L < 25)
            for (int j = 0; j < 8; j++) b[8 * L + j] = (uint8_t)(v >> (8 * j));
        if (L == 0) { b[200] = pos; b[201] = pos_begin; b[202] = cur_flags; }
    }
};

struct WMerlin {
    WStrobe128 s;
    __device__ void append_message(const uint8_t *label, size_t label_len, const uint8_t *msg, size_t len) {
        uint8_t l4[4];
        le32_bytes(l4, (uint32_t)len);
        s.meta_ad(label, label_len, false);
        s.meta_ad(l4, 4, true);
        s.ad(msg, len, false);
    }
    __device__ void append_u64(const uint8_t *label, size_t label_len, uint64_t x) {
        uint8_t b[8];
        le64_bytes(b, x);
        append_message(label, label_len, b, 8);
    }
    __device__ void challenge_bytes(const uint8_t *label, size_t label_len, uint8_t *out, size_t len) {
        uint8_t l4[4];
        le32_bytes(l4, (uint32_t)len);
        s.meta_ad(label, label_len, false);
        s.meta_ad(l4, 4, true);
        s.prf(out, len, false);
    }
};

struct WMerlinRng {
    WStrobe128 s;
    __device__ void build(const WMerlin &t, const uint8_t *witness, size_t wlen, bool have_witness, const uint8_t ext32[32]) {
        s = t.s;
        if (have_witness) {
            const uint8_t wl[7] = {'w', 'i', 't', 'n', 'e', 's', 's'};
            uint8_t l4[4];
            le32_bytes(l4, (uint32_t)wlen);
            s.meta_ad(wl, 7, false);
            s.meta_ad(l4, 4, true);
            s.key(witness, wlen, false);
        }
        const uint8_t rl[3] = {'r', 'n', 'g'};
        s.meta_ad(rl, 3, false);
        s.key(ext32, 32, false);
    }
    __device__ void fill(uint8_t *dst, size_t len) {
        uint8_t l4[4];
        le32_bytes(l4, (uint32_t)len);
        s.meta_ad(l4, 4, false);
        s.prf(dst, len, false);
    }
};

} // namespace bpp
