// Host-side Keccak-f[1600] on FOUR independent states at once (one 64-bit lane of each state per 256-bit vector), for the
// verifier-weight transcripts of the batch verifier (/root/reference/src/range_proof.rs:811-853, :894).
//
// Each reference call (chunk of <= 256 proofs) owns one sequential weight transcript: ~1.3 Keccak-f per proof, no parallelism
// inside it -- but the chunks of one bpp_verify_chunks call are independent and, when they hold the same number of proofs, walk
// exactly the same sponge positions, so four of them advance in lock-step through one vectorised permutation.  With the host
// cores shared by 8 GPU ranks this hashing is what bounded the 8-GPU throughput (0.8 ms of one core per 1024-proof step).
//
// Plain GCC vector extensions (no intrinsics); the AVX2 body is selected at run time (function multiversioning by hand), the
// generic body is the same code compiled for the baseline ISA.  Written from FIPS 202.
#include <stdint.h>

typedef uint64_t v4u __attribute__((vector_size(32)));

static const uint64_t RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, 0x0000000080000001ULL,
    0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
    0x000000000000800aULL, 0x800000008000000aULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};

#define ROL(x, n) (((x) << (n)) | ((x) >> (64 - (n))))

#define KECCAK4_BODY                                                                                                                  \
    v4u a[25];                                                                                                                        \
    for (int i = 0; i < 25; i++) a[i] = ((const v4u *)st)[i];                                                                          \
    for (int round = 0; round < 24; round++) {                                                                                        \
        v4u c0 = a[0] ^ a[5] ^ a[10] ^ a[15] ^ a[20], c1 = a[1] ^ a[6] ^ a[11] ^ a[16] ^ a[21], c2 = a[2] ^ a[7] ^ a[12] ^ a[17] ^ a[22];    \
        v4u c3 = a[3] ^ a[8] ^ a[13] ^ a[18] ^ a[23], c4 = a[4] ^ a[9] ^ a[14] ^ a[19] ^ a[24];                                            \
        v4u d0 = c4 ^ ROL(c1, 1), d1 = c0 ^ ROL(c2, 1), d2 = c1 ^ ROL(c3, 1), d3 = c2 ^ ROL(c4, 1), d4 = c3 ^ ROL(c0, 1);                  \
        v4u b0 = a[0] ^ d0;                                                                                                            \
        v4u b10 = ROL(a[1] ^ d1, 1), b20 = ROL(a[2] ^ d2, 62), b5 = ROL(a[3] ^ d3, 28), b15 = ROL(a[4] ^ d4, 27);                          \
        v4u b16 = ROL(a[5] ^ d0, 36), b1 = ROL(a[6] ^ d1, 44), b11 = ROL(a[7] ^ d2, 6), b21 = ROL(a[8] ^ d3, 55), b6 = ROL(a[9] ^ d4, 20);   \
        v4u b7 = ROL(a[10] ^ d0, 3), b17 = ROL(a[11] ^ d1, 10), b2 = ROL(a[12] ^ d2, 43), b12 = ROL(a[13] ^ d3, 25), b22 = ROL(a[14] ^ d4, 39); \
        v4u b23 = ROL(a[15] ^ d0, 41), b8 = ROL(a[16] ^ d1, 45), b18 = ROL(a[17] ^ d2, 15), b3 = ROL(a[18] ^ d3, 21), b13 = ROL(a[19] ^ d4, 8); \
        v4u b14 = ROL(a[20] ^ d0, 18), b24 = ROL(a[21] ^ d1, 2), b9 = ROL(a[22] ^ d2, 61), b19 = ROL(a[23] ^ d3, 56), b4 = ROL(a[24] ^ d4, 14); \
        a[0] = b0 ^ (~b1 & b2); a[1] = b1 ^ (~b2 & b3); a[2] = b2 ^ (~b3 & b4); a[3] = b3 ^ (~b4 & b0); a[4] = b4 ^ (~b0 & b1);             \
        a[5] = b5 ^ (~b6 & b7); a[6] = b6 ^ (~b7 & b8); a[7] = b7 ^ (~b8 & b9); a[8] = b8 ^ (~b9 & b5); a[9] = b9 ^ (~b5 & b6);             \
        a[10] = b10 ^ (~b11 & b12); a[11] = b11 ^ (~b12 & b13); a[12] = b12 ^ (~b13 & b14); a[13] = b13 ^ (~b14 & b10); a[14] = b14 ^ (~b10 & b11); \
        a[15] = b15 ^ (~b16 & b17); a[16] = b16 ^ (~b17 & b18); a[17] = b17 ^ (~b18 & b19); a[18] = b18 ^ (~b19 & b15); a[19] = b19 ^ (~b15 & b16); \
        a[20] = b20 ^ (~b21 & b22); a[21] = b21 ^ (~b22 & b23); a[22] = b22 ^ (~b23 & b24); a[23] = b23 ^ (~b24 & b20); a[24] = b24 ^ (~b20 & b21); \
        const uint64_t rc = RC[round];                                                                                                \
        a[0] ^= (v4u){rc, rc, rc, rc};                                                                                                 \
    }                                                                                                                                 \
    for (int i = 0; i < 25; i++) ((v4u *)st)[i] = a[i];

__attribute__((target("avx2"))) static void keccak4_avx2(uint64_t *st) { KECCAK4_BODY }
// AVX-512VL: 32 vector registers (no spills of the 25 + 25 live values), native 64-bit rotates and three-input logic
__attribute__((target("avx2,avx512f,avx512vl"))) static void keccak4_avx512vl(uint64_t *st) { KECCAK4_BODY }
static void keccak4_generic(uint64_t *st) { KECCAK4_BODY }

extern "C" {
// st: 25 x 4 lanes, lane k of state j at st[4 * k + j]; 32-byte aligned
void bpp_keccak_f1600_x4(uint64_t *st) {
    static const int level = (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vl")) ? 2 : __builtin_cpu_supports("avx2") ? 1 : 0;
    if (level == 2) keccak4_avx512vl(st); else if (level == 1) keccak4_avx2(st); else keccak4_generic(st);
}
int bpp_host_has_avx2(void) { return __builtin_cpu_supports("avx2") ? 1 : 0; }
}

// ------------------------------------------------------------------------------------------------ Scalar::from_bytes_mod_order_wide
// 64 bytes -> canonical scalar mod l, on 64-bit limbs (the shared 32-bit-limb arithmetic of arith.cuh is what the GPU wants; a host
// core does the same Montgomery products four times faster with 64 x 64 -> 128-bit multiplies).  One weight per proof goes through
// this (range_proof.rs:894), next to its Keccak-f.   lo + hi * 2^256 = montmul(lo, R) + montmul(hi, R^2)   (R = 2^256 mod l).
#include <string.h>
typedef unsigned __int128 u128;
static const uint64_t L64[4] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0x0ULL, 0x1000000000000000ULL};
static const uint64_t R64[4] = {0xd6ec31748d98951dULL, 0xc6ef5bf4737dcf70ULL, 0xfffffffffffffffeULL, 0x0fffffffffffffffULL};
static const uint64_t RR64[4] = {0xa40611e3449c0f01ULL, 0xd00e1ba768859347ULL, 0xceec73d217f5be65ULL, 0x0399411b7c309a3dULL};
static inline uint64_t lfactor64() {          // -l^-1 mod 2^64 by Newton iteration
    uint64_t inv = L64[0];
    for (int i = 0; i < 6; i++) inv *= 2 - L64[0] * inv;
    return (uint64_t)0 - inv;
}
static inline void cond_sub_l(uint64_t r[4]) {            // r < 2l -> r mod l
    uint64_t t[4];
    u128 bw = 0;
    for (int i = 0; i < 4; i++) { u128 d = (u128)r[i] - L64[i] - (uint64_t)bw; t[i] = (uint64_t)d; bw = (d >> 64) & 1; }
    if (!bw) for (int i = 0; i < 4; i++) r[i] = t[i];
}
static inline void montmul64(uint64_t out[4], const uint64_t a[4], const uint64_t b[4], uint64_t lf) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 4; j++) { u128 p = (u128)a[j] * b[i] + t[j] + c; t[j] = (uint64_t)p; c = (uint64_t)(p >> 64); }
        u128 s = (u128)t[4] + c;
        t[4] = (uint64_t)s; t[5] = (uint64_t)(s >> 64);
        const uint64_t m = t[0] * lf;
        u128 p = (u128)m * L64[0] + t[0];
        c = (uint64_t)(p >> 64);
        for (int j = 1; j < 4; j++) { p = (u128)m * L64[j] + t[j] + c; t[j - 1] = (uint64_t)p; c = (uint64_t)(p >> 64); }
        s = (u128)t[4] + c;
        t[3] = (uint64_t)s; t[4] = t[5] + (uint64_t)(s >> 64); t[5] = 0;
    }
    for (int i = 0; i < 4; i++) out[i] = t[i];       // < 2l, t[4] == 0
    cond_sub_l(out);
}
extern "C" void bpp_host_sc_from_wide64(const uint8_t in64[64], uint8_t out32[32]) {
    static const uint64_t lf = lfactor64();
    uint64_t lo[4], hi[4], x[4], y[4];
    memcpy(lo, in64, 32);          // little-endian host (x86-64 / aarch64), as everywhere in the host layer
    memcpy(hi, in64 + 32, 32);
    montmul64(x, lo, R64, lf);
    montmul64(y, hi, RR64, lf);
    uint64_t c = 0;
    for (int i = 0; i < 4; i++) { u128 s = (u128)x[i] + y[i] + c; x[i] = (uint64_t)s; c = (uint64_t)(s >> 64); }
    cond_sub_l(x);                 // x + y < 2l < 2^254: no carry out
    memcpy(out32, x, 32);
}

// a * b mod l on 64-bit limbs (a: any 256-bit value, b: canonical, so that a * b < 2^256 * l and one conditional subtraction after
// the Montgomery step suffices; little-endian; output canonical): the prover's host-side scalar
// bookkeeping (alpha updates, challenge powers, the final responses) is ~125 products per proof, which on the shared 32-bit-limb
// arithmetic was the largest single item of its host time.
extern "C" void bpp_host_sc_mul64(const uint8_t a32[32], const uint8_t b32[32], uint8_t out32[32]) {
    static const uint64_t lf = lfactor64();
    uint64_t a[4], b[4], t[4], r[4];
    memcpy(a, a32, 32);
    memcpy(b, b32, 32);
    montmul64(t, a, b, lf);        // a * b / R
    montmul64(r, t, RR64, lf);     // * R^2 / R
    memcpy(out32, r, 32);
}
