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
typedef uint64_t v8u __attribute__((vector_size(64)));      // eight states at once: one AVX-512 register per Keccak lane

static const uint64_t RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, 0x0000000080000001ULL,
    0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
    0x000000000000800aULL, 0x800000008000000aULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};

#define ROL(x, n) (((x) << (n)) | ((x) >> (64 - (n))))

// One round reads the 25 lanes of `a` and writes `e` plane by plane (theta, then for every output plane its five rho-pi sources and
// chi), the next round goes back from `e` to `a`: a source lane dies as soon as its plane is written, so the live set stays near the
// 32 vector registers of AVX-512VL (an "all 25 b, then all 25 a" round keeps 50 values live and spilt ~150 moves per round:
// 419 -> 370 ns per four-way permutation on a Sapphire-Rapids-class core).
#define KECCAK_BODY(KT, KB) \
    const KT *sp = (const KT *)st; \
    KT a0 = sp[0], a1 = sp[1], a2 = sp[2], a3 = sp[3], a4 = sp[4], a5 = sp[5], a6 = sp[6], a7 = sp[7], a8 = sp[8], a9 = sp[9], a10 = sp[10], a11 = sp[11], a12 = sp[12], a13 = sp[13], a14 = sp[14], a15 = sp[15], a16 = sp[16], a17 = sp[17], a18 = sp[18], a19 = sp[19], a20 = sp[20], a21 = sp[21], a22 = sp[22], a23 = sp[23], a24 = sp[24]; \
    KT e0, e1, e2, e3, e4, e5, e6, e7, e8, e9, e10, e11, e12, e13, e14, e15, e16, e17, e18, e19, e20, e21, e22, e23, e24; \
    for (int round = 0; round < 24; round += 2) { \
        const uint64_t rc0 = RC[round], rc1 = RC[round + 1]; \
        { KT c0 = a0 ^ a5 ^ a10 ^ a15 ^ a20, c1 = a1 ^ a6 ^ a11 ^ a16 ^ a21, c2 = a2 ^ a7 ^ a12 ^ a17 ^ a22, c3 = a3 ^ a8 ^ a13 ^ a18 ^ a23, c4 = a4 ^ a9 ^ a14 ^ a19 ^ a24; \
          KT d0 = c4 ^ ROL(c1, 1), d1 = c0 ^ ROL(c2, 1), d2 = c1 ^ ROL(c3, 1), d3 = c2 ^ ROL(c4, 1), d4 = c3 ^ ROL(c0, 1); \
          { KT b0 = (a0 ^ d0), b1 = ROL(a6 ^ d1, 44), b2 = ROL(a12 ^ d2, 43), b3 = ROL(a18 ^ d3, 21), b4 = ROL(a24 ^ d4, 14); \
            e0 = b0 ^ (~b1 & b2) ^ KB(rc0); e1 = b1 ^ (~b2 & b3); e2 = b2 ^ (~b3 & b4); e3 = b3 ^ (~b4 & b0); e4 = b4 ^ (~b0 & b1); } \
          { KT b0 = ROL(a3 ^ d3, 28), b1 = ROL(a9 ^ d4, 20), b2 = ROL(a10 ^ d0, 3), b3 = ROL(a16 ^ d1, 45), b4 = ROL(a22 ^ d2, 61); \
            e5 = b0 ^ (~b1 & b2); e6 = b1 ^ (~b2 & b3); e7 = b2 ^ (~b3 & b4); e8 = b3 ^ (~b4 & b0); e9 = b4 ^ (~b0 & b1); } \
          { KT b0 = ROL(a1 ^ d1, 1), b1 = ROL(a7 ^ d2, 6), b2 = ROL(a13 ^ d3, 25), b3 = ROL(a19 ^ d4, 8), b4 = ROL(a20 ^ d0, 18); \
            e10 = b0 ^ (~b1 & b2); e11 = b1 ^ (~b2 & b3); e12 = b2 ^ (~b3 & b4); e13 = b3 ^ (~b4 & b0); e14 = b4 ^ (~b0 & b1); } \
          { KT b0 = ROL(a4 ^ d4, 27), b1 = ROL(a5 ^ d0, 36), b2 = ROL(a11 ^ d1, 10), b3 = ROL(a17 ^ d2, 15), b4 = ROL(a23 ^ d3, 56); \
            e15 = b0 ^ (~b1 & b2); e16 = b1 ^ (~b2 & b3); e17 = b2 ^ (~b3 & b4); e18 = b3 ^ (~b4 & b0); e19 = b4 ^ (~b0 & b1); } \
          { KT b0 = ROL(a2 ^ d2, 62), b1 = ROL(a8 ^ d3, 55), b2 = ROL(a14 ^ d4, 39), b3 = ROL(a15 ^ d0, 41), b4 = ROL(a21 ^ d1, 2); \
            e20 = b0 ^ (~b1 & b2); e21 = b1 ^ (~b2 & b3); e22 = b2 ^ (~b3 & b4); e23 = b3 ^ (~b4 & b0); e24 = b4 ^ (~b0 & b1); } \
        } \
        { KT c0 = e0 ^ e5 ^ e10 ^ e15 ^ e20, c1 = e1 ^ e6 ^ e11 ^ e16 ^ e21, c2 = e2 ^ e7 ^ e12 ^ e17 ^ e22, c3 = e3 ^ e8 ^ e13 ^ e18 ^ e23, c4 = e4 ^ e9 ^ e14 ^ e19 ^ e24; \
          KT d0 = c4 ^ ROL(c1, 1), d1 = c0 ^ ROL(c2, 1), d2 = c1 ^ ROL(c3, 1), d3 = c2 ^ ROL(c4, 1), d4 = c3 ^ ROL(c0, 1); \
          { KT b0 = (e0 ^ d0), b1 = ROL(e6 ^ d1, 44), b2 = ROL(e12 ^ d2, 43), b3 = ROL(e18 ^ d3, 21), b4 = ROL(e24 ^ d4, 14); \
            a0 = b0 ^ (~b1 & b2) ^ KB(rc1); a1 = b1 ^ (~b2 & b3); a2 = b2 ^ (~b3 & b4); a3 = b3 ^ (~b4 & b0); a4 = b4 ^ (~b0 & b1); } \
          { KT b0 = ROL(e3 ^ d3, 28), b1 = ROL(e9 ^ d4, 20), b2 = ROL(e10 ^ d0, 3), b3 = ROL(e16 ^ d1, 45), b4 = ROL(e22 ^ d2, 61); \
            a5 = b0 ^ (~b1 & b2); a6 = b1 ^ (~b2 & b3); a7 = b2 ^ (~b3 & b4); a8 = b3 ^ (~b4 & b0); a9 = b4 ^ (~b0 & b1); } \
          { KT b0 = ROL(e1 ^ d1, 1), b1 = ROL(e7 ^ d2, 6), b2 = ROL(e13 ^ d3, 25), b3 = ROL(e19 ^ d4, 8), b4 = ROL(e20 ^ d0, 18); \
            a10 = b0 ^ (~b1 & b2); a11 = b1 ^ (~b2 & b3); a12 = b2 ^ (~b3 & b4); a13 = b3 ^ (~b4 & b0); a14 = b4 ^ (~b0 & b1); } \
          { KT b0 = ROL(e4 ^ d4, 27), b1 = ROL(e5 ^ d0, 36), b2 = ROL(e11 ^ d1, 10), b3 = ROL(e17 ^ d2, 15), b4 = ROL(e23 ^ d3, 56); \
            a15 = b0 ^ (~b1 & b2); a16 = b1 ^ (~b2 & b3); a17 = b2 ^ (~b3 & b4); a18 = b3 ^ (~b4 & b0); a19 = b4 ^ (~b0 & b1); } \
          { KT b0 = ROL(e2 ^ d2, 62), b1 = ROL(e8 ^ d3, 55), b2 = ROL(e14 ^ d4, 39), b3 = ROL(e15 ^ d0, 41), b4 = ROL(e21 ^ d1, 2); \
            a20 = b0 ^ (~b1 & b2); a21 = b1 ^ (~b2 & b3); a22 = b2 ^ (~b3 & b4); a23 = b3 ^ (~b4 & b0); a24 = b4 ^ (~b0 & b1); } \
        } \
    } \
    KT *dp = (KT *)st; \
    dp[0] = a0; dp[1] = a1; dp[2] = a2; dp[3] = a3; dp[4] = a4; dp[5] = a5; dp[6] = a6; dp[7] = a7; dp[8] = a8; dp[9] = a9; dp[10] = a10; dp[11] = a11; dp[12] = a12; dp[13] = a13; dp[14] = a14; dp[15] = a15; dp[16] = a16; dp[17] = a17; dp[18] = a18; dp[19] = a19; dp[20] = a20; dp[21] = a21; dp[22] = a22; dp[23] = a23; dp[24] = a24;

#define KB4(rc) ((v4u){rc, rc, rc, rc})
#define KB1(rc) (rc)
__attribute__((target("avx2"))) static void keccak4_avx2(uint64_t *st) { KECCAK_BODY(v4u, KB4) }
// AVX-512VL: 32 vector registers (no spills of the 25 + 25 live values), native 64-bit rotates and three-input logic
__attribute__((target("avx2,avx512f,avx512vl"))) static void keccak4_avx512vl(uint64_t *st) { KECCAK_BODY(v4u, KB4) }
static void keccak4_generic(uint64_t *st) { KECCAK_BODY(v4u, KB4) }
// EIGHT states per permutation on AVX-512F: 25 zmm registers hold the state, vprolq / vpternlogq do a rotate / a chi term in one
// instruction each (~100 instructions per round for eight permutations)
#define KB8(rc) ((v8u){rc, rc, rc, rc, rc, rc, rc, rc})
__attribute__((target("avx2,avx512f,avx512vl"))) static void keccak8_avx512(uint64_t *st) { KECCAK_BODY(v8u, KB8) }
static void keccak1_generic(uint64_t *st) { KECCAK_BODY(uint64_t, KB1) }

#include <stdlib.h>
// 2 = AVX-512VL, 1 = AVX2, 0 = baseline ISA; BPP_HOST_SIMD caps it (tests run every body on a CPU that has them all)
static int simd_level() {
    static const int level = [] {
        int l = (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vl")) ? 2 : __builtin_cpu_supports("avx2") ? 1 : 0;
        if (const char *env = getenv("BPP_HOST_SIMD")) { int cap = atoi(env); if (cap >= 0 && cap < l) l = cap; }
        return l;
    }();
    return level;
}
extern "C" {
// st: 25 x 4 lanes, lane k of state j at st[4 * k + j]; 32-byte aligned
void bpp_keccak_f1600_x4(uint64_t *st) {
    const int level = simd_level();
    if (level == 2) keccak4_avx512vl(st); else if (level == 1) keccak4_avx2(st); else keccak4_generic(st);
}
// st: 25 x 8 lanes, lane k of state j at st[8 * k + j]; 64-byte aligned.  Without AVX-512 the eight states go through two four-way
// permutations.
void bpp_keccak_f1600_x8(uint64_t *st) {
    if (simd_level() == 2) { keccak8_avx512(st); return; }
    alignas(32) uint64_t x[2][100];
    for (int k = 0; k < 25; k++)
        for (int j = 0; j < 4; j++) { x[0][4 * k + j] = st[8 * k + j]; x[1][4 * k + j] = st[8 * k + 4 + j]; }
    bpp_keccak_f1600_x4(x[0]);
    bpp_keccak_f1600_x4(x[1]);
    for (int k = 0; k < 25; k++)
        for (int j = 0; j < 4; j++) { st[8 * k + j] = x[0][4 * k + j]; st[8 * k + 4 + j] = x[1][4 * k + j]; }
}
// ONE state (25 lanes): every host-side sponge of the library (hash.cuh on the host: the prover's transcripts and TranscriptRng, SHA3 /
// SHAKE, host-mode replay) permutes through this.  With AVX2 / AVX-512VL the state rides in lane 0 of the four-way body: 16
// general-purpose registers cannot hold 25 lanes + 25 temporaries, 16 / 32 vector registers do much better -- measured 1070 ns for the
// plain 64-bit code against 390 ns through the vector body (three lanes idle) on an AVX-512 core.
void bpp_keccak_f1600_x1(uint64_t *st) {
    const int level = simd_level();
    if (level == 0) { keccak1_generic(st); return; }
    alignas(32) uint64_t x[100];
    for (int k = 0; k < 25; k++) { x[4 * k] = st[k]; x[4 * k + 1] = 0; x[4 * k + 2] = 0; x[4 * k + 3] = 0; }
    if (level == 2) keccak4_avx512vl(x); else keccak4_avx2(x);
    for (int k = 0; k < 25; k++) st[k] = x[4 * k];
}
// test hook: the plain 64-bit body whatever the CPU
void bpp_keccak_f1600_x1_generic(uint64_t *st) { keccak1_generic(st); }
int bpp_host_has_avx2(void) { return __builtin_cpu_supports("avx2") ? 1 : 0; }
int32_t bpp_host_simd_level(void) { return simd_level(); }
}

// ------------------------------------------------------------------------------------------------ Scalar::from_bytes_mod_order_wide
// 64 bytes -> canonical scalar mod l, and a * b mod l, on 64-bit limbs (the shared 32-bit-limb arithmetic of arith.cuh is what the
// GPU wants; a host core does the same with 64 x 64 -> 128-bit multiplies).  One weight per proof goes through the wide reduction
// (range_proof.rs:894), next to its Keccak-f; the prover's host-side scalar bookkeeping is ~125 products per proof.
//
// l = 2^252 + delta with delta < 2^125, so 2^252 = -delta (mod l) and a 512-bit value folds down in three short products:
//   x = a + b 2^252          (b < 2^260)   x = a - b delta,          t = b delta < 2^385
//   t = c + e 2^252          (e < 2^133)   t = c - e delta,          u = e delta < 2^258
//   u = f + g 2^252          (g < 2^6)     u = f - g delta,          v = g delta < 2^131
//   x = a - c + f - v  (mod l), every term < 2^252: r = (a + f + 2l) - (c + v) lies in (0, 4l), three conditional subtractions finish.
// 18 multiplications instead of the 64 of two Montgomery products.  Two bodies: MULX + add-with-carry intrinsics (x86-64 with BMI2,
// picked at run time; auto-vectorisation off, it turned the limb shifts into SSE code with store-forwarding stalls), and portable
// 128-bit integer code.  Measured per reduction: 110 ns (two Montgomery products) -> 45 ns.
#include <string.h>
typedef unsigned __int128 u128;
typedef unsigned long long ull;
static const ull L64[4] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0x0ULL, 0x1000000000000000ULL};
static const ull M60 = (1ULL << 60) - 1;

// ---- portable body
static inline void cond_sub_l(ull r[4]) {            // r < 2l -> r mod l (one step of it for larger r)
    ull t[4];
    u128 bw = 0;
    for (int i = 0; i < 4; i++) { u128 d = (u128)r[i] - L64[i] - (ull)bw; t[i] = (ull)d; bw = (d >> 64) & 1; }
    if (!bw) for (int i = 0; i < 4; i++) r[i] = t[i];
}
static inline void mul_delta(ull *out, const ull *b, int n) {       // out[0 .. n + 1] = b[0 .. n - 1] * delta (delta = L64[0..1])
    out[0] = 0; out[1] = 0;
    for (int i = 0; i < n; i++) {
        u128 p = (u128)b[i] * L64[0] + out[i];
        out[i] = (ull)p;
        p = (u128)b[i] * L64[1] + out[i + 1] + (ull)(p >> 64);
        out[i + 1] = (ull)p;
        out[i + 2] = (ull)(p >> 64);         // row i is the first to reach limb i + 2
    }
}
static void reduce512_generic(const ull x[8], ull r[4]) {
    ull b[5], t[7], e[3], u[5], g[1], v[3];
    for (int i = 0; i < 4; i++) b[i] = (x[3 + i] >> 60) | (x[4 + i] << 4);
    b[4] = x[7] >> 60;
    mul_delta(t, b, 5);
    for (int i = 0; i < 3; i++) e[i] = (t[3 + i] >> 60) | (t[4 + i] << 4);
    mul_delta(u, e, 3);
    g[0] = (u[3] >> 60) | (u[4] << 4);
    mul_delta(v, g, 1);
    const ull a[4] = {x[0], x[1], x[2], x[3] & M60}, f[4] = {u[0], u[1], u[2], u[3] & M60}, c[4] = {t[0], t[1], t[2], t[3] & M60};
    ull pos[4], neg[4];
    u128 cy = 0;
    for (int i = 0; i < 4; i++) { u128 s = (u128)L64[i] + L64[i] + (ull)cy; pos[i] = (ull)s; cy = s >> 64; }      // 2l
    cy = 0;
    for (int i = 0; i < 4; i++) { u128 s = (u128)pos[i] + a[i] + f[i] + (ull)cy; pos[i] = (ull)s; cy = s >> 64; }
    cy = 0;
    for (int i = 0; i < 4; i++) { u128 s = (u128)c[i] + (i < 3 ? v[i] : 0) + (ull)cy; neg[i] = (ull)s; cy = s >> 64; }
    u128 bw = 0;
    for (int i = 0; i < 4; i++) { u128 d = (u128)pos[i] - neg[i] - (ull)bw; r[i] = (ull)d; bw = (d >> 64) & 1; }
    cond_sub_l(r); cond_sub_l(r); cond_sub_l(r);
}
static void mul256_generic(const ull a[4], const ull b[4], ull out[8]) {
    for (int i = 0; i < 8; i++) out[i] = 0;
    for (int i = 0; i < 4; i++) {
        ull c = 0;
        for (int j = 0; j < 4; j++) { u128 p = (u128)a[j] * b[i] + out[i + j] + c; out[i + j] = (ull)p; c = (ull)(p >> 64); }
        out[i + 4] = c;
    }
}

// ---- MULX / ADC body
#if defined(__x86_64__)
#include <immintrin.h>
#define BPP_BMI2 __attribute__((target("bmi2"), optimize("no-tree-vectorize", "no-tree-slp-vectorize")))
template <int N> BPP_BMI2 static inline void mul_delta_bmi2(ull *out, const ull *b) {       // out[0 .. N + 1] = b[0 .. N - 1] * delta
    ull A[N + 1], B[N + 1];
    {
        ull hp = 0; unsigned char c = 0;
        for (int i = 0; i < N; i++) { ull hi, lo = _mulx_u64(b[i], L64[0], &hi); c = _addcarry_u64(c, lo, hp, &A[i]); hp = hi; }
        _addcarry_u64(c, hp, 0, &A[N]);
    }
    {
        ull hp = 0; unsigned char c = 0;
        for (int i = 0; i < N; i++) { ull hi, lo = _mulx_u64(b[i], L64[1], &hi); c = _addcarry_u64(c, lo, hp, &B[i]); hp = hi; }
        _addcarry_u64(c, hp, 0, &B[N]);
    }
    out[0] = A[0];
    unsigned char c = 0;
    for (int i = 1; i <= N; i++) c = _addcarry_u64(c, A[i], B[i - 1], &out[i]);
    _addcarry_u64(c, B[N], 0, &out[N + 1]);
}
BPP_BMI2 static inline void cond_sub_l_bmi2(ull r[4]) {
    ull t[4];
    unsigned char bw = 0;
    for (int i = 0; i < 4; i++) bw = _subborrow_u64(bw, r[i], L64[i], &t[i]);
    for (int i = 0; i < 4; i++) r[i] = bw ? r[i] : t[i];
}
BPP_BMI2 static void reduce512_bmi2(const ull x[8], ull r[4]) {
    ull b[5], t[7], e[3], u[5], g[1], v[3];
    for (int i = 0; i < 4; i++) b[i] = (x[3 + i] >> 60) | (x[4 + i] << 4);
    b[4] = x[7] >> 60;
    mul_delta_bmi2<5>(t, b);
    for (int i = 0; i < 3; i++) e[i] = (t[3 + i] >> 60) | (t[4 + i] << 4);
    mul_delta_bmi2<3>(u, e);
    g[0] = (u[3] >> 60) | (u[4] << 4);
    mul_delta_bmi2<1>(v, g);
    const ull a3 = x[3] & M60, f3 = u[3] & M60, c3 = t[3] & M60;
    ull p0, p1, p2, p3, n0, n1, n2, n3;
    unsigned char c = _addcarry_u64(0, x[0], u[0], &p0);                        // a + f
    c = _addcarry_u64(c, x[1], u[1], &p1); c = _addcarry_u64(c, x[2], u[2], &p2); _addcarry_u64(c, a3, f3, &p3);
    c = _addcarry_u64(0, p0, 0xb024c634b9eba7daULL, &p0);                        // + 2l
    c = _addcarry_u64(c, p1, 0x29bdf3bd45ef39acULL, &p1); c = _addcarry_u64(c, p2, 0, &p2); _addcarry_u64(c, p3, 0x2000000000000000ULL, &p3);
    c = _addcarry_u64(0, t[0], v[0], &n0);                                        // c + v
    c = _addcarry_u64(c, t[1], v[1], &n1); c = _addcarry_u64(c, t[2], v[2], &n2); _addcarry_u64(c, c3, 0, &n3);
    c = _subborrow_u64(0, p0, n0, &r[0]);
    c = _subborrow_u64(c, p1, n1, &r[1]); c = _subborrow_u64(c, p2, n2, &r[2]); _subborrow_u64(c, p3, n3, &r[3]);
    cond_sub_l_bmi2(r); cond_sub_l_bmi2(r); cond_sub_l_bmi2(r);
}
BPP_BMI2 static void mul256_bmi2(const ull a[4], const ull b[4], ull out[8]) {
    ull row[5];
    for (int i = 0; i < 8; i++) out[i] = 0;
    for (int i = 0; i < 4; i++) {
        ull hp = 0; unsigned char c = 0;
        for (int j = 0; j < 4; j++) { ull hi, lo = _mulx_u64(a[j], b[i], &hi); c = _addcarry_u64(c, lo, hp, &row[j]); hp = hi; }
        _addcarry_u64(c, hp, 0, &row[4]);
        c = 0;
        for (int j = 0; j < 5; j++) c = _addcarry_u64(c, out[i + j], row[j], &out[i + j]);      // out[i + 4] was 0: no carry out of the row
    }
}
static const bool have_bmi2 = __builtin_cpu_supports("bmi2") && !(getenv("BPP_HOST_SIMD") && atoi(getenv("BPP_HOST_SIMD")) == 0);
#else
static const bool have_bmi2 = false;
static void reduce512_bmi2(const ull x[8], ull r[4]) { reduce512_generic(x, r); }
static void mul256_bmi2(const ull a[4], const ull b[4], ull out[8]) { mul256_generic(a, b, out); }
#endif

extern "C" void bpp_host_sc_from_wide64(const uint8_t in64[64], uint8_t out32[32]) {
    ull x[8], r[4];
    memcpy(x, in64, 64);           // little-endian host (x86-64 / aarch64), as everywhere in the host layer
    if (have_bmi2) reduce512_bmi2(x, r); else reduce512_generic(x, r);
    memcpy(out32, r, 32);
}
// a * b mod l (any two 256-bit values, little-endian; output canonical)
extern "C" void bpp_host_sc_mul64(const uint8_t a32[32], const uint8_t b32[32], uint8_t out32[32]) {
    ull a[4], b[4], x[8], r[4];
    memcpy(a, a32, 32);
    memcpy(b, b32, 32);
    if (have_bmi2) { mul256_bmi2(a, b, x); reduce512_bmi2(x, r); } else { mul256_generic(a, b, x); reduce512_generic(x, r); }
    memcpy(out32, r, 32);
}
// test hook: the same two functions through the portable body whatever the CPU
extern "C" void bpp_host_sc_generic64(const uint8_t *a32_or_wide64, const uint8_t *b32_or_null, uint8_t out32[32]) {
    ull a[8], b[4], x[8], r[4];
    if (b32_or_null) { memcpy(a, a32_or_wide64, 32); memcpy(b, b32_or_null, 32); mul256_generic(a, b, x); }
    else memcpy(x, a32_or_wide64, 64);
    reduce512_generic(x, r);
    memcpy(out32, r, 32);
}
