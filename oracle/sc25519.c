/* ORACLE (test infrastructure only) — scalars mod l = 2^252 + 27742317777372353535851937790883648493.
 * Restates the behaviour of curve25519-dalek 4.1.3 `Scalar` as used by the reference
 * (/root/reference/src/range_proof.rs:350-392,426-435,507-537,590-594,897-1032;
 *  /root/reference/src/protocols/transcript_protocol.rs:67-78; scalar_protocol.rs:23-37).
 * Representation: canonical value in 4 x 64-bit limbs; multiplication via Montgomery (R = 2^256). */
#include "orc_internal.h"

static const uint64_t L[4] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0ULL, 0x1000000000000000ULL};
static const uint64_t RMOD[4] = {0xd6ec31748d98951dULL, 0xc6ef5bf4737dcf70ULL, 0xfffffffffffffffeULL, 0x0fffffffffffffffULL};
static const uint64_t RR[4] = {0xa40611e3449c0f01ULL, 0xd00e1ba768859347ULL, 0xceec73d217f5be65ULL, 0x0399411b7c309a3dULL};
#define LFACTOR 0xd2b51da312547e1bULL

void sc_0(sc *r) { memset(r, 0, sizeof *r); }
void sc_1(sc *r) { memset(r, 0, sizeof *r); r->v[0] = 1; }
void sc_from_u64(sc *r, uint64_t x) { memset(r, 0, sizeof *r); r->v[0] = x; }

static int geq_l(const uint64_t a[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > L[i]) return 1;
        if (a[i] < L[i]) return 0;
    }
    return 1;
}

static void sub_l(uint64_t a[4]) {
    u128 borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 t = (u128)a[i] - L[i] - (uint64_t)borrow;
        a[i] = (uint64_t)t;
        borrow = (t >> 64) & 1;
    }
}

/* Montgomery product a*b*R^-1 mod l; requires a*b < l*R; result canonical */
static void mont_mul(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 carry = 0;
        for (int j = 0; j < 4; j++) {
            u128 x = (u128)a[j] * b[i] + t[j] + (uint64_t)carry;
            t[j] = (uint64_t)x;
            carry = x >> 64;
        }
        u128 x = (u128)t[4] + (uint64_t)carry;
        t[4] = (uint64_t)x;
        t[5] = (uint64_t)(x >> 64);
        uint64_t m = t[0] * LFACTOR;
        carry = ((u128)m * L[0] + t[0]) >> 64;
        for (int j = 1; j < 4; j++) {
            u128 y = (u128)m * L[j] + t[j] + (uint64_t)carry;
            t[j - 1] = (uint64_t)y;
            carry = y >> 64;
        }
        x = (u128)t[4] + (uint64_t)carry;
        t[3] = (uint64_t)x;
        t[4] = t[5] + (uint64_t)(x >> 64);
        t[5] = 0;
    }
    uint64_t out[4] = {t[0], t[1], t[2], t[3]};
    if (t[4] || geq_l(out)) sub_l(out);
    memcpy(r, out, 32);
}

static void load256(uint64_t w[4], const uint8_t s[32]) {
    for (int i = 0; i < 4; i++) {
        uint64_t x = 0;
        for (int j = 7; j >= 0; j--) x = (x << 8) | s[8 * i + j];
        w[i] = x;
    }
}

int sc_from_canonical(sc *r, const uint8_t s[32]) {
    load256(r->v, s);
    return !geq_l(r->v);
}

void sc_from_bytes_mod_order(sc *r, const uint8_t s[32]) {
    uint64_t w[4];
    load256(w, s);
    mont_mul(r->v, w, RMOD); /* w * R * R^-1 */
}

void sc_from_wide(sc *r, const uint8_t s[64]) {
    uint64_t lo[4], hi[4], a[4], b[4];
    load256(lo, s);
    load256(hi, s + 32);
    mont_mul(a, lo, RMOD);   /* lo mod l */
    mont_mul(b, hi, RR);     /* hi * R mod l */
    sc x, y;
    memcpy(x.v, a, 32);
    memcpy(y.v, b, 32);
    sc_add(r, &x, &y);
}

void sc_tobytes(uint8_t s[32], const sc *a) {
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 8; j++) s[8 * i + j] = (uint8_t)(a->v[i] >> (8 * j));
}

void sc_add(sc *r, const sc *a, const sc *b) {
    u128 c = 0;
    uint64_t t[4];
    for (int i = 0; i < 4; i++) {
        c += (u128)a->v[i] + b->v[i];
        t[i] = (uint64_t)c;
        c >>= 64;
    }
    if (geq_l(t)) sub_l(t); /* a,b < l < 2^253 so no carry out */
    memcpy(r->v, t, 32);
}

void sc_neg(sc *r, const sc *a) {
    if (sc_iszero(a)) { sc_0(r); return; }
    uint64_t t[4];
    u128 borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 x = (u128)L[i] - a->v[i] - (uint64_t)borrow;
        t[i] = (uint64_t)x;
        borrow = (x >> 64) & 1;
    }
    memcpy(r->v, t, 32);
}

void sc_sub(sc *r, const sc *a, const sc *b) {
    sc nb;
    sc_neg(&nb, b);
    sc_add(r, a, &nb);
}

void sc_mul(sc *r, const sc *a, const sc *b) {
    uint64_t t[4];
    mont_mul(t, a->v, b->v);  /* a b R^-1 */
    mont_mul(r->v, t, RR);    /* a b */
}

void sc_pow_u64(sc *r, const sc *a, uint64_t e) {
    sc acc, base = *a;
    sc_1(&acc);
    while (e) {
        if (e & 1) sc_mul(&acc, &acc, &base);
        sc_mul(&base, &base, &base);
        e >>= 1;
    }
    *r = acc;
}

void sc_invert(sc *r, const sc *a) {
    /* a^(l-2) */
    uint64_t e[4] = {L[0] - 2, L[1], L[2], L[3]};
    sc acc, base = *a;
    sc_1(&acc);
    for (int i = 0; i < 253; i++) {
        if ((e[i / 64] >> (i % 64)) & 1) sc_mul(&acc, &acc, &base);
        sc_mul(&base, &base, &base);
    }
    *r = acc;
}

int sc_iszero(const sc *a) { return (a->v[0] | a->v[1] | a->v[2] | a->v[3]) == 0; }
int sc_eq(const sc *a, const sc *b) { return memcmp(a->v, b->v, 32) == 0; }

/* dalek Scalar::batch_invert: replaces each input by its inverse and returns the inverse of the product
 * of all inputs (relied upon at /root/reference/src/range_proof.rs:899). Inputs must be non-zero. */
void sc_batch_invert(sc *v, size_t n, sc *inv_prod) {
    sc acc, scratch[n ? n : 1];
    sc_1(&acc);
    for (size_t i = 0; i < n; i++) {
        scratch[i] = acc;
        sc_mul(&acc, &acc, &v[i]);
    }
    sc_invert(&acc, &acc);
    if (inv_prod) *inv_prod = acc;
    for (size_t i = n; i-- > 0;) {
        sc tmp;
        sc_mul(&tmp, &acc, &v[i]);
        sc_mul(&v[i], &acc, &scratch[i]);
        acc = tmp;
    }
}
