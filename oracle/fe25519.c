/* ORACLE (test infrastructure only) — GF(2^255-19) in 5 x 51-bit limbs.
 * Restates curve25519-dalek 4.1.3 `backend::serial::u64::field::FieldElement51` and `field.rs`
 * (un-vendored; the reference reaches it through every point op issued from
 * /root/reference/src/range_proof.rs:339-345,482-521,574-605,859-866,1050-1057). */
#include "orc_internal.h"

#define M51 ((uint64_t)0x7ffffffffffffULL)

void fe_0(fe *h) { memset(h, 0, sizeof *h); }
void fe_1(fe *h) { memset(h, 0, sizeof *h); h->v[0] = 1; }

static uint64_t load64(const uint8_t *p) {
    uint64_t r = 0;
    for (int i = 7; i >= 0; i--) r = (r << 8) | p[i];
    return r;
}

void fe_frombytes(fe *h, const uint8_t s[32]) {
    h->v[0] = load64(s) & M51;
    h->v[1] = (load64(s + 6) >> 3) & M51;
    h->v[2] = (load64(s + 12) >> 6) & M51;
    h->v[3] = (load64(s + 19) >> 1) & M51;
    h->v[4] = (load64(s + 24) >> 12) & M51;
}

static void fe_carry(fe *h) {
    uint64_t c;
    c = h->v[0] >> 51; h->v[0] &= M51; h->v[1] += c;
    c = h->v[1] >> 51; h->v[1] &= M51; h->v[2] += c;
    c = h->v[2] >> 51; h->v[2] &= M51; h->v[3] += c;
    c = h->v[3] >> 51; h->v[3] &= M51; h->v[4] += c;
    c = h->v[4] >> 51; h->v[4] &= M51; h->v[0] += c * 19;
    c = h->v[0] >> 51; h->v[0] &= M51; h->v[1] += c;
}

void fe_tobytes(uint8_t s[32], const fe *f) {
    fe t = *f;
    fe_carry(&t);
    fe_carry(&t);
    /* now limbs < 2^51 (+tiny); compute q = floor((t + 19) / 2^255) */
    uint64_t q = (t.v[0] + 19) >> 51;
    q = (t.v[1] + q) >> 51;
    q = (t.v[2] + q) >> 51;
    q = (t.v[3] + q) >> 51;
    q = (t.v[4] + q) >> 51;
    t.v[0] += 19 * q;
    uint64_t c;
    c = t.v[0] >> 51; t.v[0] &= M51; t.v[1] += c;
    c = t.v[1] >> 51; t.v[1] &= M51; t.v[2] += c;
    c = t.v[2] >> 51; t.v[2] &= M51; t.v[3] += c;
    c = t.v[3] >> 51; t.v[3] &= M51; t.v[4] += c;
    t.v[4] &= M51;
    uint64_t w0 = t.v[0] | (t.v[1] << 51);
    uint64_t w1 = (t.v[1] >> 13) | (t.v[2] << 38);
    uint64_t w2 = (t.v[2] >> 26) | (t.v[3] << 25);
    uint64_t w3 = (t.v[3] >> 39) | (t.v[4] << 12);
    uint64_t w[4] = {w0, w1, w2, w3};
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 8; j++) s[8 * i + j] = (uint8_t)(w[i] >> (8 * j));
}

void fe_add(fe *h, const fe *f, const fe *g) {
    for (int i = 0; i < 5; i++) h->v[i] = f->v[i] + g->v[i];
    fe_carry(h);
}

void fe_sub(fe *h, const fe *f, const fe *g) {
    /* add 16p to stay positive (limbs of g are < 2^52 after carry) */
    h->v[0] = f->v[0] + 36028797018963664ULL - g->v[0];
    h->v[1] = f->v[1] + 36028797018963952ULL - g->v[1];
    h->v[2] = f->v[2] + 36028797018963952ULL - g->v[2];
    h->v[3] = f->v[3] + 36028797018963952ULL - g->v[3];
    h->v[4] = f->v[4] + 36028797018963952ULL - g->v[4];
    fe_carry(h);
}

void fe_neg(fe *h, const fe *f) {
    fe z;
    fe_0(&z);
    fe_sub(h, &z, f);
}

void fe_mul(fe *h, const fe *f, const fe *g) {
    const uint64_t *a = f->v, *b = g->v;
    uint64_t b1_19 = b[1] * 19, b2_19 = b[2] * 19, b3_19 = b[3] * 19, b4_19 = b[4] * 19;
    u128 c0 = (u128)a[0] * b[0] + (u128)a[4] * b1_19 + (u128)a[3] * b2_19 + (u128)a[2] * b3_19 + (u128)a[1] * b4_19;
    u128 c1 = (u128)a[1] * b[0] + (u128)a[0] * b[1] + (u128)a[4] * b2_19 + (u128)a[3] * b3_19 + (u128)a[2] * b4_19;
    u128 c2 = (u128)a[2] * b[0] + (u128)a[1] * b[1] + (u128)a[0] * b[2] + (u128)a[4] * b3_19 + (u128)a[3] * b4_19;
    u128 c3 = (u128)a[3] * b[0] + (u128)a[2] * b[1] + (u128)a[1] * b[2] + (u128)a[0] * b[3] + (u128)a[4] * b4_19;
    u128 c4 = (u128)a[4] * b[0] + (u128)a[3] * b[1] + (u128)a[2] * b[2] + (u128)a[1] * b[3] + (u128)a[0] * b[4];
    c1 += (uint64_t)(c0 >> 51); uint64_t r0 = (uint64_t)c0 & M51;
    c2 += (uint64_t)(c1 >> 51); uint64_t r1 = (uint64_t)c1 & M51;
    c3 += (uint64_t)(c2 >> 51); uint64_t r2 = (uint64_t)c2 & M51;
    c4 += (uint64_t)(c3 >> 51); uint64_t r3 = (uint64_t)c3 & M51;
    uint64_t carry = (uint64_t)(c4 >> 51); uint64_t r4 = (uint64_t)c4 & M51;
    r0 += carry * 19;
    r1 += r0 >> 51; r0 &= M51;
    h->v[0] = r0; h->v[1] = r1; h->v[2] = r2; h->v[3] = r3; h->v[4] = r4;
}

void fe_sq(fe *h, const fe *f) { fe_mul(h, f, f); }

static void fe_sqn(fe *h, const fe *f, int n) {
    fe_sq(h, f);
    for (int i = 1; i < n; i++) fe_sq(h, h);
}

/* z^(2^250-1) and z^11, the shared prefix of the standard addition chains */
static void fe_pow22501(fe *t19, fe *t3, const fe *z) {
    fe t0, t1, t2, t4, t5, t6, t7, t9, t11, t13, t15, t17, tmp;
    fe_sq(&t0, z);                 /* 2 */
    fe_sqn(&tmp, &t0, 2);          /* 8 */
    fe_mul(&t1, z, &tmp);          /* 9 */
    fe_mul(&t2, &t0, &t1);         /* 11 */
    fe_sq(&tmp, &t2);              /* 22 */
    fe_mul(&t4, &t1, &tmp);        /* 31 = 2^5-1 */
    fe_sqn(&tmp, &t4, 5);
    fe_mul(&t5, &tmp, &t4);        /* 2^10-1 */
    fe_sqn(&tmp, &t5, 10);
    fe_mul(&t6, &tmp, &t5);        /* 2^20-1 */
    fe_sqn(&tmp, &t6, 20);
    fe_mul(&t7, &tmp, &t6);        /* 2^40-1 */
    fe_sqn(&tmp, &t7, 10);
    fe_mul(&t9, &tmp, &t5);        /* 2^50-1 */
    fe_sqn(&tmp, &t9, 50);
    fe_mul(&t11, &tmp, &t9);       /* 2^100-1 */
    fe_sqn(&tmp, &t11, 100);
    fe_mul(&t13, &tmp, &t11);      /* 2^200-1 */
    fe_sqn(&tmp, &t13, 50);
    fe_mul(&t15, &tmp, &t9);       /* 2^250-1 */
    (void)t17;
    *t19 = t15;
    *t3 = t2;
}

void fe_invert(fe *out, const fe *z) {
    fe t19, t3, t;
    fe_pow22501(&t19, &t3, z);
    fe_sqn(&t, &t19, 5);           /* 2^255 - 2^5 */
    fe_mul(out, &t, &t3);          /* 2^255 - 21 */
}

void fe_pow22523(fe *out, const fe *z) {
    fe t19, t3, t;
    fe_pow22501(&t19, &t3, z);
    fe_sqn(&t, &t19, 2);           /* 2^252 - 4 */
    fe_mul(out, &t, z);            /* 2^252 - 3 */
}

int fe_isnegative(const fe *f) {
    uint8_t s[32];
    fe_tobytes(s, f);
    return s[0] & 1;
}

int fe_iszero(const fe *f) {
    uint8_t s[32];
    fe_tobytes(s, f);
    uint8_t r = 0;
    for (int i = 0; i < 32; i++) r |= s[i];
    return r == 0;
}

int fe_eq(const fe *f, const fe *g) {
    uint8_t a[32], b[32];
    fe_tobytes(a, f);
    fe_tobytes(b, g);
    return memcmp(a, b, 32) == 0;
}

void fe_cmov(fe *f, const fe *g, int b) {
    if (b) *f = *g;
}

void fe_abs(fe *h, const fe *f) {
    if (fe_isnegative(f)) fe_neg(h, f); else *h = *f;
}

/* constants, little-endian byte strings of the integers in SURVEY.md Appendix A.1 (RFC 9496 §4.1) */
static fe fe_const(const char *hex) {
    /* hex is big-endian, 64 chars */
    uint8_t b[32];
    for (int i = 0; i < 32; i++) {
        unsigned hi = (unsigned char)hex[2 * i], lo = (unsigned char)hex[2 * i + 1];
        hi = hi <= '9' ? hi - '0' : hi - 'a' + 10;
        lo = lo <= '9' ? lo - '0' : lo - 'a' + 10;
        b[31 - i] = (uint8_t)(hi * 16 + lo);
    }
    fe r;
    fe_frombytes(&r, b);
    return r;
}

fe FE_D_, FE_D2_, FE_SQRTM1_, FE_SQRTADM1_, FE_INVSQRTAMD_, FE_ONEMSQD_, FE_SQDMONE_;
static int fe_consts_ready = 0;
static void fe_consts_init(void) {
    if (fe_consts_ready) return;
    FE_D_ = fe_const("52036cee2b6ffe738cc740797779e89800700a4d4141d8ab75eb4dca135978a3");
    fe_add(&FE_D2_, &FE_D_, &FE_D_);
    FE_SQRTM1_ = fe_const("2b8324804fc1df0b2b4d00993dfbd7a72f431806ad2fe478c4ee1b274a0ea0b0");
    FE_SQRTADM1_ = fe_const("376931bf2b8348ac0f3cfcc931f5d1fdaf9d8e0c1b7854bd7e97f6a0497b2e1b");
    FE_INVSQRTAMD_ = fe_const("786c8905cfaffca216c27b91fe01d8409d2f16175a4172be99c8fdaa805d40ea");
    FE_ONEMSQD_ = fe_const("029072a8b2b3e0d79994abddbe70dfe42c81a138cd5e350fe27c09c1945fc176");
    FE_SQDMONE_ = fe_const("5968b37af66c22414cdcd32f529b4eebd29e4a2cb01e199931ad5aaa44ed4d20");
    fe_consts_ready = 1;
}
__attribute__((constructor)) static void fe_ctor(void) { fe_consts_init(); }

/* RFC 9496 §4.2 SQRT_RATIO_M1 == dalek FieldElement::sqrt_ratio_i */
int fe_sqrt_ratio_i(fe *out, const fe *u, const fe *v) {
    fe_consts_init();
    fe v3, v7, r, check, t, neg_u, neg_u_i;
    fe_sq(&t, v);
    fe_mul(&v3, &t, v);            /* v^3 */
    fe_sq(&t, &v3);
    fe_mul(&v7, &t, v);            /* v^7 */
    fe_mul(&t, u, &v7);
    fe_pow22523(&t, &t);           /* (u v^7)^((p-5)/8) */
    fe_mul(&r, u, &v3);
    fe_mul(&r, &r, &t);
    fe_sq(&t, &r);
    fe_mul(&check, v, &t);
    fe_neg(&neg_u, u);
    fe_mul(&neg_u_i, &neg_u, &FE_SQRTM1_);
    int correct = fe_eq(&check, u);
    int flipped = fe_eq(&check, &neg_u);
    int flipped_i = fe_eq(&check, &neg_u_i);
    if (flipped | flipped_i) {
        fe_mul(&r, &r, &FE_SQRTM1_);
    }
    fe_abs(out, &r);
    return correct | flipped;
}
