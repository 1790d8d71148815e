/* ORACLE (test infrastructure only) — twisted Edwards curve -x^2+y^2 = 1+d x^2 y^2 in extended
 * coordinates and the Ristretto255 group (RFC 9496), restating curve25519-dalek 4.1.3
 * `edwards.rs` / `ristretto.rs` / `backend::serial::curve_models` as the reference uses them through
 * /root/reference/src/ristretto.rs:28-64 (compress, decompress, from_uniform_bytes) and
 * /root/reference/src/range_proof.rs (every +, *, ==, MSM). */
#include "orc_internal.h"

void ge_identity(ge *p) {
    fe_0(&p->X); fe_1(&p->Y); fe_1(&p->Z); fe_0(&p->T);
}

void ge_neg(ge *r, const ge *p) {
    fe_neg(&r->X, &p->X); r->Y = p->Y; r->Z = p->Z; fe_neg(&r->T, &p->T);
}

void ge_to_pniels(ge_pniels *r, const ge *p) {
    fe_add(&r->YpX, &p->Y, &p->X);
    fe_sub(&r->YmX, &p->Y, &p->X);
    fe_add(&r->Z2, &p->Z, &p->Z);
    fe_mul(&r->T2d, &p->T, &FE_D2_);
}

void ge_to_aniels(ge_aniels *r, const ge *p) {
    fe zi, x, y, xy;
    fe_invert(&zi, &p->Z);
    fe_mul(&x, &p->X, &zi);
    fe_mul(&y, &p->Y, &zi);
    fe_add(&r->ypx, &y, &x);
    fe_sub(&r->ymx, &y, &x);
    fe_mul(&xy, &x, &y);
    fe_mul(&r->xy2d, &xy, &FE_D2_);
}

/* add-2008-hwcd-3 with precomputed (Y+X, Y-X, 2Z, 2dT) */
static void add_core(ge *r, const ge *p, const fe *ypx, const fe *ymx, const fe *z2, const fe *t2d) {
    fe A, B, C, D, E, F, G, H, t;
    fe_sub(&t, &p->Y, &p->X); fe_mul(&A, &t, ymx);
    fe_add(&t, &p->Y, &p->X); fe_mul(&B, &t, ypx);
    fe_mul(&C, &p->T, t2d);
    if (z2) fe_mul(&D, &p->Z, z2); else fe_add(&D, &p->Z, &p->Z);
    fe_sub(&E, &B, &A);
    fe_sub(&F, &D, &C);
    fe_add(&G, &D, &C);
    fe_add(&H, &B, &A);
    fe_mul(&r->X, &E, &F);
    fe_mul(&r->Y, &G, &H);
    fe_mul(&r->Z, &F, &G);
    fe_mul(&r->T, &E, &H);
}

void ge_add_pniels(ge *r, const ge *p, const ge_pniels *q) { add_core(r, p, &q->YpX, &q->YmX, &q->Z2, &q->T2d); }
void ge_sub_pniels(ge *r, const ge *p, const ge_pniels *q) {
    fe nt; fe_neg(&nt, &q->T2d);
    add_core(r, p, &q->YmX, &q->YpX, &q->Z2, &nt);
}
void ge_add_aniels(ge *r, const ge *p, const ge_aniels *q) { add_core(r, p, &q->ypx, &q->ymx, NULL, &q->xy2d); }
void ge_sub_aniels(ge *r, const ge *p, const ge_aniels *q) {
    fe nt; fe_neg(&nt, &q->xy2d);
    add_core(r, p, &q->ymx, &q->ypx, NULL, &nt);
}

void ge_add(ge *r, const ge *p, const ge *q) {
    ge_pniels c;
    ge_to_pniels(&c, q);
    ge_add_pniels(r, p, &c);
}

void ge_sub(ge *r, const ge *p, const ge *q) {
    ge_pniels c;
    ge_to_pniels(&c, q);
    ge_sub_pniels(r, p, &c);
}

/* dbl-2008-hwcd, a = -1 */
void ge_dbl(ge *r, const ge *p) {
    fe XX, YY, ZZ2, XpY, XpY2, YYpXX, YYmXX, cX, cT;
    fe_sq(&XX, &p->X);
    fe_sq(&YY, &p->Y);
    fe_sq(&ZZ2, &p->Z); fe_add(&ZZ2, &ZZ2, &ZZ2);
    fe_add(&XpY, &p->X, &p->Y);
    fe_sq(&XpY2, &XpY);
    fe_add(&YYpXX, &YY, &XX);
    fe_sub(&YYmXX, &YY, &XX);
    fe_sub(&cX, &XpY2, &YYpXX);      /* 2XY */
    fe_sub(&cT, &ZZ2, &YYmXX);
    /* completed (cX : YYpXX : YYmXX : cT) -> extended */
    fe_mul(&r->X, &cX, &cT);
    fe_mul(&r->Y, &YYpXX, &YYmXX);
    fe_mul(&r->Z, &YYmXX, &cT);
    fe_mul(&r->T, &cX, &YYpXX);
}

/* RFC 9496 §4.3.1 Decode == CompressedRistretto::decompress */
int ristretto_decode(ge *p, const uint8_t s_bytes[32]) {
    fe s, ss, u1, u2, u2s, v, I, Dx, Dy, t, one;
    uint8_t chk[32];
    fe_frombytes(&s, s_bytes);
    fe_tobytes(chk, &s);
    if (memcmp(chk, s_bytes, 32) != 0) return 0;   /* non-canonical (includes bit 255 set) */
    if (s_bytes[0] & 1) return 0;                    /* negative */
    fe_1(&one);
    fe_sq(&ss, &s);
    fe_sub(&u1, &one, &ss);
    fe_add(&u2, &one, &ss);
    fe_sq(&u2s, &u2);
    fe_sq(&t, &u1);
    fe_mul(&t, &t, &FE_D_);
    fe_neg(&t, &t);
    fe_sub(&v, &t, &u2s);                            /* -(D u1^2) - u2^2 */
    fe_mul(&t, &v, &u2s);
    int ok = fe_sqrt_ratio_i(&I, &one, &t);
    fe_mul(&Dx, &I, &u2);
    fe_mul(&Dy, &I, &Dx);
    fe_mul(&Dy, &Dy, &v);
    fe x, y, tt;
    fe_add(&t, &s, &s);
    fe_mul(&x, &t, &Dx);
    fe_abs(&x, &x);
    fe_mul(&y, &u1, &Dy);
    fe_mul(&tt, &x, &y);
    if (!ok || fe_isnegative(&tt) || fe_iszero(&y)) return 0;
    p->X = x; p->Y = y; fe_1(&p->Z); p->T = tt;
    return 1;
}

/* RFC 9496 §4.3.2 Encode == RistrettoPoint::compress */
void ristretto_encode(uint8_t out[32], const ge *p) {
    fe u1, u2, t, I, D1, D2, zinv, ix, iy, eden, x, y, s, one, a, b;
    fe_1(&one);
    fe_add(&a, &p->Z, &p->Y);
    fe_sub(&b, &p->Z, &p->Y);
    fe_mul(&u1, &a, &b);
    fe_mul(&u2, &p->X, &p->Y);
    fe_sq(&t, &u2);
    fe_mul(&t, &t, &u1);
    fe_sqrt_ratio_i(&I, &one, &t);
    fe_mul(&D1, &I, &u1);
    fe_mul(&D2, &I, &u2);
    fe_mul(&zinv, &D1, &D2);
    fe_mul(&zinv, &zinv, &p->T);
    fe_mul(&ix, &p->X, &FE_SQRTM1_);
    fe_mul(&iy, &p->Y, &FE_SQRTM1_);
    fe_mul(&eden, &D1, &FE_INVSQRTAMD_);
    fe_mul(&t, &p->T, &zinv);
    int rotate = fe_isnegative(&t);
    x = p->X; y = p->Y;
    fe den = D2;
    if (rotate) { x = iy; y = ix; den = eden; }
    fe_mul(&t, &x, &zinv);
    if (fe_isnegative(&t)) fe_neg(&y, &y);
    fe_sub(&t, &p->Z, &y);
    fe_mul(&s, &den, &t);
    fe_abs(&s, &s);
    fe_tobytes(out, &s);
}

/* RFC 9496 §4.3.4 MAP == RistrettoPoint::elligator_ristretto_flavor */
static void elligator(ge *p, const fe *t0) {
    fe r, u, v, s, sp, c, N, w0, w1, w2, w3, one, t, rpd, minus_one;
    fe_1(&one);
    fe_neg(&minus_one, &one);
    fe_sq(&t, t0);
    fe_mul(&r, &FE_SQRTM1_, &t);
    fe_add(&t, &r, &one);
    fe_mul(&u, &t, &FE_ONEMSQD_);
    fe_mul(&t, &r, &FE_D_);
    fe_sub(&t, &minus_one, &t);           /* -1 - r D */
    fe_add(&rpd, &r, &FE_D_);
    fe_mul(&v, &t, &rpd);
    int was_square = fe_sqrt_ratio_i(&s, &u, &v);
    fe_mul(&sp, &s, t0);
    fe_abs(&sp, &sp);
    fe_neg(&sp, &sp);                     /* -|s t| */
    c = minus_one;
    if (!was_square) { s = sp; c = r; }
    fe_sub(&t, &r, &one);
    fe_mul(&N, &c, &t);
    fe_mul(&N, &N, &FE_SQDMONE_);
    fe_sub(&N, &N, &v);
    fe_add(&t, &s, &s);
    fe_mul(&w0, &t, &v);
    fe_mul(&w1, &N, &FE_SQRTADM1_);
    fe_sq(&t, &s);
    fe_sub(&w2, &one, &t);
    fe_add(&w3, &one, &t);
    fe_mul(&p->X, &w0, &w3);
    fe_mul(&p->Y, &w2, &w1);
    fe_mul(&p->Z, &w1, &w3);
    fe_mul(&p->T, &w0, &w2);
}

void ristretto_from_uniform(ge *p, const uint8_t b[64]) {
    fe r0, r1;
    ge p0, p1;
    fe_frombytes(&r0, b);        /* masks bit 255 */
    fe_frombytes(&r1, b + 32);
    elligator(&p0, &r0);
    elligator(&p1, &r1);
    ge_add(p, &p0, &p1);
}

int ristretto_eq(const ge *p, const ge *q) {
    fe a, b, c, d;
    fe_mul(&a, &p->X, &q->Y);
    fe_mul(&b, &p->Y, &q->X);
    fe_mul(&c, &p->X, &q->X);
    fe_mul(&d, &p->Y, &q->Y);
    return fe_eq(&a, &b) | fe_eq(&c, &d);
}

int ristretto_is_identity(const ge *p) {
    ge id;
    ge_identity(&id);
    return ristretto_eq(p, &id);
}

void ristretto_basepoint(ge *p) {
    static const uint8_t B[32] = {0xe2, 0xf2, 0xae, 0x0a, 0x6a, 0xbc, 0x4e, 0x71, 0xa8, 0x84, 0xa9, 0x61, 0xc5, 0x00, 0x51, 0x5f,
                                  0x58, 0xe3, 0x0b, 0x6a, 0xa5, 0x82, 0xdd, 0x8d, 0xb6, 0xa6, 0x59, 0x45, 0xe0, 0x8d, 0x2d, 0x76};
    ristretto_decode(p, B);
}

void ge_scalarmult(ge *r, const sc *s, const ge *p) {
    ge acc;
    ge_identity(&acc);
    for (int i = 252; i >= 0; i--) {
        ge_dbl(&acc, &acc);
        if ((s->v[i / 64] >> (i % 64)) & 1) ge_add(&acc, &acc, p);
    }
    *r = acc;
}
