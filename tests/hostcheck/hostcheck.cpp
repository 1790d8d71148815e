// Test-only: compiles the product's host/device-shared arithmetic header (csrc/arith.cuh) with the HOST
// compiler so its logic can be checked on a CPU-only box.  Not a product path: nothing in the package
// loads this library.
#include <cstring>
#include "../../bulletproofs-plus_b200/csrc/arith.cuh"
using namespace bpp;

static fe ld(const uint8_t *s) { return fe_frombytes(s, nullptr); }
extern "C" {
void hc_fe_mul(const uint8_t *a, const uint8_t *b, uint8_t *o) { fe_tobytes(o, fe_mul(ld(a), ld(b))); }
// lazy add/sub feeding a multiplication, on LOOSE (full 256-bit) inputs: o = ((a + b) * (c - d)) and ((a - b)^2)
static fe ldl(const uint8_t *s) { fe r; memcpy(r.v, s, 32); return r; }
void hc_fe_lazy_mul(const uint8_t *a, const uint8_t *b, const uint8_t *c, const uint8_t *d, uint8_t *o) {
    fe_tobytes(o, fe_mul(fe_add_l(ld(a), ld(b)), fe_sub_ll(ldl(c), ldl(d))));
}
void hc_fe_lazy_sq(const uint8_t *a, const uint8_t *b, uint8_t *o) { fe_tobytes(o, fe_sq(fe_sub_ll(ldl(a), ldl(b)))); }
void hc_fe_lazy_sq_tight(const uint8_t *a, const uint8_t *b, uint8_t *o) { fe_tobytes(o, fe_sq(fe_sub_l(ldl(a), ld(b)))); }
void hc_fe_mul_loose_raw(const uint8_t *a, const uint8_t *b, uint8_t *o) { fe r = fe_mul(ldl(a), ldl(b)); memcpy(o, r.v, 32); }
void hc_fe_sq(const uint8_t *a, uint8_t *o) { fe_tobytes(o, fe_sq(ld(a))); }
void hc_fe_add(const uint8_t *a, const uint8_t *b, uint8_t *o) { fe_tobytes(o, fe_add(ld(a), ld(b))); }
void hc_fe_sub(const uint8_t *a, const uint8_t *b, uint8_t *o) { fe_tobytes(o, fe_sub(ld(a), ld(b))); }
void hc_fe_invert(const uint8_t *a, uint8_t *o) { fe_tobytes(o, fe_invert(ld(a))); }
// raw (non-canonicalised) output limbs, to check the "< 2^255" invariant
void hc_fe_mul_raw(const uint8_t *a, const uint8_t *b, uint8_t *o) { fe r = fe_mul(ld(a), ld(b)); memcpy(o, r.v, 32); }
void hc_fe_sub_raw(const uint8_t *a, const uint8_t *b, uint8_t *o) { fe r = fe_sub(ld(a), ld(b)); memcpy(o, r.v, 32); }
void hc_fe_add_raw(const uint8_t *a, const uint8_t *b, uint8_t *o) { fe r = fe_add(ld(a), ld(b)); memcpy(o, r.v, 32); }
int hc_fe_sqrt_ratio_i(const uint8_t *u, const uint8_t *v, uint8_t *o) { fe r; bool ok = fe_sqrt_ratio_i(r, ld(u), ld(v)); fe_tobytes(o, r); return ok; }
int hc_fe_invsqrt(const uint8_t *v, uint8_t *o) { fe r; bool ok = fe_invsqrt(r, ld(v)); fe_tobytes(o, r); return ok; }
void hc_const(int which, uint8_t *o) {
    fe c;
    switch (which) {
        case 0: c = fe_const_d(); break; case 1: c = fe_const_2d(); break; case 2: c = fe_const_sqrtm1(); break;
        case 3: c = fe_const_invsqrt_a_minus_d(); break; case 4: c = fe_const_sqrt_ad_minus_one(); break;
        case 5: c = fe_const_one_minus_d_sq(); break; default: c = fe_const_d_minus_one_sq(); break;
    }
    fe_tobytes(o, c);
}
void hc_sc_const(int which, uint8_t *o) { sc c = which == 0 ? sc_const_R() : sc_const_RR(); sc_tobytes(o, c); if (which == 2) for (int i = 0; i < 8; i++) { uint32_t l = sc_l(i); memcpy(o + 4 * i, &l, 4); } }
void hc_sc_mul(const uint8_t *a, const uint8_t *b, uint8_t *o) { sc_tobytes(o, sc_mul(sc_frombytes_raw(a), sc_frombytes_raw(b))); }
void hc_sc_add(const uint8_t *a, const uint8_t *b, uint8_t *o) { sc_tobytes(o, sc_add(sc_frombytes_raw(a), sc_frombytes_raw(b))); }
void hc_sc_sub(const uint8_t *a, const uint8_t *b, uint8_t *o) { sc_tobytes(o, sc_sub(sc_frombytes_raw(a), sc_frombytes_raw(b))); }
void hc_sc_reduce256(const uint8_t *a, uint8_t *o) { sc_tobytes(o, sc_reduce256(sc_frombytes_raw(a))); }
void hc_sc_from_wide(const uint8_t *a, uint8_t *o) { uint32_t w[16]; memcpy(w, a, 64); sc_tobytes(o, sc_from_wide_words(w)); }
void hc_sc_invert(const uint8_t *a, uint8_t *o) { sc_tobytes(o, sc_from_mont(scm_invert(sc_to_mont(sc_frombytes_raw(a))))); }
void hc_sc_invert_gcd(const uint8_t *a, uint8_t *o) { sc_tobytes(o, sc_invert_gcd(sc_frombytes_raw(a))); }
void hc_sc_invert_sg(const uint8_t *a, uint8_t *o) { sc_tobytes(o, sc_invert_sg(sc_frombytes_raw(a))); }
void hc_scm_invert_gcd(const uint8_t *a, uint8_t *o) { sc_tobytes(o, sc_from_mont(scm_invert_gcd(sc_to_mont(sc_frombytes_raw(a))))); }
int hc_sc_is_canonical(const uint8_t *a) { uint32_t w[8]; memcpy(w, a, 32); return sc_is_canonical_words(w); }
int hc_decode_encode(const uint8_t *in, uint8_t *o) {
    uint32_t w[8]; memcpy(w, in, 32);
    fe x, y, t;
    bool ok = ristretto_decode(x, y, t, w);
    ge p; p.X = x; p.Y = y; p.Z = fe_one(); p.T = t;
    fe s = ristretto_encode(p); memcpy(o, s.v, 32);
    return ok;
}
void hc_from_uniform(const uint8_t *in, uint8_t *o) { uint32_t w[16]; memcpy(w, in, 64); fe s = ristretto_encode(ristretto_from_uniform_words(w)); memcpy(o, s.v, 32); }
// out = enc( 2*(P+Q) - Q + niels(P) )  exercising add, dbl, neg, madd, msub
int hc_point_ops(const uint8_t *p32, const uint8_t *q32, uint8_t *o_add, uint8_t *o_dbl, uint8_t *o_madd, uint8_t *o_msub) {
    uint32_t w[8]; fe x, y, t; ge P, Q;
    memcpy(w, p32, 32); if (!ristretto_decode(x, y, t, w)) return 0; P.X = x; P.Y = y; P.Z = fe_one(); P.T = t;
    aniels pn = ge_to_aniels_affine(x, y, t);
    memcpy(w, q32, 32); if (!ristretto_decode(x, y, t, w)) return 0; Q.X = x; Q.Y = y; Q.Z = fe_one(); Q.T = t;
    ge S = ge_add(P, Q);
    fe s = ristretto_encode(S); memcpy(o_add, s.v, 32);
    ge D = ge_dbl(S);
    s = ristretto_encode(D); memcpy(o_dbl, s.v, 32);
    s = ristretto_encode(ge_madd(D, pn)); memcpy(o_madd, s.v, 32);
    s = ristretto_encode(ge_msub(D, pn)); memcpy(o_msub, s.v, 32);
    return 1;
}
int hc_is_identity(const uint8_t *p32, const uint8_t *q32) {
    uint32_t w[8]; fe x, y, t; ge P, Q;
    memcpy(w, p32, 32); if (!ristretto_decode(x, y, t, w)) return -1; P.X = x; P.Y = y; P.Z = fe_one(); P.T = t;
    memcpy(w, q32, 32); if (!ristretto_decode(x, y, t, w)) return -1; Q.X = x; Q.Y = y; Q.Z = fe_one(); Q.T = t;
    ge R = ge_add(P, ge_neg(Q));
    return (ge_is_ristretto_identity(R) ? 1 : 0) | (ge_ristretto_eq(P, Q) ? 2 : 0);
}
}
