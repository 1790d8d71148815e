/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not a product path.
 *
 * CPU restatement (plain C11, gcc, unsigned __int128) of the algorithm the reference
 * `tari_bulletproofs_plus` 0.4.1 runs for RangeProof::prove_with_rng / RangeProof::verify_batch,
 * including the parts that live in its un-vendored dependencies:
 *   curve25519-dalek 4.1.3 (field GF(2^255-19), Scalar mod l, Edwards/Ristretto255, Straus/Pippenger MSM),
 *   merlin 3.0.0 (STROBE-128 over Keccak-f[1600]), blake2 0.10.6 (keyed/personalised BLAKE2b-512),
 *   sha3 0.10.8 (SHA3-512, SHAKE256), rand_chacha 0.3.1 (ChaCha12Rng).
 * Those crates are absent from /root/reference (no Cargo.lock, nothing vendored; versions pinned in
 * /root/reference/supply-chain/config.toml:32-34,76-78,132-134), so their PUBLISHED algorithms are
 * restated here (RFC 9496, RFC 8032 §5.1, RFC 7693, FIPS 202, the STROBE/Merlin spec, RFC 8439 core).
 *
 * PARITY STATUS: the reference's own tests pin no bytes, so proof-byte parity is
 * "parity unpinned" against the Rust crate itself.  What IS pinned (tests/test_oracle_*.py):
 *   - group ops byte-exact vs libsodium 1.0.20's crypto_core_ristretto255_* and the RFC 9496 vectors,
 *   - Merlin vs merlin's own KAT, BLAKE2b / SHA3 / SHAKE vs python hashlib, ChaCha vs RFC 7539-style KATs,
 *   - scalar arithmetic vs python big integers,
 *   - protocol algebra: prove -> verify round trips and the reference's error-path matrix.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this.
 */
#ifndef ORC_INTERNAL_H
#define ORC_INTERNAL_H

#include <stddef.h>
#include <stdint.h>
#include <string.h>

typedef unsigned __int128 u128;

/* ---------------- field GF(2^255-19), 5 x 51-bit limbs ---------------- */
typedef struct { uint64_t v[5]; } fe;

void fe_0(fe *h);
void fe_1(fe *h);
void fe_frombytes(fe *h, const uint8_t s[32]);      /* ignores bit 255 */
void fe_tobytes(uint8_t s[32], const fe *h);        /* canonical */
void fe_add(fe *h, const fe *f, const fe *g);
void fe_sub(fe *h, const fe *f, const fe *g);
void fe_neg(fe *h, const fe *f);
void fe_mul(fe *h, const fe *f, const fe *g);
void fe_sq(fe *h, const fe *f);
void fe_invert(fe *out, const fe *z);
void fe_pow22523(fe *out, const fe *z);             /* z^((p-5)/8) */
int  fe_isnegative(const fe *f);
int  fe_iszero(const fe *f);
int  fe_eq(const fe *f, const fe *g);
void fe_cmov(fe *f, const fe *g, int b);            /* f = b ? g : f */
void fe_abs(fe *h, const fe *f);
int  fe_sqrt_ratio_i(fe *r, const fe *u, const fe *v); /* RFC 9496 SQRT_RATIO_M1; returns was_square */

extern fe FE_D_, FE_D2_, FE_SQRTM1_, FE_SQRTADM1_, FE_INVSQRTAMD_, FE_ONEMSQD_, FE_SQDMONE_;

/* ---------------- scalars mod l, 4 x 64-bit limbs, always canonical ---------------- */
typedef struct { uint64_t v[4]; } sc;

void sc_0(sc *r);
void sc_1(sc *r);
void sc_from_u64(sc *r, uint64_t x);
int  sc_from_canonical(sc *r, const uint8_t s[32]);  /* 1 if s < l */
void sc_from_bytes_mod_order(sc *r, const uint8_t s[32]);
void sc_from_wide(sc *r, const uint8_t s[64]);       /* Scalar::from_bytes_mod_order_wide */
void sc_tobytes(uint8_t s[32], const sc *a);
void sc_add(sc *r, const sc *a, const sc *b);
void sc_sub(sc *r, const sc *a, const sc *b);
void sc_neg(sc *r, const sc *a);
void sc_mul(sc *r, const sc *a, const sc *b);
void sc_invert(sc *r, const sc *a);
void sc_pow_u64(sc *r, const sc *a, uint64_t e);
int  sc_iszero(const sc *a);
int  sc_eq(const sc *a, const sc *b);
void sc_batch_invert(sc *v, size_t n, sc *inv_prod); /* dalek Scalar::batch_invert: in place, returns prod^-1 */

/* ---------------- Edwards / Ristretto255 ---------------- */
typedef struct { fe X, Y, Z, T; } ge;                /* extended */
typedef struct { fe YpX, YmX, Z2, T2d; } ge_pniels;  /* projective Niels */
typedef struct { fe ypx, ymx, xy2d; } ge_aniels;     /* affine Niels */

void ge_identity(ge *p);
void ge_add(ge *r, const ge *p, const ge *q);
void ge_sub(ge *r, const ge *p, const ge *q);
void ge_neg(ge *r, const ge *p);
void ge_dbl(ge *r, const ge *p);
void ge_to_pniels(ge_pniels *r, const ge *p);
void ge_to_aniels(ge_aniels *r, const ge *p);        /* one inversion */
void ge_add_pniels(ge *r, const ge *p, const ge_pniels *q);
void ge_sub_pniels(ge *r, const ge *p, const ge_pniels *q);
void ge_add_aniels(ge *r, const ge *p, const ge_aniels *q);
void ge_sub_aniels(ge *r, const ge *p, const ge_aniels *q);
int  ristretto_decode(ge *p, const uint8_t s[32]);   /* 1 ok, 0 reject (CompressedRistretto::decompress) */
void ristretto_encode(uint8_t s[32], const ge *p);   /* RistrettoPoint::compress */
void ristretto_from_uniform(ge *p, const uint8_t b[64]); /* RistrettoPoint::from_uniform_bytes */
int  ristretto_eq(const ge *p, const ge *q);
int  ristretto_is_identity(const ge *p);
void ge_scalarmult(ge *r, const sc *s, const ge *p); /* generic variable-base, vartime */
void ristretto_basepoint(ge *p);

/* multiscalar multiplication, same algorithm classes as dalek 4.1.3 */
void msm_vartime(ge *r, const sc *scalars, const ge *points, size_t n); /* Straus NAF5 (<190) / Pippenger */
void msm_straus(ge *r, const sc *scalars, const ge *points, size_t n);
void msm_pippenger(ge *r, const sc *scalars, const ge *points, size_t n);
typedef struct { size_t n; ge_aniels *tab; } msm_precomp;   /* 64 odd multiples per point (NAF width 8) */
msm_precomp *msm_precomp_new(const ge *points, size_t n);
void msm_precomp_free(msm_precomp *pc);
/* VartimePrecomputedStraus::vartime_mixed_multiscalar_mul; static scalars beyond ns are zero */
void msm_mixed(ge *r, const msm_precomp *pc, const sc *static_scalars, size_t ns,
               const sc *dyn_scalars, const ge *dyn_points, size_t nd);

/* ---------------- hashes ---------------- */
void keccak_f1600(uint64_t st[25]);
typedef struct { uint64_t st[25]; unsigned pos, rate; } keccak_sponge;
void sha3_512(uint8_t out[64], const uint8_t *in, size_t len);
void shake256_init(keccak_sponge *s);
void shake256_absorb(keccak_sponge *s, const uint8_t *in, size_t len);
void shake256_finalize(keccak_sponge *s);
void shake256_squeeze(keccak_sponge *s, uint8_t *out, size_t len);
void blake2b_keyed_personal_512(uint8_t out[64], const uint8_t *key, size_t keylen,
                                const uint8_t *person, size_t personlen,
                                const uint8_t *msg, size_t msglen);

/* ---------------- STROBE-128 / Merlin ---------------- */
typedef struct { uint8_t st[200]; uint8_t pos, pos_begin, cur_flags; } strobe128;
typedef struct { strobe128 s; } merlin_transcript;
typedef struct { strobe128 s; } merlin_rng;

typedef struct orc_rng {
    void (*fill)(struct orc_rng *self, uint8_t *dst, size_t len);
} orc_rng;

void merlin_init(merlin_transcript *t, const uint8_t *label, size_t len);
void merlin_append_message(merlin_transcript *t, const char *label, const uint8_t *msg, size_t len);
void merlin_append_u64(merlin_transcript *t, const char *label, uint64_t x);
void merlin_challenge_bytes(merlin_transcript *t, const char *label, uint8_t *out, size_t len);
/* build_rng().rekey_with_witness_bytes("witness", w)?.finalize(rng) */
void merlin_build_rng(merlin_rng *r, const merlin_transcript *t, const uint8_t *witness, size_t wlen,
                      int have_witness, orc_rng *ext);
void merlin_rng_fill(merlin_rng *r, uint8_t *dst, size_t len);

/* ---------------- RNGs ---------------- */
typedef struct { orc_rng base; uint32_t key[8]; uint64_t counter; uint32_t buf[64]; unsigned index; } chacha12_rng;
void chacha12_seed_from_u64(chacha12_rng *r, uint64_t seed);
void chacha12_from_seed(chacha12_rng *r, const uint8_t seed[32]);
uint32_t chacha12_next_u32(chacha12_rng *r);
uint64_t chacha12_next_u64(chacha12_rng *r);
typedef struct { orc_rng base; } null_rng;
void null_rng_init(null_rng *r);
typedef struct { orc_rng base; const uint8_t *p; size_t len, off; int underflow; } buf_rng;
void buf_rng_init(buf_rng *r, const uint8_t *p, size_t len);

#endif
