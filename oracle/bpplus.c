/* ORACLE (test infrastructure only) — the Bulletproofs+ range-proof protocol exactly as
 * tari_bulletproofs_plus 0.4.1 runs it.  Every function cites the reference lines it follows.
 * The ORDER of transcript operations, RNG draws and error checks is normative. */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include "bpp_oracle.h"
#include "orc_internal.h"

struct orc_params {
    int n, M, ext;
    ge H, G[6];
    uint8_t Hc[32], Gc[6][32];
    ge *Gi, *Hi;          /* n*M each, party-major (generators/aggregated_gens_iter.rs:18-43) */
    msm_precomp *pc;      /* interleaved G,H (generators/bulletproof_gens.rs:100-103) */
};

static int is_pow2(size_t x) { return x && !(x & (x - 1)); }

/* ---- generators: src/generators/generators_chain.rs:23-49, bulletproof_gens.rs:83-112, ristretto.rs:67-112 ---- */
static void gens_chain(ge *out, int count, uint8_t tag, uint32_t party) {
    keccak_sponge s;
    uint8_t label[5] = {tag, (uint8_t)party, (uint8_t)(party >> 8), (uint8_t)(party >> 16), (uint8_t)(party >> 24)};
    shake256_init(&s);
    shake256_absorb(&s, (const uint8_t *)"GeneratorsChain", 15);
    shake256_absorb(&s, label, 5);
    shake256_finalize(&s);
    for (int i = 0; i < count; i++) {
        uint8_t u[64];
        shake256_squeeze(&s, u, 64);
        ristretto_from_uniform(&out[i], u);
    }
}

int orc_params_new(int bit_length, int max_aggregation, int extension_degree, orc_params **out) {
    *out = NULL;
    /* range_parameters.rs:32-58 */
    if (max_aggregation <= 0 || !is_pow2((size_t)max_aggregation)) return ORC_INVALID_ARGUMENT;
    if (bit_length <= 0 || !is_pow2((size_t)bit_length)) return ORC_INVALID_ARGUMENT;
    if (bit_length > 64) return ORC_INVALID_ARGUMENT;
    if (extension_degree < 1 || extension_degree > 6) return ORC_INVALID_ARGUMENT; /* pedersen_gens.rs:68-82 */
    orc_params *p = calloc(1, sizeof *p);
    p->n = bit_length; p->M = max_aggregation; p->ext = extension_degree;
    ristretto_basepoint(&p->H);                       /* ristretto.rs:70-71 */
    ristretto_encode(p->Hc, &p->H);
    for (int i = 0; i < extension_degree; i++) {      /* ristretto.rs:88-112 */
        char label[40];
        int len = snprintf(label, sizeof label, "RISTRETTO_MASKING_BASEPOINT_%d", i + 1);
        uint8_t h[64];
        sha3_512(h, (const uint8_t *)label, (size_t)len);
        ristretto_from_uniform(&p->G[i], h);
        ristretto_encode(p->Gc[i], &p->G[i]);
    }
    size_t total = (size_t)bit_length * (size_t)max_aggregation;
    p->Gi = malloc(total * sizeof(ge));
    p->Hi = malloc(total * sizeof(ge));
    for (int j = 0; j < max_aggregation; j++) {
        gens_chain(p->Gi + (size_t)j * bit_length, bit_length, 'G', (uint32_t)j);
        gens_chain(p->Hi + (size_t)j * bit_length, bit_length, 'H', (uint32_t)j);
    }
    ge *inter = malloc(2 * total * sizeof(ge));
    for (size_t i = 0; i < total; i++) { inter[2 * i] = p->Gi[i]; inter[2 * i + 1] = p->Hi[i]; }
    p->pc = msm_precomp_new(inter, 2 * total);
    free(inter);
    *out = p;
    return ORC_OK;
}

void orc_params_free(orc_params *p) {
    if (!p) return;
    free(p->Gi); free(p->Hi);
    msm_precomp_free(p->pc);
    free(p);
}

int orc_params_point(const orc_params *p, int which, size_t index, uint8_t out32[32]) {
    size_t total = (size_t)p->n * (size_t)p->M;
    switch (which) {
        case 0: memcpy(out32, p->Hc, 32); return ORC_OK;
        case 1: if (index >= (size_t)p->ext) return ORC_INVALID_ARGUMENT; memcpy(out32, p->Gc[index], 32); return ORC_OK;
        case 2: if (index >= total) return ORC_INVALID_ARGUMENT; ristretto_encode(out32, &p->Gi[index]); return ORC_OK;
        case 3: if (index >= total) return ORC_INVALID_ARGUMENT; ristretto_encode(out32, &p->Hi[index]); return ORC_OK;
    }
    return ORC_INVALID_ARGUMENT;
}
int orc_params_bit_length(const orc_params *p) { return p->n; }
int orc_params_max_aggregation(const orc_params *p) { return p->M; }
int orc_params_extension_degree(const orc_params *p) { return p->ext; }

/* ---- PedersenGens::commit: src/generators/pedersen_gens.rs:112-122 ---- */
static int commit_point(ge *out, const orc_params *p, uint64_t value, const sc *blind, int nb) {
    if (nb <= 0 || nb > p->ext) return ORC_INVALID_LENGTH;
    sc s[7]; ge pts[7];
    sc_from_u64(&s[0], value); pts[0] = p->H;
    for (int i = 0; i < nb; i++) { s[i + 1] = blind[i]; pts[i + 1] = p->G[i]; }
    msm_straus(out, s, pts, (size_t)nb + 1);
    return ORC_OK;
}

int orc_commit(const orc_params *p, uint64_t value, const uint8_t *blindings32, int nb, uint8_t out32[32]) {
    sc b[6];
    if (nb <= 0 || nb > p->ext || nb > 6) return ORC_INVALID_LENGTH;
    for (int i = 0; i < nb; i++) sc_from_bytes_mod_order(&b[i], blindings32 + 32 * i);
    ge c;
    int rc = commit_point(&c, p, value, b, nb);
    if (rc) return rc;
    ristretto_encode(out32, &c);
    return ORC_OK;
}

/* ---- RangeStatement::init: src/range_statement.rs:36-73 ---- */
int orc_statement_check(const orc_statement *st) {
    if (st->m <= 0 || !is_pow2((size_t)st->m)) return ORC_INVALID_ARGUMENT;
    if (st->n_min != st->m) return ORC_INVALID_ARGUMENT;
    if (st->params->M < st->m) return ORC_INVALID_ARGUMENT;
    if (st->seed_nonce && st->m > 1) return ORC_INVALID_ARGUMENT;
    return ORC_OK;
}

/* ---- transcripts exposed to python ---- */
void orc_transcript_new(const uint8_t *label, size_t len, uint8_t out[ORC_TRANSCRIPT_BYTES]) {
    merlin_transcript t;
    merlin_init(&t, label, len);
    memcpy(out, &t.s, ORC_TRANSCRIPT_BYTES);
}
void orc_transcript_append_message(uint8_t t[ORC_TRANSCRIPT_BYTES], const char *label, const uint8_t *msg, size_t len) {
    merlin_transcript tr;
    memcpy(&tr.s, t, ORC_TRANSCRIPT_BYTES);
    merlin_append_message(&tr, label, msg, len);
    memcpy(t, &tr.s, ORC_TRANSCRIPT_BYTES);
}
void orc_transcript_challenge_bytes(uint8_t t[ORC_TRANSCRIPT_BYTES], const char *label, uint8_t *out, size_t len) {
    merlin_transcript tr;
    memcpy(&tr.s, t, ORC_TRANSCRIPT_BYTES);
    merlin_challenge_bytes(&tr, label, out, len);
    memcpy(t, &tr.s, ORC_TRANSCRIPT_BYTES);
}

/* ---- RNG handles ---- */
orc_rng *orc_rng_chacha12_seed_from_u64(uint64_t seed) {
    chacha12_rng *r = malloc(sizeof *r);
    chacha12_seed_from_u64(r, seed);
    return &r->base;
}
orc_rng *orc_rng_chacha12_from_seed(const uint8_t seed[32]) {
    chacha12_rng *r = malloc(sizeof *r);
    chacha12_from_seed(r, seed);
    return &r->base;
}
orc_rng *orc_rng_null(void) { null_rng *r = malloc(sizeof *r); null_rng_init(r); return &r->base; }
orc_rng *orc_rng_buffer(const uint8_t *bytes, size_t len) { buf_rng *r = malloc(sizeof *r); buf_rng_init(r, bytes, len); return &r->base; }
void orc_rng_fill(orc_rng *r, uint8_t *dst, size_t len) { r->fill(r, dst, len); }
uint64_t orc_rng_next_u64(orc_rng *r) { return chacha12_next_u64((chacha12_rng *)r); }
void orc_rng_free(orc_rng *r) { free(r); }

/* Scalar::random (64 rng bytes -> wide reduce) looped until non-zero: protocols/scalar_protocol.rs:23-30 */
static void random_not_zero(sc *out, orc_rng *r) {
    uint8_t b[64];
    do { r->fill(r, b, 64); sc_from_wide(out, b); } while (sc_iszero(out));
}
void orc_random_not_zero(orc_rng *r, uint8_t out32[32]) { sc s; random_not_zero(&s, r); sc_tobytes(out32, &s); }

typedef struct { orc_rng base; merlin_rng *m; } trng_adapter;
static void trng_fill(orc_rng *self, uint8_t *dst, size_t len) { merlin_rng_fill(((trng_adapter *)self)->m, dst, len); }

/* ---- nonce: src/utils/generic.rs:30-60 ---- */
static int nonce(sc *out, const sc *seed, const char *label, int have_j, uint32_t j, int have_k, uint32_t k) {
    size_t ll = strlen(label);
    if (ll > 16) return ORC_INVALID_LENGTH;
    uint8_t key[43];
    size_t kl = 0;
    key[kl++] = 0;
    sc_tobytes(key + kl, seed); kl += 32;
    if (have_j) { key[kl++] = 'j'; for (int i = 0; i < 4; i++) key[kl++] = (uint8_t)(j >> (8 * i)); }
    if (have_k) { key[kl++] = 'k'; for (int i = 0; i < 4; i++) key[kl++] = (uint8_t)(k >> (8 * i)); }
    uint8_t h[64];
    blake2b_keyed_personal_512(h, key, kl, (const uint8_t *)label, ll, NULL, 0);
    sc_from_wide(out, h);
    return ORC_OK;
}

void orc_blake2b_nonce(const uint8_t seed32[32], const char *label, int have_j, uint32_t j, int have_k, uint32_t k,
                       uint8_t out32[32]) {
    sc seed, o;
    sc_from_bytes_mod_order(&seed, seed32);
    nonce(&o, &seed, label, have_j, j, have_k, k);
    sc_tobytes(out32, &o);
}

/* ---- RangeProofTranscript: src/transcripts.rs:59-194 ---- */
typedef struct {
    merlin_transcript *t;
    uint8_t *wit; size_t wlen; int have_wit;
    merlin_rng rng;
    orc_rng *ext;
    trng_adapter ad;
} rpt;

static int is_zero32(const uint8_t *b) { uint8_t r = 0; for (int i = 0; i < 32; i++) r |= b[i]; return r == 0; }

/* protocols/transcript_protocol.rs:49-61 */
static int validate_and_append_point(merlin_transcript *t, const char *label, const uint8_t pt[32]) {
    if (is_zero32(pt)) return ORC_VERIFICATION_FAILED;
    merlin_append_message(t, label, pt, 32);
    return ORC_OK;
}

/* protocols/transcript_protocol.rs:67-78 */
static int challenge_scalar(merlin_transcript *t, const char *label, sc *out) {
    uint8_t buf[64];
    merlin_challenge_bytes(t, label, buf, 64);
    sc_from_wide(out, buf);
    return sc_iszero(out) ? ORC_VERIFICATION_FAILED : ORC_OK;
}

static void rpt_build_rng(rpt *r) { /* transcripts.rs:185-194 */
    merlin_build_rng(&r->rng, r->t, r->wit, r->wlen, r->have_wit, r->ext);
    r->ad.base.fill = trng_fill;
    r->ad.m = &r->rng;
}

static int rpt_new(rpt *r, merlin_transcript *t, const orc_params *p, int m, const orc_statement *st,
                   const orc_witness *w, orc_rng *ext) { /* transcripts.rs:59-121 */
    int rc;
    memset(r, 0, sizeof *r);
    r->t = t; r->ext = ext;
    merlin_append_message(t, "dom-sep", (const uint8_t *)"Bulletproofs+ Range Proof", 25);
    if ((rc = validate_and_append_point(t, "H", p->Hc))) return rc;
    for (int i = 0; i < p->ext; i++)
        if ((rc = validate_and_append_point(t, "G", p->Gc[i]))) return rc;
    merlin_append_u64(t, "N", (uint64_t)p->n);
    merlin_append_u64(t, "T", (uint64_t)p->ext);
    merlin_append_u64(t, "M", (uint64_t)m);
    for (int i = 0; i < st->m; i++) merlin_append_message(t, "Ci", st->commitments + 32 * i, 32);
    for (int i = 0; i < st->n_min; i++)
        merlin_append_u64(t, "vi - minimum_value", st->min_present[i] ? st->min_values[i] : 0);
    if (w) {
        r->wlen = (size_t)w->n_openings * (8 + (size_t)w->r_len * 32);
        r->wit = malloc(r->wlen ? r->wlen : 1);
        size_t o = 0;
        for (int i = 0; i < w->n_openings; i++) {
            for (int b = 0; b < 8; b++) r->wit[o++] = (uint8_t)(w->values[i] >> (8 * b));
            memcpy(r->wit + o, w->blindings + (size_t)i * w->r_len * 32, (size_t)w->r_len * 32);
            o += (size_t)w->r_len * 32;
        }
        r->have_wit = 1;
    }
    rpt_build_rng(r);
    return ORC_OK;
}

static void rpt_free(rpt *r) {
    if (r->wit) { memset(r->wit, 0, r->wlen); free(r->wit); r->wit = NULL; }
}

static int rpt_challenges_y_z(rpt *r, const uint8_t a[32], sc *y, sc *z) { /* transcripts.rs:124-136 */
    int rc;
    if ((rc = validate_and_append_point(r->t, "A", a))) return rc;
    rpt_build_rng(r);
    if ((rc = challenge_scalar(r->t, "y", y))) return rc;
    return challenge_scalar(r->t, "z", z);
}

static int rpt_challenge_round_e(rpt *r, const uint8_t l[32], const uint8_t rr[32], sc *e) { /* :139-149 */
    int rc;
    if ((rc = validate_and_append_point(r->t, "L", l))) return rc;
    if ((rc = validate_and_append_point(r->t, "R", rr))) return rc;
    rpt_build_rng(r);
    return challenge_scalar(r->t, "e", e);
}

static int rpt_challenge_final_e(rpt *r, const uint8_t a1[32], const uint8_t b[32], sc *e) { /* :152-162 */
    int rc;
    if ((rc = validate_and_append_point(r->t, "A1", a1))) return rc;
    if ((rc = validate_and_append_point(r->t, "B", b))) return rc;
    rpt_build_rng(r);
    return challenge_scalar(r->t, "e", e);
}

static void rpt_to_verifier_rng(rpt *r, const uint8_t r1[32], const uint8_t s1[32], const uint8_t (*d1)[32], int nd1) {
    /* transcripts.rs:166-179 */
    merlin_append_message(r->t, "r1", r1, 32);
    merlin_append_message(r->t, "s1", s1, 32);
    for (int i = 0; i < nd1; i++) merlin_append_message(r->t, "d1", d1[i], 32);
    rpt_build_rng(r);
}

/* ---- RangeProof::prove_with_rng: src/range_proof.rs:232-608 ---- */
int orc_prove(uint8_t transcript[ORC_TRANSCRIPT_BYTES], const orc_statement *st, const orc_witness *w,
              orc_rng *rng, orc_proof *out) {
    const orc_params *p = st->params;
    int rc = ORC_OK;
    size_t n = (size_t)p->n, m = (size_t)st->m, ext = (size_t)p->ext;
    size_t N = n * m;                                                  /* :239-244 */
    if ((size_t)w->n_openings != m) return ORC_INVALID_LENGTH;         /* :248-252 */
    if (w->r_len != p->ext) return ORC_INVALID_LENGTH;                 /* :256-260 */
    for (size_t j = 0; j < m; j++)                                     /* :264-271 */
        if (n < 64 && (w->values[j] >> n) > 0) return ORC_INVALID_LENGTH;

    sc *blind = malloc(sizeof(sc) * m * ext);
    for (size_t i = 0; i < m * ext; i++) sc_from_bytes_mod_order(&blind[i], w->blindings + 32 * i);
    ge *V = malloc(sizeof(ge) * m);
    for (size_t j = 0; j < m; j++) {                                   /* :275-284 */
        ge c;
        if (!ristretto_decode(&V[j], st->commitments + 32 * j)) { rc = ORC_INVALID_ARGUMENT; goto out0; }
        if ((rc = commit_point(&c, p, w->values[j], blind + j * ext, (int)ext))) goto out0;
        if (!ristretto_eq(&c, &V[j])) { rc = ORC_INVALID_ARGUMENT; goto out0; }
    }

    merlin_transcript T;
    memcpy(&T.s, transcript, ORC_TRANSCRIPT_BYTES);
    rpt R;
    if ((rc = rpt_new(&R, &T, p, (int)m, st, w, rng))) { rpt_free(&R); goto out0; }   /* :287-297 */

    sc *aL = malloc(sizeof(sc) * N), *aR = malloc(sizeof(sc) * N);
    sc *ypow = malloc(sizeof(sc) * (N + 2)), *d = malloc(sizeof(sc) * N);
    ge *Gi = malloc(sizeof(ge) * N), *Hi = malloc(sizeof(ge) * N);
    sc *mscal = malloc(sizeof(sc) * (2 * N + 2 * p->n * p->M + 8));
    ge *mpts = malloc(sizeof(ge) * (2 * N + 8));
    sc one, alpha[6], seed;
    sc_1(&one);
    if (st->seed_nonce) sc_from_bytes_mod_order(&seed, st->seed_nonce);

    for (size_t j = 0; j < m; j++) {                                   /* :300-322 */
        uint64_t v = w->values[j];
        if (st->min_present[j]) {
            if (v < st->min_values[j]) { rc = ORC_INVALID_ARGUMENT; goto out1; }
            v -= st->min_values[j];
        }
        for (size_t i = 0; i < n; i++) {
            sc_from_u64(&aL[j * n + i], (v >> i) & 1);
            sc_sub(&aR[j * n + i], &aL[j * n + i], &one);
        }
    }
    for (size_t k = 0; k < ext; k++) {                                 /* :325-333 */
        if (st->seed_nonce) { if ((rc = nonce(&alpha[k], &seed, "alpha", 0, 0, 1, (uint32_t)k))) goto out1; }
        else random_not_zero(&alpha[k], &R.ad.base);
    }
    ge A;
    {                                                                  /* :334-345 */
        for (size_t i = 0; i < N; i++) { mscal[2 * i] = aL[i]; mscal[2 * i + 1] = aR[i]; }
        msm_mixed(&A, p->pc, mscal, 2 * N, alpha, p->G, ext);
    }
    uint8_t Ac[32];
    ristretto_encode(Ac, &A);
    sc y, z, z2;
    if ((rc = rpt_challenges_y_z(&R, Ac, &y, &z))) goto out1;         /* :348 */
    sc_mul(&z2, &z, &z);
    sc_1(&ypow[0]);                                                    /* :353-359 */
    for (size_t i = 1; i < N + 2; i++) sc_mul(&ypow[i], &ypow[i - 1], &y);
    d[0] = z2;                                                         /* :362-373 */
    for (size_t i = 1; i < n; i++) sc_add(&d[i], &d[i - 1], &d[i - 1]);
    for (size_t j = 1; j < m; j++)
        for (size_t i = 0; i < n; i++) sc_mul(&d[j * n + i], &d[(j - 1) * n + i], &z2);
    for (size_t i = 0; i < N; i++) sc_sub(&aL[i], &aL[i], &z);         /* :376-378 */
    for (size_t i = 0; i < N; i++) {                                   /* :379-381 */
        sc t;
        sc_mul(&t, &d[i], &ypow[N - i]);
        sc_add(&t, &t, &z);
        sc_add(&aR[i], &aR[i], &t);
    }
    {                                                                  /* :382-392 */
        sc zeven;
        sc_1(&zeven);
        for (size_t j = 0; j < m; j++) {
            sc_mul(&zeven, &zeven, &z2);
            for (size_t k = 0; k < ext; k++) {
                sc t;
                sc_mul(&t, &zeven, &blind[j * ext + k]);
                sc_mul(&t, &t, &ypow[N + 1]);
                sc_add(&alpha[k], &alpha[k], &t);
            }
        }
    }
    memcpy(Gi, p->Gi, sizeof(ge) * N);                                 /* :395-396 */
    memcpy(Hi, p->Hi, sizeof(ge) * N);

    size_t rounds = 0;
    while (((size_t)1 << rounds) < N) rounds++;
    memset(out, 0, sizeof *out);
    size_t nn = N, round = 0;
    while (nn > 1) {                                                   /* :409-538 */
        nn /= 2;
        sc *a_lo = aL, *a_hi = aL + nn, *b_lo = aR, *b_hi = aR + nn;
        if (sc_iszero(&ypow[nn])) { rc = ORC_INVALID_ARGUMENT; goto out1; }
        sc yninv;
        sc_invert(&yninv, &ypow[nn]);
        sc *a_lo_off = malloc(sizeof(sc) * nn), *a_hi_off = malloc(sizeof(sc) * nn);
        for (size_t i = 0; i < nn; i++) { sc_mul(&a_lo_off[i], &a_lo[i], &yninv); sc_mul(&a_hi_off[i], &a_hi[i], &ypow[nn]); }
        sc dL[6], dR[6];
        for (size_t k = 0; k < ext; k++) {                             /* :437-450 */
            if (st->seed_nonce) nonce(&dL[k], &seed, "dL", 1, (uint32_t)round, 1, (uint32_t)k);
            else random_not_zero(&dL[k], &R.ad.base);
        }
        for (size_t k = 0; k < ext; k++) {                             /* :451-464 */
            if (st->seed_nonce) nonce(&dR[k], &seed, "dR", 1, (uint32_t)round, 1, (uint32_t)k);
            else random_not_zero(&dR[k], &R.ad.base);
        }
        round++;
        sc cL, cR;
        sc_0(&cL); sc_0(&cR);
        for (size_t i = 0; i < nn; i++) {                              /* :468-479 */
            sc t;
            sc_mul(&t, &a_lo[i], &ypow[i + 1]); sc_mul(&t, &t, &b_hi[i]); sc_add(&cL, &cL, &t);
            sc_mul(&t, &a_hi[i], &ypow[nn + 1 + i]); sc_mul(&t, &t, &b_lo[i]); sc_add(&cR, &cR, &t);
        }
        ge Lp, Rp;
        {                                                              /* :482-495 */
            size_t c = 0;
            mscal[c] = cL; mpts[c++] = p->H;
            for (size_t k = 0; k < ext; k++) { mscal[c] = dL[k]; mpts[c++] = p->G[k]; }
            for (size_t i = 0; i < nn; i++) { mscal[c] = a_lo_off[i]; mpts[c++] = Gi[nn + i]; }
            for (size_t i = 0; i < nn; i++) { mscal[c] = b_hi[i]; mpts[c++] = Hi[i]; }
            msm_vartime(&Lp, mscal, mpts, c);
            c = 0;
            mscal[c] = cR; mpts[c++] = p->H;
            for (size_t k = 0; k < ext; k++) { mscal[c] = dR[k]; mpts[c++] = p->G[k]; }
            for (size_t i = 0; i < nn; i++) { mscal[c] = a_hi_off[i]; mpts[c++] = Gi[i]; }
            for (size_t i = 0; i < nn; i++) { mscal[c] = b_lo[i]; mpts[c++] = Hi[nn + i]; }
            msm_vartime(&Rp, mscal, mpts, c);
        }
        ristretto_encode(out->li[out->n_li++], &Lp);
        ristretto_encode(out->ri[out->n_ri++], &Rp);
        sc e, e2, einv, einv2, eyninv;
        if ((rc = rpt_challenge_round_e(&R, out->li[out->n_li - 1], out->ri[out->n_ri - 1], &e))) {   /* :498-505 */
            free(a_lo_off); free(a_hi_off); goto out1;
        }
        sc_mul(&e2, &e, &e);
        sc_invert(&einv, &e);
        sc_mul(&einv2, &einv, &einv);
        sc_mul(&eyninv, &e, &yninv);
        for (size_t i = 0; i < nn; i++) {                              /* :511-533 */
            sc s2[2]; ge p2[2]; ge g, h;
            s2[0] = einv; s2[1] = eyninv; p2[0] = Gi[i]; p2[1] = Gi[nn + i];
            msm_straus(&g, s2, p2, 2);
            s2[0] = e; s2[1] = einv; p2[0] = Hi[i]; p2[1] = Hi[nn + i];
            msm_straus(&h, s2, p2, 2);
            Gi[i] = g; Hi[i] = h;
            sc t, u;
            sc_mul(&t, &a_lo[i], &e); sc_mul(&u, &a_hi_off[i], &einv); sc_add(&t, &t, &u);
            sc_mul(&u, &b_lo[i], &einv);
            sc v2; sc_mul(&v2, &b_hi[i], &e); sc_add(&u, &u, &v2);
            aL[i] = t; aR[i] = u;
        }
        for (size_t k = 0; k < ext; k++) {                             /* :535-537 */
            sc t, u;
            sc_mul(&t, &dL[k], &e2); sc_mul(&u, &dR[k], &einv2); sc_add(&t, &t, &u);
            sc_add(&alpha[k], &alpha[k], &t);
        }
        free(a_lo_off); free(a_hi_off);
    }
    (void)rounds;
    sc r, s, dd[6], eta[6];
    random_not_zero(&r, &R.ad.base);                                   /* :542-543 */
    random_not_zero(&s, &R.ad.base);
    for (size_t k = 0; k < ext; k++) {                                 /* :544-557 */
        if (st->seed_nonce) nonce(&dd[k], &seed, "d", 0, 0, 1, (uint32_t)k);
        else random_not_zero(&dd[k], &R.ad.base);
    }
    for (size_t k = 0; k < ext; k++) {                                 /* :558-571 */
        if (st->seed_nonce) nonce(&eta[k], &seed, "eta", 0, 0, 1, (uint32_t)k);
        else random_not_zero(&eta[k], &R.ad.base);
    }
    ge A1, B;
    {                                                                  /* :574-584 */
        sc t, u;
        size_t c = 0;
        mscal[c] = r; mpts[c++] = Gi[0];
        mscal[c] = s; mpts[c++] = Hi[0];
        sc_mul(&t, &r, &ypow[1]); sc_mul(&t, &t, &aR[0]);
        sc_mul(&u, &s, &ypow[1]); sc_mul(&u, &u, &aL[0]);
        sc_add(&t, &t, &u);
        mscal[c] = t; mpts[c++] = p->H;
        for (size_t k = 0; k < ext; k++) { mscal[c] = dd[k]; mpts[c++] = p->G[k]; }
        msm_straus(&A1, mscal, mpts, c);
        c = 0;
        sc_mul(&t, &r, &ypow[1]); sc_mul(&t, &t, &s);
        mscal[c] = t; mpts[c++] = p->H;
        for (size_t k = 0; k < ext; k++) { mscal[c] = eta[k]; mpts[c++] = p->G[k]; }
        msm_straus(&B, mscal, mpts, c);
    }
    ristretto_encode(out->a1, &A1);
    ristretto_encode(out->b, &B);
    sc e, e2, t, u;
    if ((rc = rpt_challenge_final_e(&R, out->a1, out->b, &e))) goto out1;   /* :587 */
    sc_mul(&e2, &e, &e);
    sc_mul(&t, &aL[0], &e); sc_add(&t, &t, &r); sc_tobytes(out->r1, &t);    /* :590-594 */
    sc_mul(&t, &aR[0], &e); sc_add(&t, &t, &s); sc_tobytes(out->s1, &t);
    for (size_t k = 0; k < ext; k++) {
        sc_mul(&t, &dd[k], &e); sc_mul(&u, &alpha[k], &e2);
        sc_add(&t, &t, &u); sc_add(&t, &t, &eta[k]);
        sc_tobytes(out->d1[k], &t);
    }
    out->n_d1 = (int32_t)ext;
    out->extension_degree = p->ext;
    memcpy(out->a, Ac, 32);
    memcpy(transcript, &T.s, ORC_TRANSCRIPT_BYTES);
out1:
    rpt_free(&R);
    free(aL); free(aR); free(ypow); free(d); free(Gi); free(Hi); free(mscal); free(mpts);
out0:
    free(blind); free(V);
    return rc;
}

/* ---- to_bytes / from_bytes: src/range_proof.rs:1120-1257 ---- */
int orc_proof_to_bytes(const orc_proof *p, uint8_t *out, size_t cap, size_t *len) {
    size_t pairs = (size_t)(p->n_li < p->n_ri ? p->n_li : p->n_ri);
    size_t need = 1 + 32 * ((size_t)p->n_d1 + 5 + 2 * pairs);
    *len = need;
    if (cap < need) return ORC_INVALID_LENGTH;
    size_t o = 0;
    out[o++] = (uint8_t)p->extension_degree;
    for (int i = 0; i < p->n_d1; i++) { memcpy(out + o, p->d1[i], 32); o += 32; }
    memcpy(out + o, p->a, 32); o += 32;
    memcpy(out + o, p->a1, 32); o += 32;
    memcpy(out + o, p->b, 32); o += 32;
    memcpy(out + o, p->r1, 32); o += 32;
    memcpy(out + o, p->s1, 32); o += 32;
    for (size_t i = 0; i < pairs; i++) {
        memcpy(out + o, p->li[i], 32); o += 32;
        memcpy(out + o, p->ri[i], 32); o += 32;
    }
    return ORC_OK;
}

int orc_proof_from_bytes(const uint8_t *in, size_t len, orc_proof *out) {
    memset(out, 0, sizeof *out);
    if (len < 1) return ORC_INVALID_LENGTH;                            /* :1184-1188 */
    if (in[0] < 1 || in[0] > 6) return ORC_INVALID_ARGUMENT;
    int ext = in[0];
    const uint8_t *q = in + 1;
    size_t chunks = (len - 1) / 32, rem = (len - 1) % 32, c = 0;
    sc tmp;
    for (int i = 0; i < ext; i++) {                                    /* :1197-1199 */
        if (c >= chunks) return ORC_INVALID_LENGTH;
        if (!sc_from_canonical(&tmp, q + 32 * c)) return ORC_INVALID_ARGUMENT;
        memcpy(out->d1[i], q + 32 * c, 32); c++;
    }
    out->n_d1 = ext; out->extension_degree = ext;
    if (c >= chunks) return ORC_INVALID_LENGTH;
    memcpy(out->a, q + 32 * c, 32); c++;                               /* :1202-1206 */
    if (c >= chunks) return ORC_INVALID_LENGTH;
    memcpy(out->a1, q + 32 * c, 32); c++;
    if (c >= chunks) return ORC_INVALID_LENGTH;
    memcpy(out->b, q + 32 * c, 32); c++;
    if (c >= chunks) return ORC_INVALID_LENGTH;
    if (!sc_from_canonical(&tmp, q + 32 * c)) return ORC_INVALID_ARGUMENT;
    memcpy(out->r1, q + 32 * c, 32); c++;
    if (c >= chunks) return ORC_INVALID_LENGTH;
    if (!sc_from_canonical(&tmp, q + 32 * c)) return ORC_INVALID_ARGUMENT;
    memcpy(out->s1, q + 32 * c, 32); c++;
    size_t left = chunks - c;
    size_t pairs = left / 2;
    if (pairs > ORC_MAX_ROUNDS) return ORC_INVALID_LENGTH;
    for (size_t i = 0; i < pairs; i++) {
        memcpy(out->li[i], q + 32 * (c + 2 * i), 32);
        memcpy(out->ri[i], q + 32 * (c + 2 * i + 1), 32);
    }
    out->n_li = out->n_ri = (int32_t)pairs;
    if (pairs == 0) return ORC_INVALID_LENGTH;                         /* :1232-1234 */
    if ((left & 1) || rem) return ORC_INVALID_LENGTH;                  /* :1240-1244 */
    return ORC_OK;
}

/* ---- verify_statements_and_generators_consistency: src/range_proof.rs:610-709 ---- */
static int params_same_pc(const orc_params *a, const orc_params *b, int *g_ok, int *h_ok) {
    *g_ok = (a->ext == b->ext);
    if (*g_ok) for (int i = 0; i < a->ext; i++) if (!ristretto_eq(&a->G[i], &b->G[i])) *g_ok = 0;
    *h_ok = ristretto_eq(&a->H, &b->H);
    return 0;
}

static int consistency(const orc_statement *st, const orc_proof *pr, size_t n, size_t *max_mn, size_t *max_index) {
    if (n == 0) return ORC_INVALID_ARGUMENT;
    const orc_params *p0 = st[0].params;
    size_t mm = (size_t)st[0].m * (size_t)p0->n, mi = 0;
    if (pr[0].n_d1 < 1 || pr[0].n_d1 > 6) return ORC_INVALID_ARGUMENT;   /* ExtensionDegree::try_from */
    if (p0->ext != pr[0].n_d1) return ORC_INVALID_ARGUMENT;             /* :637-639 */
    for (size_t i = 1; i < n; i++) {                                    /* :640-670 */
        const orc_params *p = st[i].params;
        int g_ok, h_ok;
        params_same_pc(p0, p, &g_ok, &h_ok);
        if (!g_ok) return ORC_INVALID_ARGUMENT;
        if (!h_ok) return ORC_INVALID_ARGUMENT;
        if (p0->n != p->n) return ORC_INVALID_ARGUMENT;
        if (pr[i].n_d1 < 1 || pr[i].n_d1 > 6) return ORC_INVALID_ARGUMENT;
        if (p0->ext != p->ext || p0->ext != pr[i].n_d1) return ORC_INVALID_ARGUMENT;
        size_t full = (size_t)st[i].m * (size_t)p->n;
        if (full > mm) { mm = full; mi = i; }
    }
    const orc_params *pm = st[mi].params;
    for (size_t i = 0; i < n; i++) {                                    /* :674-706 */
        for (int j = 0; j < st[i].n_min; j++)
            if (st[i].min_present[j] && p0->n < 64 && (st[i].min_values[j] >> p0->n) > 0) return ORC_INVALID_LENGTH;
        if (i == mi) continue;
        const orc_params *p = st[i].params;
        if (p == pm) continue; /* same object: trivially equal */
        size_t cnt = (size_t)p->n * (size_t)p->M, cm = (size_t)pm->n * (size_t)pm->M;
        if (cm < cnt) cnt = cm;                                         /* zip() truncates */
        for (size_t k = 0; k < cnt; k++) if (!ristretto_eq(&p->Gi[k], &pm->Gi[k])) return ORC_INVALID_ARGUMENT;
        for (size_t k = 0; k < cnt; k++) if (!ristretto_eq(&p->Hi[k], &pm->Hi[k])) return ORC_INVALID_ARGUMENT;
    }
    *max_mn = mm; *max_index = mi;
    return ORC_OK;
}

typedef struct { sc y, z, e; sc round_e[ORC_MAX_ROUNDS]; int n_round; } challenges;

/* ---- RangeProof::verify: src/range_proof.rs:756-1065 ---- */
static int verify_inner(uint8_t *transcripts, size_t n_transcripts, const orc_statement *st, const orc_proof *pr,
                        size_t n, int action, uint8_t *out_masks, uint8_t *out_mask_present) {
    size_t max_mn, max_index;
    int rc = consistency(st, pr, n, &max_mn, &max_index);               /* :763 */
    if (rc) return rc;
    const orc_params *p0 = st[0].params, *pm = st[max_index].params;
    size_t bit_length = (size_t)p0->n, ext = (size_t)p0->ext;
    sc one, two, two_n_minus_one;
    sc_1(&one); sc_from_u64(&two, 2);
    sc_pow_u64(&two_n_minus_one, &two, (uint64_t)bit_length);          /* :781-782 */
    sc_sub(&two_n_minus_one, &two_n_minus_one, &one);

    sc g_scal[6], h_scal;
    for (int i = 0; i < 6; i++) sc_0(&g_scal[i]);
    sc_0(&h_scal);
    sc *gi_scal = calloc(max_mn, sizeof(sc)), *hi_scal = calloc(max_mn, sizeof(sc));
    size_t dyn_cap = ext + 1;                                           /* :794-803 */
    for (size_t i = 0; i < n; i++) dyn_cap += (size_t)st[i].m + 3 + 2 * (size_t)pr[i].n_li;
    sc *dyn_s = malloc(sizeof(sc) * (dyn_cap + 2 * ORC_MAX_ROUNDS));
    ge *dyn_p = malloc(sizeof(ge) * (dyn_cap + 2 * ORC_MAX_ROUNDS));
    size_t nd = 0;
    challenges *ch = malloc(sizeof(challenges) * n);
    sc *s_vec = malloc(sizeof(sc) * (max_mn ? max_mn : 1)), *d = malloc(sizeof(sc) * (max_mn ? max_mn : 1));
    size_t n_masks = 0;

    merlin_transcript weight_t;
    merlin_init(&weight_t, (const uint8_t *)"Bulletproofs+ verifier weights", 30);   /* :811 */
    null_rng nrng;
    null_rng_init(&nrng);

    size_t loop1 = n < n_transcripts ? n : n_transcripts;               /* izip! truncation, :816 */
    for (size_t i = 0; i < loop1; i++) {                                /* :816-850 */
        merlin_transcript T;
        memcpy(&T.s, transcripts + ORC_TRANSCRIPT_BYTES * i, ORC_TRANSCRIPT_BYTES);
        rpt R;
        /* NB the verifier passes first_statement's compressed bases and ITS OWN commitments.len() */
        rc = rpt_new(&R, &T, p0, st[i].m, &st[i], NULL, &nrng.base);
        if (!rc) rc = rpt_challenges_y_z(&R, pr[i].a, &ch[i].y, &ch[i].z);
        size_t pairs = (size_t)(pr[i].n_li < pr[i].n_ri ? pr[i].n_li : pr[i].n_ri);
        ch[i].n_round = (int)pairs;
        for (size_t j = 0; !rc && j < pairs; j++) rc = rpt_challenge_round_e(&R, pr[i].li[j], pr[i].ri[j], &ch[i].round_e[j]);
        if (!rc) rc = rpt_challenge_final_e(&R, pr[i].a1, pr[i].b, &ch[i].e);
        if (!rc) {
            rpt_to_verifier_rng(&R, pr[i].r1, pr[i].s1, pr[i].d1, pr[i].n_d1);
            uint8_t bytes[32];
            merlin_rng_fill(&R.rng, bytes, 32);
            merlin_append_message(&weight_t, "proof", bytes, 32);
        }
        memcpy(transcripts + ORC_TRANSCRIPT_BYTES * i, &T.s, ORC_TRANSCRIPT_BYTES);
        rpt_free(&R);
        if (rc) goto done;
    }
    merlin_rng wrng;                                                    /* :853 */
    merlin_build_rng(&wrng, &weight_t, NULL, 0, 0, &nrng.base);
    trng_adapter wad;
    wad.base.fill = trng_fill; wad.m = &wrng;

    for (size_t pi = 0; pi < loop1; pi++) {                             /* :856-1033 */
        const orc_proof *proof = &pr[pi];
        const orc_statement *s = &st[pi];
        ge A, A1, B, Lp[ORC_MAX_ROUNDS], Rp[ORC_MAX_ROUNDS];
        if (!ristretto_decode(&A, proof->a)) { rc = ORC_INVALID_ARGUMENT; goto done; }     /* :859-866 */
        if (!ristretto_decode(&A1, proof->a1)) { rc = ORC_INVALID_ARGUMENT; goto done; }
        if (!ristretto_decode(&B, proof->b)) { rc = ORC_INVALID_ARGUMENT; goto done; }
        sc r1, s1, d1[6];
        sc_from_bytes_mod_order(&r1, proof->r1);
        sc_from_bytes_mod_order(&s1, proof->s1);
        for (int k = 0; k < proof->n_d1; k++) sc_from_bytes_mod_order(&d1[k], proof->d1[k]);
        for (int j = 0; j < proof->n_li; j++)
            if (!ristretto_decode(&Lp[j], proof->li[j])) { rc = ORC_INVALID_ARGUMENT; goto done; }
        for (int j = 0; j < proof->n_ri; j++)
            if (!ristretto_decode(&Rp[j], proof->ri[j])) { rc = ORC_INVALID_ARGUMENT; goto done; }
        size_t m = (size_t)s->m, N = m * bit_length, rounds = (size_t)proof->n_li;
        if (proof->n_li != proof->n_ri) { rc = ORC_INVALID_LENGTH; goto done; }             /* :875-879 */
        if (rounds >= 32) { rc = ORC_SIZE_OVERFLOW; goto done; }
        if (((size_t)1 << rounds) != N) { rc = ORC_INVALID_LENGTH; goto done; }             /* :886-888 */
        sc y = ch[pi].y, z = ch[pi].z, e = ch[pi].e;
        sc weight;
        random_not_zero(&weight, &wad.base);                                                /* :894 */
        sc cinv[ORC_MAX_ROUNDS + 2], cinv_prod, ym1, y_inv, y_1_inv;
        for (size_t j = 0; j < rounds; j++) cinv[j] = ch[pi].round_e[j];                    /* :897-905 */
        sc_sub(&ym1, &y, &one);
        cinv[rounds] = y; cinv[rounds + 1] = ym1;
        /* dalek's batch_invert asserts non-zero in debug only; y == 1 makes (y-1)^-1 = 0-ish garbage there.
         * y == 1 has probability 2^-252; treat as a failed verification to stay total. */
        if (sc_iszero(&ym1)) { rc = ORC_VERIFICATION_FAILED; goto done; }
        sc_batch_invert(cinv, rounds + 2, &cinv_prod);
        sc_mul(&cinv_prod, &cinv_prod, &y); sc_mul(&cinv_prod, &cinv_prod, &ym1);
        y_1_inv = cinv[rounds + 1]; y_inv = cinv[rounds];
        sc z2, e2, csq[ORC_MAX_ROUNDS], csqinv[ORC_MAX_ROUNDS], y_nm, y_nm_1, y_sum, t, u;
        sc_mul(&z2, &z, &z); sc_mul(&e2, &e, &e);                                           /* :908-916 */
        for (size_t j = 0; j < rounds; j++) { sc_mul(&csq[j], &ch[pi].round_e[j], &ch[pi].round_e[j]); sc_mul(&csqinv[j], &cinv[j], &cinv[j]); }
        sc_pow_u64(&y_nm, &y, (uint64_t)N);
        sc_mul(&y_nm_1, &y_nm, &y);
        sc_sub(&t, &y_nm, &one); sc_mul(&t, &t, &y); sc_mul(&y_sum, &t, &y_1_inv);
        d[0] = z2;                                                                          /* :919-929 */
        for (size_t i = 1; i < bit_length; i++) sc_add(&d[i], &d[i - 1], &d[i - 1]);
        for (size_t j = 1; j < m; j++)
            for (size_t i = 0; i < bit_length; i++) sc_mul(&d[j * bit_length + i], &d[(j - 1) * bit_length + i], &z2);
        sc d_sum = z2, d_sum_temp_z = z2;                                                   /* :932-938 */
        for (size_t mm = m; mm > 1; mm >>= 1) {
            sc_mul(&t, &d_sum, &d_sum_temp_z); sc_add(&d_sum, &d_sum, &t);
            sc_mul(&d_sum_temp_z, &d_sum_temp_z, &d_sum_temp_z);
        }
        sc_mul(&d_sum, &d_sum, &two_n_minus_one);

        if (action == ORC_VERIFY_ONLY) {                                                    /* :941-969 */
            out_mask_present[n_masks++] = 0;
        } else {
            if (s->seed_nonce) {
                sc seed;
                sc_from_bytes_mod_order(&seed, s->seed_nonce);
                sc e2inv, zy_inv;
                sc_invert(&e2inv, &e2);
                sc_mul(&zy_inv, &z2, &y_nm_1); sc_invert(&zy_inv, &zy_inv);
                for (size_t k = 0; k < ext && k < (size_t)proof->n_d1; k++) {
                    sc mask, nn;
                    nonce(&nn, &seed, "eta", 0, 0, 1, (uint32_t)k); sc_sub(&mask, &d1[k], &nn);
                    nonce(&nn, &seed, "d", 0, 0, 1, (uint32_t)k); sc_mul(&nn, &nn, &e); sc_sub(&mask, &mask, &nn);
                    sc_mul(&mask, &mask, &e2inv);
                    nonce(&nn, &seed, "alpha", 0, 0, 1, (uint32_t)k); sc_sub(&mask, &mask, &nn);
                    for (size_t j = 0; j < rounds; j++) {
                        nonce(&nn, &seed, "dL", 1, (uint32_t)j, 1, (uint32_t)k); sc_mul(&nn, &nn, &csq[j]); sc_sub(&mask, &mask, &nn);
                        nonce(&nn, &seed, "dR", 1, (uint32_t)j, 1, (uint32_t)k); sc_mul(&nn, &nn, &csqinv[j]); sc_sub(&mask, &mask, &nn);
                    }
                    sc_mul(&mask, &mask, &zy_inv);
                    sc_tobytes(out_masks + (n_masks * ext + k) * 32, &mask);
                }
                out_mask_present[n_masks++] = 1;
            } else {
                out_mask_present[n_masks++] = 0;
            }
            if (action == ORC_RECOVER_ONLY) continue;
        }

        sc y_inv_i = one, y_nm_i = y_nm;                                                    /* :972-1003 */
        s_vec[0] = cinv_prod;
        for (size_t i = 1; i < N; i++) {
            size_t log_i = 63 - (size_t)__builtin_clzll((unsigned long long)i);
            size_t j = (size_t)1 << log_i;
            sc_mul(&s_vec[i], &s_vec[i - j], &csq[rounds - log_i - 1]);
        }
        sc r1_e, s1_e, e2z;
        sc_mul(&r1_e, &r1, &e); sc_mul(&s1_e, &s1, &e); sc_mul(&e2z, &e2, &z);
        for (size_t i = 0; i < N && i < max_mn; i++) {
            sc g, h;
            sc_mul(&g, &r1_e, &y_inv_i); sc_mul(&g, &g, &s_vec[i]);
            sc_mul(&h, &s1_e, &s_vec[N - 1 - i]);
            sc_add(&t, &g, &e2z); sc_mul(&t, &t, &weight); sc_add(&gi_scal[i], &gi_scal[i], &t);
            sc_mul(&u, &d[i], &y_nm_i); sc_add(&u, &u, &z); sc_mul(&u, &u, &e2); sc_sub(&u, &h, &u);
            sc_mul(&u, &u, &weight); sc_add(&hi_scal[i], &hi_scal[i], &u);
            sc_mul(&y_inv_i, &y_inv_i, &y_inv);
            sc_mul(&y_nm_i, &y_nm_i, &y_inv);
        }
        sc zeven = one, neg_e2;                                                             /* :1006-1015 */
        sc_neg(&neg_e2, &e2);
        for (size_t j = 0; j < (size_t)s->n_min; j++) {
            sc weighted;
            sc_mul(&zeven, &zeven, &z2);
            sc_mul(&weighted, &neg_e2, &zeven); sc_mul(&weighted, &weighted, &y_nm_1); sc_mul(&weighted, &weighted, &weight);
            dyn_s[nd + j] = weighted;
            if (s->min_present[j]) {
                sc mv;
                sc_from_u64(&mv, s->min_values[j]);
                sc_mul(&t, &weighted, &mv); sc_sub(&h_scal, &h_scal, &t);
            }
        }
        for (size_t j = 0; j < m; j++)
            if (!ristretto_decode(&dyn_p[nd + j], s->commitments + 32 * j)) { rc = ORC_INVALID_ARGUMENT; goto done; }
        nd += m;
        /* :1017-1020 */
        sc_mul(&t, &r1, &y); sc_mul(&t, &t, &s1);
        sc_mul(&u, &y_nm_1, &z); sc_mul(&u, &u, &d_sum);
        sc zz; sc_sub(&zz, &z2, &z); sc_mul(&zz, &zz, &y_sum);
        sc_add(&u, &u, &zz); sc_mul(&u, &u, &e2);
        sc_add(&t, &t, &u); sc_mul(&t, &t, &weight); sc_add(&h_scal, &h_scal, &t);
        for (size_t k = 0; k < ext && k < (size_t)proof->n_d1; k++) { sc_mul(&t, &weight, &d1[k]); sc_add(&g_scal[k], &g_scal[k], &t); }
        /* :1022-1032 */
        sc neg_e, neg_w, w_neg_e2;
        sc_neg(&neg_e, &e); sc_mul(&dyn_s[nd], &weight, &neg_e); dyn_p[nd++] = A1;
        sc_neg(&neg_w, &weight); dyn_s[nd] = neg_w; dyn_p[nd++] = B;
        sc_mul(&w_neg_e2, &weight, &neg_e2); dyn_s[nd] = w_neg_e2; dyn_p[nd++] = A;
        for (size_t j = 0; j < rounds; j++) { sc_mul(&dyn_s[nd], &w_neg_e2, &csq[j]); dyn_p[nd++] = Lp[j]; }
        for (size_t j = 0; j < rounds; j++) { sc_mul(&dyn_s[nd], &w_neg_e2, &csqinv[j]); dyn_p[nd++] = Rp[j]; }
    }
    if (action == ORC_RECOVER_ONLY) { rc = ORC_OK; goto done; }          /* :1034-1036 */
    for (size_t k = 0; k < ext; k++) { dyn_s[nd] = g_scal[k]; dyn_p[nd++] = p0->G[k]; }     /* :1039-1042 */
    dyn_s[nd] = h_scal; dyn_p[nd++] = p0->H;
    {                                                                    /* :1045-1062 */
        sc *stat = malloc(sizeof(sc) * (2 * max_mn + 1));
        for (size_t i = 0; i < max_mn; i++) { stat[2 * i] = gi_scal[i]; stat[2 * i + 1] = hi_scal[i]; }
        ge res;
        msm_mixed(&res, pm->pc, stat, 2 * max_mn, dyn_s, dyn_p, nd);
        free(stat);
        if (!ristretto_is_identity(&res)) rc = ORC_VERIFICATION_FAILED;
    }
done:
    free(gi_scal); free(hi_scal); free(dyn_s); free(dyn_p); free(ch); free(s_vec); free(d);
    return rc;
}

/* ---- RangeProof::verify_batch: src/range_proof.rs:712-752 ---- */
int orc_verify_batch(uint8_t *transcripts, size_t n_transcripts, const orc_statement *statements, size_t n_statements,
                     const orc_proof *proofs, size_t n_proofs, int action,
                     uint8_t *out_masks, uint8_t *out_mask_present, size_t *n_results) {
    if (n_results) *n_results = 0;
    if (n_statements == 0 || n_proofs == 0 || n_transcripts == 0) return ORC_INVALID_ARGUMENT;   /* :719-723 */
    if (n_statements != n_proofs) return ORC_INVALID_ARGUMENT;                                  /* :725-729 */
    if (n_transcripts != n_statements) return ORC_INVALID_ARGUMENT;                             /* :730-734 */
    size_t n = n_statements < 256 ? n_statements : 256;                                         /* :739-749 */
    int rc = verify_inner(transcripts, n_transcripts, statements, proofs, n, action, out_masks, out_mask_present);
    if (!rc && n_results) *n_results = n;
    return rc;
}

/* ---- multi-threaded CPU baseline driver ---- */
typedef struct {
    const uint8_t *transcripts; const orc_statement *st; const orc_proof *pr; const size_t *off;
    size_t n_chunks; int action; int32_t *codes; size_t next; pthread_mutex_t mu;
} mt_job;

static void *mt_worker(void *arg) {
    mt_job *job = arg;
    for (;;) {
        pthread_mutex_lock(&job->mu);
        size_t c = job->next++;
        pthread_mutex_unlock(&job->mu);
        if (c >= job->n_chunks) break;
        size_t lo = job->off[c], hi = job->off[c + 1], cnt = hi - lo;
        uint8_t *tr = malloc(cnt * ORC_TRANSCRIPT_BYTES);
        memcpy(tr, job->transcripts + lo * ORC_TRANSCRIPT_BYTES, cnt * ORC_TRANSCRIPT_BYTES);
        uint8_t *masks = malloc(cnt * 6 * 32 + 1), *present = malloc(cnt + 1);
        size_t nres;
        job->codes[c] = orc_verify_batch(tr, cnt, job->st + lo, cnt, job->pr + lo, cnt, job->action, masks, present, &nres);
        free(tr); free(masks); free(present);
    }
    return NULL;
}

double orc_verify_chunks_mt(const uint8_t *transcripts, const orc_statement *statements, const orc_proof *proofs,
                            const size_t *chunk_offsets, size_t n_chunks, int action, int threads, int32_t *out_codes) {
    mt_job job = {transcripts, statements, proofs, chunk_offsets, n_chunks, action, out_codes, 0, PTHREAD_MUTEX_INITIALIZER};
    if (threads < 1) threads = 1;
    pthread_t *th = malloc(sizeof(pthread_t) * (size_t)threads);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int i = 0; i < threads; i++) pthread_create(&th[i], NULL, mt_worker, &job);
    for (int i = 0; i < threads; i++) pthread_join(th[i], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ---- primitive wrappers for cross-checks ---- */
int orc_ristretto_decode_encode(const uint8_t in32[32], uint8_t out32[32]) {
    ge p;
    if (!ristretto_decode(&p, in32)) return 0;
    ristretto_encode(out32, &p);
    return 1;
}
void orc_ristretto_from_uniform(const uint8_t in64[64], uint8_t out32[32]) {
    ge p; ristretto_from_uniform(&p, in64); ristretto_encode(out32, &p);
}
int orc_ristretto_add(const uint8_t a32[32], const uint8_t b32[32], uint8_t out32[32]) {
    ge a, b, r;
    if (!ristretto_decode(&a, a32) || !ristretto_decode(&b, b32)) return 0;
    ge_add(&r, &a, &b); ristretto_encode(out32, &r);
    return 1;
}
int orc_ristretto_scalarmult(const uint8_t s32[32], const uint8_t p32[32], uint8_t out32[32]) {
    ge p, r; sc s;
    if (!ristretto_decode(&p, p32)) return 0;
    sc_from_bytes_mod_order(&s, s32);
    ge_scalarmult(&r, &s, &p); ristretto_encode(out32, &r);
    return 1;
}
int orc_msm(const uint8_t *scalars32, const uint8_t *points32, size_t n, int algo, uint8_t out32[32]) {
    sc *s = malloc(sizeof(sc) * (n ? n : 1)); ge *p = malloc(sizeof(ge) * (n ? n : 1)), r;
    int ok = 1;
    for (size_t i = 0; i < n; i++) {
        sc_from_bytes_mod_order(&s[i], scalars32 + 32 * i);
        if (!ristretto_decode(&p[i], points32 + 32 * i)) ok = 0;
    }
    if (ok) {
        if (algo == 1) msm_straus(&r, s, p, n); else if (algo == 2) msm_pippenger(&r, s, p, n); else msm_vartime(&r, s, p, n);
        ristretto_encode(out32, &r);
    }
    free(s); free(p);
    return ok;
}
void orc_sc_mul(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) { sc x, y, r; sc_from_bytes_mod_order(&x, a); sc_from_bytes_mod_order(&y, b); sc_mul(&r, &x, &y); sc_tobytes(out, &r); }
void orc_sc_add(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) { sc x, y, r; sc_from_bytes_mod_order(&x, a); sc_from_bytes_mod_order(&y, b); sc_add(&r, &x, &y); sc_tobytes(out, &r); }
void orc_sc_sub(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) { sc x, y, r; sc_from_bytes_mod_order(&x, a); sc_from_bytes_mod_order(&y, b); sc_sub(&r, &x, &y); sc_tobytes(out, &r); }
void orc_sc_invert(const uint8_t a[32], uint8_t out[32]) { sc x, r; sc_from_bytes_mod_order(&x, a); sc_invert(&r, &x); sc_tobytes(out, &r); }
void orc_sc_from_wide(const uint8_t in[64], uint8_t out[32]) { sc r; sc_from_wide(&r, in); sc_tobytes(out, &r); }
int orc_sc_is_canonical(const uint8_t a[32]) { sc r; return sc_from_canonical(&r, a); }
void orc_fe_mul(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) { fe x, y, r; fe_frombytes(&x, a); fe_frombytes(&y, b); fe_mul(&r, &x, &y); fe_tobytes(out, &r); }
void orc_fe_invert(const uint8_t a[32], uint8_t out[32]) { fe x, r; fe_frombytes(&x, a); fe_invert(&r, &x); fe_tobytes(out, &r); }
int orc_fe_sqrt_ratio_i(const uint8_t u[32], const uint8_t v[32], uint8_t out[32]) { fe x, y, r; fe_frombytes(&x, u); fe_frombytes(&y, v); int ok = fe_sqrt_ratio_i(&r, &x, &y); fe_tobytes(out, &r); return ok; }
void orc_sha3_512(const uint8_t *in, size_t len, uint8_t out[64]) { sha3_512(out, in, len); }
void orc_shake256(const uint8_t *in, size_t len, uint8_t *out, size_t outlen) { keccak_sponge s; shake256_init(&s); shake256_absorb(&s, in, len); shake256_finalize(&s); shake256_squeeze(&s, out, outlen); }
void orc_keccak_f1600(uint64_t st[25]) { keccak_f1600(st); }
/* The verifier-weight transcript of one verify_batch call on its own (range_proof.rs:811, :849, :853, :894): wbytes32 = the 32 rng
 * bytes each proof contributes, in proof order; weights32 = the batch weight drawn for each proof.  Same statements as orc_verify. */
void orc_verifier_weights(const uint8_t *wbytes32, size_t n, uint8_t *weights32) {
    merlin_transcript weight_t;
    merlin_init(&weight_t, (const uint8_t *)"Bulletproofs+ verifier weights", 30);
    null_rng nrng;
    null_rng_init(&nrng);
    for (size_t i = 0; i < n; i++) merlin_append_message(&weight_t, "proof", wbytes32 + 32 * i, 32);
    merlin_rng wrng;
    merlin_build_rng(&wrng, &weight_t, NULL, 0, 0, &nrng.base);
    trng_adapter wad;
    wad.base.fill = trng_fill; wad.m = &wrng;
    for (size_t i = 0; i < n; i++) {
        sc w;
        random_not_zero(&w, &wad.base);
        sc_tobytes(weights32 + 32 * i, &w);
    }
}
