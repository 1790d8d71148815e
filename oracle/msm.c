/* ORACLE (test infrastructure only) — multiscalar multiplication in the same algorithm classes that
 * curve25519-dalek 4.1.3 selects for the reference's call sites:
 *   RistrettoPoint::vartime_multiscalar_mul  -> Straus (width-5 NAF) below 190 points, Pippenger above
 *       (/root/reference/src/range_proof.rs:482-495, 512-521)
 *   VartimeRistrettoPrecomputation::vartime_mixed_multiscalar_mul -> Straus with width-8 NAF affine-Niels
 *       tables for the static points and width-5 NAF projective-Niels tables for the dynamic points
 *       (/root/reference/src/range_proof.rs:339-345, 1050-1057; tables built at
 *        /root/reference/src/generators/bulletproof_gens.rs:100-103).
 * The group element returned is algorithm-independent; the algorithm choice matters only for the CPU
 * baseline timing. */
#include <stdlib.h>
#include "orc_internal.h"

static void naf(int8_t out[256], const sc *s, int w) {
    uint64_t x[5] = {s->v[0], s->v[1], s->v[2], s->v[3], 0};
    uint64_t width = 1ULL << w, mask = width - 1;
    memset(out, 0, 256);
    unsigned pos = 0;
    uint64_t carry = 0;
    while (pos < 256) {
        unsigned idx = pos / 64, bit = pos % 64;
        uint64_t buf = (bit < 64 - (unsigned)w) ? (x[idx] >> bit) : ((x[idx] >> bit) | (x[idx + 1] << (64 - bit)));
        uint64_t window = carry + (buf & mask);
        if ((window & 1) == 0) { pos += 1; continue; }
        if (window < width / 2) { carry = 0; out[pos] = (int8_t)window; }
        else { carry = 1; out[pos] = (int8_t)((int64_t)window - (int64_t)width); }
        pos += w;
    }
}

static void odd_multiples_pniels(ge_pniels *tab, int count, const ge *p) {
    ge p2, cur = *p;
    ge_dbl(&p2, p);
    ge_to_pniels(&tab[0], &cur);
    for (int i = 1; i < count; i++) {
        ge_add(&cur, &cur, &p2);
        ge_to_pniels(&tab[i], &cur);
    }
}

void msm_straus(ge *r, const sc *scalars, const ge *points, size_t n) {
    int8_t (*nafs)[256] = malloc(n ? n * 256 : 1);
    ge_pniels (*tabs)[8] = malloc(n ? n * sizeof(ge_pniels[8]) : 1);
    for (size_t k = 0; k < n; k++) {
        naf(nafs[k], &scalars[k], 5);
        odd_multiples_pniels(tabs[k], 8, &points[k]);
    }
    ge acc;
    ge_identity(&acc);
    int started = 0;
    for (int i = 255; i >= 0; i--) {
        if (started) ge_dbl(&acc, &acc);
        for (size_t k = 0; k < n; k++) {
            int d = nafs[k][i];
            if (d > 0) { ge_add_pniels(&acc, &acc, &tabs[k][d / 2]); started = 1; }
            else if (d < 0) { ge_sub_pniels(&acc, &acc, &tabs[k][(-d) / 2]); started = 1; }
        }
    }
    *r = acc;
    free(nafs);
    free(tabs);
}

/* signed radix-2^w digits, dalek Scalar::as_radix_2w */
static int radix_2w(int8_t *digits, const sc *s, int w) {
    int count = (256 + w - 1) / w;
    if (w == 8) count += 1;
    uint64_t x[5] = {s->v[0], s->v[1], s->v[2], s->v[3], 0};
    uint64_t radix = 1ULL << w, mask = radix - 1;
    uint64_t carry = 0;
    for (int i = 0; i < count; i++) {
        unsigned pos = (unsigned)(i * w), idx = pos / 64, bit = pos % 64;
        uint64_t buf;
        if (idx >= 4) buf = 0;
        else buf = (bit < 64 - (unsigned)w || idx == 3) ? (x[idx] >> bit) : ((x[idx] >> bit) | (x[idx + 1] << (64 - bit)));
        uint64_t coef = carry + (buf & mask);
        carry = (coef + radix / 2) >> w;
        digits[i] = (int8_t)((int64_t)coef - (int64_t)(carry << w));
    }
    return count;
}

void msm_pippenger(ge *r, const sc *scalars, const ge *points, size_t n) {
    int w = n < 500 ? 6 : (n < 800 ? 7 : 8);
    int max_digit = 1 << w;
    int buckets_count = max_digit / 2;
    int8_t *digits = malloc(n ? n * 48 : 1);
    ge_pniels *pn = malloc(n ? n * sizeof(ge_pniels) : 1);
    int count = 0;
    for (size_t k = 0; k < n; k++) {
        count = radix_2w(digits + 48 * k, &scalars[k], w);
        ge_to_pniels(&pn[k], &points[k]);
    }
    if (n == 0) count = (256 + w - 1) / w + (w == 8);
    ge *buckets = malloc(sizeof(ge) * (size_t)buckets_count);
    ge total;
    ge_identity(&total);
    for (int col = count - 1; col >= 0; col--) {
        for (int b = 0; b < buckets_count; b++) ge_identity(&buckets[b]);
        for (size_t k = 0; k < n; k++) {
            int d = digits[48 * k + col];
            if (d > 0) ge_add_pniels(&buckets[d - 1], &buckets[d - 1], &pn[k]);
            else if (d < 0) ge_sub_pniels(&buckets[-d - 1], &buckets[-d - 1], &pn[k]);
        }
        ge inter = buckets[buckets_count - 1], sum = buckets[buckets_count - 1];
        for (int b = buckets_count - 2; b >= 0; b--) {
            ge_add(&inter, &inter, &buckets[b]);
            ge_add(&sum, &sum, &inter);
        }
        if (col != count - 1)
            for (int j = 0; j < w; j++) ge_dbl(&total, &total);
        ge_add(&total, &total, &sum);
    }
    *r = total;
    free(digits);
    free(pn);
    free(buckets);
}

void msm_vartime(ge *r, const sc *scalars, const ge *points, size_t n) {
    if (n < 190) msm_straus(r, scalars, points, n);
    else msm_pippenger(r, scalars, points, n);
}

msm_precomp *msm_precomp_new(const ge *points, size_t n) {
    msm_precomp *pc = malloc(sizeof *pc);
    pc->n = n;
    pc->tab = malloc(n ? n * 64 * sizeof(ge_aniels) : 1);
    for (size_t k = 0; k < n; k++) {
        ge p2, cur = points[k];
        ge_dbl(&p2, &points[k]);
        ge_to_aniels(&pc->tab[64 * k], &cur);
        for (int i = 1; i < 64; i++) {
            ge_add(&cur, &cur, &p2);
            ge_to_aniels(&pc->tab[64 * k + i], &cur);
        }
    }
    return pc;
}

void msm_precomp_free(msm_precomp *pc) {
    if (!pc) return;
    free(pc->tab);
    free(pc);
}

void msm_mixed(ge *r, const msm_precomp *pc, const sc *static_scalars, size_t ns,
               const sc *dyn_scalars, const ge *dyn_points, size_t nd) {
    if (ns > pc->n) ns = pc->n;
    int8_t (*snaf)[256] = malloc(ns ? ns * 256 : 1);
    int8_t (*dnaf)[256] = malloc(nd ? nd * 256 : 1);
    ge_pniels (*dtab)[8] = malloc(nd ? nd * sizeof(ge_pniels[8]) : 1);
    for (size_t k = 0; k < ns; k++) naf(snaf[k], &static_scalars[k], 8);
    for (size_t k = 0; k < nd; k++) {
        naf(dnaf[k], &dyn_scalars[k], 5);
        odd_multiples_pniels(dtab[k], 8, &dyn_points[k]);
    }
    ge acc;
    ge_identity(&acc);
    int started = 0;
    for (int i = 255; i >= 0; i--) {
        if (started) ge_dbl(&acc, &acc);
        for (size_t k = 0; k < nd; k++) {
            int d = dnaf[k][i];
            if (d > 0) { ge_add_pniels(&acc, &acc, &dtab[k][d / 2]); started = 1; }
            else if (d < 0) { ge_sub_pniels(&acc, &acc, &dtab[k][(-d) / 2]); started = 1; }
        }
        for (size_t k = 0; k < ns; k++) {
            int d = snaf[k][i];
            if (d > 0) { ge_add_aniels(&acc, &acc, &pc->tab[64 * k + d / 2]); started = 1; }
            else if (d < 0) { ge_sub_aniels(&acc, &acc, &pc->tab[64 * k + (-d) / 2]); started = 1; }
        }
    }
    *r = acc;
    free(snaf);
    free(dnaf);
    free(dtab);
}
