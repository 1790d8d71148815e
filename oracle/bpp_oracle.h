/* ORACLE — TEST INFRASTRUCTURE ONLY (see orc_internal.h for scope and parity status).
 * Public C surface of the CPU restatement of tari_bulletproofs_plus 0.4.1, driven from python/ctypes by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs. */
#ifndef BPP_ORACLE_H
#define BPP_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* /root/reference/src/errors.rs:12-28 */
enum { ORC_OK = 0, ORC_VERIFICATION_FAILED = 1, ORC_INVALID_ARGUMENT = 2, ORC_INVALID_LENGTH = 3,
       ORC_INVALID_BLAKE2B = 4, ORC_SIZE_OVERFLOW = 5 };
/* /root/reference/src/range_proof.rs VerifyAction */
enum { ORC_RECOVER_ONLY = 0, ORC_RECOVER_AND_VERIFY = 1, ORC_VERIFY_ONLY = 2 };

#define ORC_MAX_ROUNDS 32
#define ORC_TRANSCRIPT_BYTES 203 /* 200 B Strobe state + pos + pos_begin + cur_flags */

typedef struct orc_params orc_params;
typedef struct orc_rng orc_rng;

/* parsed RangeProof (/root/reference/src/range_proof.rs:58-68); scalars canonical little-endian */
typedef struct {
    int32_t extension_degree;
    int32_t n_d1, n_li, n_ri;
    uint8_t d1[6][32];
    uint8_t a[32], a1[32], b[32], r1[32], s1[32];
    uint8_t li[ORC_MAX_ROUNDS][32], ri[ORC_MAX_ROUNDS][32];
} orc_proof;

/* RangeStatement (/root/reference/src/range_statement.rs:21-73) as plain data */
typedef struct {
    const orc_params *params;
    int32_t m;
    const uint8_t *commitments;     /* m x 32 B Ristretto encodings */
    const uint64_t *min_values;     /* m */
    const uint8_t *min_present;     /* m, 0/1 */
    const uint8_t *seed_nonce;      /* NULL or 32 B canonical scalar */
    int32_t n_min;                  /* length of the promises vector (== m unless testing the error) */
} orc_statement;

/* RangeWitness (/root/reference/src/range_witness.rs:15-41) */
typedef struct {
    int32_t n_openings;
    const uint64_t *values;         /* n_openings */
    const uint8_t *blindings;       /* n_openings x r_len x 32 B */
    int32_t r_len;                  /* blinding count per opening (extension degree of the witness) */
} orc_witness;

/* RangeParameters::init + ristretto::create_pedersen_gens_with_extension_degree */
int orc_params_new(int bit_length, int max_aggregation, int extension_degree, orc_params **out);
void orc_params_free(orc_params *p);
/* which: 0 = h_base, 1 = g_base[index], 2 = gi_base flat party-major, 3 = hi_base */
int orc_params_point(const orc_params *p, int which, size_t index, uint8_t out32[32]);
int orc_params_bit_length(const orc_params *p);
int orc_params_max_aggregation(const orc_params *p);
int orc_params_extension_degree(const orc_params *p);

/* PedersenGens::commit */
int orc_commit(const orc_params *p, uint64_t value, const uint8_t *blindings32, int n_blindings, uint8_t out32[32]);
int orc_statement_check(const orc_statement *st);   /* RangeStatement::init validation */

void orc_transcript_new(const uint8_t *label, size_t len, uint8_t out[ORC_TRANSCRIPT_BYTES]);
void orc_transcript_append_message(uint8_t t[ORC_TRANSCRIPT_BYTES], const char *label, const uint8_t *msg, size_t len);
void orc_transcript_challenge_bytes(uint8_t t[ORC_TRANSCRIPT_BYTES], const char *label, uint8_t *out, size_t len);

orc_rng *orc_rng_chacha12_seed_from_u64(uint64_t seed);
orc_rng *orc_rng_chacha12_from_seed(const uint8_t seed[32]);
orc_rng *orc_rng_null(void);
orc_rng *orc_rng_buffer(const uint8_t *bytes, size_t len); /* bytes must outlive the rng */
void orc_rng_fill(orc_rng *r, uint8_t *dst, size_t len);
uint64_t orc_rng_next_u64(orc_rng *r);                      /* chacha only */
void orc_rng_free(orc_rng *r);
void orc_random_not_zero(orc_rng *r, uint8_t out32[32]);    /* Scalar::random_not_zero */

/* RangeProof::prove_with_rng; transcript is advanced in place */
int orc_prove(uint8_t transcript[ORC_TRANSCRIPT_BYTES], const orc_statement *st, const orc_witness *w,
              orc_rng *rng, orc_proof *out);
int orc_proof_to_bytes(const orc_proof *p, uint8_t *out, size_t cap, size_t *len);
int orc_proof_from_bytes(const uint8_t *in, size_t len, orc_proof *out);

/* RangeProof::verify_batch.  transcripts: n x 203 B, advanced in place.  Looks only at the first 256 entries
 * (range_proof.rs:739-751).  out_masks: min(n,256) x ext x 32 B; out_mask_present: min(n,256).
 * Returns an ORC_* code; *n_results = number of mask entries written. */
int orc_verify_batch(uint8_t *transcripts, size_t n_transcripts, const orc_statement *statements, size_t n_statements,
                     const orc_proof *proofs, size_t n_proofs, int action,
                     uint8_t *out_masks, uint8_t *out_mask_present, size_t *n_results);

/* primitives exposed for cross-checks */
int orc_ristretto_decode_encode(const uint8_t in32[32], uint8_t out32[32]);       /* 1 if decoded */
void orc_ristretto_from_uniform(const uint8_t in64[64], uint8_t out32[32]);
int orc_ristretto_add(const uint8_t a32[32], const uint8_t b32[32], uint8_t out32[32]);
int orc_ristretto_scalarmult(const uint8_t s32[32], const uint8_t p32[32], uint8_t out32[32]);
/* algo: 0 = dalek dispatch (Straus <190 / Pippenger), 1 = Straus, 2 = Pippenger */
int orc_msm(const uint8_t *scalars32, const uint8_t *points32, size_t n, int algo, uint8_t out32[32]);
void orc_sc_mul(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]);
void orc_sc_add(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]);
void orc_sc_sub(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]);
void orc_sc_invert(const uint8_t a[32], uint8_t out[32]);
void orc_sc_from_wide(const uint8_t in[64], uint8_t out[32]);
int orc_sc_is_canonical(const uint8_t a[32]);
void orc_fe_mul(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]);
void orc_fe_invert(const uint8_t a[32], uint8_t out[32]);
int orc_fe_sqrt_ratio_i(const uint8_t u[32], const uint8_t v[32], uint8_t out[32]);
void orc_sha3_512(const uint8_t *in, size_t len, uint8_t out[64]);
void orc_shake256(const uint8_t *in, size_t len, uint8_t *out, size_t outlen);
void orc_blake2b_nonce(const uint8_t seed32[32], const char *label, int have_j, uint32_t j, int have_k, uint32_t k,
                       uint8_t out32[32]);
void orc_keccak_f1600(uint64_t st[25]);
void orc_verifier_weights(const uint8_t *wbytes32, size_t n, uint8_t *weights32);   /* range_proof.rs:811-853, :894 on its own */

/* CPU baseline: verify `n_chunks` independent batches (each <= 256 proofs, laid out contiguously with
 * chunk_offsets[n_chunks+1]) on `threads` pthreads; per-chunk result codes in out_codes. Returns seconds. */
double orc_verify_chunks_mt(const uint8_t *transcripts, const orc_statement *statements, const orc_proof *proofs,
                            const size_t *chunk_offsets, size_t n_chunks, int action, int threads, int32_t *out_codes);

#ifdef __cplusplus
}
#endif
#endif
