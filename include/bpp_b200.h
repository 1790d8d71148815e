/*
 * bpp_b200.h — C ABI of the B200-native Bulletproofs+ engine (libbpp_b200.so).
 *
 * This is the drop-in boundary for the MSM / inner-product-argument hot path of tari_bulletproofs_plus 0.4.1.
 * The reference has no FFI: its seam is the trait bundle `P: CurvePointProtocol + Precomputable +
 * MultiscalarMul` (/root/reference/src/range_proof.rs:207-213, src/traits.rs:7-43,
 * src/protocols/curve_point_protocol.rs:18-36) instantiated once for Ristretto (src/ristretto.rs:26-64).
 * Each entry point below names the reference call(s) it replaces.  INTEGRATION.md shows the Rust `extern "C"`
 * block a maintainer would add.
 *
 * Conventions: all buffers caller-owned host memory unless a name ends in `_dev`; little-endian; scalars are
 * 32-byte canonical encodings (values >= l are reduced where the reference would reduce, rejected where it
 * would reject); points are 32-byte Ristretto255 encodings (RFC 9496).  Every function returns a bpp_status.
 * No pointer is retained after return.  One bpp_ctx per (host thread, device); calls on one ctx are serialised
 * on its CUDA stream.  There is NO CPU fallback: without a usable CUDA device every call fails with
 * BPP_ERR_CUDA and bpp_last_error() says why.
 */
#ifndef BPP_B200_H
#define BPP_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes 0..5 mirror ProofError (/root/reference/src/errors.rs:12-28). */
typedef enum {
    BPP_OK = 0,
    BPP_VERIFICATION_FAILED = 1,
    BPP_INVALID_ARGUMENT = 2,
    BPP_INVALID_LENGTH = 3,
    BPP_INVALID_BLAKE2B = 4,
    BPP_SIZE_OVERFLOW = 5,
    BPP_ERR_CUDA = 100,      /* CUDA runtime / no device / launch failure */
    BPP_ERR_INTERNAL = 101
} bpp_status;

/* VerifyAction (/root/reference/src/range_proof.rs:47-55) */
typedef enum { BPP_RECOVER_ONLY = 0, BPP_RECOVER_AND_VERIFY = 1, BPP_VERIFY_ONLY = 2 } bpp_verify_action;

#define BPP_MAX_BATCH 256          /* MAX_RANGE_PROOF_BATCH_SIZE, range_proof.rs:76 */
#define BPP_MAX_BIT_LENGTH 64      /* MAX_RANGE_PROOF_BIT_LENGTH, range_proof.rs:71 */
#define BPP_MAX_EXT 6              /* ExtensionDegree::AddFiveBasePoints, generators/pedersen_gens.rs:40-66 */
#define BPP_TRANSCRIPT_BYTES 203   /* Merlin/STROBE-128 state: 200 B Keccak state, pos, pos_begin, cur_flags */

typedef struct bpp_ctx bpp_ctx;
typedef struct bpp_gens bpp_gens;
typedef struct bpp_vbatch bpp_vbatch;
typedef struct bpp_msm_plan bpp_msm_plan;

/* ---------------------------------------------------------------- context */
/* Many ctxs per device (lanes): their streams only overlap if they map to different hardware work queues; bpp_ctx_create sets
 * CUDA_DEVICE_MAX_CONNECTIONS=32 when it is unset, which takes effect if no CUDA context exists yet in the process (otherwise the
 * caller sets it before initialising CUDA; the driver default is 8). */
int32_t bpp_ctx_create(int32_t device_ordinal, bpp_ctx **out);
void bpp_ctx_destroy(bpp_ctx *ctx);
const char *bpp_last_error(const bpp_ctx *ctx);      /* valid until the next call on ctx */
int32_t bpp_ctx_sync(bpp_ctx *ctx);
/* number of engine kernels launched on this ctx since creation (bench.py's gpu_launches) */
uint64_t bpp_ctx_launch_count(const bpp_ctx *ctx);
void *bpp_ctx_stream(bpp_ctx *ctx);                   /* cudaStream_t, for event timing by the caller */
/* measurement aid: writes `bytes` (> L2 size) of scratch on the ctx stream before the next call */
int32_t bpp_ctx_l2_flush(bpp_ctx *ctx, size_t bytes);
/* CUDA-event timing on the ctx stream: start records an event; stop records one, waits for it, returns milliseconds */
int32_t bpp_ctx_timer_start(bpp_ctx *ctx);
int32_t bpp_ctx_timer_stop(bpp_ctx *ctx, float *ms);
/* per-phase device times of the last bpp_vbatch_run / bpp_msm_plan_run (events between the kernels, when enabled):
 * ms11 = {transcript replay, decompress, verifier prep per proof, per (proof, i), wait for host weight transcripts, weighting +
 *         column sums, MSM sort, MSM bucket sums, MSM window reduction, MSM Horner, encode/identity} */
int32_t bpp_ctx_phase_timing(bpp_ctx *ctx, int32_t enable);
int32_t bpp_ctx_phase_ms(bpp_ctx *ctx, float *ms11);
/* where loop 1 of the verifier (the per-proof Merlin transcript replay, range_proof.rs:816-850) runs: 1 = CUDA kernel (default),
 * 0 = host worker threads (BASELINE.json north_star's host/device split); 2 / 3 = CUDA kernel forced to one thread per proof /
 * one warp per proof (1 = one thread per proof).  Results are bit-identical. */
int32_t bpp_ctx_set_replay_mode(bpp_ctx *ctx, int32_t on_device);
/* verification passes are replayed as captured CUDA graphs (default on; BPP_NO_GRAPHS=1 or enable = 0 issues every kernel and
 * copy with its own driver call, which is also what bpp_ctx_phase_timing(ctx, 1) forces).  Results are identical. */
int32_t bpp_ctx_set_graphs(bpp_ctx *ctx, int32_t enable);
uint64_t bpp_ctx_graph_launch_count(const bpp_ctx *ctx);
/* test hooks (results must not change): bit 0 = every verification pass is repeated through the zero-weight fallback (the path taken
 * when a batch weight reduces to zero and Scalar::random_not_zero would draw again, range_proof.rs:894); bit 1 = a merged check
 * (bpp_ctx_set_merged_check) is always followed by the call-by-call pass */
int32_t bpp_ctx_set_test_hooks(bpp_ctx *ctx, uint32_t flags);
/* throughput mode (default 0), for callers that keep several verification calls in flight from several ctxs:
 *   1 = the calling thread sleeps between polls of its events (60 us naps, BPP_NAP_US) instead of spinning while the device works:
 *       with many ctxs per GPU spinning threads starve each other and the host hashing.  (cudaEventBlockingSync waits were tried
 *       first: they cost the process 0.24 ms of CPU in driver threads per pass and 0.4 ms of wake-up latency; sleeping polls cost
 *       0.03 ms and 0.1 ms.);
 *   2 = additionally the verifier-weight transcripts (range_proof.rs:811-853, :894) run on the device (k_weights, one warp per
 *       chunk) and the whole pass is ONE graph launch with no host step in the middle.  Measured on B200 this is slower (a
 *       256-proof chunk is a chain of ~330 dependent Keccak-f permutations: 4.5 ms on one warp against ~0.2 ms on a host core),
 *       so it is not what api.VerifierPool selects; it stays as a tested option for hosts with no cycles to spare.
 * Results are identical in every mode.  Mode 2 needs device replay and graphs (the defaults), else it behaves like mode 1. */
int32_t bpp_ctx_set_throughput_mode(bpp_ctx *ctx, int32_t enable);
/* wall-clock milliseconds of the host phases of the last bpp_vbatch_create / bpp_verify_chunks on this ctx:
 * ms6 = {parse + statement checks, layout + buffers, blob fill (+ loop-1 replay in host mode), weight transcripts (host mode),
 *        H2D + sync, unused} */
int32_t bpp_ctx_host_ms(bpp_ctx *ctx, double *ms6);
/* bytes moved host->device ([0]) and device->host ([1]) by the last bpp_vbatch_create + bpp_vbatch_run / bpp_verify_chunks */
int32_t bpp_ctx_io_bytes(bpp_ctx *ctx, uint64_t *h2d_d2h);
/* host threads used for the Fiat-Shamir replay of bpp_verify_chunks (default: min(64, hardware threads); the
 * reference is single-threaded, the harness supplies parallelism -- see BASELINE.md) */
int32_t bpp_ctx_set_host_threads(bpp_ctx *ctx, int32_t n);

/* ---------------------------------------------------------------- batched point primitives
 * replace CompressedRistretto::decompress / RistrettoPoint::compress / from_uniform_bytes as issued from
 * range_proof.rs:859-866,1067-1109 (decompress), :289,:348,:499-504,:587,:598-605 and range_statement.rs:62-65
 * (compress), ristretto.rs:48-52 <- generators_chain.rs:44-49 (one-way map). */
/* host only: Scalar::from_bytes_mod_order_wide on 64-bit limbs (the verifier-weight path); checked against bpp_scalar_from_wide */
void bpp_host_sc_from_wide64(const uint8_t in64[64], uint8_t out32[32]);
/* host only: a * b mod l on 64-bit limbs; a, b: any 256-bit values; output canonical (< l).  The prover's host-side scalar
 * bookkeeping. */
void bpp_host_sc_mul64(const uint8_t a32[32], const uint8_t b32[32], uint8_t out32[32]);
/* test hook, host only: the portable (non-MULX) body of the two functions above: b32_or_null != NULL -> a32 * b32 mod l,
 * else the wide reduction of 64 bytes */
void bpp_host_sc_generic64(const uint8_t *a32_or_wide64, const uint8_t *b32_or_null, uint8_t out32[32]);
/* test hook, host only: the verifier weights (range_proof.rs:811-853, :894) of n_chunks <= 8 chunks of `len` proofs each from the 32
 * bytes every proof feeds into the weight transcript; lockstep = 0: one transcript at a time, 1: n_chunks <= 4 through the four-way
 * vectorised sponge, 2: through the eight-way sponge (AVX-512; two four-way permutations elsewhere) */
/* test hook, host only: eight transcripts (203-byte states with equal position bytes) through the prover's lock-step sponge and through the
 * one-at-a-time Merlin: append_message("L", msg_j) -> challenge_bytes("e", 64) -> TranscriptRng keyed with witness_j and ext32_j -> two
 * fill_bytes(64).  Per lane 598 bytes out: [transcript | challenge | rng state | 128 rng bytes]; the two outputs must be equal. */
int32_t bpp_host_lockstep_selftest(const uint8_t *states203, const uint8_t *msgs, size_t msg_len, const uint8_t *witness, size_t wlen,
                                   const uint8_t *ext32, uint8_t *out_scalar, uint8_t *out_lockstep);
int32_t bpp_host_verifier_weights(const uint8_t *wbytes32, size_t len, size_t n_chunks, int32_t lockstep, uint8_t *weights32);
/* Host-side sum of n <= 64 points (32-byte encodings): the last step of a multi-GPU MSM, each GPU having reduced its shard to one
 * partial result.  BPP_INVALID_ARGUMENT if an encoding does not decode. */
int32_t bpp_points_sum_host(size_t n, const uint8_t *in32, uint8_t out32[32]);
/* ok[i] = 1 iff in32[i] is a canonical encoding; out32[i] = compress(decompress(in32[i])) (== in32[i] when ok) */
int32_t bpp_decompress_check(bpp_ctx *ctx, size_t n, const uint8_t *in32, uint8_t *ok, uint8_t *out32_or_null);
int32_t bpp_from_uniform_batch(bpp_ctx *ctx, size_t n, const uint8_t *in64, uint8_t *out32);

/* ---------------------------------------------------------------- multiscalar multiplication
 * replaces P::vartime_multiscalar_mul (range_proof.rs:482-495,512-521), P::multiscalar_mul
 * (generators/pedersen_gens.rs:117-120) and Precomputation::vartime_mixed_multiscalar_mul
 * (range_proof.rs:339-345,1050-1057).  k independent MSMs per call: segment s covers entries
 * [offsets[s], offsets[s+1]).  out32: k encodings.  A non-canonical point encoding => BPP_INVALID_ARGUMENT. */
int32_t bpp_msm(bpp_ctx *ctx, size_t n, const uint8_t *scalars32, const uint8_t *points32, uint8_t *out32);
int32_t bpp_msm_segmented(bpp_ctx *ctx, size_t k, const uint64_t *offsets, const uint8_t *scalars32,
                          const uint8_t *points32, uint8_t *out32);
/* device-resident variant for throughput measurement (BASELINE.json configs[4]): points are decompressed once */
int32_t bpp_msm_plan_create(bpp_ctx *ctx, size_t n, const uint8_t *points32, int32_t window_bits_or_0,
                            bpp_msm_plan **out);
int32_t bpp_msm_plan_set_scalars(bpp_msm_plan *plan, const uint8_t *scalars32);
int32_t bpp_msm_plan_run(bpp_msm_plan *plan, uint8_t *out32_or_null);   /* async unless out32 given */
int32_t bpp_msm_plan_window_bits(const bpp_msm_plan *plan);
int32_t bpp_msm_window_bits(size_t n_entries, size_t n_seg);   /* the window width chosen for n_seg sums over n_entries entries in all */
void bpp_msm_plan_destroy(bpp_msm_plan *plan);

/* ---------------------------------------------------------------- generators
 * replaces RangeParameters::init -> BulletproofGens::new (range_parameters.rs:32-58,
 * generators/bulletproof_gens.rs:83-112) + ristretto::create_pedersen_gens_with_extension_degree
 * (ristretto.rs:67-112).  SHAKE256 / SHA3-512 run on the host, the one-way map on the device; the tables stay
 * resident in HBM for the lifetime of the handle. */
int32_t bpp_gens_create(bpp_ctx *ctx, int32_t bit_length, int32_t max_aggregation, int32_t extension_degree,
                        bpp_gens **out);
/* the same with caller-made PedersenGens (RangeParameters::init takes any pc_gens: range_parameters.rs:32-58,
 * generators/pedersen_gens.rs:25-36): h_base32 = the value base, g_bases32 = extension_degree masking bases, as Ristretto encodings;
 * NULL = the reference's constants for that part.  BPP_INVALID_ARGUMENT if an encoding does not decode. */
int32_t bpp_gens_create_with_bases(bpp_ctx *ctx, int32_t bit_length, int32_t max_aggregation, int32_t extension_degree,
                                   const uint8_t *h_base32_or_null, const uint8_t *g_bases32_or_null, bpp_gens **out);
void bpp_gens_destroy(bpp_gens *g);
/* Fixed-base multiscalar multiplication over the generator set (the static half of Precomputation::vartime_mixed_multiscalar_mul,
 * generators/bulletproof_gens.rs:103, range_proof.rs:339-345): n_seg sums of seg_len terms, out32[s] = encode(sum_e scalars32[s][e] *
 * P[gidx[e]]), generator order Gi(n*M) | Hi(n*M) | G(ext) | H.  Window tables are built on first use (BPP_FB_MAX_MB, default 2048,
 * bounds them; BPP_SIZE_OVERFLOW if they cannot fit). */
int32_t bpp_gens_fixed_base_msm(bpp_gens *gens, size_t n_seg, size_t seg_len, const uint8_t *scalars32, const uint32_t *gidx, uint8_t *out32);
/* which: 0 = h_base, 1 = g_base[index], 2 = gi_base (flat, party-major), 3 = hi_base */
int32_t bpp_gens_get(const bpp_gens *g, int32_t which, size_t index, uint8_t out32[32]);
/* PedersenGens::commit for `count` openings: values[count], blindings32[count * n_blindings] (generators/pedersen_gens.rs:112-122).
 * Batches of >= 256 openings, and every batch once the fixed-base window tables of `g` exist, are summed from those tables; smaller first
 * batches go through the general MSM.  Same bytes either way. */
int32_t bpp_pedersen_commit_batch(bpp_gens *g, size_t count, const uint64_t *values, const uint8_t *blindings32,
                                  int32_t n_blindings, uint8_t *out32);

/* ---------------------------------------------------------------- batch verification
 * replaces RangeProof::verify_batch -> verify (range_proof.rs:712-1065), one call for K independent
 * reference calls ("chunks", each <= 256 proofs as verify_batch itself enforces at :739-751).
 * All proofs of the call share `gens` (bit length, extension degree, H, G, Gi, Hi): the reference's
 * consistency scan (:610-709) reduces to that identity.
 *
 * Layout: n proofs total; chunk c = proofs [chunk_offsets[c], chunk_offsets[c+1]).
 *   proof_bytes / proof_offsets[n+1]   serialised proofs (to_bytes layout, range_proof.rs:1120-1150)
 *   commit_offsets[n+1]                first commitment of proof i in commitments32 / min_values / min_present
 *   seed_nonces32 (n x 32) + seed_present (n)   optional (NULL = none)
 *   transcripts (n x 203 B)            Merlin state of each caller transcript BEFORE the call; bpp_verify_chunks advances
 *                                      them in place exactly as `&mut Transcript` is in the reference (the split form
 *                                      reads them in create and hands the advanced states out via bpp_vbatch_transcripts)
 * Results: chunk_status[K] (bpp_status per reference call), masks32 (n x ext x 32) and mask_present (n) as
 * Vec<Option<ExtendedMask>>.
 *
 * proof_bytes in PAGE-LOCKED host memory (bpp_host_alloc, cudaHostAlloc, cudaHostRegister) are not staged: the copy engine reads them
 * where they are, after the create / submit call has returned -- keep them valid and unchanged until the call's results are back
 * (bpp_verify_chunks returned, bpp_vbatch_run returned, bpp_vqueue_wait returned).  Pageable buffers are copied inside create. */
int32_t bpp_host_alloc(size_t bytes, void **out);
void bpp_host_free(void *p);
typedef struct {
    size_t n_proofs;
    size_t n_chunks;
    const uint64_t *chunk_offsets;
    const uint8_t *proof_bytes;
    const uint64_t *proof_offsets;
    const uint8_t *commitments32;
    const uint64_t *commit_offsets;
    const uint64_t *min_values;
    const uint8_t *min_present;
    const uint8_t *seed_nonces32;
    const uint8_t *seed_present;
    uint8_t *transcripts;
    int32_t action;
} bpp_verify_args;

int32_t bpp_verify_chunks(bpp_gens *g, const bpp_verify_args *args, int32_t *chunk_status,
                          uint8_t *masks32, uint8_t *mask_present);

/* Challenge-input form, for a host that keeps `merlin::Transcript` itself (its STROBE state is private, so a stock Rust host cannot
 * hand over the 203-byte state above): the caller runs loop 1 of RangeProof::verify (range_proof.rs:816-850) with the reference's
 * own src/transcripts.rs and draws the batch weights (:853, :894), then passes, per proof i,
 *   challenges32[challenge_offsets[i] .. challenge_offsets[i+1])  = y, z, e, e_0 .. e_{rounds-1}   (32-byte canonical, non-zero)
 *   weights32[i]                                                  = the proof's batch weight          (32-byte canonical, non-zero)
 * and this call does everything from :856 on (decompression, scalar synthesis, the merged multiscalar check, mask recovery).
 * args->transcripts is not read (may be NULL).  Identity-point / zero-challenge rejections of loop 1 (VerificationFailed) are the
 * caller's, as they happen inside its transcript code; the status precedence from :859 on is reproduced here. */
typedef struct {
    const uint8_t *challenges32;
    const uint64_t *challenge_offsets;     /* n_proofs + 1, in scalars */
    const uint8_t *weights32;              /* n_proofs x 32 */
} bpp_verify_challenges;
int32_t bpp_verify_chunks_ch(bpp_gens *g, const bpp_verify_args *args, const bpp_verify_challenges *ch, int32_t *chunk_status,
                             uint8_t *masks32, uint8_t *mask_present);

/* Split form used to time the device path with inputs resident in HBM (bench.py `value`):
 * create = host parsing + Fiat-Shamir + upload; run = all device work + verdict readback. */
int32_t bpp_vbatch_create(bpp_gens *g, const bpp_verify_args *args, bpp_vbatch **out);
/* ONE device pass over the calls of several callers (same action): proofs, chunks, statuses and masks of the pass are the calls'
 * concatenated in order.  This is what the coalescing queue below builds from the calls waiting in it. */
int32_t bpp_vbatch_create_multi(bpp_gens *g, size_t n_calls, const bpp_verify_args *const *calls, bpp_vbatch **out);
size_t bpp_vbatch_call_count(const bpp_vbatch *vb);
/* the advanced Merlin states of call `call` of the pass (that call's n_proofs x 203 B) */
int32_t bpp_vbatch_transcripts_call(const bpp_vbatch *vb, size_t call, uint8_t *transcripts);
int32_t bpp_vbatch_run(bpp_vbatch *vb, int32_t *chunk_status, uint8_t *masks32, uint8_t *mask_present);
/* the same with one output buffer set per call of the pass (masks32 / mask_present may be NULL, as may their entries) */
int32_t bpp_vbatch_run_multi(bpp_vbatch *vb, int32_t *const *chunk_status, uint8_t *const *masks32, uint8_t *const *mask_present);
/* after bpp_vbatch_run: write the advanced Merlin states (n_proofs x 203 B) -- what `&mut [Transcript]` holds after the
 * reference call; bpp_verify_chunks does this into args->transcripts itself */
int32_t bpp_vbatch_transcripts(const bpp_vbatch *vb, uint8_t *transcripts);
void bpp_vbatch_destroy(bpp_vbatch *vb);

/* ---------------------------------------------------------------- coalescing queue
 * One reference call (<= 256 proofs looked at) is far too little work for a B200.  Callers submit bpp_verify_args from any number of
 * threads; each of `lanes` lanes (a bpp_ctx + generator tables + host thread, all on `device_ordinal`) takes what is waiting -- up to
 * max_calls_per_pass calls with the same action -- and verifies it as ONE device pass, then writes every call's statuses, masks and
 * advanced transcripts to that call's own buffers.  Results are those of bpp_verify_chunks on each call alone.
 *   submit: returns at once; the buffers named by `args` and the output buffers must stay valid until wait(ticket) returns
 *   wait:   blocks until the call is done; returns BPP_OK when its chunk_status was written, else the engine error of its pass
 *   verify: submit + wait (a synchronous verify_batch that coalesces with the calls of other threads)
 *   stats:  {passes, calls, proofs, engine kernels launched, graph launches} since creation (read while idle) */
typedef struct bpp_vqueue bpp_vqueue;
int32_t bpp_vqueue_create(int32_t device_ordinal, int32_t bit_length, int32_t max_aggregation, int32_t extension_degree,
                          const uint8_t *h_base32_or_null, const uint8_t *g_bases32_or_null, int32_t lanes, int32_t max_calls_per_pass,
                          int32_t host_threads_per_lane, bpp_vqueue **out);
void bpp_vqueue_destroy(bpp_vqueue *q);
int32_t bpp_vqueue_submit(bpp_vqueue *q, const bpp_verify_args *args, int32_t *chunk_status, uint8_t *masks32, uint8_t *mask_present,
                          uint64_t *ticket);
int32_t bpp_vqueue_wait(bpp_vqueue *q, uint64_t ticket);
int32_t bpp_vqueue_verify(bpp_vqueue *q, const bpp_verify_args *args, int32_t *chunk_status, uint8_t *masks32, uint8_t *mask_present);
int32_t bpp_vqueue_stats(bpp_vqueue *q, uint64_t out5[5]);
int32_t bpp_vqueue_lanes(const bpp_vqueue *q);
/* wall time of the lane threads since creation, summed over lanes, in ms: {waiting for calls, building passes (parse + staging),
 * running them (device + weight hashing), handing results back}; read while idle */
int32_t bpp_vqueue_lane_ms(bpp_vqueue *q, double out4[4]);
/* where the verifier-weight transcripts (range_proof.rs:811-853, :894) of a pass are hashed: 0 = on the lane's host threads between two
 * graph launches (shortest pass), 1 = by k_weights on the device next to the scalar prep, the pass being ONE graph launch with no host
 * step in the middle (least host work per proof: what a host with few cores per GPU wants).  Call while the queue is idle. */
int32_t bpp_vqueue_set_device_weights(bpp_vqueue *q, int32_t enable);
/* Merged check (off by default): ONE multiscalar check for all reference calls of a device pass instead of one per call.  The reference
 * folds the <= 256 proofs of a call into one sum with random weights w_p drawn from the call's weight transcript (range_proof.rs:811-853,
 * :894, :1050-1057); with this switch the call sums S_c of a pass are folded once more, sum_c rho_c S_c, rho_c being the next value of
 * call c's own weight-transcript rng.  If the merged sum is the identity every call passed its check (a non-zero S_c survives with
 * probability 2^-252, the reference's own argument one level up); otherwise -- some proof of the pass is invalid -- the pass is settled
 * call by call with the same scalars, so every call's status is exactly what the reference returns for it.  It makes the pass ~15 %
 * cheaper (one large sum uses 14-bit windows, 64 small ones 9-bit) and a pass that holds an invalid proof ~40 % dearer.
 * bpp_ctx_set_merged_check does the same for bpp_verify_chunks / bpp_vbatch_* on a ctx (passes of four or more reference calls). */
int32_t bpp_vqueue_set_merged_check(bpp_vqueue *q, int32_t enable);
int32_t bpp_ctx_set_merged_check(bpp_ctx *ctx, int32_t enable);
uint64_t bpp_ctx_merged_fallbacks(const bpp_ctx *ctx);

/* ---------------------------------------------------------------- batched proving
 * replaces P calls of RangeProof::prove_with_rng (range_proof.rs:232-608) for statements of ONE shape (same bit length,
 * extension degree and aggregation factor m), advancing in lock-step: A, every L / R, the generator / scalar folding, A1 and B
 * run on the device; transcripts, nonces and random draws stay on the host in the reference's exact order.
 *   commitments32 (P x m x 32), values (P x m), blindings32 (P x m x ext x 32, canonical), min_values / min_present (P x m),
 *   seed_nonces32 / seed_present (P; a seed needs m == 1), transcripts (P x 203 B, advanced in place),
 *   rng_bytes (P x rng_stride): the bytes the caller's external RNG yields, 32 per TranscriptRng rebuild
 *   (transcripts.rs:185-194), log2(n*m) + 3 rebuilds per proof; with a seeded RNG the proofs are byte-identical to the
 *   reference's.  proofs_out: P x proof_stride (>= bpp_proof_size), status: ProofError code per proof (0 = proof written). */
typedef struct {
    size_t n_proofs;
    int32_t aggregation;
    const uint8_t *commitments32;
    const uint64_t *values;
    const uint8_t *blindings32;
    const uint64_t *min_values;
    const uint8_t *min_present;
    const uint8_t *seed_nonces32;
    const uint8_t *seed_present;
    uint8_t *transcripts;
    const uint8_t *rng_bytes;
    size_t rng_stride;
} bpp_prove_args;
size_t bpp_proof_size(int32_t extension_degree, int32_t rounds);
int32_t bpp_prove_batch(bpp_gens *g, const bpp_prove_args *args, uint8_t *proofs_out, size_t proof_stride, int32_t *status);

/* ---------------------------------------------------------------- proof bytes (host)
 * RangeProof::from_bytes validation (range_proof.rs:1155-1257): returns BPP_OK and the number of (L,R) rounds. */
int32_t bpp_proof_check_bytes(const uint8_t *bytes, size_t len, int32_t *extension_degree, int32_t *rounds);

/* ---------------------------------------------------------------- Merlin transcripts (host)
 * merlin::Transcript::new / append_message / challenge_bytes on the 203-byte state used above. */
void bpp_transcript_new(const uint8_t *label, size_t len, uint8_t out[BPP_TRANSCRIPT_BYTES]);
void bpp_transcript_append_message(uint8_t t[BPP_TRANSCRIPT_BYTES], const uint8_t *label, size_t label_len,
                                   const uint8_t *msg, size_t len);
void bpp_transcript_challenge_bytes(uint8_t t[BPP_TRANSCRIPT_BYTES], const uint8_t *label, size_t label_len,
                                    uint8_t *out, size_t len);

/* host hash layer, exposed so it can be pinned against external KATs (python hashlib) without a GPU:
 * SHA3-512 (ristretto.rs:92-95), SHAKE256 (generators_chain.rs:23-49), BLAKE2b-512 keyed+personalised with an empty
 * message (utils/generic.rs:56-57), Scalar::from_bytes_mod_order_wide (transcript_protocol.rs:70) */
/* host only: Keccak-f[1600] on one 25-lane state -- what every host-side sponge of the library permutes through (vector registers when
 * the CPU has AVX2 / AVX-512VL), and the plain 64-bit body for comparison (test hook) */
void bpp_keccak_f1600_x1(uint64_t *state25);
void bpp_keccak_f1600_x1_generic(uint64_t *state25);
/* host only: which body the host-side hashing and scalar arithmetic run through: 2 = AVX-512VL, 1 = AVX2, 0 = baseline ISA (no MULX either);
 * the environment variable BPP_HOST_SIMD caps it (the tests run every body on one CPU) */
int32_t bpp_host_simd_level(void);
void bpp_hash_sha3_512(const uint8_t *in, size_t len, uint8_t out[64]);
void bpp_hash_shake256(const uint8_t *in, size_t len, uint8_t *out, size_t outlen);
int32_t bpp_hash_blake2b_nonce_bytes(const uint8_t *key, size_t keylen, const uint8_t *personal, size_t plen, uint8_t out[64]);
void bpp_scalar_from_wide(const uint8_t in64[64], uint8_t out32[32]);

/* ---------------------------------------------------------------- measurement
 * Integer-pipe microbenchmarks: which = 0 IMAD.lo, 1 IMAD.HI, 2 IMAD.WIDE, 3 IADD3, 4 field-mul,
 * 5 field-square, 6 point madd, 7 point dbl, 8 scalar montmul.  Returns operations per second (per lane). */
int32_t bpp_microbench(bpp_ctx *ctx, int32_t which, int32_t iters, double *ops_per_sec, double *seconds);

#ifdef __cplusplus
}
#endif
#endif
