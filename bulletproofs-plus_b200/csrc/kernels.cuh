// Launch-side declarations of the bpp-b200 device kernels.  Definitions live in k_*.cu; engine.cu drives them.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "arith.cuh"

namespace bpp {

// ---------------------------------------------------------------- k_point.cu
// in: n x 8 words (Ristretto encodings).  out_tab (optional): affine-Niels table entry per point (identity when
// invalid); ok: 1/0; out_enc (optional): re-encoding of the decoded point (8 words); bad_count (optional): += #invalid.
void launch_decompress(cudaStream_t s, size_t n, const uint32_t *in, aniels *out_tab, uint8_t *ok, uint32_t *out_enc,
                       uint32_t *bad_count);
// the points of a verification pass straight from the uploaded proof bytes and commitments: point i belongs to the proof p with
// pt_offsets[p] <= i < pt_offsets[p + 1], slot i - pt_offsets[p] of [A, A1, B, L_0.., R_0.., V_0..] (kernels.cuh VProof)
struct VProof;
void launch_decompress_proofs(cudaStream_t s, uint32_t n_pts, uint32_t n_proofs, uint32_t ext, const VProof *proofs, const uint32_t *pt_offsets,
                              const uint8_t *blob, const uint8_t *commitments32, aniels *out_tab, uint8_t *ok);
// extended points -> encodings (8 words each) and/or identity flags
void launch_encode(cudaStream_t s, size_t n, const ge *in, uint32_t *out_enc, uint8_t *is_identity);
// 64-byte uniform strings -> points: encodings and/or affine-Niels table entries
void launch_from_uniform(cudaStream_t s, size_t n, const uint32_t *in16, uint32_t *out_enc, aniels *out_tab);

// ---------------------------------------------------------------- k_msm.cu
struct MsmShape {
    uint32_t n_entries;   // (scalar, point) pairs over all segments
    uint32_t n_seg;       // independent MSMs
    int c;                // window bits
    int W;                // windows = ceil(252 / c)
    uint32_t B;           // buckets per window = 2^(c-1)
    uint32_t max_seg_entries;   // largest segment, when the caller knows it (0 = unknown): small segments sort in shared memory
};
MsmShape msm_shape(uint32_t n_entries, uint32_t n_seg, int forced_c);
// bytes of scratch needed for a shape
size_t msm_scratch_bytes(const MsmShape &sh);
// experiment / test switches of launch_msm (BPP_MSM_BUCKET, BPP_MSM_SPLIT, BPP_MSM_REDUCE, BPP_MSM_REDUCE_PARTS), read from the environment
// ONCE per process; callers that cache launch sequences (CUDA graphs) key them on these values
void msm_knobs(int32_t out[4]);
// scalars: n_entries x 8 words, canonical.  seg_offsets: n_seg + 1 entry offsets (nullptr when n_seg == 1).
// pidx: per-entry index into `gens` (bit 31 set), `dync` (bit 30 set: projective "cached" points) or `dyn`; nullptr = identity
// mapping into dyn.
// result: n_seg extended points.
void launch_msm(cudaStream_t s, const MsmShape &sh, const uint32_t *scalars, const uint32_t *seg_offsets, const uint32_t *pidx,
                const aniels *dyn, const aniels *gens, void *scratch, ge *result, uint64_t *launches, cudaEvent_t *marks = nullptr,
                const cached *dync = nullptr);
// marks (optional, 4 events): recorded after the sort phase (digits+scan+scatter), bucket sums, window reduction, Horner

// ---------------------------------------------------------------- k_fb.cu
// fixed-base window tables over a generator set: table[g][w][d-1] = d * 2^(c*w) * P_g (affine Niels), W = ceil(252/c), B = 2^(c-1)
struct FbShape { uint32_t n_gens; int c, W; uint32_t B; };
FbShape fb_shape(uint32_t n_gens, int forced_c, size_t max_bytes);
size_t fb_table_bytes(const FbShape &sh);
// builds the table (synchronises the stream; temporary device scratch is allocated and freed inside); 0 = ok
int fb_build(cudaStream_t s, const FbShape &sh, const aniels *gens, aniels *tab, uint64_t *launches);
// n_seg independent sums of seg_len terms each: scalars canonical, segment-major; gidx[kinds][seg_len] generator index per entry
// (segment s uses row s % kinds); out: n_seg extended points
void launch_fb_msm(cudaStream_t s, const FbShape &sh, uint32_t n_seg, uint32_t seg_len, uint32_t kinds, const uint32_t *scalars, const uint32_t *gidx,
                   const aniels *tab, ge *out, uint64_t *launches);

// ---------------------------------------------------------------- k_verify.cu
#define BPP_TSTATE_BYTES 203 // merlin STROBE-128 state on the wire: 200 B Keccak state, pos, pos_begin, cur_flags
#define BPP_MAX_ROUNDS 24  // log2(n * m) <= 24 (generator sets are capped at 2^24 points)
struct VProof {            // per-proof metadata, device-resident
    uint32_t m;            // aggregation factor (commitments)
    uint32_t rounds;       // log2(n * m)
    uint32_t raw_off;      // byte offset of the serialised proof (to_bytes layout, range_proof.rs:1120-1150) from the start of the blob
    uint32_t ch_off;       // first challenge: [y, z, e, e_0..e_{r-1}]
    uint32_t entry_off;    // first dynamic MSM entry of this proof: [A1, B, A, L.., R.., V..]
    uint32_t commit_off;   // first commitment (min_values / min_present index)
    uint32_t nonce_off;    // first nonce: [eta_k, d_k, alpha_k (ext each), dL_jk, dR_jk (rounds*ext each, j-major)] or 0xffffffff
    uint32_t contrib_off;  // first slot in the gi/hi contribution array: [gi(N) | hi(N)]
    uint32_t pv_off;       // first slot of the stage A -> B hand-off vector (8 + 3*rounds + m scalars)
    uint32_t active;       // 1 = contributes to its chunk's MSM
    uint32_t pt_off;       // first slot in the point table / encoding array: [A, A1, B, L_0.., R_0.., V_0..]
    uint32_t replay;       // 1 = its transcript is replayed (loop 1) and its scalars are prepared
    uint32_t ts_idx;       // which uploaded transcript state it starts from (calls whose transcripts are all equal upload ONE)
    uint32_t chunk;        // the reference call (chunk) the proof belongs to
};
struct VChunk {
    uint32_t proof_lo, proof_hi;
    uint32_t max_mn;       // largest n*m in the chunk
    uint32_t entry_off;    // first MSM entry of the chunk: [Gi(max_mn) | Hi(max_mn) | G(ext) | H | dynamic...]
    uint32_t active;
};
// gens_nm = bit_length * max_aggregation; merged = 1: ONE multiscalar check for all chunks of the pass (engine_verify.cu, "merged check"):
// chunk c's terms carry an extra factor rho_c, the wide value at weights[16 * (n_proofs + c)], drawn from the chunk's weight transcript
struct VDims { uint32_t n_proofs, n_chunks, bit_length, ext; int action; uint32_t gens_nm; uint32_t merged; };
// field offsets inside a serialised proof: [ext:u8] d1[ext] a a1 b r1 s1 (L_j R_j)*
#define BPP_RAW_D1(ext, k) (1u + 32u * (k))
#define BPP_RAW_A(ext) (1u + 32u * (ext))
#define BPP_RAW_R1(ext) (1u + 32u * ((ext) + 3u))
#define BPP_RAW_S1(ext) (1u + 32u * ((ext) + 4u))
#define BPP_RAW_L(ext, j) (1u + 32u * ((ext) + 5u) + 64u * (j))
#define BPP_RAW_R(ext, j) (1u + 32u * ((ext) + 6u) + 64u * (j))
struct VBuffers {
    const VProof *proofs; const VChunk *chunks;
    const uint32_t *pt_offsets;      // n_proofs + 1: first slot of every proof in the point table (prefix sums; proofs without device work are empty)
    const uint8_t *blob;             // uploaded bytes; the proof scalars r1, s1, d1 are read from VProof::raw_off
    const uint32_t *challenges;      // words
    const uint32_t *weights;         // n_proofs x 16 words: the 512-bit value whose reduction mod l is the batch weight; consumed by k_vprep_weight only
    uint8_t *weight_zero;            // out: n_chunks flags, set when a weight of the chunk reduced to zero
    uint32_t *weights_mont;          // scratch: n_proofs x 8 words, Montgomery form
    const uint64_t *min_values; const uint8_t *min_present;
    const uint32_t *nonces;          // words, may be null
    uint32_t *msm_scalars;           // out: n_entries x 8 words
    uint32_t *msm_pidx;              // out: n_entries point indices (bit 31 = generator table), written by k_vprep_weight / k_vprep_reduce
    uint32_t *contrib;               // scratch: gi/hi contributions (Montgomery form)
    uint32_t *hg_contrib;            // scratch: per proof (1 + ext) scalars (Montgomery form): h, g_k
    uint32_t *pervec;                // scratch: stage A -> B hand-off
    uint32_t *masks;                 // out: n_proofs x ext x 8 words (plain), may be null
};
// weight-free part (per proof, per (proof, i)); marks (optional, 2 events): after each of the two kernels
void launch_verify_prep(cudaStream_t s, const VDims &d, const VBuffers &b, uint32_t total_vec, uint32_t max_rounds, uint64_t *launches,
                        cudaEvent_t *marks = nullptr);
// weight application + column sums into the MSM entry lists (needs b.weights)
void launch_verify_weigh(cudaStream_t s, const VDims &d, const VBuffers &b, uint32_t max_static, uint64_t *launches);

// ---------------------------------------------------------------- k_replay.cu
struct RBuffers {
    const VProof *proofs;
    const uint8_t *tstates_in;       // 203-byte initial states, indexed by VProof::ts_idx
    const uint8_t *hg32;             // compressed H, then G[0..ext)
    const uint8_t *blob;             // uploaded bytes: serialised proofs at VProof::raw_off
    const uint8_t *commitments32;    // 32 x (VProof::commit_off + j)
    const uint64_t *min_values; const uint8_t *min_present;
    uint8_t *challenges;             // out (VProof::ch_off): y, z, e, e_0..
    uint8_t *wbytes;                 // out: n_proofs x 32
    uint8_t *tstates_out;            // out: n_proofs x 203
    uint8_t *flags;                  // out: n_proofs; bit 0 = loop-1 VerificationFailed, bit 1 = y == 1
};
// kernel: 0 = one thread per proof, sponge state in shared memory and the permutation in registers (default); 1 = one thread per
// proof, state in local memory (the round-1 kernel); 2 = one warp per proof (wstrobe.cuh)
void launch_replay(cudaStream_t s, const VDims &d, const RBuffers &b, int kernel, uint64_t *launches);
// verifier weights of every active chunk (one warp per chunk); wt_init = 203-byte state of the weight transcript after its
// domain separator; weights: n_proofs x 16 words (canonical weight, upper half zero)
void launch_weights(cudaStream_t s, const VDims &d, const VChunk *chunks, const uint8_t *wt_init, const uint8_t *wbytes, const uint8_t *flags,
                    uint32_t *weights, uint64_t *launches);

// ---------------------------------------------------------------- k_prove.cu
struct PDims { uint32_t P, n, m, N, ext, rounds, gens_nm; };   // P proofs of one shape; N = n*m; gens_nm = n * max_aggregation
struct PBuffers {
    const uint64_t *offset_values;   // P x m: value - minimum_value_promise
    uint32_t *a, *b;                 // P x N scalars, Montgomery form (a_L / a_R, folded in place)
    uint32_t *ypow;                  // P x (N + 2): y^0 .. y^(N+1), Montgomery
    uint32_t *yinv2;                 // P x BPP_MAX_ROUNDS: y^-(2^k), Montgomery
    const uint32_t *yz;              // P x 2 canonical: y, z; then P canonical: y^-1
    const uint32_t *dlr;             // P x 2 x ext canonical: d_L[k], d_R[k] of the current round
    const uint32_t *e;               // 2P canonical: round challenges e[0..P), their inverses e^-1[P..2P)
    uint32_t *fsc;                   // P x 6: fold scalars (see k_prove_round_inv)
    cached *folded;                  // [Gi: P x N | Hi: P x N] folded generator vectors (valid from round 1 on; folding path)
    uint32_t *sg, *sh;               // P x N scalars, Montgomery: the folding as coefficients over the original generators (fixed-base path)
    uint32_t *msm_scalars;           // entry scalars of the MSM being assembled (canonical)
    uint32_t *msm_pidx;              // entry point indices
};
void launch_prove_bits(cudaStream_t s, const PDims &d, const PBuffers &b);
void launch_prove_init(cudaStream_t s, const PDims &d, const PBuffers &b);
void launch_prove_round_pre(cudaStream_t s, const PDims &d, const PBuffers &b, uint32_t nn, uint32_t round);
void launch_prove_fold(cudaStream_t s, const PDims &d, const PBuffers &b, uint32_t nn, uint32_t round, const aniels *gens);
void launch_prove_final_ab(cudaStream_t s, const PDims &d, const PBuffers &b, uint32_t *out);
// fixed-base path (no generator folding; see k_prove.cu)
void launch_prove_bits_fb(cudaStream_t s, const PDims &d, const PBuffers &b);
void launch_prove_round_pre_fb(cudaStream_t s, const PDims &d, const PBuffers &b, uint32_t nn);
void launch_prove_fold_fb(cudaStream_t s, const PDims &d, const PBuffers &b, uint32_t nn);      // e^-1, sG / sH update, a / b fold: 2 kernels
void launch_prove_final_fb(cudaStream_t s, const PDims &d, const PBuffers &b, const uint32_t *rs);

// ---------------------------------------------------------------- k_bench.cu
// returns elapsed seconds for `iters` dependent ops in each of `threads_total` lanes; ops counted by caller
int microbench_run(cudaStream_t s, int which, int iters, double *ops_per_sec, double *seconds, uint64_t *launches);

} // namespace bpp
