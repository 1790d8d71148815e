// Batch verification entry points: bpp_vbatch_create[_multi] / bpp_vbatch_run / bpp_vbatch_transcripts / bpp_verify_chunks[_ch].
//
// Restates the control flow of RangeProof::verify_batch -> verify (/root/reference/src/range_proof.rs:712-1065):
// argument checks (:719-734), first-256 truncation (:739-751), consistency (:610-709), loop 1 = Fiat-Shamir replay of every
// proof's transcript (:816-850), the sequential verifier-weight transcript (:811, :849-853, :894), loop 2 (:856-1033) and the
// single merged multiscalar check (:1039-1062).  One PASS verifies any number of reference calls ("chunks") of any number of
// callers (bpp_vbatch_create_multi: the coalescing front end engine_queue.cpp merges queued calls into one pass):
//
//   host    :  header parse + statement checks -> memcpy of the callers' raw byte arrays into ONE pinned blob -> one H2D copy
//   stream A:  K-REPLAY -> D2H(wbytes, flags, transcripts) | K-VPREP proof, vector (weight-free) .... K-VPREP weight, reduce -> K-MSM -> verdicts
//   stream B:  K-DECOMPRESS (straight from the proof bytes) ..........................................^ (joins before the bucket sums)
//   host    :  ............ verifier-weight transcripts per chunk (4 chunks in lock-step) -> weights H2D ^
//
// The device reads proofs, commitments and transcripts where the callers' buffers put them (rawld.cuh): the host neither gathers
// points nor builds index lists (round 1 did both: 0.12 ms of host time per 1024 proofs, the limiter of the 8-GPU runs).
// Loop 1 runs on the device by default (k_replay.cu); `bpp_ctx_set_replay_mode(ctx, 0)` keeps it on host threads (BASELINE.json
// north_star's split) and bpp_verify_chunks_ch takes the challenges and weights from a caller that keeps merlin itself; all
// three produce bit-identical results.  Error precedence of the reference is reproduced when the per-chunk status is resolved
// after the device returns.  The pass is replayed as three captured CUDA graphs (see "CUDA graphs" below).
#include <algorithm>
#include <array>
#include <chrono>
#include <cstring>
#include <ctime>
#include <cstdio>
#include "engine.hpp"
#include "hash.cuh"
#include "replay.cuh"

using namespace bpp;

namespace {

struct HProof {
    int32_t pre_rc = 0;        // from_bytes / RangeStatement::init class errors
    int32_t loop2_rc = 0;      // host-known loop-2 error: InvalidLength / SizeOverflow when 2^rounds != n*m (:875-888)
    int ext = 0, rounds = 0;
    uint32_t m = 0;
    const uint8_t *bytes = nullptr;   // serialised proof in the caller's buffer (valid during create only)
    bool has_seed = false, looked = false;
    uint32_t pt_off = 0, n_pts = 0;   // slots in the point table: [A, A1, B, L.., R.., V..]
    uint32_t call = 0;                // which call of the pass
    size_t local = 0;                 // index inside its call
};

struct HChunk {
    size_t lo = 0, hi = 0, end = 0;   // proofs looked at: [lo, hi) (hi - lo <= 256); the caller's chunk is [lo, end)
    int32_t pre_rc = 0;        // empty batch / from_bytes / statement / consistency errors
    bool computable = false;   // no host-known error: device prep + MSM (or mask recovery) runs
    uint32_t max_mn = 0;
    uint32_t entry_off = 0, n_entries = 0;
};

// one caller's bpp_verify_args inside a pass
struct HCall {
    bpp_verify_args a;                // copy of the struct; its pointers are only dereferenced inside create
    const bpp_verify_challenges *ch = nullptr;
    size_t proof0 = 0, chunk0 = 0, commit0 = 0, raw0 = 0;      // first proof / chunk / commitment / raw byte of this call in the pass
    size_t raw_base = 0;              // a.proof_offsets[0]
    bool raw_pinned = false;          // the call's proof bytes lie in page-locked memory: the copy engine reads them where they are
    bool same_transcripts = false;    // every transcript of the call holds the same state (one label for all proofs: the usual case)
    size_t ts0 = 0;                   // first uploaded transcript state of this call
};

inline bool is_zero32(const uint8_t *p) { return replay_is_zero32(p); }
// [p, p + n) inside page-locked host memory (cudaHostAlloc / cudaHostRegister / bpp_host_alloc)?  Called on the thread that owns the ctx.
bool host_range_pinned(const void *p, size_t n) {
    static const bool off = [] { const char *e = getenv("BPP_NO_DIRECT_DMA"); return e && atoi(e) != 0; }();
    if (off || !p || !n) return false;
    cudaPointerAttributes at;
    for (const uint8_t *q : {(const uint8_t *)p, (const uint8_t *)p + (n - 1)}) {
        if (cudaPointerGetAttributes(&at, q) != cudaSuccess) { cudaGetLastError(); return false; }
        if (at.type != cudaMemoryTypeHost) return false;
    }
    return true;
}
#define LBL(s) BPP_LBL(s)

// utils/generic.rs:30-60
void nonce(const uint8_t seed[32], const char *label, bool have_j, uint32_t j, bool have_k, uint32_t k, uint8_t out32[32]) {
    uint8_t key[43];
    size_t kl = 0;
    key[kl++] = 0;
    memcpy(key + kl, seed, 32); kl += 32;
    if (have_j) { key[kl++] = 'j'; for (int i = 0; i < 4; i++) key[kl++] = (uint8_t)(j >> (8 * i)); }
    if (have_k) { key[kl++] = 'k'; for (int i = 0; i < 4; i++) key[kl++] = (uint8_t)(k >> (8 * i)); }
    uint8_t h[64];
    blake2b_keyed_personal_empty(h, key, kl, (const uint8_t *)label, strlen(label));
    host_sc_from_wide(h, out32);
}

} // namespace

// device + pinned buffers of one verification pass; pooled per ctx so that repeated calls do not pay cudaMalloc /
// cudaMallocHost / cudaFree every time (grow-only, returned to the pool by bpp_vbatch_destroy)
struct VWork {
    DevBuf d_blob, d_tab, d_ok, d_mscal, d_pidx, d_chal, d_contrib, d_hg, d_pervec, d_masks, d_scratch, d_res, d_ident, d_weights, d_wmont, d_mid;
    PinBuf h_blob, h_out, h_mid, h_weights, h_chal;
    void release() {
        for (DevBuf *b : {&d_blob, &d_tab, &d_ok, &d_mscal, &d_pidx, &d_chal, &d_contrib, &d_hg, &d_pervec, &d_masks, &d_scratch, &d_res, &d_ident,
                          &d_weights, &d_wmont, &d_mid})
            b->release();
        h_blob.release(); h_out.release(); h_mid.release(); h_weights.release(); h_chal.release();
    }
};

struct bpp_vbatch {
    bpp_gens *g = nullptr;
    int32_t action = BPP_VERIFY_ONLY;
    bool device_replay = true;
    bool caller_challenges = false;      // bpp_verify_chunks_ch: loop 1 and the weights were made by the caller
    size_t n_proofs = 0, n_chunks = 0, n_commit = 0;
    std::vector<HProof> hp;
    std::vector<HChunk> hc;
    std::vector<HCall> calls;
    uint32_t n_pts = 0, n_entries = 0, n_chal = 0, max_static = 0, max_rounds = 0;
    bool any_msm = false, any_masks = false, any_replay = false, any_vec = false;
    bool ran = false;
    bool merged = false;                 // ONE multiscalar check for all chunks of the pass (segmented check as the fall-back), see bpp_vbatch_run_multi
    MsmShape shape;
    MsmShape shape_m;                    // the merged check: all entries as one sum
    VWork *w = nullptr;
    // all inputs travel as ONE pinned blob -> ONE H2D copy; these are the section offsets inside it
    size_t o_proofs = 0, o_chunks = 0, o_ptoff = 0, o_segoff = 0, o_hg = 0, o_wtinit = 0, o_minv = 0, o_minp = 0, o_commit = 0, o_raw = 0,
           o_tstate = 0, o_nonces = 0, blob_bytes = 0;
    // mid-pipeline results of loop 1: [wbytes n x 32 | flags n | tstates n x 203]; same layout on device and host
    size_t mo_wbytes = 0, mo_flags = 0, mo_tstate = 0, mid_bytes = 0;
    size_t ho_ok = 0, ho_ident = 0, ho_masks = 0, hout_bytes = 0;
    template <class T> T *dev(size_t off) const { return reinterpret_cast<T *>(w->d_blob.as<uint8_t>() + off); }
    uint8_t *mid() const { return w->h_mid.as<uint8_t>(); }
    uint8_t *wbytes(size_t i) const { return mid() + mo_wbytes + 32 * i; }
    uint8_t &flag(size_t i) const { return mid()[mo_flags + i]; }
    uint8_t *tstate(size_t i) const { return mid() + mo_tstate + BPP_TRANSCRIPT_BYTES * i; }
};

namespace bpp {
void vwork_pool_free(bpp_ctx *ctx) {
    for (void *p : ctx->vwork_pool) { VWork *w = (VWork *)p; w->release(); delete w; }
    ctx->vwork_pool.clear();
}
}
static VWork *vwork_acquire(bpp_ctx *ctx) {
    if (!ctx->vwork_pool.empty()) { VWork *w = (VWork *)ctx->vwork_pool.back(); ctx->vwork_pool.pop_back(); return w; }
    return new VWork();
}
static void vwork_return(bpp_ctx *ctx, VWork *w) {
    if (ctx->vwork_pool.size() < 4) ctx->vwork_pool.push_back(w);
    else { w->release(); delete w; }
}

// ------------------------------------------------------------------------------------------------ CUDA graphs
// One verification pass is ~20 kernels, 2 memsets and 5 copies.  Issued one by one that is ~35 driver calls per pass, and with
// several lanes (one bpp_ctx + host thread each) verifying concurrently the driver's submission path, not the GPU, capped a
// B200 at ~6.5 k passes/s whatever their size (measured in round 1: 1.7 M proofs/s with 256-proof passes, 4.4 M with 1024).
// The pass is therefore captured once per (workspace, layout) as three graphs -- A: transcript replay + D2H of its results,
// B: point decompression || weight-free scalar prep, C: weights H2D, weighting, MSM, verdict D2H -- split where the host hashes
// the verifier-weight transcript, and replayed with three cudaGraphLaunch calls afterwards.
struct VGraphKey {
    const void *bufs[21];
    const void *gens_table;
    size_t n_proofs, n_chunks, n_commit;
    size_t off[21];
    uint32_t n_pts, n_entries, n_chal, max_static, max_rounds;
    int32_t action, ext, bit_length;
    MsmShape shape;
    int32_t msm_knobs[4];
    uint8_t any_msm, any_masks, any_replay, any_vec, device_replay, replay_kernel, fused, caller_challenges;
};
struct VGraph {
    VGraphKey key;
    cudaGraphExec_t ex[3] = {nullptr, nullptr, nullptr};
    uint64_t kernels[3] = {0, 0, 0};
    uint64_t last_use = 0;
};
static void vgraph_free(VGraph *g) {
    for (auto &e : g->ex) if (e) cudaGraphExecDestroy(e);
    delete g;
}
namespace bpp {
void vgraph_cache_free(bpp_ctx *ctx) {
    for (void *p : ctx->vgraphs) vgraph_free((VGraph *)p);
    ctx->vgraphs.clear();
}
}

// ------------------------------------------------------------------------------------------------ verifier weights (host)
// LANES (4 or 8) STROBE-128 sponges advancing in lock-step (host_keccak4.cpp permutes the states with one vectorised Keccak-f): the
// weight transcripts of chunks that hold the same number of proofs perform the same operations at the same sponge positions, only
// the absorbed bytes differ.  Mirrors Strobe128 / Merlin / MerlinRng of hash.cuh operation by operation.
// The 64 bytes drawn per weight (Scalar::random, range_proof.rs:894) are handed out as they are: the wide reduction mod l runs on the
// device (k_vprep_weight), which also reports the zero weight random_not_zero would redraw (probability 2^-252; handled by rerunning
// the chunk through weights_scalar, see bpp_vbatch_run).
#include "strobe_n.hpp"
extern "C" int32_t bpp_host_simd_level(void);
extern "C" void bpp_host_sc_from_wide64(const uint8_t in64[64], uint8_t out32[32]);      // host_keccak4.cpp: 64-bit-limb wide reduction
namespace {
const std::array<uint8_t, BPP_TRANSCRIPT_BYTES> &weight_transcript_init() {      // Transcript::new("Bulletproofs+ verifier weights") (:811)
    static const std::array<uint8_t, BPP_TRANSCRIPT_BYTES> wt0 = [] {
        std::array<uint8_t, BPP_TRANSCRIPT_BYTES> st;
        Merlin wt;
        wt.init(LBL("Bulletproofs+ verifier weights"));
        wt.s.store(st.data());
        return st;
    }();
    return wt0;
}
} // namespace

// Sequential part of loop 1 + the weight draw of loop 2 for every chunk: the verifier-weight transcript (range_proof.rs:811,
// :849, :853) and random_not_zero per proof (:894).  Needs wbytes / flags of all proofs.  Chunks are independent: one at a time
// through the scalar sponge, or LANES chunks of equal length at a time through StrobeN.
// wb: len x 32 bytes (what every proof of the chunk feeds into the transcript, in proof order); out: len x 64, the weight of proof k as a
// 512-bit little-endian value whose reduction mod l is the weight.  weights_scalar is the reference statement by statement (it
// redraws a zero weight, so what it writes is the final weight, canonical, upper half zero).
// rho (may be null): 64 bytes for the chunk's factor of the merged check, the next value of the same rng (canonical, upper half zero)
static void weights_scalar(const uint8_t *wb, size_t len, uint8_t *out, uint8_t *rho = nullptr) {
    Merlin wt;
    wt.s.load(weight_transcript_init().data());
    for (size_t k = 0; k < len; k++) wt.append_message(LBL("proof"), wb + 32 * k, 32);   // :849
    MerlinRng wr;
    const uint8_t zeros[32] = {0};
    wr.build(wt, nullptr, 0, false, zeros);                                           // :853
    for (size_t k = 0; k < len; k++) {
        uint8_t wide[64], *wgt = out + 64 * k;
        do { wr.fill(wide, 64); bpp_host_sc_from_wide64(wide, wgt); } while (is_zero32(wgt));   // :894 random_not_zero
        memset(wgt + 32, 0, 32);
    }
    if (rho) {
        uint8_t wide[64];
        do { wr.fill(wide, 64); bpp_host_sc_from_wide64(wide, rho); } while (is_zero32(rho));
        memset(rho + 32, 0, 32);
    }
}
// LANES chunks with the same number of proofs (pointers may repeat: padding of an incomplete group)
template <int LANES> static void weights_xn(const uint8_t *const wb[LANES], size_t len, uint8_t *const out[LANES], uint8_t *const rho[LANES] = nullptr) {
    StrobeN<LANES> s;
    s.load_all(weight_transcript_init().data());
    const uint8_t l32[4] = {32, 0, 0, 0}, l64[4] = {64, 0, 0, 0};
    for (size_t k = 0; k < len; k++) {                                                // append_message("proof", wbytes, 32)
        const uint8_t *d[LANES];
        for (int j = 0; j < LANES; j++) d[j] = wb[j] + 32 * k;
        s.meta_ad_same(LBL("proof"), false);
        s.meta_ad_same(l32, 4, true);
        s.ad_each(d, 32);
    }
    const uint8_t zeros[32] = {0};
    s.meta_ad_same(LBL("rng"), false);                                                // build_rng().finalize(NullRng)
    s.key_same(zeros, 32);
    uint8_t spare[LANES][64];
    for (size_t k = 0; k < len; k++) {                                                // fill_bytes(64): the device reduces mod l
        uint8_t *d[LANES];
        // padded lanes repeat a chunk: only the first lane that names a destination writes it
        for (int j = 0; j < LANES; j++) {
            d[j] = out[j] + 64 * k;
            for (int i = 0; i < j; i++) if (out[i] == out[j]) d[j] = spare[j];
        }
        s.meta_ad_same(l64, 4, false);
        s.prf_each(d, 64);
    }
    if (rho) {                                                                        // one more fill_bytes(64) per chunk, unreduced
        uint8_t *d[LANES];
        for (int j = 0; j < LANES; j++) {
            d[j] = rho[j];
            for (int i = 0; i < j; i++) if (rho[i] == rho[j]) d[j] = spare[j];
        }
        s.meta_ad_same(l64, 4, false);
        s.prf_each(d, 64);
    }
}
extern "C" {
// test hook (host only): verifier weights of n_chunks (1..8) chunks of `len` proofs each, wbytes / weights chunk-major (weights canonical,
// 32 bytes each); lockstep = 0: one transcript at a time, 1: through the four-way sponge, 2: through the eight-way sponge
int32_t bpp_host_verifier_weights(const uint8_t *wbytes32, size_t len, size_t n_chunks, int32_t lockstep, uint8_t *weights32) {
    if (!wbytes32 || !weights32 || n_chunks < 1 || n_chunks > 8 || (lockstep == 1 && n_chunks > 4)) return BPP_INVALID_ARGUMENT;
    std::vector<uint8_t> wide(64 * len * n_chunks);
    if (!lockstep) {
        for (size_t c = 0; c < n_chunks; c++) weights_scalar(wbytes32 + 32 * len * c, len, wide.data() + 64 * len * c);
    } else if (lockstep == 1) {
        const uint8_t *wb[4];
        uint8_t *out[4];
        for (size_t j = 0; j < 4; j++) { size_t c = j < n_chunks ? j : n_chunks - 1; wb[j] = wbytes32 + 32 * len * c; out[j] = wide.data() + 64 * len * c; }
        weights_xn<4>(wb, len, out);
    } else {
        const uint8_t *wb[8];
        uint8_t *out[8];
        for (size_t j = 0; j < 8; j++) { size_t c = j < n_chunks ? j : n_chunks - 1; wb[j] = wbytes32 + 32 * len * c; out[j] = wide.data() + 64 * len * c; }
        weights_xn<8>(wb, len, out);
    }
    for (size_t i = 0; i < len * n_chunks; i++) bpp_host_sc_from_wide64(wide.data() + 64 * i, weights32 + 32 * i);
    return BPP_OK;
}
}
// weights of the chunks listed in `only` (or of every chunk whose loop 1 succeeded when `only` is null); scalar = one transcript at a
// time with the reference's redraw of a zero weight
static void compute_weights(bpp_vbatch *vb, const std::vector<size_t> *only = nullptr, bool scalar = false) {
    bpp_ctx *ctx = vb->g->ctx;
    const int width = (ctx->scalar_weights || scalar) ? 1 : bpp_host_simd_level() == 2 ? 8 : 4;
    struct Task { size_t c[8]; int n; };
    std::vector<Task> tasks;
    std::vector<std::pair<size_t, size_t>> todo;       // (length, chunk)
    auto consider = [&](size_t c) {
        const HChunk &hc = vb->hc[c];
        if (hc.pre_rc) return;
        bool failed = false;
        for (size_t i = hc.lo; i < hc.hi && !failed; i++) failed = (vb->flag(i) & 1) != 0;      // loop 1 failed: the call ends there
        if (!failed) todo.emplace_back(hc.hi - hc.lo, c);
    };
    if (only) for (size_t c : *only) consider(c);
    else for (size_t c = 0; c < vb->n_chunks; c++) consider(c);
    std::sort(todo.begin(), todo.end());
    for (size_t i = 0; i < todo.size();) {
        size_t j = i;
        while (j < todo.size() && todo[j].first == todo[i].first && j - i < (size_t)width) j++;
        Task t;
        t.n = (int)(j - i);
        for (int k = 0; k < 8; k++) t.c[k] = todo[i + (size_t)std::min<int>(k, t.n - 1)].second;
        tasks.push_back(t);
        i = j;
    }
    ctx->workers().run(tasks.size(), 1, [&](size_t ti) {
        const Task &t = tasks[ti];
        uint8_t *wts = vb->w->h_weights.as<uint8_t>();
        uint8_t *rhos = vb->merged ? wts + 64 * vb->n_proofs : nullptr;          // one 64-byte value per chunk behind the weights
        const size_t len = vb->hc[t.c[0]].hi - vb->hc[t.c[0]].lo;
        if (t.n >= 2 && width > 1) {       // 2 or more chunks: padded lanes still beat scalar passes
            const uint8_t *wb[8];
            uint8_t *out[8], *rho[8];
            for (int k = 0; k < 8; k++) { wb[k] = vb->wbytes(vb->hc[t.c[k]].lo); out[k] = wts + 64 * vb->hc[t.c[k]].lo; rho[k] = rhos ? rhos + 64 * t.c[k] : nullptr; }
            if (t.n > 4) weights_xn<8>(wb, len, out, rhos ? rho : nullptr);
            else weights_xn<4>(wb, len, out, rhos ? rho : nullptr);
        } else {
            for (int k = 0; k < t.n; k++) weights_scalar(vb->wbytes(vb->hc[t.c[k]].lo), len, wts + 64 * vb->hc[t.c[k]].lo, rhos ? rhos + 64 * t.c[k] : nullptr);
        }
    });
}

static int32_t vbatch_create_impl(bpp_gens *g, size_t n_calls, const bpp_verify_args *const *args, const bpp_verify_challenges *const *chs,
                                  bpp_vbatch **out) {
    if (!g || !args || !out || n_calls == 0) return BPP_INVALID_ARGUMENT;
    bpp_ctx *ctx = g->ctx;
    *out = nullptr;
    // ---- argument checks per call, totals
    uint64_t tot_proofs = 0, tot_chunks = 0, tot_commit = 0, tot_raw = 0;
    for (size_t ci = 0; ci < n_calls; ci++) {
        const bpp_verify_args *a = args[ci];
        if (!a) return fail(ctx, BPP_INVALID_ARGUMENT, "null argument");
        if (a->n_chunks == 0 || !a->chunk_offsets) return fail(ctx, BPP_INVALID_ARGUMENT, "Range statements or proofs length empty");
        if (a->chunk_offsets[0] != 0 || a->chunk_offsets[a->n_chunks] != a->n_proofs) return fail(ctx, BPP_INVALID_ARGUMENT, "bad chunk offsets");
        for (size_t c = 0; c < a->n_chunks; c++)
            if (a->chunk_offsets[c + 1] < a->chunk_offsets[c]) return fail(ctx, BPP_INVALID_ARGUMENT, "bad chunk offsets");
        const bool need_t = !(chs && chs[ci]);
        if (a->n_proofs && (!a->proof_bytes || !a->proof_offsets || !a->commitments32 || !a->commit_offsets || !a->min_values ||
                            !a->min_present || (need_t && !a->transcripts)))
            return fail(ctx, BPP_INVALID_ARGUMENT, "null argument");
        if (chs && chs[ci] && a->n_proofs && (!chs[ci]->challenges32 || !chs[ci]->challenge_offsets || !chs[ci]->weights32))
            return fail(ctx, BPP_INVALID_ARGUMENT, "null argument");
        if ((chs && chs[ci]) != (chs && chs[0])) return fail(ctx, BPP_INVALID_ARGUMENT, "calls with and without caller challenges in one pass");
        if (a->action < BPP_RECOVER_ONLY || a->action > BPP_VERIFY_ONLY) return fail(ctx, BPP_INVALID_ARGUMENT, "bad action");
        if (a->action != args[0]->action) return fail(ctx, BPP_INVALID_ARGUMENT, "calls of one pass must share the action");
        tot_proofs += a->n_proofs; tot_chunks += a->n_chunks;
        if (a->n_proofs) {
            if (a->proof_offsets[a->n_proofs] < a->proof_offsets[0] || a->commit_offsets[a->n_proofs] < a->commit_offsets[0])
                return fail(ctx, BPP_INVALID_ARGUMENT, "bad offsets");
            tot_raw += a->proof_offsets[a->n_proofs] - a->proof_offsets[0];
            tot_commit += a->commit_offsets[a->n_proofs] - a->commit_offsets[0];
        }
    }
    // the reference counts with checked_add / checked_mul and returns SizeOverflow; device offsets are 32-bit
    if (tot_proofs >= (1u << 24) || tot_chunks >= (1u << 24) || tot_commit >= (1u << 26) || tot_raw >= (1ull << 31))
        return fail(ctx, BPP_SIZE_OVERFLOW, "too many proofs in one pass");
    cudaSetDevice(ctx->device);

    auto t_prev = std::chrono::steady_clock::now();
    int t_slot = 0;
    for (double &x : ctx->host_ms) x = 0;
    auto lap = [&]() {
        auto now = std::chrono::steady_clock::now();
        ctx->host_ms[t_slot++] = std::chrono::duration<double, std::milli>(now - t_prev).count();
        t_prev = now;
    };
    bpp_vbatch *vb = new bpp_vbatch();
    vb->g = g; vb->action = args[0]->action; vb->n_proofs = (size_t)tot_proofs; vb->n_chunks = (size_t)tot_chunks; vb->n_commit = (size_t)tot_commit;
    vb->caller_challenges = chs && chs[0];
    vb->device_replay = ctx->device_replay && !vb->caller_challenges;
    vb->hp.resize(vb->n_proofs);
    vb->hc.resize(vb->n_chunks);
    vb->calls.resize(n_calls);
    const int n = g->n, ext = g->ext;
    const bool want_masks = vb->action != BPP_VERIFY_ONLY;
    const size_t NP = vb->n_proofs, NC = vb->n_chunks;

    // ---- calls, chunks, which proofs are looked at (range_proof.rs:739-751: the first 256 of a call)
    {
        size_t p0 = 0, c0 = 0, cm0 = 0, r0 = 0;
        for (size_t ci = 0; ci < n_calls; ci++) {
            HCall &hcall = vb->calls[ci];
            hcall.a = *args[ci];
            hcall.ch = chs ? chs[ci] : nullptr;
            hcall.proof0 = p0; hcall.chunk0 = c0; hcall.commit0 = cm0; hcall.raw0 = r0;
            const bpp_verify_args &a = hcall.a;
            hcall.raw_base = a.n_proofs ? a.proof_offsets[0] : 0;
            for (size_t c = 0; c < a.n_chunks; c++) {
                HChunk &hc = vb->hc[c0 + c];
                hc.lo = p0 + a.chunk_offsets[c];
                hc.end = p0 + a.chunk_offsets[c + 1];
                hc.hi = std::min<size_t>(hc.end, hc.lo + BPP_MAX_BATCH);
                if (hc.hi == hc.lo) hc.pre_rc = BPP_INVALID_ARGUMENT;                           // :719-723
                for (size_t i = hc.lo; i < hc.hi; i++) vb->hp[i].looked = true;
            }
            for (size_t i = 0; i < a.n_proofs; i++) { vb->hp[p0 + i].call = (uint32_t)ci; vb->hp[p0 + i].local = i; }
            if (a.n_proofs) {
                r0 += a.proof_offsets[a.n_proofs] - a.proof_offsets[0];
                cm0 += a.commit_offsets[a.n_proofs] - a.commit_offsets[0];
            }
            p0 += a.n_proofs; c0 += a.n_chunks;
        }
    }
    // ---- per-proof header parse + statement checks (parallel for large passes)
    ctx->workers().run(NP, 512, [&](size_t gi) {
        HProof &p = vb->hp[gi];
        if (!p.looked) return;
        const bpp_verify_args &a = vb->calls[p.call].a;
        const size_t i = p.local;
        const size_t plen = a.proof_offsets[i + 1] - a.proof_offsets[i];
        p.bytes = a.proof_bytes + a.proof_offsets[i];
        int32_t pext = 0, rounds = 0;
        p.pre_rc = bpp_proof_check_bytes(p.bytes, plen, &pext, &rounds);                  // RangeProof::from_bytes
        p.ext = pext; p.rounds = rounds;
        const uint64_t m64 = a.commit_offsets[i + 1] - a.commit_offsets[i];
        p.m = (uint32_t)m64;
        p.has_seed = a.seed_present && a.seed_nonces32 && a.seed_present[i];
        if (!p.pre_rc) {                                                                   // RangeStatement::init, range_statement.rs:42-61
            if (m64 == 0 || (m64 & (m64 - 1)) || m64 > (uint64_t)g->M) p.pre_rc = BPP_INVALID_ARGUMENT;
            else if (p.has_seed && m64 > 1) p.pre_rc = BPP_INVALID_ARGUMENT;
        }
        if (p.pre_rc) return;
        // :875-888 -- 2^rounds must equal n*m; checked_shl overflows from 64 rounds on (usize is 64 bits wide)
        if (p.rounds >= 64) p.loop2_rc = BPP_SIZE_OVERFLOW;
        else if (p.rounds >= 32 || (1ull << p.rounds) != (uint64_t)p.m * (uint64_t)n) p.loop2_rc = BPP_INVALID_LENGTH;
    });
    // ---- transcripts of a call that are all in the same state (same label, nothing appended yet) travel once: 203 of the ~940 bytes a
    // proof costs on the host-to-device link
    size_t n_tstates = 0;
    if (vb->device_replay) {
        ctx->workers().run(n_calls, 1, [&](size_t ci) {
            HCall &call = vb->calls[ci];
            const uint8_t *t = call.a.transcripts;
            bool same = call.a.n_proofs > 1;
            for (size_t i = 1; i < call.a.n_proofs && same; i++) same = memcmp(t, t + BPP_TRANSCRIPT_BYTES * i, BPP_TRANSCRIPT_BYTES) == 0;
            call.same_transcripts = same;
        });
        for (HCall &call : vb->calls) { call.ts0 = n_tstates; n_tstates += call.same_transcripts ? 1 : call.a.n_proofs; }
    }
    // ---- per-chunk consistency (:610-709) and the totals of the device layout
    uint64_t n_pts = 0, n_entries = 0, contrib = 0, pv = 0, n_chal = 0, n_nonce = 0;
    uint32_t max_static = 0, max_seg_entries = 0;
    for (size_t c = 0; c < NC; c++) {
        HChunk &hc = vb->hc[c];
        if (hc.pre_rc) continue;
        const uint64_t entries_before = n_entries;
        int32_t ext_rc = 0, promise_rc = 0;
        bool rounds_ok = true;
        uint32_t max_mn = 0;
        for (size_t i = hc.lo; i < hc.hi; i++) {
            const HProof &p = vb->hp[i];
            if (p.pre_rc) { if (!hc.pre_rc) hc.pre_rc = p.pre_rc; continue; }
            if (p.ext != ext && !ext_rc) ext_rc = BPP_INVALID_ARGUMENT;                    // :637-660
            if (n < 64) {
                const bpp_verify_args &a = vb->calls[p.call].a;
                for (uint32_t j = 0; j < p.m; j++) {
                    const size_t ci = a.commit_offsets[p.local] + j;
                    if (a.min_present[ci] && (a.min_values[ci] >> n) > 0 && !promise_rc) promise_rc = BPP_INVALID_LENGTH;     // :675-682
                }
            }
            if (p.loop2_rc) rounds_ok = false;
            max_mn = std::max<uint32_t>(max_mn, p.m * (uint32_t)n);
        }
        if (!hc.pre_rc) hc.pre_rc = ext_rc ? ext_rc : promise_rc;
        hc.computable = !hc.pre_rc && rounds_ok;
        hc.max_mn = max_mn;
        if (hc.pre_rc) continue;
        const bool msm = hc.computable && vb->action != BPP_RECOVER_ONLY;
        if (msm) n_entries += 2 * (uint64_t)max_mn + (uint64_t)ext + 1;
        for (size_t i = hc.lo; i < hc.hi; i++) {
            const HProof &p = vb->hp[i];
            const uint64_t R = (uint64_t)p.rounds;
            n_pts += 3 + 2 * R + p.m;
            n_chal += 3 + R;
            if (hc.computable && want_masks && p.has_seed) n_nonce += (uint64_t)ext * (3 + 2 * R);
            if (msm) { contrib += 2ull << R; pv += 8 + 3 * R + p.m; n_entries += 3 + 2 * R + p.m; }
        }
        max_seg_entries = (uint32_t)std::max<uint64_t>(max_seg_entries, std::min<uint64_t>(n_entries - entries_before, 0xffffffffu));
    }
    if (n_pts >= (1ull << 30) || n_entries >= (1ull << 30) || contrib >= (1ull << 31) || pv >= (1ull << 31) || n_chal >= (1ull << 31) ||
        n_nonce >= (1ull << 31)) {
        delete vb;
        return fail(ctx, BPP_SIZE_OVERFLOW, "verification pass too large");
    }
    lap();   // [0] parse

    // ---- blob sections
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off = (off + bytes + 16 + 255) & ~(size_t)255; return o; };      // >= 16 bytes of padding
    vb->o_proofs = carve(sizeof(VProof) * NP);
    vb->o_chunks = carve(sizeof(VChunk) * NC);
    vb->o_ptoff = carve(4 * (NP + 1));
    vb->o_segoff = carve(4 * (NC + 1));
    vb->o_hg = carve(32 * ((size_t)ext + 1));
    vb->o_wtinit = carve(BPP_TRANSCRIPT_BYTES);
    vb->o_minv = carve(8 * vb->n_commit);
    vb->o_minp = carve(vb->n_commit);
    vb->o_commit = carve(32 * vb->n_commit);
    vb->o_raw = carve((size_t)tot_raw);
    vb->o_tstate = carve(BPP_TRANSCRIPT_BYTES * n_tstates);
    vb->o_nonces = carve(32 * (size_t)n_nonce);
    vb->blob_bytes = off;
    if (vb->blob_bytes >= (1ull << 32)) { delete vb; return fail(ctx, BPP_SIZE_OVERFLOW, "verification pass too large"); }
    vb->n_pts = (uint32_t)n_pts; vb->n_entries = (uint32_t)n_entries; vb->n_chal = (uint32_t)n_chal;
    vb->mo_wbytes = 0;
    vb->mo_flags = 32 * NP;
    vb->mo_tstate = (vb->mo_flags + NP + 255) & ~(size_t)255;
    vb->mid_bytes = vb->mo_tstate + BPP_TRANSCRIPT_BYTES * NP;
    vb->ho_ok = 0;
    vb->ho_ident = ((size_t)n_pts + 255) & ~(size_t)255;
    vb->ho_masks = vb->ho_ident + ((2 * NC + 255) & ~(size_t)255);      // [identity flags | zero-weight flags]
    vb->hout_bytes = vb->ho_masks + 32 * std::max<size_t>(NP, 1) * (size_t)ext;

    // ---- buffers (pooled)
    VWork *w = vwork_acquire(ctx);
    vb->w = w;
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
    const size_t np1 = std::max<size_t>(NP, 1);
    // static entries of a chunk are bounded by the generator set; the MSM shape needs the entry count only
    vb->shape = msm_shape(vb->n_entries, (uint32_t)NC, 0);
    vb->shape.max_seg_entries = max_seg_entries;
    {   // merged check: worth it from two chunks with multiscalar work on; needs the weight transcripts (not the caller-challenge form)
        size_t msm_chunks = 0;
        for (size_t c = 0; c < NC; c++) msm_chunks += (!vb->hc[c].pre_rc && vb->hc[c].computable && vb->action != BPP_RECOVER_ONLY) ? 1 : 0;
        // (from four chunks on: two chunks as one sum take the large-sum kernels' longer chains -- 4096 proofs over 8 GPUs, 512 each: 0.82 ms
        // merged against 0.71 ms chunk by chunk -- while four chunks, one 1024-proof job, already gain: 0.75 against 0.80 ms;
        // BPP_MERGED_MIN_CHUNKS overrides, the tests use 2)
        static const size_t merged_min = [] { const char *e = getenv("BPP_MERGED_MIN_CHUNKS"); return e && atoi(e) >= 2 ? (size_t)atoi(e) : (size_t)4; }();
        vb->merged = ctx->merged_check && !vb->caller_challenges && msm_chunks >= merged_min && vb->n_entries > 0;
        if (vb->merged) vb->shape_m = msm_shape(vb->n_entries, 1, 0);
    }
    ok(w->h_blob.ensure(vb->blob_bytes));
    ok(w->d_blob.ensure(vb->blob_bytes));
    ok(w->h_out.ensure(vb->hout_bytes));
    ok(w->h_mid.ensure(vb->mid_bytes + 256));
    ok(w->d_mid.ensure(vb->mid_bytes + 256));
    ok(w->h_weights.ensure(64 * (np1 + NC)));          // [weight per proof | rho per chunk (merged check)]
    ok(w->d_weights.ensure(64 * (np1 + NC)));
    ok(w->d_wmont.ensure(32 * np1));
    ok(w->d_chal.ensure(32 * std::max<size_t>(n_chal, 1)));
    if (!vb->device_replay) ok(w->h_chal.ensure(32 * std::max<size_t>(n_chal, 1)));
    ok(w->d_tab.ensure(sizeof(aniels) * std::max<size_t>(n_pts, 1)));
    ok(w->d_ok.ensure(std::max<size_t>(n_pts, 1)));
    ok(w->d_mscal.ensure(32 * std::max<size_t>(n_entries, 1)));
    ok(w->d_pidx.ensure(4 * std::max<size_t>(n_entries, 1)));
    ok(w->d_contrib.ensure(32 * std::max<size_t>(contrib, 1)));
    ok(w->d_hg.ensure(32 * np1 * (1 + (size_t)ext)));
    ok(w->d_pervec.ensure(32 * std::max<size_t>(pv, 1)));
    ok(w->d_masks.ensure(32 * np1 * (size_t)ext));
    ok(w->d_scratch.ensure(std::max(msm_scratch_bytes(vb->shape), vb->merged ? msm_scratch_bytes(vb->shape_m) : (size_t)0)));
    ok(w->d_res.ensure(sizeof(ge) * NC));
    ok(w->d_ident.ensure(2 * NC));
    if (e != cudaSuccess) { vwork_return(ctx, w); delete vb; return cuda_fail(ctx, e, "vbatch buffers"); }

    // ---- device layout, written straight into the pinned blob
    uint8_t *hb = w->h_blob.as<uint8_t>();
    VProof *dp = (VProof *)(hb + vb->o_proofs);
    VChunk *dc = (VChunk *)(hb + vb->o_chunks);
    uint32_t *ptoff = (uint32_t *)(hb + vb->o_ptoff), *segoff = (uint32_t *)(hb + vb->o_segoff);
    {
        uint32_t pts = 0, entries = 0, cb = 0, pvo = 0, chal = 0, nonces = 0;
        memset(dp, 0, sizeof(VProof) * NP);
        size_t next_proof = 0;      // ptoff is filled for every proof, also those no chunk looks at
        for (size_t c = 0; c < NC; c++) {
            HChunk &hc = vb->hc[c];
            VChunk &ch = dc[c];
            ch.proof_lo = (uint32_t)hc.lo; ch.proof_hi = (uint32_t)hc.hi; ch.max_mn = hc.max_mn;
            const bool msm = hc.computable && vb->action != BPP_RECOVER_ONLY;
            ch.active = msm ? 1 : 0;
            ch.entry_off = entries;
            hc.entry_off = entries;
            segoff[c] = entries;
            if (msm) {
                vb->any_msm = true;
                const uint32_t n_static = 2 * hc.max_mn + (uint32_t)ext + 1;
                max_static = std::max(max_static, n_static);
                entries += n_static;
            }
            for (; next_proof < hc.lo; next_proof++) ptoff[next_proof] = pts;
            for (size_t i = hc.lo; i < hc.end; i++) {
                HProof &p = vb->hp[i];
                VProof &v = dp[i];
                ptoff[i] = pts;
                v.nonce_off = 0xffffffffu;
                if (i >= hc.hi || hc.pre_rc) continue;          // not looked at, or the call failed on the host: no device work
                const HCall &call = vb->calls[p.call];
                const bpp_verify_args &a = call.a;
                const uint32_t R = (uint32_t)p.rounds;
                p.pt_off = pts; p.n_pts = 3 + 2 * R + p.m;
                pts += p.n_pts;
                // loop 1 runs over every proof of a call that passed the consistency checks (:816-850)
                v.replay = 1; vb->any_replay = true;
                v.pt_off = p.pt_off;
                v.m = p.m; v.rounds = R;
                v.raw_off = (uint32_t)(vb->o_raw + call.raw0 + (a.proof_offsets[p.local] - call.raw_base));
                v.commit_off = (uint32_t)(call.commit0 + (a.commit_offsets[p.local] - a.commit_offsets[0]));
                v.ch_off = chal; chal += 3 + R;
                v.ts_idx = (uint32_t)(call.ts0 + (call.same_transcripts ? 0 : p.local));
                v.chunk = (uint32_t)c;
                if (!hc.computable) continue;
                if (want_masks && p.has_seed) {
                    v.nonce_off = nonces; nonces += (uint32_t)ext * (3 + 2 * R);
                    vb->any_masks = true;
                }
                if (msm) {
                    v.active = 1; vb->any_vec = true;
                    v.entry_off = entries;
                    v.contrib_off = cb; cb += 2u << R;
                    v.pv_off = pvo; pvo += 8 + 3 * R + p.m;
                    vb->max_rounds = std::max(vb->max_rounds, R);
                    entries += 3 + 2 * R + p.m;
                }
            }
            next_proof = hc.end;
            hc.n_entries = entries - hc.entry_off;
        }
        for (; next_proof <= NP; next_proof++) ptoff[next_proof] = pts;
        segoff[NC] = entries;
        vb->max_static = max_static;
    }
    lap();   // [1] layout

    // ---- fill: the callers' byte arrays as they are, one memcpy each
    // (the callers' buffers are cold in the cache more often than not: one thread copies ~7 GB/s, so the calls are spread over the workers)
    // Proof bytes that already lie in page-locked memory are not staged at all: the upload below reads them where they are (78 % of the
    // bytes of a pass; on an 8-GPU host the staging copies of eight ranks were what saturated the host's memory system).
    bool any_pinned = false;
    for (HCall &call : vb->calls) {
        const bpp_verify_args &a = call.a;
        call.raw_pinned = a.n_proofs && host_range_pinned(a.proof_bytes + a.proof_offsets[0], a.proof_offsets[a.n_proofs] - a.proof_offsets[0]);
        any_pinned = any_pinned || call.raw_pinned;
    }
    ctx->workers().run(vb->calls.size(), 1, [&](size_t ci) {
        const HCall &call = vb->calls[ci];
        const bpp_verify_args &a = call.a;
        if (!a.n_proofs) return;
        const size_t raw_len = a.proof_offsets[a.n_proofs] - a.proof_offsets[0];
        const size_t c_lo = a.commit_offsets[0], c_n = a.commit_offsets[a.n_proofs] - c_lo;
        if (!call.raw_pinned) memcpy(hb + vb->o_raw + call.raw0, a.proof_bytes + a.proof_offsets[0], raw_len);
        memcpy(hb + vb->o_commit + 32 * call.commit0, a.commitments32 + 32 * c_lo, 32 * c_n);
        memcpy(hb + vb->o_minv + 8 * call.commit0, a.min_values + c_lo, 8 * c_n);
        memcpy(hb + vb->o_minp + call.commit0, a.min_present + c_lo, c_n);
        if (vb->device_replay) memcpy(hb + vb->o_tstate + BPP_TRANSCRIPT_BYTES * call.ts0, a.transcripts, BPP_TRANSCRIPT_BYTES * (call.same_transcripts ? 1 : a.n_proofs));
    });
    memcpy(hb + vb->o_hg, g->h(), 32);
    memcpy(hb + vb->o_hg + 32, g->g(0), 32 * (size_t)ext);
    memcpy(hb + vb->o_wtinit, weight_transcript_init().data(), BPP_TRANSCRIPT_BYTES);      // starting state of k_weights
    memset(vb->mid(), 0, vb->mo_tstate);             // wbytes + flags (the transcript states are written by whoever replays)
    memset(w->h_weights.p, 0, 64 * (np1 + NC));
    const bool host_replay = !vb->device_replay && !vb->caller_challenges;
    if (vb->caller_challenges) {
        // the caller ran loop 1 with its own merlin (src/transcripts.rs unmodified): challenges [y, z, e, e_0..e_{r-1}] per proof and
        // the batch weights (:894) come in; zero challenges / identity points were the caller's to reject (VerificationFailed)
        for (size_t gi = 0; gi < NP; gi++) {
            const HProof &p = vb->hp[gi];
            const VProof &v = dp[gi];
            if (!v.replay) continue;
            const HCall &call = vb->calls[p.call];
            const uint64_t co = call.ch->challenge_offsets[p.local], cn = call.ch->challenge_offsets[p.local + 1] - co;
            uint8_t *dst = w->h_chal.as<uint8_t>() + 32 * (size_t)v.ch_off;
            if (cn != 3 + (uint64_t)p.rounds) { vwork_return(ctx, w); delete vb; return fail(ctx, BPP_INVALID_ARGUMENT, "challenge count does not match the proof"); }
            memcpy(dst, call.ch->challenges32 + 32 * co, 32 * cn);
            bool canon = true;
            for (uint64_t k = 0; k < cn; k++) canon = canon && host_sc_is_canonical(dst + 32 * k) && !is_zero32(dst + 32 * k);
            if (!canon || !host_sc_is_canonical(call.ch->weights32 + 32 * p.local) || is_zero32(call.ch->weights32 + 32 * p.local)) {
                vwork_return(ctx, w); delete vb;
                return fail(ctx, BPP_INVALID_ARGUMENT, "challenges and weights must be canonical non-zero scalars");
            }
            memcpy(w->h_weights.as<uint8_t>() + 64 * gi, call.ch->weights32 + 32 * p.local, 32);      // canonical: the upper half stays zero
            uint8_t one[32] = {1};
            vb->flag(gi) = memcmp(dst, one, 32) ? 0 : 2;      // y == 1
        }
    }
    if (host_replay || n_nonce) {
        ctx->workers().run(NP, host_replay ? 8 : 64, [&](size_t gi) {
            const HProof &p = vb->hp[gi];
            const VProof &v = dp[gi];
            if (!v.replay) return;
            const HCall &call = vb->calls[p.call];
            const bpp_verify_args &a = call.a;
            const size_t i = p.local;
            if (host_replay) {              // loop 1 on this host thread (north_star's split); same statements as k_replay.cu
                const uint8_t *b = p.bytes;
                uint8_t *chp = w->h_chal.as<uint8_t>() + 32 * (size_t)v.ch_off;
                ReplayIn in;
                in.tstate = a.transcripts + BPP_TRANSCRIPT_BYTES * i;
                in.h32 = g->h(); in.g32 = g->g(0);
                in.bit_length = (uint32_t)n; in.ext = (uint32_t)ext; in.m = p.m; in.rounds = (uint32_t)p.rounds;
                in.commitments32 = a.commitments32 + 32 * a.commit_offsets[i];
                in.min_values = a.min_values + a.commit_offsets[i]; in.min_present = a.min_present + a.commit_offsets[i];
                in.a = b + BPP_RAW_A(ext); in.a1 = in.a + 32; in.b = in.a + 64;
                in.l_base = b + BPP_RAW_L(ext, 0); in.r_base = b + BPP_RAW_R(ext, 0); in.lr_stride = 64;
                in.r1 = b + BPP_RAW_R1(ext); in.s1 = b + BPP_RAW_S1(ext); in.d1 = b + BPP_RAW_D1(ext, 0);
                ReplayOut o;
                o.y = chp; o.z = chp + 32; o.e = chp + 64; o.ej = chp + 96;
                o.wbytes = vb->wbytes(gi); o.tstate = vb->tstate(gi);
                int rc = replay_transcript_core(in, o);
                uint8_t flag = rc ? 1 : 0;
                uint8_t one[32] = {1};
                if (!rc && !memcmp(chp, one, 32)) flag |= 2;
                vb->flag(gi) = flag;
            }
            if (v.nonce_off != 0xffffffffu) {
                uint8_t seed[32];            // Scalar::from_bytes_mod_order; a seed that is already canonical (the usual case) needs no reduction
                const uint8_t *sb = a.seed_nonces32 + 32 * i;
                if (host_sc_is_canonical(sb)) memcpy(seed, sb, 32);
                else {
                    uint32_t ww[8];
                    memcpy(ww, sb, 32);
                    sc s; for (int k = 0; k < 8; k++) s.v[k] = ww[k];
                    sc_tobytes(seed, sc_reduce256(s));
                }
                uint8_t *nn = hb + vb->o_nonces + 32 * (size_t)v.nonce_off;
                for (int k = 0; k < ext; k++) {
                    nonce(seed, "eta", false, 0, true, (uint32_t)k, nn + 32 * k);
                    nonce(seed, "d", false, 0, true, (uint32_t)k, nn + 32 * (ext + k));
                    nonce(seed, "alpha", false, 0, true, (uint32_t)k, nn + 32 * (2 * ext + k));
                    for (int j = 0; j < p.rounds; j++) {
                        nonce(seed, "dL", true, (uint32_t)j, true, (uint32_t)k, nn + 32 * (3 * ext + j * ext + k));
                        nonce(seed, "dR", true, (uint32_t)j, true, (uint32_t)k, nn + 32 * (3 * ext + p.rounds * ext + j * ext + k));
                    }
                }
                memset(seed, 0, sizeof seed);
            }
        });
    }
    lap();   // [2] fill (+ host transcript replay in host mode)
    if (host_replay) compute_weights(vb);
    lap();   // [3] weight transcripts (host mode; in device mode they run inside bpp_vbatch_run)
    cudaStream_t st = ctx->stream;
    if (vb->blob_bytes && !any_pinned) {
        ok(cudaMemcpyAsync(w->d_blob.p, hb, vb->blob_bytes, cudaMemcpyHostToDevice, st));
    } else if (vb->blob_bytes) {
        // [descriptors .. commitments) and [transcript states .. end) from the staging blob, the raw proofs call by call from where they
        // are (staged calls that follow each other travel as one copy)
        uint8_t *db = w->d_blob.as<uint8_t>();
        ok(cudaMemcpyAsync(db, hb, vb->o_raw, cudaMemcpyHostToDevice, st));
        size_t run_lo = 0, run_len = 0;            // pending run of staged raw bytes, relative to o_raw
        auto flush = [&]() {
            if (run_len) ok(cudaMemcpyAsync(db + vb->o_raw + run_lo, hb + vb->o_raw + run_lo, run_len, cudaMemcpyHostToDevice, st));
            run_len = 0;
        };
        for (const HCall &call : vb->calls) {
            const bpp_verify_args &a = call.a;
            if (!a.n_proofs) continue;
            const size_t raw_len = a.proof_offsets[a.n_proofs] - a.proof_offsets[0];
            if (call.raw_pinned) {
                flush();
                ok(cudaMemcpyAsync(db + vb->o_raw + call.raw0, a.proof_bytes + a.proof_offsets[0], raw_len, cudaMemcpyHostToDevice, st));
            } else {
                if (!run_len) run_lo = call.raw0;
                run_len = call.raw0 + raw_len - run_lo;
            }
        }
        flush();
        if (vb->blob_bytes > vb->o_tstate) ok(cudaMemcpyAsync(db + vb->o_tstate, hb + vb->o_tstate, vb->blob_bytes - vb->o_tstate, cudaMemcpyHostToDevice, st));
    }
    if (!vb->device_replay && n_chal) ok(cudaMemcpyAsync(w->d_chal.p, w->h_chal.p, 32 * (size_t)n_chal, cudaMemcpyHostToDevice, st));
    // The upload is ordered before the kernels of bpp_vbatch_run on the same stream and the pinned blob belongs to this vbatch's
    // workspace until bpp_vbatch_destroy (which drains the stream), so nothing needs the host to wait here (an upload error
    // surfaces in bpp_vbatch_run).
    static const bool always_sync = getenv("BPP_CREATE_SYNC") != nullptr;      // debugging: upload errors surface here, host_ms[4] = H2D time
    if (e == cudaSuccess && always_sync) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        cudaStreamSynchronize(st);          // copies that did get queued may still be reading the staging blob or page-locked caller buffers
        vwork_return(ctx, w); delete vb;
        return cuda_fail(ctx, e, "vbatch upload");
    }
    lap();   // [4] H2D
    ctx->io_bytes[0] = vb->blob_bytes + (vb->device_replay ? 0 : 32 * (size_t)n_chal); ctx->io_bytes[1] = 0;
    for (HProof &p : vb->hp) p.bytes = nullptr;      // the host does not read the callers' buffers after this point (the copy engine still
                                                     // reads page-locked proof bytes until the pass has run, see bpp_b200.h)
    *out = vb;
    return BPP_OK;
}

extern "C" {

void bpp_vbatch_destroy(bpp_vbatch *vb) {
    if (!vb) return;
    cudaSetDevice(vb->g->ctx->device);
    cudaStreamSynchronize(vb->g->ctx->stream);
    if (vb->w) {
        if (vb->any_masks && vb->w->h_blob.p) memset(vb->w->h_blob.as<uint8_t>() + vb->o_nonces, 0, vb->blob_bytes - vb->o_nonces);   // seed-derived nonces
        vwork_return(vb->g->ctx, vb->w);
    }
    delete vb;
}

int32_t bpp_vbatch_create(bpp_gens *g, const bpp_verify_args *a, bpp_vbatch **out) {
    return vbatch_create_impl(g, 1, &a, nullptr, out);
}
int32_t bpp_vbatch_create_multi(bpp_gens *g, size_t n_calls, const bpp_verify_args *const *calls, bpp_vbatch **out) {
    return vbatch_create_impl(g, n_calls, calls, nullptr, out);
}

// Wait for an event without spinning and without the driver's blocking-sync machinery: poll it between sleeps.  Measured per
// 1024-proof pass (one lane, scripts/e2e_cpu_cost.py): cudaEventBlockingSync waits cost the process ~0.24 ms of CPU in driver
// threads on top of the calling thread's own work; a spin wait costs the whole pass (1 ms).
// Option (BPP_ADAPTIVE_WAIT=1, `ema_ns` != nullptr): `ema_ns` remembers how long this wait took recently, the first sleep covers 3/4
// of that in one go and the naps after it are 1/20 of it (~6 wake-ups per pass instead of ~220).  MEASURED, NOT A WIN: passes of a
// busy queue vary too much for the estimate (oversleeping holds results back): 8.1 / 7.8 M proofs/s end to end against 9.6 / 9.7 M with
// fixed 60 us naps (4 / 16 host cores); the fixed naps cost little -- two host cores sustain 8.8 M proofs/s with device-side weights.
// The estimate is kept per proof of the pass (floor: 1024 proofs, below that a pass is its latency chains whatever it holds), so that a
// small pass after large ones is not overslept.
static cudaError_t wait_sleeping(cudaEvent_t ev, long nap_ns, double *ema_ns, size_t n_proofs) {
    cudaError_t e = cudaEventQuery(ev);
    if (e != cudaErrorNotReady) { if (ema_ns) *ema_ns *= 0.7; return e; }
    const auto t0 = std::chrono::steady_clock::now();
    const double scale = (double)std::max<size_t>(n_proofs, 1024);
    const double est = ema_ns ? *ema_ns * scale : 0.0;
    long nap = nap_ns;
    if (ema_ns && est > 4.0 * (double)nap_ns) {
        const long ns = (long)(0.75 * est);
        timespec ts = {ns / 1000000000L, ns % 1000000000L};
        nanosleep(&ts, nullptr);
        e = cudaEventQuery(ev);
        if (e != cudaErrorNotReady) { *ema_ns *= 0.7; return e; }
        nap = std::max(nap_ns, (long)(0.05 * est));
    }
    for (;;) {
        timespec ts = {nap / 1000000000L, nap % 1000000000L};
        nanosleep(&ts, nullptr);
        e = cudaEventQuery(ev);
        if (e != cudaErrorNotReady) break;
    }
    if (ema_ns) {
        const double el = std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t0).count() / scale;
        *ema_ns = *ema_ns > 0 ? 0.75 * *ema_ns + 0.25 * el : el;
    }
    return e;
}

// kernel arguments of one pass
struct VLaunch {
    VDims d;
    VBuffers b;
    RBuffers rb;
    bool dev_replay;
    int replay_kernel;
};
static VLaunch make_launch(bpp_vbatch *vb) {
    bpp_gens *g = vb->g;
    bpp_ctx *ctx = g->ctx;
    VWork *w = vb->w;
    VLaunch L;
    VDims &d = L.d;
    d.n_proofs = (uint32_t)vb->n_proofs; d.n_chunks = (uint32_t)vb->n_chunks; d.bit_length = (uint32_t)g->n; d.ext = (uint32_t)g->ext;
    d.action = vb->action; d.gens_nm = (uint32_t)g->nm; d.merged = vb->merged ? 1u : 0u;
    VBuffers &b = L.b;
    b.proofs = vb->dev<VProof>(vb->o_proofs); b.chunks = vb->dev<VChunk>(vb->o_chunks); b.pt_offsets = vb->dev<uint32_t>(vb->o_ptoff);
    b.blob = w->d_blob.as<uint8_t>(); b.challenges = w->d_chal.as<uint32_t>();
    b.weights = w->d_weights.as<uint32_t>(); b.weights_mont = w->d_wmont.as<uint32_t>();
    b.weight_zero = w->d_ident.as<uint8_t>() + vb->n_chunks;
    b.min_values = vb->dev<uint64_t>(vb->o_minv); b.min_present = vb->dev<uint8_t>(vb->o_minp); b.nonces = vb->dev<uint32_t>(vb->o_nonces);
    b.msm_scalars = w->d_mscal.as<uint32_t>(); b.msm_pidx = w->d_pidx.as<uint32_t>();
    b.contrib = w->d_contrib.as<uint32_t>(); b.hg_contrib = w->d_hg.as<uint32_t>();
    b.pervec = w->d_pervec.as<uint32_t>(); b.masks = vb->any_masks ? w->d_masks.as<uint32_t>() : nullptr;
    L.dev_replay = vb->device_replay && vb->any_replay;
    RBuffers &rb = L.rb;
    rb.proofs = b.proofs; rb.tstates_in = vb->dev<uint8_t>(vb->o_tstate); rb.hg32 = vb->dev<uint8_t>(vb->o_hg);
    rb.blob = b.blob; rb.commitments32 = vb->dev<uint8_t>(vb->o_commit);
    rb.min_values = b.min_values; rb.min_present = b.min_present;
    rb.challenges = w->d_chal.as<uint8_t>();
    uint8_t *dm = w->d_mid.as<uint8_t>();
    rb.wbytes = dm + vb->mo_wbytes; rb.flags = dm + vb->mo_flags; rb.tstates_out = dm + vb->mo_tstate;
    L.replay_kernel = ctx->replay_kernel;
    return L;
}

static void enqueue_decompress(bpp_vbatch *vb, const VLaunch &L, cudaStream_t s, uint64_t *kernels) {
    VWork *w = vb->w;
    launch_decompress_proofs(s, vb->n_pts, L.d.n_proofs, L.d.ext, L.b.proofs, L.b.pt_offsets, L.b.blob, L.rb.commitments32, w->d_tab.as<aniels>(),
                             w->d_ok.as<uint8_t>());
    (*kernels)++;
}
// merged = true: every entry of the pass as ONE sum (result and identity flag in slot 0); false: one sum per chunk
static void enqueue_msm(bpp_vbatch *vb, const VLaunch &L, cudaStream_t st, uint64_t *kernels, cudaEvent_t *marks, bool merged) {
    VWork *w = vb->w;
    if (merged) {
        launch_msm(st, vb->shape_m, w->d_mscal.as<uint32_t>(), nullptr, w->d_pidx.as<uint32_t>(), w->d_tab.as<aniels>(), vb->g->d_table.as<aniels>(),
                   w->d_scratch.p, w->d_res.as<ge>(), kernels, marks);
        launch_encode(st, 1, w->d_res.as<ge>(), nullptr, w->d_ident.as<uint8_t>());
    } else {
        launch_msm(st, vb->shape, w->d_mscal.as<uint32_t>(), vb->n_chunks > 1 ? vb->dev<uint32_t>(vb->o_segoff) : nullptr, w->d_pidx.as<uint32_t>(),
                   w->d_tab.as<aniels>(), vb->g->d_table.as<aniels>(), w->d_scratch.p, w->d_res.as<ge>(), kernels, marks);
        launch_encode(st, vb->n_chunks, w->d_res.as<ge>(), nullptr, w->d_ident.as<uint8_t>());
    }
    (*kernels)++;
}

// the three sections of a pass, enqueued on the ctx streams (directly, or into a stream capture); no host synchronisation
static cudaError_t enqueue_section(bpp_vbatch *vb, const VLaunch &L, int section, uint64_t *kernels) {
    bpp_gens *g = vb->g;
    bpp_ctx *ctx = g->ctx;
    VWork *w = vb->w;
    cudaStream_t st = ctx->stream;
    const size_t n = vb->n_proofs;
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) { if (e == cudaSuccess && r != cudaSuccess) e = r; };
    const bool prep = vb->any_msm || vb->any_masks;
    if (section == 0) {
        if (L.dev_replay) {
            launch_replay(st, L.d, L.rb, L.replay_kernel, kernels);
            ok(cudaMemcpyAsync(vb->mid(), w->d_mid.p, vb->mid_bytes, cudaMemcpyDeviceToHost, st));
        }
    } else if (section == 1) {
        const bool fork = vb->n_pts && prep;          // decompression next to the scalar prep chain, joined at the end of the section
        if (vb->n_pts) {
            if (fork) { ok(cudaEventRecord(ctx->ev_fork, st)); ok(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0)); }
            enqueue_decompress(vb, L, fork ? ctx->stream2 : st, kernels);
            if (fork) ok(cudaEventRecord(ctx->ev_join, ctx->stream2));
        }
        if (prep) launch_verify_prep(st, L.d, L.b, vb->any_vec ? 1u : 0u, vb->max_rounds, kernels, nullptr);
        if (fork) ok(cudaStreamWaitEvent(st, ctx->ev_join, 0));
    } else if (section == 3) {
        // throughput mode 2: the whole pass without a host step.  st: replay -> D2H(flags, transcripts) -> scalar prep -> [weights]
        // -> weighting -> [points] -> MSM -> verdicts;  stream2: decompression (from the start);  stream3: weight transcripts
        const bool fork_pts = vb->n_pts != 0;
        if (fork_pts) {
            ok(cudaEventRecord(ctx->ev_fork, st)); ok(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
            enqueue_decompress(vb, L, ctx->stream2, kernels);
            ok(cudaEventRecord(ctx->ev_join, ctx->stream2));
        }
        launch_replay(st, L.d, L.rb, L.replay_kernel, kernels);
        if (vb->any_msm) {
            ok(cudaEventRecord(ctx->ev_fork2, st)); ok(cudaStreamWaitEvent(ctx->stream3, ctx->ev_fork2, 0));
            launch_weights(ctx->stream3, L.d, L.b.chunks, vb->dev<uint8_t>(vb->o_wtinit), L.rb.wbytes, L.rb.flags, w->d_weights.as<uint32_t>(), kernels);
            ok(cudaEventRecord(ctx->ev_join2, ctx->stream3));
        }
        ok(cudaMemcpyAsync(vb->mid(), w->d_mid.p, vb->mid_bytes, cudaMemcpyDeviceToHost, st));
        if (prep) launch_verify_prep(st, L.d, L.b, vb->any_vec ? 1u : 0u, vb->max_rounds, kernels, nullptr);
        if (vb->any_msm) {
            ok(cudaStreamWaitEvent(st, ctx->ev_join2, 0));
            ok(cudaMemsetAsync(L.b.weight_zero, 0, vb->n_chunks, st));
            launch_verify_weigh(st, L.d, L.b, vb->max_static, kernels);
            if (fork_pts) ok(cudaStreamWaitEvent(st, ctx->ev_join, 0));
            enqueue_msm(vb, L, st, kernels, nullptr, vb->merged);
            ok(cudaMemcpyAsync(w->h_out.as<uint8_t>() + vb->ho_ident, w->d_ident.p, 2 * vb->n_chunks, cudaMemcpyDeviceToHost, st));
        } else if (fork_pts) {
            ok(cudaStreamWaitEvent(st, ctx->ev_join, 0));
        }
        if (vb->n_pts) ok(cudaMemcpyAsync(w->h_out.as<uint8_t>() + vb->ho_ok, w->d_ok.p, vb->n_pts, cudaMemcpyDeviceToHost, st));
        if (vb->any_masks) ok(cudaMemcpyAsync(w->h_out.as<uint8_t>() + vb->ho_masks, w->d_masks.p, 32 * n * (size_t)g->ext, cudaMemcpyDeviceToHost, st));
    } else {
        if (vb->any_msm) {
            ok(cudaMemcpyAsync(w->d_weights.p, w->h_weights.p, 64 * (n + (vb->merged ? vb->n_chunks : 0)), cudaMemcpyHostToDevice, st));
            ok(cudaMemsetAsync(L.b.weight_zero, 0, vb->n_chunks, st));
            launch_verify_weigh(st, L.d, L.b, vb->max_static, kernels);
            enqueue_msm(vb, L, st, kernels, nullptr, vb->merged);
            ok(cudaMemcpyAsync(w->h_out.as<uint8_t>() + vb->ho_ident, w->d_ident.p, 2 * vb->n_chunks, cudaMemcpyDeviceToHost, st));
        }
        if (vb->n_pts) ok(cudaMemcpyAsync(w->h_out.as<uint8_t>() + vb->ho_ok, w->d_ok.p, vb->n_pts, cudaMemcpyDeviceToHost, st));
        if (vb->any_masks) ok(cudaMemcpyAsync(w->h_out.as<uint8_t>() + vb->ho_masks, w->d_masks.p, 32 * n * (size_t)g->ext, cudaMemcpyDeviceToHost, st));
    }
    ok(cudaGetLastError());
    return e;
}

static VGraphKey make_graph_key(const bpp_vbatch *vb, const VLaunch &L, bool fused) {
    VGraphKey k;
    memset(&k, 0, sizeof k);
    const VWork *w = vb->w;
    const void *bufs[21] = {w->d_blob.p, w->d_tab.p, w->d_ok.p, w->d_mscal.p, w->d_pidx.p, w->d_chal.p, w->d_contrib.p, w->d_hg.p, w->d_pervec.p,
                            w->d_masks.p, w->d_scratch.p, w->d_res.p, w->d_ident.p, w->d_weights.p, w->d_wmont.p, w->d_mid.p, w->h_blob.p, w->h_out.p,
                            w->h_mid.p, w->h_weights.p, w->h_chal.p};
    memcpy(k.bufs, bufs, sizeof bufs);
    k.gens_table = vb->g->d_table.p;
    k.n_proofs = vb->n_proofs; k.n_chunks = vb->n_chunks; k.n_commit = vb->n_commit;
    const size_t off[21] = {vb->o_proofs, vb->o_chunks, vb->o_ptoff, vb->o_segoff, vb->o_hg, vb->o_wtinit, vb->o_minv, vb->o_minp, vb->o_commit, vb->o_raw,
                            vb->o_tstate, vb->o_nonces, vb->blob_bytes, vb->mo_wbytes, vb->mo_flags, vb->mo_tstate, vb->mid_bytes,
                            vb->ho_ok, vb->ho_ident, vb->ho_masks, vb->hout_bytes};
    memcpy(k.off, off, sizeof off);
    k.n_pts = vb->n_pts; k.n_entries = vb->n_entries; k.n_chal = vb->n_chal; k.max_static = vb->max_static; k.max_rounds = vb->max_rounds;
    k.action = vb->action; k.ext = vb->g->ext; k.bit_length = vb->g->n;
    k.shape.n_entries = vb->shape.n_entries; k.shape.n_seg = vb->shape.n_seg; k.shape.c = vb->shape.c; k.shape.W = vb->shape.W; k.shape.B = vb->shape.B;
    k.shape.max_seg_entries = vb->shape.max_seg_entries;
    msm_knobs(k.msm_knobs);
    k.any_msm = vb->any_msm; k.any_masks = vb->any_masks; k.any_replay = vb->any_replay; k.any_vec = vb->any_vec;
    k.device_replay = L.dev_replay; k.replay_kernel = (uint8_t)L.replay_kernel;
    k.fused = (uint8_t)((fused ? 1 : 0) | (vb->merged ? 2 : 0)); k.caller_challenges = vb->caller_challenges;      // (the merged shape follows from n_entries and the knobs)
    return k;
}

// finds or captures the three graphs of this pass; nullptr (with *err set) if a capture failed
static VGraph *vgraph_get(bpp_vbatch *vb, const VLaunch &L, bool fused, cudaError_t *err) {
    bpp_ctx *ctx = vb->g->ctx;
    const VGraphKey key = make_graph_key(vb, L, fused);
    for (void *p : ctx->vgraphs) {
        VGraph *g = (VGraph *)p;
        if (!memcmp(&g->key, &key, sizeof key)) { g->last_use = ++ctx->vgraph_clock; return g; }
    }
    VGraph *g = new VGraph();
    g->key = key;
    for (int sct = 0; sct < 3; sct++) {
        if (fused ? sct != 0 : (sct == 0 && !L.dev_replay)) continue;
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal);
        if (e == cudaSuccess) {
            cudaError_t e1 = enqueue_section(vb, L, fused ? 3 : sct, &g->kernels[sct]);
            e = cudaStreamEndCapture(ctx->stream, &graph);
            if (e == cudaSuccess) e = e1;
        }
        if (e == cudaSuccess) e = cudaGraphInstantiate(&g->ex[sct], graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (e != cudaSuccess) { *err = e; vgraph_free(g); cudaGetLastError(); return nullptr; }
    }
    if (ctx->vgraphs.size() >= 8) {          // evict the least recently used layout
        size_t victim = 0;
        for (size_t i = 1; i < ctx->vgraphs.size(); i++)
            if (((VGraph *)ctx->vgraphs[i])->last_use < ((VGraph *)ctx->vgraphs[victim])->last_use) victim = i;
        vgraph_free((VGraph *)ctx->vgraphs[victim]);
        ctx->vgraphs.erase(ctx->vgraphs.begin() + (long)victim);
    }
    g->last_use = ++ctx->vgraph_clock;
    ctx->vgraphs.push_back(g);
    return g;
}

// per-call output buffers: chunk_status[c] (that call's n_chunks), masks32[c] (its n_proofs x ext x 32, may be null), mask_present[c]
int32_t bpp_vbatch_run_multi(bpp_vbatch *vb, int32_t *const *chunk_status, uint8_t *const *masks32, uint8_t *const *mask_present) {
    if (!vb || !chunk_status) return BPP_INVALID_ARGUMENT;
    for (size_t c = 0; c < vb->calls.size(); c++) if (!chunk_status[c]) return BPP_INVALID_ARGUMENT;
    bpp_gens *g = vb->g;
    bpp_ctx *ctx = g->ctx;
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    VWork *w = vb->w;
    const int ext = g->ext;
    const size_t n = vb->n_proofs;
    const VLaunch L = make_launch(vb);
    const VDims &d = L.d;
    const VBuffers &b = L.b;
    const bool dev_replay = L.dev_replay;
    const bool prep = vb->any_msm || vb->any_masks;

    ctx->clear_marks();
    const bool fused = ctx->device_weights && dev_replay && ctx->use_graphs && !ctx->phase_timing;
    if (fused) {
        cudaError_t ge = cudaSuccess;
        VGraph *vg = vgraph_get(vb, L, true, &ge);
        if (!vg) return cuda_fail(ctx, ge, "verification graph capture");
        BPP_CUDA(ctx, cudaGraphLaunch(vg->ex[0], st));
        ctx->launches += vg->kernels[0];
        ctx->graph_launches += 1;
    } else if (ctx->use_graphs && !ctx->phase_timing) {
        cudaError_t ge = cudaSuccess;
        VGraph *vg = vgraph_get(vb, L, false, &ge);
        if (!vg) return cuda_fail(ctx, ge, "verification graph capture");
        if (vg->ex[0]) {
            BPP_CUDA(ctx, cudaGraphLaunch(vg->ex[0], st));
            BPP_CUDA(ctx, cudaEventRecord(ctx->ev_mid, st));
        }
        BPP_CUDA(ctx, cudaGraphLaunch(vg->ex[1], st));
        if (dev_replay) {       // the host hashes the weight transcripts while the device runs section B
            if (ctx->throughput_mode) BPP_CUDA(ctx, wait_sleeping(ctx->ev_mid, ctx->nap_ns, ctx->adaptive_wait ? &ctx->wait_ema_ns[0] : nullptr, n));
            else BPP_CUDA(ctx, cudaEventSynchronize(ctx->ev_mid));
            auto tw = std::chrono::steady_clock::now();
            compute_weights(vb);
            ctx->host_ms[3] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tw).count();
        }
        BPP_CUDA(ctx, cudaGraphLaunch(vg->ex[2], st));
        ctx->launches += vg->kernels[0] + vg->kernels[1] + vg->kernels[2];
        ctx->graph_launches += vg->ex[0] ? 3 : 2;
    } else {
    // The decompression (k_point.cu) only feeds the bucket sums, so it runs on the side stream next to the transcript
    // replay and the scalar prep chain; with phase timing on everything stays on one stream so that the per-phase events
    // mean what they say (replay, then decompress, then the prep).
    const bool overlap = !ctx->phase_timing && vb->n_pts && prep;
    if (overlap) {
        BPP_CUDA(ctx, cudaEventRecord(ctx->ev_fork, st));
        BPP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
        enqueue_decompress(vb, L, ctx->stream2, &ctx->launches);
        BPP_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->stream2));
    }
    ctx->mark(0);
    if (dev_replay) {
        launch_replay(st, d, L.rb, L.replay_kernel, &ctx->launches);
        BPP_CUDA(ctx, cudaMemcpyAsync(vb->mid(), w->d_mid.p, vb->mid_bytes, cudaMemcpyDeviceToHost, st));
        BPP_CUDA(ctx, cudaEventRecord(ctx->ev_mid, st));
    }
    ctx->mark(1);
    if (!overlap && vb->n_pts) enqueue_decompress(vb, L, st, &ctx->launches);
    ctx->mark(2);
    if (prep) {
        launch_verify_prep(st, d, b, vb->any_vec ? 1u : 0u, vb->max_rounds, &ctx->launches, ctx->phase_timing ? &ctx->ph[3] : nullptr);
        if (ctx->phase_timing) { ctx->ph_set[3] = true; ctx->ph_set[4] = true; }
    }
    if (dev_replay) {       // the host hashes the weight transcripts while the device runs the weight-free scalar prep
        BPP_CUDA(ctx, cudaEventSynchronize(ctx->ev_mid));
        compute_weights(vb);
    }
    if (vb->any_msm) {
        BPP_CUDA(ctx, cudaMemcpyAsync(w->d_weights.p, w->h_weights.p, 64 * (n + (vb->merged ? vb->n_chunks : 0)), cudaMemcpyHostToDevice, st));
        ctx->mark(5);
        BPP_CUDA(ctx, cudaMemsetAsync(b.weight_zero, 0, vb->n_chunks, st));
        launch_verify_weigh(st, d, b, vb->max_static, &ctx->launches);
        ctx->mark(6);
        if (overlap) BPP_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_join, 0));
        enqueue_msm(vb, L, st, &ctx->launches, ctx->phase_timing ? &ctx->ph[7] : nullptr, vb->merged);
        if (ctx->phase_timing) for (int i = 7; i <= 10; i++) ctx->ph_set[i] = true;
        ctx->mark(11);
        BPP_CUDA(ctx, cudaMemcpyAsync(w->h_out.as<uint8_t>() + vb->ho_ident, w->d_ident.p, 2 * vb->n_chunks, cudaMemcpyDeviceToHost, st));
    } else if (overlap) {
        BPP_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_join, 0));
    }
    BPP_CUDA(ctx, cudaGetLastError());
    if (vb->n_pts) BPP_CUDA(ctx, cudaMemcpyAsync(w->h_out.as<uint8_t>() + vb->ho_ok, w->d_ok.p, vb->n_pts, cudaMemcpyDeviceToHost, st));
    if (vb->any_masks) BPP_CUDA(ctx, cudaMemcpyAsync(w->h_out.as<uint8_t>() + vb->ho_masks, w->d_masks.p, 32 * n * (size_t)ext, cudaMemcpyDeviceToHost, st));
    }
    if (ctx->throughput_mode) {         // sleep until the pass is done: with many lanes per GPU spinning threads starve each other
        BPP_CUDA(ctx, cudaEventRecord(ctx->ev_done, st));
        BPP_CUDA(ctx, wait_sleeping(ctx->ev_done, ctx->nap_ns, ctx->adaptive_wait ? &ctx->wait_ema_ns[1] : nullptr, n));
    } else {
        BPP_CUDA(ctx, cudaStreamSynchronize(st));
    }
    // A weight that reduced to zero (the reference's random_not_zero draws again, range_proof.rs:894 / scalar_protocol.rs:23-30): the
    // chunk's weight transcript is redone one transcript at a time with the redraw, and sections B + C are repeated kernel by kernel.
    // Probability 2^-252 per weight; BPP test hook 1 forces the path so that it stays tested.
    if (vb->any_msm && !fused) {
        const uint8_t *wz = w->h_out.as<uint8_t>() + vb->ho_ident + vb->n_chunks;
        std::vector<size_t> redo;
        for (size_t c = 0; c < vb->n_chunks; c++)
            if (!vb->hc[c].pre_rc && (wz[c] || (ctx->test_hooks & 1))) redo.push_back(c);
        if (!redo.empty() && !vb->caller_challenges) {
            compute_weights(vb, &redo, true);
            cudaError_t e1 = enqueue_section(vb, L, 1, &ctx->launches);
            cudaError_t e2 = enqueue_section(vb, L, 2, &ctx->launches);
            BPP_CUDA(ctx, e1);
            BPP_CUDA(ctx, e2);
            BPP_CUDA(ctx, cudaStreamSynchronize(st));
        }
    }
    // Merged check: the identity means that every chunk's sum vanishes (a non-zero chunk sum survives the random combination with
    // probability 2^-252); anything else is settled chunk by chunk with the scalars as they are (chunk c's sum is rho_c times the
    // reference's, rho_c != 0).  Test hook 2 forces the chunk-by-chunk pass.
    if (vb->any_msm && vb->merged) {
        uint8_t *id = w->h_out.as<uint8_t>() + vb->ho_ident;
        if (id[0] && !(ctx->test_hooks & 2)) {
            memset(id, 1, vb->n_chunks);
        } else {
            enqueue_msm(vb, L, st, &ctx->launches, nullptr, false);
            BPP_CUDA(ctx, cudaGetLastError());
            BPP_CUDA(ctx, cudaMemcpyAsync(id, w->d_ident.p, vb->n_chunks, cudaMemcpyDeviceToHost, st));
            BPP_CUDA(ctx, cudaStreamSynchronize(st));
            ctx->merged_fallbacks++;
        }
    }
    vb->ran = true;
    ctx->io_bytes[0] = vb->blob_bytes + (vb->device_replay ? 0 : 32 * (size_t)vb->n_chal) + (vb->any_msm && !fused ? 64 * n : 0);
    ctx->io_bytes[1] = (dev_replay ? vb->mid_bytes : 0) + vb->n_pts + (vb->any_msm ? vb->n_chunks : 0) + (vb->any_masks ? 32 * n * (size_t)ext : 0);

    // ---- resolve per-chunk status with the reference's precedence
    const uint8_t *okf = w->h_out.as<uint8_t>() + vb->ho_ok;
    const uint8_t *ident = w->h_out.as<uint8_t>() + vb->ho_ident;
    const uint8_t *hmasks = w->h_out.as<uint8_t>() + vb->ho_masks;
    size_t call = 0;
    for (size_t c = 0; c < vb->n_chunks; c++) {
        const HChunk &hc = vb->hc[c];
        while (c >= vb->calls[call].chunk0 + vb->calls[call].a.n_chunks) call++;
        const HCall &hcall = vb->calls[call];
        uint8_t *c_masks = masks32 ? masks32[call] : nullptr, *c_present = mask_present ? mask_present[call] : nullptr;
        int32_t rc = hc.pre_rc;
        if (!rc) {   // a commitment that is not a valid encoding can not be a RangeStatement commitment (a point)
            for (size_t i = hc.lo; i < hc.hi && !rc; i++) {
                const HProof &p = vb->hp[i];
                for (uint32_t j = 0; j < p.m; j++)
                    if (!okf[p.pt_off + 3 + 2 * p.rounds + j]) { rc = BPP_INVALID_ARGUMENT; break; }
            }
        }
        for (size_t i = hc.lo; i < hc.hi && !rc; i++)                                       // loop 1 (:816-850)
            if (vb->flag(i) & 1) rc = BPP_VERIFICATION_FAILED;
        for (size_t i = hc.lo; i < hc.hi && !rc; i++) {                                     // loop 2, proof order
            const HProof &p = vb->hp[i];
            for (uint32_t j = 0; j < 3 + 2 * (uint32_t)p.rounds; j++)
                if (!okf[p.pt_off + j]) { rc = BPP_INVALID_ARGUMENT; break; }             // :859-866
            if (!rc) rc = p.loop2_rc;                                                       // :875-888
            if (!rc && (vb->flag(i) & 2)) rc = BPP_VERIFICATION_FAILED;                     // y == 1: (y - 1) is not invertible
        }
        if (!rc && vb->action != BPP_RECOVER_ONLY && !ident[c]) rc = BPP_VERIFICATION_FAILED;   // :1057-1061
        chunk_status[call][c - hcall.chunk0] = rc;
        // Vec<Option<ExtendedMask>>
        for (size_t i = hc.lo; i < hc.end; i++) {
            const size_t li = i - hcall.proof0;
            bool have = !rc && i < hc.hi && vb->action != BPP_VERIFY_ONLY && vb->hp[i].has_seed;
            if (c_present) c_present[li] = have ? 1 : 0;
            if (c_masks) {
                if (have) memcpy(c_masks + 32 * li * (size_t)ext, hmasks + 32 * i * (size_t)ext, 32 * (size_t)ext);
                else memset(c_masks + 32 * li * (size_t)ext, 0, 32 * (size_t)ext);
            }
        }
    }
    return BPP_OK;
}

// the calls' outputs concatenated in order: chunk_status[n_chunks], masks32[n_proofs x ext x 32], mask_present[n_proofs]
int32_t bpp_vbatch_run(bpp_vbatch *vb, int32_t *chunk_status, uint8_t *masks32, uint8_t *mask_present) {
    if (!vb || !chunk_status) return BPP_INVALID_ARGUMENT;
    const size_t nc = vb->calls.size();
    std::vector<int32_t *> st(nc);
    std::vector<uint8_t *> mk(nc), mp(nc);
    for (size_t c = 0; c < nc; c++) {
        const HCall &h = vb->calls[c];
        st[c] = chunk_status + h.chunk0;
        mk[c] = masks32 ? masks32 + 32 * h.proof0 * (size_t)vb->g->ext : nullptr;
        mp[c] = mask_present ? mask_present + h.proof0 : nullptr;
    }
    return bpp_vbatch_run_multi(vb, st.data(), masks32 ? mk.data() : nullptr, mask_present ? mp.data() : nullptr);
}

size_t bpp_vbatch_call_count(const bpp_vbatch *vb) { return vb ? vb->calls.size() : 0; }

// `&mut Transcript` semantics of the reference: every transcript of a call that reached loop 1 is advanced, up to and
// including the first proof whose replay failed.  transcripts: the proofs of call `call` x 203 bytes, updated in place.
int32_t bpp_vbatch_transcripts_call(const bpp_vbatch *vb, size_t call, uint8_t *transcripts) {
    if (!vb || !transcripts || call >= vb->calls.size()) return BPP_INVALID_ARGUMENT;
    if (vb->caller_challenges) return fail(vb->g->ctx, BPP_INVALID_ARGUMENT, "the caller keeps the transcripts in the challenge-input form");
    if (vb->device_replay && !vb->ran) return fail(vb->g->ctx, BPP_INVALID_ARGUMENT, "bpp_vbatch_run has not been called");
    const HCall &hcall = vb->calls[call];
    for (size_t c = hcall.chunk0; c < hcall.chunk0 + hcall.a.n_chunks; c++) {
        const HChunk &hc = vb->hc[c];
        if (hc.pre_rc) continue;
        for (size_t i = hc.lo; i < hc.hi; i++) {
            memcpy(transcripts + BPP_TRANSCRIPT_BYTES * (i - hcall.proof0), vb->tstate(i), BPP_TRANSCRIPT_BYTES);
            if (vb->flag(i) & 1) break;
        }
    }
    return BPP_OK;
}
// all calls of the pass, concatenated (n_proofs x 203 B)
int32_t bpp_vbatch_transcripts(const bpp_vbatch *vb, uint8_t *transcripts) {
    if (!vb || !transcripts) return BPP_INVALID_ARGUMENT;
    for (size_t c = 0; c < vb->calls.size(); c++) {
        int32_t rc = bpp_vbatch_transcripts_call(vb, c, transcripts + BPP_TRANSCRIPT_BYTES * vb->calls[c].proof0);
        if (rc) return rc;
    }
    return BPP_OK;
}

int32_t bpp_verify_chunks(bpp_gens *g, const bpp_verify_args *args, int32_t *chunk_status, uint8_t *masks32, uint8_t *mask_present) {
    bpp_vbatch *vb = nullptr;
    int32_t rc = bpp_vbatch_create(g, args, &vb);
    if (rc) return rc;
    auto t0 = std::chrono::steady_clock::now();
    rc = bpp_vbatch_run(vb, chunk_status, masks32, mask_present);
    g->ctx->host_ms[5] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (!rc) rc = bpp_vbatch_transcripts(vb, args->transcripts);
    bpp_vbatch_destroy(vb);
    return rc;
}

// Challenge-input form: loop 1 of RangeProof::verify (:816-850) and the weight draw (:894) stay with a caller that
// keeps merlin::Transcript itself; everything from the point decompression on runs here.
int32_t bpp_verify_chunks_ch(bpp_gens *g, const bpp_verify_args *args, const bpp_verify_challenges *ch, int32_t *chunk_status, uint8_t *masks32,
                             uint8_t *mask_present) {
    if (!ch) return BPP_INVALID_ARGUMENT;
    bpp_vbatch *vb = nullptr;
    int32_t rc = vbatch_create_impl(g, 1, &args, &ch, &vb);
    if (rc) return rc;
    rc = bpp_vbatch_run(vb, chunk_status, masks32, mask_present);
    bpp_vbatch_destroy(vb);
    return rc;
}

} // extern "C"
