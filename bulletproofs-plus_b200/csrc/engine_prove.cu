// Batched prover entry point: bpp_prove_batch = P calls of RangeProof::prove_with_rng
// (/root/reference/src/range_proof.rs:232-608) for statements of one shape, advancing in lock-step.
//
// Host (this file): argument checks (:239-284), RangeProofTranscript (transcripts.rs:59-194) with the witness-keyed
// TranscriptRng, nonces (utils/generic.rs:30-60), every random draw in the reference's order, the alpha bookkeeping and the
// final responses (:590-594), proof serialisation (:1120-1150).  Device (k_prove.cu + k_msm.cu): A, every L / R, the generator
// and scalar folding, A1 and B.  One host<->device meeting per round: 64 bytes (L, R) down, d_L / d_R / e (<= 13 scalars) up
// per proof.  The caller supplies the bytes its external RNG would have delivered (32 per TranscriptRng rebuild,
// log2(n*m) + 3 rebuilds), so a seeded RNG reproduces the reference's proof bytes.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <functional>
#include "engine.hpp"
#include "hash.cuh"
#include "strobe_n.hpp"

using namespace bpp;

// host scalar arithmetic on 64-bit limbs (host_keccak4.cpp); results are canonical, identical to the shared 32-bit-limb code
extern "C" void bpp_host_sc_mul64(const uint8_t a32[32], const uint8_t b32[32], uint8_t out32[32]);
extern "C" void bpp_host_sc_from_wide64(const uint8_t in64[64], uint8_t out32[32]);

namespace {

#define LBL(s) (const uint8_t *)(s), (sizeof(s) - 1)

inline sc sc_load(const uint8_t *b) { return sc_frombytes_raw(b); }
inline void sc_store(uint8_t *b, const sc &a) { sc_tobytes(b, a); }
inline sc sc_reduce_bytes(const uint8_t *b) { return sc_reduce256(sc_frombytes_raw(b)); }
inline bool zero32(const uint8_t *p) { uint8_t r = 0; for (int i = 0; i < 32; i++) r |= p[i]; return r == 0; }

inline sc hmul(const sc &a, const sc &b) {
    sc r;
    bpp_host_sc_mul64(reinterpret_cast<const uint8_t *>(a.v), reinterpret_cast<const uint8_t *>(b.v), reinterpret_cast<uint8_t *>(r.v));
    return r;
}
inline sc hpow(sc base, uint64_t e) {          // square-and-multiply
    sc r = sc_one();
    while (e) {
        if (e & 1) r = hmul(r, base);
        base = hmul(base, base);
        e >>= 1;
    }
    return r;
}
inline sc wide_to_sc(const uint8_t in[64]) {
    sc r;
    bpp_host_sc_from_wide64(in, reinterpret_cast<uint8_t *>(r.v));
    return r;
}

// memset that the optimiser may not drop (the buffers are dead afterwards)
inline void secure_zero(void *p, size_t n) {
    if (!p || !n) return;
    memset(p, 0, n);
    __asm__ __volatile__("" : : "r"(p) : "memory");
}
// runs a wipe on every exit path of bpp_prove_batch (normal return, argument errors discovered late, CUDA failures)
template <class F> struct ScopeExit {
    F f;
    explicit ScopeExit(F fn) : f(fn) {}
    ~ScopeExit() { f(); }
    ScopeExit(const ScopeExit &) = delete;
    ScopeExit &operator=(const ScopeExit &) = delete;
};

struct PProof {
    int32_t rc = 0;
    bool live = false;             // takes part in the device batch
    uint32_t slot = 0;             // index inside the device batch
    Merlin t;
    MerlinRng rng;
    std::vector<uint8_t> witness;  // LE64(v) || r[0..ext) per opening (transcripts.rs:91-109)
    bool has_seed = false;
    uint8_t seed[32];
    const uint8_t *rng_bytes = nullptr;
    size_t rng_used = 0;
    sc alpha[BPP_MAX_EXT];
    sc y, z, e_final;
    std::vector<sc> e_round, einv_round, dL, dR;   // per round (dL/dR: rounds x ext)
    sc r, s, d[BPP_MAX_EXT], eta[BPP_MAX_EXT];
    uint8_t A[32], A1[32], B[32];
    std::vector<uint8_t> LR;           // rounds x 64
};

// utils/generic.rs:30-60
sc nonce(const uint8_t seed[32], const char *label, bool have_j, uint32_t j, bool have_k, uint32_t k) {
    uint8_t key[43];
    size_t kl = 0;
    key[kl++] = 0;
    memcpy(key + kl, seed, 32); kl += 32;
    if (have_j) { key[kl++] = 'j'; for (int i = 0; i < 4; i++) key[kl++] = (uint8_t)(j >> (8 * i)); }
    if (have_k) { key[kl++] = 'k'; for (int i = 0; i < 4; i++) key[kl++] = (uint8_t)(k >> (8 * i)); }
    uint8_t h[64];
    blake2b_keyed_personal_empty(h, key, kl, (const uint8_t *)label, strlen(label));
    return wide_to_sc(h);
}
// Scalar::random_not_zero over the TranscriptRng (protocols/scalar_protocol.rs:20-30)
sc random_not_zero(MerlinRng &rng) {
    for (;;) {
        uint8_t wide[64];
        rng.fill(wide, 64);
        sc v = wide_to_sc(wide);
        if (!sc_is_zero(v)) return v;
    }
}
// transcripts.rs:185-194: transcript.build_rng().rekey_with_witness_bytes("witness", w).finalize(external rng)
void rebuild_rng(PProof &p) {
    p.rng.build(p.t, p.witness.data(), p.witness.size(), true, p.rng_bytes + 32 * p.rng_used);
    p.rng_used++;
}
bool append_point(Merlin &t, const uint8_t *label, size_t ll, const uint8_t pt[32]) {
    if (zero32(pt)) return false;
    t.append_message(label, ll, pt, 32);
    return true;
}
bool challenge(Merlin &t, const uint8_t *label, size_t ll, sc &out) {
    uint8_t buf[64];
    t.challenge_bytes(label, ll, buf, 64);
    out = wide_to_sc(buf);
    return !sc_is_zero(out);
}

// ---- eight proofs of a call through ONE vectorised sponge (strobe_n.hpp).  The proofs of a bpp_prove_batch call have one shape, so their
// transcripts and TranscriptRngs run the same operations at the same sponge positions; ~45 Keccak-f per proof (the largest host cost of
// the prover) become ~45 eight-way permutations per eight proofs.  A stage takes this path only when its eight lanes agree on everything
// that steers the operation sequence (alive, sponge positions, rebuild count, nonce source) and nothing exceptional can happen in it
// (identity points are seen up front; a zero challenge or a zero random scalar -- probability 2^-252 -- is seen afterwards, BEFORE anything
// is written back); otherwise the stage returns false with the proofs untouched and the caller runs the one-at-a-time code on each lane.
// The canonical state stays in PProof between stages.  BPP_PROVE_LOCKSTEP=0 turns the path off (tests compare both byte for byte).
constexpr int LK = 8;
bool lockstep_enabled() {
    static const bool on = [] { const char *e = getenv("BPP_PROVE_LOCKSTEP"); return !(e && atoi(e) == 0); }();
    return on;
}
struct Lock8 {
    PProof *p[LK];
    MerlinN<LK> t;
    StrobeN<LK> r;
    static bool same_pos(const Strobe128 &a, const Strobe128 &b) { return a.pos == b.pos && a.pos_begin == b.pos_begin && a.cur_flags == b.cur_flags; }
    bool uniform(bool with_rng) const {
        const PProof &q0 = *p[0];
        for (int j = 0; j < LK; j++) {
            const PProof &q = *p[j];
            if (q.rc || q.has_seed != q0.has_seed || q.rng_used != q0.rng_used || q.witness.size() != q0.witness.size()) return false;
            if (!same_pos(q.t.s, q0.t.s) || (with_rng && !same_pos(q.rng.s, q0.rng.s))) return false;
        }
        return true;
    }
    void load_t() { for (int j = 0; j < LK; j++) t.s.load_lane(j, p[j]->t.s); }
    void load_r() { for (int j = 0; j < LK; j++) r.load_lane(j, p[j]->rng.s); }
    void store_t() const { for (int j = 0; j < LK; j++) t.s.store_lane(j, p[j]->t.s); }
    void store_r() const { for (int j = 0; j < LK; j++) r.store_lane(j, p[j]->rng.s); }
    // rebuild_rng of every lane (from the lock-step transcript as it stands); the callers bump rng_used when they commit
    void rebuild() {
        const uint8_t *w[LK], *e[LK];
        for (int j = 0; j < LK; j++) { w[j] = p[j]->witness.data(); e[j] = p[j]->rng_bytes + 32 * p[j]->rng_used; }
        t.build_rng(r, w, p[0]->witness.size(), e);
    }
    bool draw(sc out[LK]) {                        // Scalar::random_not_zero per lane; false = some lane would have to draw again
        uint8_t buf[LK][64], *ptr[LK];
        for (int j = 0; j < LK; j++) ptr[j] = buf[j];
        rng_fill_each(r, ptr, 64);
        bool ok = true;
        for (int j = 0; j < LK; j++) { out[j] = wide_to_sc(buf[j]); ok = ok && !sc_is_zero(out[j]); }
        secure_zero(buf, sizeof buf);
        return ok;
    }
    bool challenge(const uint8_t *label, size_t ll, sc out[LK]) {
        uint8_t buf[LK][64], *ptr[LK];
        for (int j = 0; j < LK; j++) ptr[j] = buf[j];
        t.challenge_each(label, ll, ptr, 64);
        bool ok = true;
        for (int j = 0; j < LK; j++) { out[j] = wide_to_sc(buf[j]); ok = ok && !sc_is_zero(out[j]); }
        return ok;
    }
    void wipe() { r.wipe(); }
};

// device + pinned buffers of the prover, kept per ctx across calls (grow-only; cudaMalloc / cudaFree of the ~200 MB bucket
// scratch cost more than the proving itself); the secrets in them are wiped at the end of every call
struct ProveWS {
    DevBuf d_offs, d_a, d_b, d_ypow, d_yinv2, d_yz, d_dlr, d_e, d_fsc, d_folded, d_mscal, d_pidx, d_segoff, d_scratch, d_res, d_enc, d_ab;
    DevBuf d_sg, d_sh, d_gidx, d_rs;      // fixed-base path
    PinBuf h_io;
    void release() {
        for (DevBuf *b : {&d_offs, &d_a, &d_b, &d_ypow, &d_yinv2, &d_yz, &d_dlr, &d_e, &d_fsc, &d_folded, &d_mscal, &d_pidx, &d_segoff, &d_scratch,
                          &d_res, &d_enc, &d_ab, &d_sg, &d_sh, &d_gidx, &d_rs})
            b->release();
        h_io.release();
    }
};

} // namespace

namespace bpp {
void prove_ws_free(bpp_ctx *ctx) {
    if (ctx->prove_ws) { ((ProveWS *)ctx->prove_ws)->release(); delete (ProveWS *)ctx->prove_ws; ctx->prove_ws = nullptr; }
}
}

extern "C" {

// test hook, host only: the lock-step sponge (strobe_n.hpp) against the one-at-a-time Merlin of hash.cuh on eight transcripts in the same
// state positions.  Script per lane: append_message("L", msg) -> challenge_bytes("e", 64) -> TranscriptRng(witness, ext32) -> fill_bytes(64)
// -> fill_bytes(64).  out_*: per lane [203 transcript | 64 challenge | 203 rng state | 128 rng bytes] (598 bytes).
int32_t bpp_host_lockstep_selftest(const uint8_t *states203, const uint8_t *msgs, size_t msg_len, const uint8_t *witness, size_t wlen,
                                   const uint8_t *ext32, uint8_t *out_scalar, uint8_t *out_lockstep) {
    if (!states203 || !msgs || !witness || !ext32 || !out_scalar || !out_lockstep) return BPP_INVALID_ARGUMENT;
    const size_t rec = 2 * BPP_TRANSCRIPT_BYTES + 64 + 128;
    for (int j = 0; j < LK; j++) {
        Merlin t;
        t.s.load(states203 + BPP_TRANSCRIPT_BYTES * j);
        if (t.s.pos != states203[200] || t.s.pos_begin != states203[201] || t.s.cur_flags != states203[202]) return BPP_INVALID_ARGUMENT;   // lanes must agree
        uint8_t *o = out_scalar + rec * j;
        t.append_message(LBL("L"), msgs + msg_len * j, msg_len);
        t.challenge_bytes(LBL("e"), o + BPP_TRANSCRIPT_BYTES, 64);
        t.s.store(o);
        MerlinRng r;
        r.build(t, witness + wlen * j, wlen, true, ext32 + 32 * j);
        r.fill(o + 2 * BPP_TRANSCRIPT_BYTES + 64, 64);
        r.fill(o + 2 * BPP_TRANSCRIPT_BYTES + 128, 64);
        r.s.store(o + BPP_TRANSCRIPT_BYTES + 64);
    }
    MerlinN<LK> t;
    StrobeN<LK> r;
    const uint8_t *pm[LK], *pw[LK], *pe[LK];
    uint8_t *pc[LK], *pf0[LK], *pf1[LK];
    for (int j = 0; j < LK; j++) {
        Strobe128 s1;
        s1.load(states203 + BPP_TRANSCRIPT_BYTES * j);
        t.s.load_lane(j, s1);
        uint8_t *o = out_lockstep + rec * j;
        pm[j] = msgs + msg_len * j; pw[j] = witness + wlen * j; pe[j] = ext32 + 32 * j;
        pc[j] = o + BPP_TRANSCRIPT_BYTES; pf0[j] = o + 2 * BPP_TRANSCRIPT_BYTES + 64; pf1[j] = pf0[j] + 64;
    }
    t.append_each(LBL("L"), pm, msg_len);
    t.challenge_each(LBL("e"), pc, 64);
    t.build_rng(r, pw, wlen, pe);
    rng_fill_each(r, pf0, 64);
    rng_fill_each(r, pf1, 64);
    for (int j = 0; j < LK; j++) {
        Strobe128 s1;
        t.s.store_lane(j, s1);
        s1.store(out_lockstep + rec * j);
        r.store_lane(j, s1);
        s1.store(out_lockstep + rec * j + BPP_TRANSCRIPT_BYTES + 64);
    }
    return BPP_OK;
}

size_t bpp_proof_size(int32_t extension_degree, int32_t rounds) {
    return 1 + 32 * ((size_t)extension_degree + 5 + 2 * (size_t)rounds);
}

int32_t bpp_prove_batch(bpp_gens *g, const bpp_prove_args *a, uint8_t *proofs_out, size_t proof_stride, int32_t *status) {
    if (!g || !a || !status) return BPP_INVALID_ARGUMENT;
    bpp_ctx *ctx = g->ctx;
    const size_t P0 = a->n_proofs;
    if (P0 == 0) return BPP_OK;
    if (!proofs_out || !a->commitments32 || !a->values || !a->blindings32 || !a->min_values || !a->min_present || !a->transcripts || !a->rng_bytes)
        return fail(ctx, BPP_INVALID_ARGUMENT, "null argument");
    const uint32_t n = (uint32_t)g->n, ext = (uint32_t)g->ext;
    const int64_t m64 = a->aggregation;
    // RangeStatement::init (range_statement.rs:42-61)
    if (m64 <= 0 || (m64 & (m64 - 1))) return fail(ctx, BPP_INVALID_ARGUMENT, "Number of commitments must be a power of two");
    if (m64 > g->M) return fail(ctx, BPP_INVALID_ARGUMENT, "Not enough generators for this statement");
    const uint32_t m = (uint32_t)m64, N = n * m;
    uint32_t rounds = 0;
    while ((1u << rounds) < N) rounds++;
    if (rounds > BPP_MAX_ROUNDS) return fail(ctx, BPP_SIZE_OVERFLOW, "vector too long");
    const size_t need_rng = 32 * ((size_t)rounds + 3), plen = bpp_proof_size((int32_t)ext, (int32_t)rounds);
    if (a->rng_stride < need_rng) return fail(ctx, BPP_INVALID_LENGTH, "rng_stride must cover 32 bytes per TranscriptRng rebuild (log2(n*m) + 3)");
    if (proof_stride < plen) return fail(ctx, BPP_INVALID_LENGTH, "proof_stride too small");
    if (P0 * (uint64_t)N >= (1u << 28)) return fail(ctx, BPP_SIZE_OVERFLOW, "batch too large");
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    // debug aid (BPP_PROVE_TRACE=1): wall time spent in the host stages (Fiat-Shamir, nonces, scalar bookkeeping) of this call
    static const bool prove_trace = getenv("BPP_PROVE_TRACE") != nullptr;
    double host_stage_ms = 0;
    const auto t_call0 = std::chrono::steady_clock::now();
    auto host_stage = [&](size_t n, size_t grain, const std::function<void(size_t)> &f) {
        const auto t0 = std::chrono::steady_clock::now();
        ctx->workers().run(n, grain, f);
        host_stage_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };

    // a stage over n proofs: groups of LK through the lock-step sponge when `lock(first)` takes them, else one at a time
    auto host_stage_lk = [&](size_t n, const std::function<bool(size_t)> &lock, const std::function<void(size_t)> &one) {
        host_stage((n + LK - 1) / LK, 1, [&](size_t gi) {
            const size_t lo = LK * gi, hi = std::min(n, lo + LK);
            if (hi - lo == LK && lockstep_enabled() && lock(lo)) return;
            for (size_t s = lo; s < hi; s++) one(s);
        });
    };

    std::vector<PProof> pp(P0);
    // Everything below that holds witness-derived data is wiped when this function is left, whichever way (the reference keeps these in
    // Zeroizing<..>: range_proof.rs:300-301, :325, :438-464, :542-571): the host copies (per-proof state, bit offsets, a[0] / b[0], the
    // pinned staging buffer that carried alpha, d_L / d_R, r, s, d, eta) and the device copies.
    std::vector<uint64_t> offs;
    std::vector<uint8_t> ab;
    struct { uint8_t *hio = nullptr; size_t hio_bytes = 0; std::vector<std::pair<void *, size_t>> dev; cudaStream_t st = nullptr; } wipe;
    ScopeExit wipe_guard([&]() {
        for (PProof &p : pp) {
            secure_zero(p.witness.data(), p.witness.size());
            secure_zero(p.seed, sizeof p.seed);
            secure_zero(p.alpha, sizeof p.alpha); secure_zero(p.d, sizeof p.d); secure_zero(p.eta, sizeof p.eta);
            secure_zero(&p.r, sizeof p.r); secure_zero(&p.s, sizeof p.s);
            secure_zero(p.dL.data(), sizeof(sc) * p.dL.size()); secure_zero(p.dR.data(), sizeof(sc) * p.dR.size());
            secure_zero(&p.rng, sizeof p.rng);               // TranscriptRng keyed with the witness bytes
        }
        secure_zero(offs.data(), 8 * offs.size());
        secure_zero(ab.data(), ab.size());
        if (!wipe.dev.empty()) {
            for (auto &d : wipe.dev) if (d.first && d.second) cudaMemsetAsync(d.first, 0, d.second, wipe.st);
            cudaStreamSynchronize(wipe.st);                 // also orders the staging buffer's last DMA before its wipe
            cudaGetLastError();
        }
        secure_zero(wipe.hio, wipe.hio_bytes);
    });
    // ---- :264-271 value range, :275-284 opening == commitment (device commit, compared as canonical encodings)
    for (size_t i = 0; i < P0; i++) {
        PProof &p = pp[i];
        p.has_seed = a->seed_present && a->seed_nonces32 && a->seed_present[i];
        if (p.has_seed && m > 1) p.rc = BPP_INVALID_ARGUMENT;     // "Mask recovery is not supported with an aggregated statement"
        if (!p.rc && n < 64)
            for (uint32_t j = 0; j < m; j++)
                if ((a->values[i * m + j] >> n) > 0) p.rc = BPP_INVALID_LENGTH;
        for (size_t k = 0; k < (size_t)m * ext && !p.rc; k++)
            if (!host_sc_is_canonical(a->blindings32 + 32 * (i * m * ext + k))) p.rc = BPP_INVALID_ARGUMENT;
    }
    {
        std::vector<uint8_t> recommit(32 * P0 * m);
        std::vector<uint8_t> bl(a->blindings32, a->blindings32 + 32 * P0 * m * ext);
        for (size_t i = 0; i < P0; i++)
            if (pp[i].rc) memset(bl.data() + 32 * i * m * ext, 0, 32 * (size_t)m * ext);
        int32_t rc = bpp_pedersen_commit_batch(g, P0 * m, a->values, bl.data(), (int32_t)ext, recommit.data());
        secure_zero(bl.data(), bl.size());
        if (rc) return rc;
        for (size_t i = 0; i < P0; i++)
            if (!pp[i].rc && memcmp(recommit.data() + 32 * i * m, a->commitments32 + 32 * i * m, 32 * (size_t)m)) pp[i].rc = BPP_INVALID_ARGUMENT;
    }

    // ---- RangeProofTranscript::new (:287-297), bit offsets (:300-322), alpha (:325-333)
    std::vector<size_t> live;
    auto fill_witness = [&](PProof &p, size_t i) {
        p.rng_bytes = a->rng_bytes + a->rng_stride * i;
        p.witness.resize((size_t)m * (8 + 32 * ext));
        for (uint32_t j = 0; j < m; j++) {
            uint8_t *w = p.witness.data() + (size_t)j * (8 + 32 * ext);
            le64_bytes(w, a->values[i * m + j]);
            memcpy(w + 8, a->blindings32 + 32 * ((i * m + j) * ext), 32 * (size_t)ext);
        }
    };
    auto below_minimum = [&](size_t i) {                                                    // :309-311
        for (uint32_t j = 0; j < m; j++)
            if (a->min_present[i * m + j] && a->values[i * m + j] < a->min_values[i * m + j]) return true;
        return false;
    };
    auto init_finish = [&](PProof &p, size_t i) {        // what follows the rng build; alpha from the rng was drawn by the caller
        if (p.has_seed) {
            sc_store(p.seed, sc_reduce_bytes(a->seed_nonces32 + 32 * i));
            for (uint32_t k = 0; k < ext; k++) p.alpha[k] = nonce(p.seed, "alpha", false, 0, true, k);
        }
        p.e_round.resize(rounds); p.einv_round.resize(rounds); p.dL.resize((size_t)rounds * ext); p.dR.resize((size_t)rounds * ext); p.LR.resize(64 * (size_t)rounds);
    };
    bool gens_nonzero = !zero32(g->h());
    for (uint32_t k = 0; k < ext; k++) gens_nonzero = gens_nonzero && !zero32(g->g(k));
    auto init_one = [&](size_t i) {
        PProof &p = pp[i];
        if (p.rc) return;
        p.t.s.load(a->transcripts + BPP_TRANSCRIPT_BYTES * i);
        p.t.append_message(LBL("dom-sep"), LBL("Bulletproofs+ Range Proof"));
        bool ok = append_point(p.t, LBL("H"), g->h());
        for (uint32_t k = 0; k < ext && ok; k++) ok = append_point(p.t, LBL("G"), g->g(k));
        if (!ok) { p.rc = BPP_VERIFICATION_FAILED; return; }
        p.t.append_u64(LBL("N"), n);
        p.t.append_u64(LBL("T"), ext);
        p.t.append_u64(LBL("M"), m);
        for (uint32_t j = 0; j < m; j++) p.t.append_message(LBL("Ci"), a->commitments32 + 32 * (i * m + j), 32);
        for (uint32_t j = 0; j < m; j++) p.t.append_u64(LBL("vi - minimum_value"), a->min_present[i * m + j] ? a->min_values[i * m + j] : 0);
        fill_witness(p, i);
        rebuild_rng(p);
        if (below_minimum(i)) p.rc = BPP_INVALID_ARGUMENT;
        if (p.rc) { p.t.s.store(a->transcripts + BPP_TRANSCRIPT_BYTES * i); return; }
        if (!p.has_seed)
            for (uint32_t k = 0; k < ext; k++) p.alpha[k] = random_not_zero(p.rng);
        init_finish(p, i);
    };
    auto init_lock = [&](size_t i0) -> bool {
        if (!gens_nonzero) return false;
        Lock8 L;
        for (int j = 0; j < LK; j++) {
            PProof &p = pp[i0 + j];
            L.p[j] = &p;
            if (p.rc || below_minimum(i0 + j)) return false;
            p.t.s.load(a->transcripts + BPP_TRANSCRIPT_BYTES * (i0 + j));
            fill_witness(p, i0 + j);
        }
        if (!L.uniform(false)) return false;
        L.load_t();
        L.t.append_same(LBL("dom-sep"), LBL("Bulletproofs+ Range Proof"));
        L.t.append_same(LBL("H"), g->h(), 32);
        for (uint32_t k = 0; k < ext; k++) L.t.append_same(LBL("G"), g->g(k), 32);
        L.t.append_u64_same(LBL("N"), n);
        L.t.append_u64_same(LBL("T"), ext);
        L.t.append_u64_same(LBL("M"), m);
        const uint8_t *ptr[LK];
        for (uint32_t j = 0; j < m; j++) {
            for (int l = 0; l < LK; l++) ptr[l] = a->commitments32 + 32 * ((i0 + l) * m + j);
            L.t.append_each(LBL("Ci"), ptr, 32);
        }
        for (uint32_t j = 0; j < m; j++) {
            uint8_t mv[LK][8];
            for (int l = 0; l < LK; l++) { le64_bytes(mv[l], a->min_present[(i0 + l) * m + j] ? a->min_values[(i0 + l) * m + j] : 0); ptr[l] = mv[l]; }
            L.t.append_each(LBL("vi - minimum_value"), ptr, 8);
        }
        L.rebuild();
        sc al[BPP_MAX_EXT][LK];
        bool ok = true;
        if (!L.p[0]->has_seed)
            for (uint32_t k = 0; k < ext && ok; k++) ok = L.draw(al[k]);
        if (ok) {
            L.store_t(); L.store_r();
            for (int l = 0; l < LK; l++) {
                PProof &p = *L.p[l];
                p.rng_used++;
                if (!p.has_seed)
                    for (uint32_t k = 0; k < ext; k++) p.alpha[k] = al[k][l];
                init_finish(p, i0 + l);
            }
        }
        secure_zero(al, sizeof al);
        L.wipe();
        return ok;
    };
    host_stage_lk(P0, init_lock, init_one);
    for (size_t i = 0; i < P0; i++)
        if (!pp[i].rc) { pp[i].live = true; pp[i].slot = (uint32_t)live.size(); live.push_back(i); }
    const uint32_t P = (uint32_t)live.size();
    for (size_t i = 0; i < P0; i++) status[i] = pp[i].rc;
    if (P == 0) return BPP_OK;
    offs.resize((size_t)P * m);
    for (uint32_t s = 0; s < P; s++)
        for (uint32_t j = 0; j < m; j++) {
            size_t i = live[s];
            offs[(size_t)s * m + j] = a->values[i * m + j] - (a->min_present[i * m + j] ? a->min_values[i * m + j] : 0);
        }

    // ---- device state
    PDims d;
    d.P = P; d.n = n; d.m = m; d.N = N; d.ext = ext; d.rounds = rounds; d.gens_nm = (uint32_t)g->nm;
    // Fixed-base path (default): every commitment is a sum over the static generators, evaluated from window tables (k_fb.cu); the
    // generator folding of the reference survives as two scalar vectors.  The folding path (k_prove_fold_pts + K-MSM) remains
    // for generator sets whose tables would not fit BPP_FB_MAX_MB, and under BPP_PROVE_FOLD=1 for the parity tests.
    static const bool force_fold = getenv("BPP_PROVE_FOLD") != nullptr && atoi(getenv("BPP_PROVE_FOLD")) != 0;
    const bool fb = !force_fold && gens_fb_ensure(g);
    const size_t max_entries = fb ? std::max<size_t>((size_t)2 * P * (1 + ext + N), (size_t)P * (2 * N + 1 + ext) + (size_t)P * (1 + ext))
                                  : std::max<size_t>((size_t)P * (N + ext), (size_t)2 * P * (1 + ext + N));
    if (!ctx->prove_ws) ctx->prove_ws = new ProveWS();
    ProveWS &ws = *(ProveWS *)ctx->prove_ws;
    DevBuf &d_offs = ws.d_offs, &d_a = ws.d_a, &d_b = ws.d_b, &d_ypow = ws.d_ypow, &d_yinv2 = ws.d_yinv2, &d_yz = ws.d_yz, &d_dlr = ws.d_dlr,
           &d_e = ws.d_e, &d_fsc = ws.d_fsc, &d_folded = ws.d_folded, &d_mscal = ws.d_mscal, &d_pidx = ws.d_pidx, &d_segoff = ws.d_segoff,
           &d_scratch = ws.d_scratch, &d_res = ws.d_res, &d_enc = ws.d_enc, &d_ab = ws.d_ab;
    PinBuf &h_io = ws.h_io;
    auto release = [&]() {};       // buffers stay with the ctx
    size_t scratch_bytes = 256;
    if (!fb) {
        scratch_bytes = msm_scratch_bytes(msm_shape((uint32_t)((size_t)P * (N + ext)), P, 0));
        for (uint32_t r = 0; r < rounds; r++) {
            uint32_t nn = N >> (r + 1);
            scratch_bytes = std::max(scratch_bytes, msm_scratch_bytes(msm_shape(2 * P * (1 + ext + 2 * nn), 2 * P, 0)));
        }
        scratch_bytes = std::max(scratch_bytes, msm_scratch_bytes(msm_shape(P * (4 + 2 * ext), 2 * P, 0)));
    }
    // generator-index rows of the fixed-base sums (k_prove.cu, "fixed-base path"): A | (L, R) per round | A1 | B
    std::vector<uint32_t> gidx;
    size_t gi_A = 0, gi_round0 = 0, gi_A1 = 0, gi_B = 0;
    const uint32_t segA = 2 * N + ext, segLR = 1 + ext + N, segA1 = 2 * N + 1 + ext, segB = 1 + ext;
    if (fb) {
        const uint32_t nm = (uint32_t)g->nm, iG = 2 * nm, iH = 2 * nm + ext;
        gi_A = gidx.size();
        for (uint32_t j = 0; j < N; j++) gidx.push_back(j);
        for (uint32_t j = 0; j < N; j++) gidx.push_back(nm + j);
        for (uint32_t k = 0; k < ext; k++) gidx.push_back(iG + k);
        gi_round0 = gidx.size();
        for (uint32_t r = 0; r < rounds; r++) {
            const uint32_t nn = N >> (r + 1), half = N / 2;
            for (int side = 0; side < 2; side++) {                   // 0 = L, 1 = R
                gidx.push_back(iH);
                for (uint32_t k = 0; k < ext; k++) gidx.push_back(iG + k);
                for (uint32_t t = 0; t < half; t++) { uint32_t jl = (t / nn) * 2 * nn + t % nn; gidx.push_back(side ? jl : jl + nn); }
                for (uint32_t t = 0; t < half; t++) { uint32_t jl = (t / nn) * 2 * nn + t % nn; gidx.push_back(nm + (side ? jl + nn : jl)); }
            }
        }
        gi_A1 = gidx.size();
        for (uint32_t j = 0; j < N; j++) gidx.push_back(j);
        for (uint32_t j = 0; j < N; j++) gidx.push_back(nm + j);
        for (uint32_t k = 0; k < ext; k++) gidx.push_back(iG + k);
        gidx.push_back(iH);
        gi_B = gidx.size();
        gidx.push_back(iH);
        for (uint32_t k = 0; k < ext; k++) gidx.push_back(iG + k);
    }
    cudaError_t ce = cudaSuccess;
    auto ok = [&](cudaError_t x) { if (ce == cudaSuccess) ce = x; };
    ok(d_offs.ensure(8 * (size_t)P * m));
    ok(d_a.ensure(32 * (size_t)P * N)); ok(d_b.ensure(32 * (size_t)P * N));
    ok(d_ypow.ensure(32 * (size_t)P * (N + 2))); ok(d_yinv2.ensure(32 * (size_t)P * BPP_MAX_ROUNDS));
    ok(d_yz.ensure(96 * (size_t)P)); ok(d_dlr.ensure(64 * (size_t)P * ext)); ok(d_e.ensure(64 * (size_t)P)); ok(d_fsc.ensure(32 * 6 * (size_t)P));
    if (fb) {
        ok(ws.d_sg.ensure(32 * (size_t)P * N)); ok(ws.d_sh.ensure(32 * (size_t)P * N));
        ok(ws.d_gidx.ensure(4 * gidx.size())); ok(ws.d_rs.ensure(64 * (size_t)P));
    } else {
        ok(d_folded.ensure(sizeof(cached) * 2 * (size_t)P * N));
    }
    ok(d_mscal.ensure(32 * max_entries)); ok(d_pidx.ensure(4 * max_entries)); ok(d_segoff.ensure(4 * (2 * (size_t)P + 1)));
    ok(d_scratch.ensure(scratch_bytes)); ok(d_res.ensure(sizeof(ge) * 2 * (size_t)P)); ok(d_enc.ensure(64 * (size_t)P)); ok(d_ab.ensure(64 * (size_t)P));
    const size_t io_bytes = std::max<size_t>(32 * (size_t)P * (4 + 2 * ext) + 4 * (size_t)P * (4 + 2 * ext), 64 * (size_t)P * std::max<uint32_t>(ext, 1) + 64 * (size_t)P) + 4 * (2 * (size_t)P + 1) + 1024;
    static_assert(BPP_MAX_EXT >= 1, "extension degree");
    ok(h_io.ensure(io_bytes));
    if (ce != cudaSuccess) { release(); return cuda_fail(ctx, ce, "prover buffers"); }
    PBuffers b;
    b.offset_values = d_offs.as<uint64_t>(); b.a = d_a.as<uint32_t>(); b.b = d_b.as<uint32_t>(); b.ypow = d_ypow.as<uint32_t>();
    b.yinv2 = d_yinv2.as<uint32_t>(); b.yz = d_yz.as<uint32_t>(); b.dlr = d_dlr.as<uint32_t>(); b.e = d_e.as<uint32_t>(); b.fsc = d_fsc.as<uint32_t>();
    b.folded = d_folded.as<cached>(); b.msm_scalars = d_mscal.as<uint32_t>(); b.msm_pidx = d_pidx.as<uint32_t>();
    b.sg = ws.d_sg.as<uint32_t>(); b.sh = ws.d_sh.as<uint32_t>();
    uint8_t *hio = h_io.as<uint8_t>();
    wipe.hio = hio; wipe.hio_bytes = io_bytes; wipe.st = st;
    wipe.dev = {{d_a.p, 32 * (size_t)P * N}, {d_b.p, 32 * (size_t)P * N}, {d_offs.p, 8 * (size_t)P * m}, {d_mscal.p, 32 * max_entries},
                {d_dlr.p, 64 * (size_t)P * ext}, {d_ab.p, 64 * (size_t)P}, {d_yz.p, 96 * (size_t)P}, {d_e.p, 64 * (size_t)P}};
    if (fb) { wipe.dev.push_back({ws.d_rs.p, 64 * (size_t)P}); wipe.dev.push_back({ws.d_sg.p, 32 * (size_t)P * N}); wipe.dev.push_back({ws.d_sh.p, 32 * (size_t)P * N}); }
#define PCUDA(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { release(); return cuda_fail(ctx, _e, #call); } } while (0)
    auto upload_offsets = [&](uint32_t n_seg, const std::vector<uint32_t> &off) -> cudaError_t {
        memcpy(hio + io_bytes - 4 * (2 * (size_t)P + 1) - 16, off.data(), 4 * (n_seg + 1));
        return cudaMemcpyAsync(d_segoff.p, hio + io_bytes - 4 * (2 * (size_t)P + 1) - 16, 4 * (n_seg + 1), cudaMemcpyHostToDevice, st);
    };
    auto run_msm = [&](uint32_t n_entries, uint32_t n_seg, const std::vector<uint32_t> &off) -> cudaError_t {
        cudaError_t e1 = upload_offsets(n_seg, off);
        if (e1 != cudaSuccess) return e1;
        MsmShape sh = msm_shape(n_entries, n_seg, 0);
        launch_msm(st, sh, d_mscal.as<uint32_t>(), n_seg > 1 ? d_segoff.as<uint32_t>() : nullptr, d_pidx.as<uint32_t>(), nullptr, g->d_table.as<aniels>(),
                   d_scratch.p, d_res.as<ge>(), &ctx->launches, nullptr, d_folded.as<cached>());
        launch_encode(st, n_seg, d_res.as<ge>(), d_enc.as<uint32_t>(), nullptr);
        ctx->launches++;
        return cudaGetLastError();
    };

    // fixed-base sums: n_seg segments of seg_len entries starting at entry `first`, results (and their encodings) from slot `res0` on
    auto run_fb = [&](uint32_t n_seg, uint32_t seg_len, uint32_t kinds, size_t gi_off, size_t first, uint32_t res0) -> cudaError_t {
        launch_fb_msm(st, g->fb, n_seg, seg_len, kinds, d_mscal.as<uint32_t>() + 8 * first, ws.d_gidx.as<uint32_t>() + gi_off, g->d_fb.as<aniels>(),
                      d_res.as<ge>() + res0, &ctx->launches);
        launch_encode(st, n_seg, d_res.as<ge>() + res0, d_enc.as<uint32_t>() + 8 * (size_t)res0, nullptr);
        ctx->launches++;
        return cudaGetLastError();
    };
    if (fb) PCUDA(cudaMemcpyAsync(ws.d_gidx.p, gidx.data(), 4 * gidx.size(), cudaMemcpyHostToDevice, st));

    // ---- A (:334-345)
    PCUDA(cudaMemcpyAsync(d_offs.p, offs.data(), 8 * (size_t)P * m, cudaMemcpyHostToDevice, st));
    if (fb) launch_prove_bits_fb(st, d, b); else launch_prove_bits(st, d, b);
    ctx->launches++;
    for (uint32_t s = 0; s < P; s++)
        for (uint32_t k = 0; k < ext; k++) sc_store(hio + 32 * ((size_t)s * ext + k), pp[live[s]].alpha[k]);
    if (fb) {
        PCUDA(cudaMemcpy2DAsync(d_mscal.as<uint8_t>() + 32 * (size_t)(2 * N), 32 * (size_t)segA, hio, 32 * (size_t)ext, 32 * (size_t)ext, P,
                                cudaMemcpyHostToDevice, st));
        PCUDA(run_fb(P, segA, 1, gi_A, 0, 0));
    } else {
        PCUDA(cudaMemcpy2DAsync(d_mscal.as<uint8_t>() + 32 * (size_t)N, 32 * (size_t)(N + ext), hio, 32 * (size_t)ext, 32 * (size_t)ext, P,
                                cudaMemcpyHostToDevice, st));
        std::vector<uint32_t> off(P + 1);
        for (uint32_t s = 0; s <= P; s++) off[s] = s * (N + ext);
        PCUDA(run_msm(P * (N + ext), P, off));
    }
    PCUDA(cudaMemcpyAsync(hio, d_enc.p, 32 * (size_t)P, cudaMemcpyDeviceToHost, st));
    PCUDA(cudaStreamSynchronize(st));
    // ---- challenges y, z (:348, transcripts.rs:124-136), alpha update (:382-392)
    auto yz_finish = [&](PProof &p, size_t i) {
        const sc z2 = hmul(p.z, p.z), yN1 = hpow(p.y, (uint64_t)N + 1);
        sc zeven = sc_one();
        for (uint32_t j = 0; j < m; j++) {
            zeven = hmul(zeven, z2);
            for (uint32_t k = 0; k < ext; k++) {
                sc r = sc_load(a->blindings32 + 32 * ((i * m + j) * ext + k));
                p.alpha[k] = sc_add(p.alpha[k], hmul(hmul(zeven, r), yN1));
            }
        }
    };
    auto yz_one = [&](size_t s) {
        PProof &p = pp[live[s]];
        memcpy(p.A, hio + 32 * s, 32);
        bool good = append_point(p.t, LBL("A"), p.A);
        if (good) { rebuild_rng(p); good = challenge(p.t, LBL("y"), p.y) && challenge(p.t, LBL("z"), p.z); }
        if (!good) { p.rc = BPP_VERIFICATION_FAILED; p.y = sc_one(); p.z = sc_one(); }
        yz_finish(p, live[s]);
    };
    auto yz_lock = [&](size_t s0) -> bool {
        Lock8 L;
        const uint8_t *ptr[LK];
        for (int l = 0; l < LK; l++) {
            L.p[l] = &pp[live[s0 + l]];
            ptr[l] = hio + 32 * (s0 + l);
            if (zero32(ptr[l])) return false;
        }
        if (!L.uniform(false)) return false;
        L.load_t();
        L.t.append_each(LBL("A"), ptr, 32);
        L.rebuild();
        sc y[LK], z[LK];
        const bool ok = L.challenge(LBL("y"), y) && L.challenge(LBL("z"), z);
        if (ok) {
            L.store_t(); L.store_r();
            for (int l = 0; l < LK; l++) {
                PProof &p = *L.p[l];
                memcpy(p.A, ptr[l], 32);
                p.rng_used++;
                p.y = y[l]; p.z = z[l];
                yz_finish(p, live[s0 + l]);
            }
        }
        L.wipe();
        return ok;
    };
    host_stage_lk(P, yz_lock, yz_one);
    for (uint32_t s = 0; s < P; s++) { sc_store(hio + 64 * (size_t)s, pp[live[s]].y); sc_store(hio + 64 * (size_t)s + 32, pp[live[s]].z); }
    // y^-1 by Montgomery's trick over chunks of 64 proofs (y is never zero: challenge() rejects it above)
    host_stage((P + 63) / 64, 1, [&](size_t c) {
        const size_t lo = 64 * c, hi = std::min<size_t>(P, lo + 64);
        sc pre[64], acc = sc_one();
        for (size_t s = lo; s < hi; s++) { pre[s - lo] = acc; acc = hmul(acc, pp[live[s]].y); }
        sc inv = sc_invert_gcd(acc);
        for (size_t s = hi; s-- > lo;) {
            sc_store(hio + 64 * (size_t)P + 32 * s, hmul(inv, pre[s - lo]));
            inv = hmul(inv, pp[live[s]].y);
        }
    });
    PCUDA(cudaMemcpyAsync(d_yz.p, hio, 96 * (size_t)P, cudaMemcpyHostToDevice, st));
    launch_prove_init(st, d, b);
    ctx->launches += 2;

    // ---- rounds (:409-538)
    for (uint32_t round = 0; round < rounds; round++) {
        const uint32_t nn = N >> (round + 1);
        PCUDA(cudaStreamSynchronize(st));      // hio is about to be rewritten
        auto dlr_put = [&](PProof &p, size_t s) {
            for (uint32_t k = 0; k < ext; k++) {
                sc_store(hio + 32 * ((size_t)s * 2 * ext + k), p.dL[(size_t)round * ext + k]);
                sc_store(hio + 32 * ((size_t)s * 2 * ext + ext + k), p.dR[(size_t)round * ext + k]);
            }
        };
        auto dlr_one = [&](size_t s) {
            PProof &p = pp[live[s]];
            for (uint32_t k = 0; k < ext; k++) p.dL[(size_t)round * ext + k] = p.has_seed ? nonce(p.seed, "dL", true, round, true, k) : random_not_zero(p.rng);
            for (uint32_t k = 0; k < ext; k++) p.dR[(size_t)round * ext + k] = p.has_seed ? nonce(p.seed, "dR", true, round, true, k) : random_not_zero(p.rng);
            dlr_put(p, s);
        };
        auto dlr_lock = [&](size_t s0) -> bool {       // only the rng draws go through the sponge; seed nonces are BLAKE2b
            Lock8 L;
            for (int l = 0; l < LK; l++) L.p[l] = &pp[live[s0 + l]];
            if (L.p[0]->has_seed || !L.uniform(true)) return false;
            L.load_r();
            sc dl[BPP_MAX_EXT][LK], dr[BPP_MAX_EXT][LK];
            bool ok = true;
            for (uint32_t k = 0; k < ext && ok; k++) ok = L.draw(dl[k]);
            for (uint32_t k = 0; k < ext && ok; k++) ok = L.draw(dr[k]);
            if (ok) {
                L.store_r();
                for (int l = 0; l < LK; l++) {
                    PProof &p = *L.p[l];
                    for (uint32_t k = 0; k < ext; k++) { p.dL[(size_t)round * ext + k] = dl[k][l]; p.dR[(size_t)round * ext + k] = dr[k][l]; }
                    dlr_put(p, s0 + l);
                }
            }
            secure_zero(dl, sizeof dl); secure_zero(dr, sizeof dr);
            L.wipe();
            return ok;
        };
        host_stage_lk(P, dlr_lock, dlr_one);
        PCUDA(cudaMemcpyAsync(d_dlr.p, hio, 64 * (size_t)P * ext, cudaMemcpyHostToDevice, st));
        if (fb) {
            launch_prove_round_pre_fb(st, d, b, nn);
            ctx->launches++;
            PCUDA(run_fb(2 * P, segLR, 2, gi_round0 + (size_t)round * 2 * segLR, 0, 0));
        } else {
            launch_prove_round_pre(st, d, b, nn, round);
            ctx->launches++;
            const uint32_t seg_len = 1 + ext + 2 * nn;
            std::vector<uint32_t> off(2 * P + 1);
            for (uint32_t s = 0; s <= 2 * P; s++) off[s] = s * seg_len;
            PCUDA(cudaStreamSynchronize(st));  // hio (d_L / d_R) consumed before the offsets share the staging buffer's tail
            PCUDA(run_msm(2 * P * seg_len, 2 * P, off));
        }
        PCUDA(cudaMemcpyAsync(hio, d_enc.p, 64 * (size_t)P, cudaMemcpyDeviceToHost, st));
        PCUDA(cudaStreamSynchronize(st));
        auto e_one = [&](size_t s) {       // transcripts.rs:139-149
            PProof &p = pp[live[s]];
            memcpy(p.LR.data() + 64 * (size_t)round, hio + 64 * s, 64);
            bool good = append_point(p.t, LBL("L"), hio + 64 * s) && append_point(p.t, LBL("R"), hio + 64 * s + 32);
            sc e = sc_one();
            if (good) { rebuild_rng(p); good = challenge(p.t, LBL("e"), e); }
            if (!good) { if (!p.rc) p.rc = BPP_VERIFICATION_FAILED; e = sc_one(); }
            p.e_round[round] = e;
        };
        auto e_lock = [&](size_t s0) -> bool {
            Lock8 L;
            const uint8_t *pl[LK], *pr[LK];
            for (int l = 0; l < LK; l++) {
                L.p[l] = &pp[live[s0 + l]];
                pl[l] = hio + 64 * (s0 + l); pr[l] = pl[l] + 32;
                if (zero32(pl[l]) || zero32(pr[l])) return false;
            }
            if (!L.uniform(false)) return false;
            L.load_t();
            L.t.append_each(LBL("L"), pl, 32);
            L.t.append_each(LBL("R"), pr, 32);
            L.rebuild();
            sc e[LK];
            const bool ok = L.challenge(LBL("e"), e);
            if (ok) {
                L.store_t(); L.store_r();
                for (int l = 0; l < LK; l++) {
                    PProof &p = *L.p[l];
                    memcpy(p.LR.data() + 64 * (size_t)round, pl[l], 64);
                    p.rng_used++;
                    p.e_round[round] = e[l];
                }
            }
            L.wipe();
            return ok;
        };
        host_stage_lk(P, e_lock, e_one);
        // e^-1 of every proof by Montgomery's trick over chunks of 64 proofs (one inversion + 3 products per proof): on the device
        // it was one binary-Euclid inversion per proof per round on the critical path (0.13 ms per round), and the responses at
        // the end needed one more inversion per proof.  Challenges are never zero (challenge() rejects them above).
        host_stage((P + 63) / 64, 1, [&](size_t c) {
            const size_t lo = 64 * c, hi = std::min<size_t>(P, lo + 64);
            sc pre[64], acc = sc_one();
            for (size_t s = lo; s < hi; s++) { pre[s - lo] = acc; acc = hmul(acc, pp[live[s]].e_round[round]); }
            sc inv = sc_invert_gcd(acc);
            for (size_t s = hi; s-- > lo;) {
                PProof &p = pp[live[s]];
                p.einv_round[round] = hmul(inv, pre[s - lo]);
                inv = hmul(inv, p.e_round[round]);
            }
        });
        for (uint32_t s = 0; s < P; s++) {
            sc_store(hio + 32 * (size_t)s, pp[live[s]].e_round[round]);
            sc_store(hio + 32 * ((size_t)P + s), pp[live[s]].einv_round[round]);
        }
        PCUDA(cudaMemcpyAsync(d_e.p, hio, 64 * (size_t)P, cudaMemcpyHostToDevice, st));
        if (fb) { launch_prove_fold_fb(st, d, b, nn); ctx->launches += 2; }
        else { launch_prove_fold(st, d, b, nn, round, g->d_table.as<aniels>()); ctx->launches += 3; }
    }

    // ---- final (:542-594)
    launch_prove_final_ab(st, d, b, d_ab.as<uint32_t>());
    ctx->launches++;
    PCUDA(cudaStreamSynchronize(st));
    ab.resize(64 * (size_t)P);
    PCUDA(cudaMemcpyAsync(ab.data(), d_ab.p, 64 * (size_t)P, cudaMemcpyDeviceToHost, st));
    PCUDA(cudaStreamSynchronize(st));
    auto draw_final = [&](PProof &p) {
        p.r = random_not_zero(p.rng);                 // always from the rng, even with a seed nonce (:542-543)
        p.s = random_not_zero(p.rng);
        for (uint32_t k = 0; k < ext; k++) p.d[k] = p.has_seed ? nonce(p.seed, "d", false, 0, true, k) : random_not_zero(p.rng);
        for (uint32_t k = 0; k < ext; k++) p.eta[k] = p.has_seed ? nonce(p.seed, "eta", false, 0, true, k) : random_not_zero(p.rng);
    };
    auto draw_lock = [&](size_t s0) -> bool {
        Lock8 L;
        for (int l = 0; l < LK; l++) L.p[l] = &pp[live[s0 + l]];
        if (!L.uniform(true)) return false;
        L.load_r();
        const bool seeded = L.p[0]->has_seed;
        sc r[LK], sv[LK], dd[BPP_MAX_EXT][LK], ee[BPP_MAX_EXT][LK];
        bool ok = L.draw(r) && L.draw(sv);
        if (!seeded) {
            for (uint32_t k = 0; k < ext && ok; k++) ok = L.draw(dd[k]);
            for (uint32_t k = 0; k < ext && ok; k++) ok = L.draw(ee[k]);
        }
        if (ok) {
            L.store_r();
            for (int l = 0; l < LK; l++) {
                PProof &p = *L.p[l];
                p.r = r[l]; p.s = sv[l];
                for (uint32_t k = 0; k < ext; k++) {
                    p.d[k] = seeded ? nonce(p.seed, "d", false, 0, true, k) : dd[k][l];
                    p.eta[k] = seeded ? nonce(p.seed, "eta", false, 0, true, k) : ee[k][l];
                }
            }
        }
        secure_zero(r, sizeof r); secure_zero(sv, sizeof sv); secure_zero(dd, sizeof dd); secure_zero(ee, sizeof ee);
        L.wipe();
        return ok;
    };
    host_stage_lk(P, draw_lock, [&](size_t s) { draw_final(pp[live[s]]); });
    if (fb) {
        // A1 = sum_j r*sG[j]*G_j + sum_j s*sH[j]*H_j + sum d[k]*G[k] + (r*y*b[0] + s*y*a[0])*H   (:574-580, Gi[0] / Hi[0] unfolded)
        // B  = (r*y*s)*H + sum eta[k]*G[k]                                                        (:581-584)
        uint8_t *h_rs = hio, *h_tail = hio + 64 * (size_t)P, *h_b = h_tail + 32 * (size_t)P * (1 + ext);
        host_stage(P, 8, [&](size_t s) {
            PProof &p = pp[live[s]];
            const sc a0 = sc_load(ab.data() + 64 * s), b0 = sc_load(ab.data() + 64 * s + 32);
            const sc ry = hmul(p.r, p.y), sy = hmul(p.s, p.y);
            sc_store(h_rs + 64 * s, p.r); sc_store(h_rs + 64 * s + 32, p.s);
            uint8_t *tl = h_tail + 32 * s * (1 + ext);
            for (uint32_t k = 0; k < ext; k++) sc_store(tl + 32 * k, p.d[k]);
            sc_store(tl + 32 * ext, sc_add(hmul(ry, b0), hmul(sy, a0)));
            uint8_t *bv = h_b + 32 * s * segB;
            sc_store(bv, hmul(ry, p.s));
            for (uint32_t k = 0; k < ext; k++) sc_store(bv + 32 * (1 + k), p.eta[k]);
        });
        const size_t firstB = (size_t)P * segA1;
        PCUDA(cudaMemcpyAsync(ws.d_rs.p, h_rs, 64 * (size_t)P, cudaMemcpyHostToDevice, st));
        PCUDA(cudaMemcpy2DAsync(d_mscal.as<uint8_t>() + 32 * (size_t)(2 * N), 32 * (size_t)segA1, h_tail, 32 * (size_t)(1 + ext), 32 * (size_t)(1 + ext), P,
                                cudaMemcpyHostToDevice, st));
        PCUDA(cudaMemcpyAsync(d_mscal.as<uint8_t>() + 32 * firstB, h_b, 32 * (size_t)P * segB, cudaMemcpyHostToDevice, st));
        launch_prove_final_fb(st, d, b, ws.d_rs.as<uint32_t>());
        ctx->launches++;
        PCUDA(run_fb(P, segA1, 1, gi_A1, 0, 0));
        PCUDA(run_fb(P, segB, 1, gi_B, firstB, P));
        PCUDA(cudaStreamSynchronize(st));      // hio is about to receive the encodings
        PCUDA(cudaMemsetAsync(ws.d_rs.p, 0, 64 * (size_t)P, st));
        PCUDA(cudaMemsetAsync(ws.d_sg.p, 0, 32 * (size_t)P * N, st));
        PCUDA(cudaMemsetAsync(ws.d_sh.p, 0, 32 * (size_t)P * N, st));
    } else {
    const uint32_t len1 = 3 + ext, len2 = 1 + ext, per = len1 + len2;
    uint8_t *h_sc = hio;
    uint32_t *h_px = reinterpret_cast<uint32_t *>(hio + 32 * (size_t)P * per);
    const uint32_t GEN = 0x80000000u, CACHED = 0x40000000u;
    host_stage(P, 8, [&](size_t s) {
        PProof &p = pp[live[s]];
        const sc a0 = sc_load(ab.data() + 64 * s), b0 = sc_load(ab.data() + 64 * s + 32);
        const sc ry = hmul(p.r, p.y), sy = hmul(p.s, p.y);
        uint8_t *sv = h_sc + 32 * s * per;
        uint32_t *px = h_px + s * per;
        // A1 = r*Gi[0] + s*Hi[0] + (r*y*b[0] + s*y*a[0])*H + sum d[k]*G[k]      (:574-580)
        sc_store(sv, p.r);                       px[0] = CACHED | (uint32_t)(s * N);
        sc_store(sv + 32, p.s);                  px[1] = CACHED | (uint32_t)((size_t)P * N + s * N);
        sc_store(sv + 64, sc_add(hmul(ry, b0), hmul(sy, a0)));   px[2] = GEN | (uint32_t)(2 * g->nm + ext);
        for (uint32_t k = 0; k < ext; k++) { sc_store(sv + 32 * (3 + k), p.d[k]); px[3 + k] = GEN | (uint32_t)(2 * g->nm + k); }
        // B = (r*y*s)*H + sum eta[k]*G[k]                                        (:581-584)
        sc_store(sv + 32 * len1, hmul(ry, p.s)); px[len1] = GEN | (uint32_t)(2 * g->nm + ext);
        for (uint32_t k = 0; k < ext; k++) { sc_store(sv + 32 * (len1 + 1 + k), p.eta[k]); px[len1 + 1 + k] = GEN | (uint32_t)(2 * g->nm + k); }
    });
    PCUDA(cudaMemcpyAsync(d_mscal.p, h_sc, 32 * (size_t)P * per, cudaMemcpyHostToDevice, st));
    PCUDA(cudaMemcpyAsync(d_pidx.p, h_px, 4 * (size_t)P * per, cudaMemcpyHostToDevice, st));
    {
        std::vector<uint32_t> off(2 * P + 1);
        for (uint32_t s = 0; s < P; s++) { off[2 * s] = s * per; off[2 * s + 1] = s * per + len1; }
        off[2 * P] = P * per;
        PCUDA(cudaStreamSynchronize(st));
        PCUDA(run_msm(P * per, 2 * P, off));
    }
    }
    PCUDA(cudaMemcpyAsync(hio, d_enc.p, 64 * (size_t)P, cudaMemcpyDeviceToHost, st));
    // the device copies of the secrets are wiped by wipe_guard when the function is left
    PCUDA(cudaStreamSynchronize(st));
    const uint8_t *encA1 = hio, *encB = fb ? hio + 32 * (size_t)P : hio + 32;      // fixed-base path: [A1 x P | B x P], else [A1 | B] x P
    const size_t enc_stride = fb ? 32 : 64;
    auto last_lock = [&](size_t s0) -> bool {       // transcripts.rs:152-162
        Lock8 L;
        const uint8_t *pa[LK], *pb[LK];
        for (int l = 0; l < LK; l++) {
            L.p[l] = &pp[live[s0 + l]];
            pa[l] = encA1 + enc_stride * (s0 + l); pb[l] = encB + enc_stride * (s0 + l);
            if (zero32(pa[l]) || zero32(pb[l])) return false;
        }
        if (!L.uniform(false)) return false;
        L.load_t();
        L.t.append_each(LBL("A1"), pa, 32);
        L.t.append_each(LBL("B"), pb, 32);
        L.rebuild();
        sc e[LK];
        const bool ok = L.challenge(LBL("e"), e);
        if (ok) {
            L.store_t(); L.store_r();
            for (int l = 0; l < LK; l++) {
                PProof &p = *L.p[l];
                memcpy(p.A1, pa[l], 32); memcpy(p.B, pb[l], 32);
                p.rng_used++;
                p.e_final = e[l];
            }
        }
        L.wipe();
        return ok;
    };
    auto last_one = [&](size_t s) {
        PProof &p = pp[live[s]];
        memcpy(p.A1, encA1 + enc_stride * s, 32);
        memcpy(p.B, encB + enc_stride * s, 32);
        bool good = append_point(p.t, LBL("A1"), p.A1) && append_point(p.t, LBL("B"), p.B);
        if (good) { rebuild_rng(p); good = challenge(p.t, LBL("e"), p.e_final); }
        if (!good && !p.rc) p.rc = BPP_VERIFICATION_FAILED;
    };
    host_stage_lk(P, last_lock, last_one);
    host_stage(P, 8, [&](size_t s) {
        const size_t i = live[s];
        PProof &p = pp[i];
        p.t.s.store(a->transcripts + BPP_TRANSCRIPT_BYTES * i);
        if (p.rc) return;
        const sc e = p.e_final, e2 = hmul(e, e);
        const sc a0 = sc_load(ab.data() + 64 * s), b0 = sc_load(ab.data() + 64 * s + 32);
        // alpha += sum_rounds d_L*e_r^2 + d_R*e_r^-2 (:535-537), with the inverses the rounds already made
        for (int r = (int)rounds - 1; r >= 0; r--) {
            const sc einv = p.einv_round[r];
            sc er2 = hmul(p.e_round[r], p.e_round[r]), einv2 = hmul(einv, einv);
            for (uint32_t k = 0; k < ext; k++)
                p.alpha[k] = sc_add(p.alpha[k], sc_add(hmul(p.dL[(size_t)r * ext + k], er2), hmul(p.dR[(size_t)r * ext + k], einv2)));
        }
        uint8_t *out = proofs_out + proof_stride * i;        // to_bytes layout (:1120-1150)
        out[0] = (uint8_t)ext;
        for (uint32_t k = 0; k < ext; k++)
            sc_store(out + 1 + 32 * k, sc_add(sc_add(p.eta[k], hmul(p.d[k], e)), hmul(p.alpha[k], e2)));    // d1 (:592-594)
        uint8_t *q = out + 1 + 32 * ext;
        memcpy(q, p.A, 32); memcpy(q + 32, p.A1, 32); memcpy(q + 64, p.B, 32);
        sc_store(q + 96, sc_add(p.r, hmul(a0, e)));           // r1 (:590)
        sc_store(q + 128, sc_add(p.s, hmul(b0, e)));          // s1 (:591)
        memcpy(q + 160, p.LR.data(), 64 * (size_t)rounds);
        // wipe host secrets
        memset(p.witness.data(), 0, p.witness.size());
        for (uint32_t k = 0; k < BPP_MAX_EXT; k++) { p.alpha[k] = sc_zero(); p.d[k] = sc_zero(); p.eta[k] = sc_zero(); }
        p.r = sc_zero(); p.s = sc_zero();
    });
    if (prove_trace)
        fprintf(stderr, "bpp_prove_batch P=%u: %.2f ms, host stages %.2f ms\n", P,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call0).count(), host_stage_ms);
    for (size_t i = 0; i < P0; i++) status[i] = pp[i].rc;
    release();
#undef PCUDA
    return BPP_OK;
}

} // extern "C"
