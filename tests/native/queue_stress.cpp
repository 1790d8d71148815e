// Native stress driver for the threaded parts of libbpp_b200.so, meant to run under ASan / TSan / compute-sanitizer on a GPU box
// (scripts/sanitize.sh; VERDICT r1 "what's weak" 6: a 16-lane run once died inside ncu with glibc's `unaligned tcache chunk`).
//   1. P proofs are made by the device prover (bpp_prove_batch);
//   2. T submitter threads push calls of 1-3 chunks each through ONE bpp_vqueue (3 lanes, <= 8 calls per pass), some calls corrupted,
//      and check every status;
//   3. the round-1 pattern: L threads with one bpp_ctx each calling bpp_verify_chunks concurrently (lanes), same checks.
// Exit code 0 = every verdict as expected.  No oracle involved: valid proofs must verify, corrupted ones must not.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include "../../include/bpp_b200.h"

#define CHECK(x) do { int32_t _rc = (x); if (_rc != 0) { fprintf(stderr, "%s:%d: %s -> %d\n", __FILE__, __LINE__, #x, _rc); exit(2); } } while (0)

static const int N = 64, EXT = 1, ROUNDS = 6;

struct Workload {
    size_t P;
    size_t plen;
    std::vector<uint8_t> proofs, commits, tstate0;
    std::vector<uint64_t> mins;
};

static Workload make_workload(size_t P) {
    Workload w;
    w.P = P;
    bpp_ctx *ctx = nullptr;
    bpp_gens *g = nullptr;
    CHECK(bpp_ctx_create(0, &ctx));
    CHECK(bpp_gens_create(ctx, N, 1, EXT, &g));
    std::vector<uint64_t> values(P);
    std::vector<uint8_t> blind(32 * P), seeds(32 * P), present(P, 1), rng(32 * (ROUNDS + 3) * P), trs(BPP_TRANSCRIPT_BYTES * P);
    w.mins.resize(P);
    uint64_t x = 88172645463325252ull;
    auto next = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    for (size_t i = 0; i < P; i++) {
        values[i] = next() >> 1;
        w.mins[i] = values[i] / 3;
        for (int k = 0; k < 31; k++) { blind[32 * i + k] = (uint8_t)next(); seeds[32 * i + k] = (uint8_t)next(); }
        blind[32 * i + 31] = 0x05; seeds[32 * i + 31] = 0x03;          // < 2^252: canonical, non-zero
    }
    for (auto &b : rng) b = (uint8_t)next();
    w.tstate0.resize(BPP_TRANSCRIPT_BYTES);
    bpp_transcript_new((const uint8_t *)"stress", 6, w.tstate0.data());
    for (size_t i = 0; i < P; i++) memcpy(trs.data() + BPP_TRANSCRIPT_BYTES * i, w.tstate0.data(), BPP_TRANSCRIPT_BYTES);
    w.commits.resize(32 * P);
    CHECK(bpp_pedersen_commit_batch(g, P, values.data(), blind.data(), EXT, w.commits.data()));
    w.plen = bpp_proof_size(EXT, ROUNDS);
    w.proofs.resize(w.plen * P);
    std::vector<int32_t> st(P);
    bpp_prove_args a;
    memset(&a, 0, sizeof a);
    a.n_proofs = P; a.aggregation = 1; a.commitments32 = w.commits.data(); a.values = values.data(); a.blindings32 = blind.data();
    a.min_values = w.mins.data(); a.min_present = present.data(); a.seed_nonces32 = seeds.data(); a.seed_present = present.data();
    a.transcripts = trs.data(); a.rng_bytes = rng.data(); a.rng_stride = 32 * (ROUNDS + 3);
    CHECK(bpp_prove_batch(g, &a, w.proofs.data(), w.plen, st.data()));
    for (size_t i = 0; i < P; i++) if (st[i]) { fprintf(stderr, "proof %zu: status %d\n", i, st[i]); exit(2); }
    bpp_gens_destroy(g);
    bpp_ctx_destroy(ctx);
    return w;
}

// one call: `chunks` chunks of `per` proofs starting at proof `first`; corrupt >= 0 flips a byte of r1 of that proof of the call
struct Call {
    std::vector<uint64_t> chunk_off, proof_off, commit_off;
    std::vector<uint8_t> proofs, commits, present, trs;
    std::vector<uint64_t> mins;
    std::vector<int32_t> status;
    bpp_verify_args a;
    int bad_chunk = -1;
    void *pinned = nullptr;          // proof bytes in page-locked memory: uploaded in place by the engine (no staging copy)
    ~Call() { if (pinned) bpp_host_free(pinned); }
    Call(const Call &) = delete;
    Call(const Workload &w, size_t first, size_t chunks, size_t per, int corrupt, bool pin = false) {
        const size_t n = chunks * per;
        for (size_t c = 0; c <= chunks; c++) chunk_off.push_back(c * per);
        for (size_t i = 0; i <= n; i++) { proof_off.push_back(i * w.plen); commit_off.push_back(i); }
        proofs.resize(w.plen * n); commits.resize(32 * n); mins.resize(n); present.assign(n, 1); trs.resize(BPP_TRANSCRIPT_BYTES * n);
        for (size_t i = 0; i < n; i++) {
            const size_t src = (first + i) % w.P;
            memcpy(proofs.data() + w.plen * i, w.proofs.data() + w.plen * src, w.plen);
            memcpy(commits.data() + 32 * i, w.commits.data() + 32 * src, 32);
            mins[i] = w.mins[src];
            memcpy(trs.data() + BPP_TRANSCRIPT_BYTES * i, w.tstate0.data(), BPP_TRANSCRIPT_BYTES);
        }
        if (corrupt >= 0) { proofs[w.plen * (size_t)corrupt + 1 + 32 * (EXT + 3) + 2] ^= 0x40; bad_chunk = (int)((size_t)corrupt / per); }
        status.assign(chunks, -1);
        memset(&a, 0, sizeof a);
        const uint8_t *pb = proofs.data();
        if (pin && bpp_host_alloc(proofs.size(), &pinned) == BPP_OK) { memcpy(pinned, proofs.data(), proofs.size()); pb = (const uint8_t *)pinned; }
        a.n_proofs = n; a.n_chunks = chunks; a.chunk_offsets = chunk_off.data(); a.proof_bytes = pb; a.proof_offsets = proof_off.data();
        a.commitments32 = commits.data(); a.commit_offsets = commit_off.data(); a.min_values = mins.data(); a.min_present = present.data();
        a.transcripts = trs.data(); a.action = BPP_VERIFY_ONLY;
    }
    bool ok() const {
        for (size_t c = 0; c < status.size(); c++)
            if (status[c] != ((int)c == bad_chunk ? BPP_VERIFICATION_FAILED : BPP_OK)) return false;
        return true;
    }
};

int main(int argc, char **argv) {
    const int T = argc > 1 ? atoi(argv[1]) : 8, per_thread = argc > 2 ? atoi(argv[2]) : 24, L = argc > 3 ? atoi(argv[3]) : 8;
    Workload w = make_workload(256);
    std::atomic<int> failures{0};
    // ---- 2. the coalescing queue: host-side weight transcripts, then device-side ones (one graph per pass); a third of the calls keep
    // their proof bytes in page-locked memory
    for (int device_weights = 0; device_weights < 2; device_weights++) {
        bpp_vqueue *q = nullptr;
        CHECK(bpp_vqueue_create(0, N, 1, EXT, nullptr, nullptr, 3, 8, 1, &q));
        CHECK(bpp_vqueue_set_device_weights(q, device_weights));
        std::vector<std::thread> ths;
        for (int t = 0; t < T; t++)
            ths.emplace_back([&, t]() {
                std::vector<Call *> calls;
                std::vector<uint64_t> tickets;
                for (int i = 0; i < per_thread; i++) {
                    const size_t chunks = 1 + (size_t)((t + i) % 3), per = 8 + 4 * (size_t)(i % 3);
                    const int corrupt = (i % 5 == 3) ? (int)((t * 7 + i) % (chunks * per)) : -1;
                    calls.push_back(new Call(w, (size_t)(t * 31 + i * 13), chunks, per, corrupt, (t + i) % 3 == 0));
                    uint64_t tk = 0;
                    if (i % 2) {
                        CHECK(bpp_vqueue_verify(q, &calls.back()->a, calls.back()->status.data(), nullptr, nullptr));
                        tickets.push_back(0);
                    } else {
                        CHECK(bpp_vqueue_submit(q, &calls.back()->a, calls.back()->status.data(), nullptr, nullptr, &tk));
                        tickets.push_back(tk);
                    }
                }
                for (size_t i = 0; i < calls.size(); i++) {
                    if (tickets[i]) CHECK(bpp_vqueue_wait(q, tickets[i]));
                    if (!calls[i]->ok()) failures++;
                    delete calls[i];
                }
            });
        for (auto &t : ths) t.join();
        uint64_t st[5];
        bpp_vqueue_stats(q, st);
        fprintf(stderr, "queue: %llu passes, %llu calls, %llu proofs, %llu kernels\n", (unsigned long long)st[0], (unsigned long long)st[1],
                (unsigned long long)st[2], (unsigned long long)st[3]);
        bpp_vqueue_destroy(q);
    }
    // ---- 3. lanes: one ctx per thread, plain bpp_verify_chunks
    {
        std::vector<std::thread> ths;
        for (int t = 0; t < L; t++)
            ths.emplace_back([&, t]() {
                bpp_ctx *ctx = nullptr;
                bpp_gens *g = nullptr;
                CHECK(bpp_ctx_create(0, &ctx));
                bpp_ctx_set_host_threads(ctx, 2);
                bpp_ctx_set_throughput_mode(ctx, t % 2);
                CHECK(bpp_gens_create(ctx, N, 1, EXT, &g));
                for (int i = 0; i < per_thread; i++) {
                    Call c(w, (size_t)(t * 17 + i * 5), 1 + (size_t)(i % 4), 16, (i % 4 == 1) ? 3 : -1, i % 3 == 1);
                    CHECK(bpp_verify_chunks(g, &c.a, c.status.data(), nullptr, nullptr));
                    if (!c.ok()) failures++;
                }
                bpp_gens_destroy(g);
                bpp_ctx_destroy(ctx);
            });
        for (auto &t : ths) t.join();
    }
    fprintf(stderr, "queue_stress: %d failures\n", failures.load());
    return failures.load() ? 1 : 0;
}
