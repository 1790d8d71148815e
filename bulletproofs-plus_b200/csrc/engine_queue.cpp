// bpp_vqueue: the coalescing front end of the batch verifier (include/bpp_b200.h).
//
// One RangeProof::verify_batch call (<= 256 proofs looked at, /root/reference/src/range_proof.rs:739-751) is far too little work
// for a B200: alone it is a chain of latency-bound kernels on a handful of SMs (round 1: 1.03 ms per 1024 proofs, 20-32 host
// threads with one CUDA context each were needed to overlap enough of them, which is what broke down on an 8-GPU box with 4 host
// cores per GPU).  The queue turns it round: callers submit calls from any number of threads, each of a few LANES (one bpp_ctx +
// generator tables + host thread) takes whatever is waiting -- up to max_calls_per_pass calls -- and verifies it as ONE device pass
// (bpp_vbatch_create_multi: every kernel of the pass then runs over 16-64 chunks instead of 1-4), writes each call's statuses,
// masks and advanced transcripts back to that call's own buffers and signals its ticket.  Results are exactly those of
// bpp_verify_chunks on every call alone (tests/test_gpu_queue.py).
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include "../../include/bpp_b200.h"

namespace {

struct QCall {
    bpp_verify_args args;
    int32_t *chunk_status = nullptr;
    uint8_t *masks32 = nullptr, *mask_present = nullptr;
    uint64_t ticket = 0;
    int32_t rc = BPP_OK;
    bool done = false;
};

struct QLane {
    bpp_ctx *ctx = nullptr;
    bpp_gens *gens = nullptr;
    std::thread th;
    double ms[4] = {0, 0, 0, 0};          // wall time of this lane's thread: waiting for calls, building passes, running them, handing results back
};
inline double ms_since(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

} // namespace

struct bpp_vqueue {
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::deque<QCall *> pending;
    std::unordered_map<uint64_t, QCall *> calls;       // every submitted call until its ticket has been waited for
    std::vector<QLane> lanes;
    bool stopping = false;
    uint64_t next_ticket = 1;
    size_t max_calls = 16;
    std::atomic<uint64_t> n_passes{0}, n_calls{0}, n_proofs{0};
    std::string err;
};

static void run_pass(bpp_vqueue *q, QLane &lane, std::vector<QCall *> &batch) {
    const size_t n = batch.size();
    std::vector<const bpp_verify_args *> ptrs(n);
    std::vector<int32_t *> st(n);
    std::vector<uint8_t *> mk(n), mp(n);
    for (size_t i = 0; i < n; i++) { ptrs[i] = &batch[i]->args; st[i] = batch[i]->chunk_status; mk[i] = batch[i]->masks32; mp[i] = batch[i]->mask_present; }
    bpp_vbatch *vb = nullptr;
    auto t0 = std::chrono::steady_clock::now();
    int32_t rc = bpp_vbatch_create_multi(lane.gens, n, ptrs.data(), &vb);
    lane.ms[1] += ms_since(t0);
    if (rc == BPP_OK) {
        t0 = std::chrono::steady_clock::now();
        rc = bpp_vbatch_run_multi(vb, st.data(), mk.data(), mp.data());
        lane.ms[2] += ms_since(t0);
        t0 = std::chrono::steady_clock::now();
        for (size_t i = 0; i < n && rc == BPP_OK; i++)
            if (batch[i]->args.transcripts) rc = bpp_vbatch_transcripts_call(vb, i, batch[i]->args.transcripts);
        bpp_vbatch_destroy(vb);
        lane.ms[3] += ms_since(t0);
        for (QCall *c : batch) c->rc = rc;
    } else if (n == 1) {
        batch[0]->rc = rc;
    } else {
        // an argument-level failure of one call (null pointers, bad offsets, size overflow) must not take the others down: one by one
        for (QCall *c : batch) {
            std::vector<QCall *> one(1, c);
            run_pass(q, lane, one);
        }
        return;
    }
    q->n_passes++;
    q->n_calls += n;
    for (QCall *c : batch) q->n_proofs += c->args.n_proofs;
}

static void lane_loop(bpp_vqueue *q, size_t li) {
    QLane &lane = q->lanes[li];
    std::vector<QCall *> batch;
    for (;;) {
        batch.clear();
        {
            const auto t_idle = std::chrono::steady_clock::now();
            std::unique_lock<std::mutex> lk(q->mu);
            q->cv_work.wait(lk, [&] { return q->stopping || !q->pending.empty(); });
            lane.ms[0] += ms_since(t_idle);
            if (q->pending.empty()) return;            // stopping and drained
            const int32_t action = q->pending.front()->args.action;
            while (!q->pending.empty() && batch.size() < q->max_calls && q->pending.front()->args.action == action) {
                batch.push_back(q->pending.front());
                q->pending.pop_front();
            }
        }
        run_pass(q, lane, batch);
        {
            std::lock_guard<std::mutex> lk(q->mu);
            for (QCall *c : batch) c->done = true;
        }
        q->cv_done.notify_all();
    }
}

extern "C" {

int32_t bpp_vqueue_create(int32_t device_ordinal, int32_t bit_length, int32_t max_aggregation, int32_t extension_degree,
                          const uint8_t *h_base32_or_null, const uint8_t *g_bases32_or_null, int32_t lanes, int32_t max_calls_per_pass,
                          int32_t host_threads_per_lane, bpp_vqueue **out) {
    if (!out || lanes < 1 || lanes > 64 || max_calls_per_pass < 1) return BPP_INVALID_ARGUMENT;
    *out = nullptr;
    bpp_vqueue *q = new bpp_vqueue();
    q->max_calls = (size_t)max_calls_per_pass;
    q->lanes.resize((size_t)lanes);
    int32_t rc = BPP_OK;
    for (QLane &l : q->lanes) {
        rc = bpp_ctx_create(device_ordinal, &l.ctx);
        if (rc) break;
        if (host_threads_per_lane > 0) bpp_ctx_set_host_threads(l.ctx, host_threads_per_lane);
        bpp_ctx_set_throughput_mode(l.ctx, 1);           // lane threads sleep while their pass runs
        rc = bpp_gens_create_with_bases(l.ctx, bit_length, max_aggregation, extension_degree, h_base32_or_null, g_bases32_or_null, &l.gens);
        if (rc) break;
    }
    if (rc) {
        for (QLane &l : q->lanes) { if (l.gens) bpp_gens_destroy(l.gens); if (l.ctx) bpp_ctx_destroy(l.ctx); }
        delete q;
        return rc;
    }
    for (size_t i = 0; i < q->lanes.size(); i++) q->lanes[i].th = std::thread(lane_loop, q, i);
    *out = q;
    return BPP_OK;
}

void bpp_vqueue_destroy(bpp_vqueue *q) {
    if (!q) return;
    {
        std::lock_guard<std::mutex> lk(q->mu);
        q->stopping = true;
    }
    q->cv_work.notify_all();
    for (QLane &l : q->lanes) if (l.th.joinable()) l.th.join();
    for (QLane &l : q->lanes) { bpp_gens_destroy(l.gens); bpp_ctx_destroy(l.ctx); }
    for (auto &kv : q->calls) delete kv.second;
    delete q;
}

int32_t bpp_vqueue_submit(bpp_vqueue *q, const bpp_verify_args *args, int32_t *chunk_status, uint8_t *masks32, uint8_t *mask_present,
                          uint64_t *ticket) {
    if (!q || !args || !chunk_status || !ticket) return BPP_INVALID_ARGUMENT;
    QCall *c = new QCall();
    c->args = *args;
    c->chunk_status = chunk_status; c->masks32 = masks32; c->mask_present = mask_present;
    {
        std::lock_guard<std::mutex> lk(q->mu);
        if (q->stopping) { delete c; return BPP_INVALID_ARGUMENT; }
        c->ticket = q->next_ticket++;
        q->calls[c->ticket] = c;
        q->pending.push_back(c);
        *ticket = c->ticket;
    }
    q->cv_work.notify_one();
    return BPP_OK;
}

int32_t bpp_vqueue_wait(bpp_vqueue *q, uint64_t ticket) {
    if (!q) return BPP_INVALID_ARGUMENT;
    QCall *c = nullptr;
    {
        std::unique_lock<std::mutex> lk(q->mu);
        auto it = q->calls.find(ticket);
        if (it == q->calls.end()) return BPP_INVALID_ARGUMENT;
        c = it->second;
        q->cv_done.wait(lk, [&] { return c->done; });
        q->calls.erase(it);
    }
    const int32_t rc = c->rc;
    delete c;
    return rc;
}

int32_t bpp_vqueue_verify(bpp_vqueue *q, const bpp_verify_args *args, int32_t *chunk_status, uint8_t *masks32, uint8_t *mask_present) {
    uint64_t t = 0;
    int32_t rc = bpp_vqueue_submit(q, args, chunk_status, masks32, mask_present, &t);
    if (rc) return rc;
    return bpp_vqueue_wait(q, t);
}

int32_t bpp_vqueue_stats(bpp_vqueue *q, uint64_t out5[5]) {
    if (!q || !out5) return BPP_INVALID_ARGUMENT;
    out5[0] = q->n_passes; out5[1] = q->n_calls; out5[2] = q->n_proofs;
    uint64_t k = 0, g = 0;
    for (QLane &l : q->lanes) { k += bpp_ctx_launch_count(l.ctx); g += bpp_ctx_graph_launch_count(l.ctx); }
    out5[3] = k; out5[4] = g;
    return BPP_OK;
}

int32_t bpp_vqueue_lane_ms(bpp_vqueue *q, double out4[4]) {
    if (!q || !out4) return BPP_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(q->mu);
    for (int k = 0; k < 4; k++) out4[k] = 0;
    for (QLane &l : q->lanes) for (int k = 0; k < 4; k++) out4[k] += l.ms[k];
    return BPP_OK;
}

int32_t bpp_vqueue_set_device_weights(bpp_vqueue *q, int32_t enable) {
    if (!q) return BPP_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(q->mu);             // lanes read their mode at the start of a pass
    for (QLane &l : q->lanes) bpp_ctx_set_throughput_mode(l.ctx, enable ? 2 : 1);
    return BPP_OK;
}

int32_t bpp_vqueue_set_merged_check(bpp_vqueue *q, int32_t enable) {
    if (!q) return BPP_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(q->mu);
    for (QLane &l : q->lanes) bpp_ctx_set_merged_check(l.ctx, enable);
    return BPP_OK;
}

int32_t bpp_vqueue_lanes(const bpp_vqueue *q) { return q ? (int32_t)q->lanes.size() : 0; }

} // extern "C"
