// Internal host-side definitions shared by the C-ABI translation units (engine.cu, engine_verify.cu, engine_prove.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/bpp_b200.h"
#include "kernels.cuh"
#include "hostpool.hpp"

namespace bpp {

// grow-only device / pinned-host buffers: the hot entry points are called repeatedly with similar sizes
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

} // namespace bpp

struct bpp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;     // side stream: point decompression overlaps the scalar prep chain
    cudaStream_t stream3 = nullptr;     // side stream: device-side verifier-weight transcripts (throughput mode)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_mid = nullptr, ev_fork2 = nullptr, ev_join2 = nullptr;
    cudaEvent_t ev_done = nullptr;      // end of a pass; polled between short sleeps in throughput mode
    long nap_ns = 60000;                // sleep between event polls in throughput mode (BPP_NAP_US)
    bool throughput_mode = false;       // sleeping polls instead of spinning: see bpp_ctx_set_throughput_mode
    bool adaptive_wait = false;         // BPP_ADAPTIVE_WAIT=1: first sleep of a wait = 3/4 of what the same wait took last time, longer naps after it (wait_sleeping)
    double wait_ema_ns[2] = {0, 0};     // running estimate of the two waits of a pass (replay results, end of pass)
    uint32_t test_hooks = 0;            // bpp_ctx_set_test_hooks: bit 0 = repeat every pass through the zero-weight fallback
    bool scalar_weights = false;        // test hook (BPP_SCALAR_WEIGHTS=1): one weight transcript at a time instead of four in lock-step
    bool merged_check = false;          // bpp_ctx_set_merged_check: one multiscalar check per PASS, chunk by chunk only when it fails (engine_verify.cu)
    uint64_t merged_fallbacks = 0;      // passes whose merged check failed and were settled chunk by chunk
    bool device_weights = false;        // whole pass as ONE graph with the verifier-weight transcripts on the device (k_weights)
    bool device_replay = true;          // loop 1 (transcript replay) on the device (k_replay.cu) or on host threads
    int replay_kernel = 0;              // 0 = by batch size, 1 = one thread per proof, 2 = one warp per proof
    std::string err;
    uint64_t launches = 0;
    int host_threads = 1;
    uint64_t io_bytes[2] = {0, 0};      // host->device / device->host bytes moved by the last verification call
    double host_ms[8] = {};             // wall-clock breakdown of the last bpp_vbatch_create (see bpp_ctx_host_ms)
    bpp::HostPool *pool = nullptr;      // lazily created with host_threads workers
    bpp::HostPool &workers() {
        if (!pool || pool->size() != host_threads) { delete pool; pool = new bpp::HostPool(host_threads); }
        return *pool;
    }
    // measurement: wall timer and per-phase marks on `stream` (bench.py)
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    bool phase_timing = false;
    // 11 phases between 12 marks: replay, decompress, vprep_proof, vprep_vector, (host weight transcripts: device idle),
    // vprep_weigh, msm sort / bucket / reduce / combine, encode
    static constexpr int N_MARKS = 12;
    cudaEvent_t ph[N_MARKS] = {};
    bool ph_set[N_MARKS] = {};
    void mark(int i) { if (phase_timing && ph[i]) { cudaEventRecord(ph[i], stream); ph_set[i] = true; } }
    void clear_marks() { for (int i = 0; i < N_MARKS; i++) ph_set[i] = false; }
    // reusable scratch for the one-shot entry points
    bpp::DevBuf d_in, d_in2, d_tab, d_flags, d_out, d_scratch, d_res, d_misc, d_flush;
    bpp::PinBuf h_stage, h_stage2;
    std::vector<void *> vwork_pool;     // pooled verification workspaces (engine_verify.cu)
    std::vector<void *> vgraphs;        // captured verification passes, keyed by workspace + layout (engine_verify.cu)
    uint64_t vgraph_clock = 0, graph_launches = 0;
    bool use_graphs = true;             // replay captured CUDA graphs instead of issuing the ~35 driver calls of a pass one by one
    void *prove_ws = nullptr;           // persistent prover workspace (engine_prove.cu)
};

struct bpp_gens {
    bpp_ctx *ctx = nullptr;
    int n = 0, M = 0, ext = 0;
    size_t nm = 0;                       // n * M
    bpp::DevBuf d_table;                 // aniels[2*nm + ext + 1]: Gi | Hi | G | H
    bpp::DevBuf d_fb;                    // fixed-base window tables over the same generators (k_fb.cu), built on first use
    bpp::FbShape fb = {0, 0, 0, 0};
    bool custom_bases = false;           // Pedersen bases supplied by the caller (bpp_gens_create_with_bases)
    int fb_state = 0;                    // 0 = not built, 1 = ready, -1 = over the memory budget (callers use the folding path)
    std::vector<uint8_t> enc;            // (2*nm + ext + 1) x 32 B compressed, same order
    const uint8_t *gi(size_t i) const { return enc.data() + 32 * i; }
    const uint8_t *hi(size_t i) const { return enc.data() + 32 * (nm + i); }
    const uint8_t *g(size_t k) const { return enc.data() + 32 * (2 * nm + k); }
    const uint8_t *h() const { return enc.data() + 32 * (2 * nm + ext); }
    size_t table_len() const { return 2 * nm + (size_t)ext + 1; }
};

namespace bpp {
void vwork_pool_free(bpp_ctx *ctx);
// builds g->d_fb on first use; false if the tables would exceed the budget (BPP_FB_MAX_MB, default 2048) or the build failed
bool gens_fb_ensure(bpp_gens *g);
void vgraph_cache_free(bpp_ctx *ctx);
void prove_ws_free(bpp_ctx *ctx);
int32_t fail(bpp_ctx *ctx, int32_t code, const char *what);
int32_t cuda_fail(bpp_ctx *ctx, cudaError_t e, const char *where);
#define BPP_CUDA(ctx, call)                                             \
    do {                                                                \
        cudaError_t _e = (call);                                        \
        if (_e != cudaSuccess) return bpp::cuda_fail((ctx), _e, #call); \
    } while (0)

// host scalar helpers over the shared arithmetic header (portable path)
bool host_sc_is_canonical(const uint8_t *b32);
void host_sc_from_wide(const uint8_t in64[64], uint8_t out32[32]);
} // namespace bpp
