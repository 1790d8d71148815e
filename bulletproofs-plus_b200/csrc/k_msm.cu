// K-MSM: signed-window Pippenger multiscalar multiplication over Ristretto255, single or segmented
// (n_seg independent MSMs per launch sequence).
//
// Replaces P::vartime_multiscalar_mul / Precomputation::vartime_mixed_multiscalar_mul as issued from
// /root/reference/src/range_proof.rs:339-345, :482-495, :1050-1057 (the reference reaches dalek's Straus /
// Pippenger there; the algorithm is free because only the compressed result / the identity test is observable).
//
// Pipeline (all on one stream):
//   1. k_msm_count    one thread per (scalar): signed c-bit digits -> histogram over keys (seg, window, bucket)
//   2. scan           exclusive prefix sum of the histogram (three small kernels)
//   3. k_msm_scatter  counting-sort scatter of (entry | sign) into bucket order  (HBM-bound phase)
//   4. k_msm_bucket   one thread per bucket: sum its points (affine-Niels mixed adds)   (IMAD-bound phase)
//   5. k_msm_reduce   one block per (seg, window): sum_b (b+1)*bucket[b] by chunked running sums + tree reduce
//   6. k_msm_combine  one thread per segment: Horner over the windows
// field multiplications are inlined in this translation unit: every kernel here is a short loop around two or three of them
#define BPP_INLINE_MUL
#include <stdlib.h>
#include "kernels.cuh"
#include "quad.cuh"

namespace bpp {

// ------------------------------------------------------------------------------------------------ helpers
static __device__ __forceinline__ aniels ld_aniels(const aniels *p) {
    aniels q;
    ld8(q.ypx.v, p->ypx.v); ld8(q.ymx.v, p->ymx.v); ld8(q.t2d.v, p->t2d.v);
    return q;
}
static __device__ __forceinline__ ge ld_ge(const ge *p) {
    ge r;
    ld8(r.X.v, p->X.v); ld8(r.Y.v, p->Y.v); ld8(r.Z.v, p->Z.v); ld8(r.T.v, p->T.v);
    return r;
}
static __device__ __forceinline__ void st_ge(ge *p, const ge &r) {
    st8(p->X.v, r.X.v); st8(p->Y.v, r.Y.v); st8(p->Z.v, r.Z.v); st8(p->T.v, r.T.v);
}
// c-bit field of a 256-bit little-endian scalar at bit offset `off`
static __device__ __forceinline__ uint32_t bits_at(const uint32_t s[8], int off, int c) {
    int wi = off >> 5, sh = off & 31;
    if (wi > 7) return 0;
    uint64_t lo = s[wi], hi = (wi + 1 <= 7) ? s[wi + 1] : 0;
    uint64_t v = (lo | (hi << 32)) >> sh;
    return (uint32_t)v & ((1u << c) - 1u);
}

static __device__ __forceinline__ uint32_t seg_of(const uint32_t *seg_offsets, uint32_t n_seg, uint32_t i) {
    if (n_seg <= 1) return 0;
    uint32_t lo = 0, hi = n_seg;   // invariant: seg_offsets[lo] <= i < seg_offsets[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (seg_offsets[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// the switches below select kernel variants (tests / experiments); getenv is neither cheap nor safe against a concurrent setenv, so
// they are read once
namespace {
struct MsmKnobs { int bucket, split, reduce, reduce_parts, scan_sort, no_heavy, window_sort, local_rank; };
const MsmKnobs &knobs() {
    static const MsmKnobs k = [] {
        auto geti = [](const char *n, int dflt) { const char *e = getenv(n); return e ? atoi(e) : dflt; };
        MsmKnobs r;
        r.bucket = geti("BPP_MSM_BUCKET", 0);                 // 1 = quads, 2 = threads, 3 = split threads
        r.split = geti("BPP_MSM_SPLIT", 4);
        r.reduce = geti("BPP_MSM_REDUCE", 0);                 // 1 = CTA of quads, 2 = warp of threads
        r.reduce_parts = geti("BPP_MSM_REDUCE_PARTS", 0);
        r.window_sort = geti("BPP_MSM_WINDOW_SORT", 0);         // 1 = shared-memory sort one (segment, window) per CTA (the first form)
        r.local_rank = geti("BPP_MSM_LOCAL_RANK", 0);           // 1 = buckets ranked by size inside each CTA only (the round-1 form)
        r.no_heavy = geti("BPP_MSM_NO_HEAVY", 0);               // 1 = over-full buckets stay with the ordinary bucket kernels (comparison)
        r.scan_sort = geti("BPP_MSM_SCAN_SORT", 0);             // 1 = always the scan-based counting sort (tests: both sorts give the same sums)
        return r;
    }();
    return k;
}
}
void msm_knobs(int32_t out[4]) { const MsmKnobs &k = knobs(); out[0] = k.bucket; out[1] = k.split; out[2] = k.reduce; out[3] = k.reduce_parts | (k.scan_sort << 8) | (k.no_heavy << 9) | (k.window_sort << 10) | (k.local_rank << 11); }

// ------------------------------------------------------------------------------------------------ shape
MsmShape msm_shape(uint32_t n_entries, uint32_t n_seg, int forced_c) {
    MsmShape sh;
    sh.n_entries = n_entries;
    sh.n_seg = n_seg ? n_seg : 1;
    sh.max_seg_entries = 0;
    int c = forced_c;
    static const int env_c = getenv("BPP_MSM_C") ? atoi(getenv("BPP_MSM_C")) : 0;       // experiments only
    if (c <= 0 && env_c > 0 && n_entries / sh.n_seg < 16384) c = env_c;
    if (c <= 0) {
        // Small segments (the verifier's 4226-entry chunks, the prover's L / R): cost in quad stages (one warp-wide field
        // multiplication each) per segment -- a bucket add is 2 stages, a bucket in the running sums 5 (add + re-cache +
        // add), a window of the Horner combine 2c + 3.  Large segments are throughput-bound in the bucket sums
        // (~0.136 ns per add measured at 2^22) and latency-bound in the running sums (~48 ns per bucket of one window).
        // Recoded scalars are < 2^251, so the top window only has t = 251 - c(W-1) magnitude bits: windows whose top is less
        // than a quarter full are skipped once segments are large (their few non-empty buckets would each get n / 2^t
        // entries: c = 13 at 2^20 points measured 120 ms against 3.9 ms for c = 14).
        double per = (double)n_entries / (double)sh.n_seg;
        double best = 1e300;
        for (int cc = 2; cc <= 16; cc++) {
            int W = (252 + cc - 1) / cc;
            int top_bits = 251 - cc * (W - 1);
            // (over-full buckets are split since round 2 -- section 4b -- so a sparse top window no longer serialises; measured, the
            // wider windows it would allow still lose: c = 16 at 2^24 points 567 vs 575 M points/s for c = 14, c = 15 at 2^20 382 vs 474:
            // fewer additions, but shorter buckets per thread, a slower scatter over 4x the keys and a longer window reduction)
            if (per >= 1024 && top_bits < cc - 3) continue;
            double B = (double)(1u << (cc - 1));
            double cost;
            if (per >= 16384 && sh.n_seg == 1) {
                // one large MSM, in ns (scripts/msm_sweep.py, msm_bucket_probe.py): digit sort 0.02 per (entry, window); bucket sums
                // 0.09 per addition with whole threads per bucket (enough buckets for that: launch_msm), 0.15 with split threads;
                // window reduction, split over up to 8 CTAs per window: 54 us at 256 buckets, 112 at 1024, 138 at 8192, 250 / 340 at
                // 2^14 / 2^15 buckets.  Picks c = 14 from 2^16 points on (2^16: 127 M points/s against 106 M at c = 11; 2^18: 304 M
                // against 253 M at c = 12)
                const double keys = (double)W * B;
                const double red = B <= 1024 ? 35e3 + 75.0 * B : 110e3 + 3.4 * B + (B > 8192 ? 7.0 * (B - 8192) : 0.0);
                cost = per * W * ((keys >= 4.0 * 148 * 4 * 32 ? 0.09 : 0.15) + 0.02) + red;
            } else if (per >= 16384) {
                cost = per * W * 0.136 + B * 48.0 * ((double)sh.n_seg * W / 148.0);
            } else {
                cost = 2.0 * per * W + 5.0 * B * W + (2.0 * cc + 3.0) * W;
            }
            if (cost < best) { best = cost; c = cc; }
        }
    }
    if (c < 2) c = 2;
    if (c > 16) c = 16;
    sh.c = c;
    sh.W = (252 + c - 1) / c;    // scalars are recoded to |s| < 2^251 (see k_msm_digits)
    sh.B = 1u << (c - 1);
    return sh;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
#define SCAN_TILE 4096u
#define REDUCE_PARTS_MAX 8u
// over-full buckets (section 4b): a bucket with more than 2 * part entries is cut into parts of `part` entries, part = max(64, twice
// the average bucket size of the sum)
static inline uint32_t heavy_part_size(const MsmShape &sh) {
    const size_t n_keys = (size_t)sh.n_seg * sh.W * sh.B;
    const size_t avg = n_keys ? ((size_t)sh.n_entries * sh.W) / n_keys : 0;
    return (uint32_t)(avg * 2 > 64 ? avg * 2 : 64);
}

struct MsmScratch {
    uint32_t *starts;   // n_keys + 1
    uint32_t *cursor;   // n_keys (scan-based sort: scatter cursors; shared-memory sort: per-bucket counts)
    uint32_t *tile_sums;
    uint32_t *sorted;   // n_entries * W
    cached *buckets;    // n_keys
    ge *windows;        // n_seg * W
    ge *wparts;         // n_seg * W * REDUCE_PARTS_MAX: partial window sums of k_msm_reduce when a window is split over several CTAs
    uint32_t *perm;     // n_keys: buckets ranked by size (k_msm_size_*)
    uint32_t *size_bins;// 256
    uint32_t *heavy_n;  // number of work items of the over-full buckets (see "4b")
    uint2 *heavy_items; // (key, part)
    ge *heavy_parts;    // partial sum of every item
    uint32_t heavy_cap;
    size_t total;
};
static MsmScratch msm_carve(const MsmShape &sh, void *base) {
    size_t n_keys = (size_t)sh.n_seg * sh.W * sh.B;
    size_t n_tiles = (n_keys + SCAN_TILE - 1) / SCAN_TILE;
    char *p = (char *)base;
    size_t off = 0;
    MsmScratch s;
    s.starts = (uint32_t *)(p + off); off = align_up(off + (n_keys + 1) * 4, 256);
    s.cursor = (uint32_t *)(p + off); off = align_up(off + n_keys * 4, 256);
    s.tile_sums = (uint32_t *)(p + off); off = align_up(off + (n_tiles + 1) * 4, 256);
    s.sorted = (uint32_t *)(p + off); off = align_up(off + (size_t)sh.n_entries * sh.W * 4 + 4, 256);
    s.buckets = (cached *)(p + off); off = align_up(off + n_keys * sizeof(cached), 256);
    s.windows = (ge *)(p + off); off = align_up(off + (size_t)sh.n_seg * sh.W * sizeof(ge), 256);
    s.wparts = (ge *)(p + off); off = align_up(off + (size_t)sh.n_seg * sh.W * REDUCE_PARTS_MAX * sizeof(ge), 256);
    // sum over the over-full buckets of ceil(count / part) <= total / part + (number of them) <= 1.5 * total / part
    s.heavy_cap = (uint32_t)((3 * ((size_t)sh.n_entries * sh.W)) / (2 * (size_t)heavy_part_size(sh)) + 1024);
    s.perm = (uint32_t *)(p + off); off = align_up(off + n_keys * 4, 256);
    s.size_bins = (uint32_t *)(p + off); off = align_up(off + 1024, 256);
    s.heavy_n = (uint32_t *)(p + off); off = align_up(off + 256, 256);
    s.heavy_items = (uint2 *)(p + off); off = align_up(off + (size_t)s.heavy_cap * sizeof(uint2), 256);
    s.heavy_parts = (ge *)(p + off); off = align_up(off + (size_t)s.heavy_cap * sizeof(ge), 256);
    s.total = off;
    return s;
}
size_t msm_scratch_bytes(const MsmShape &sh) { return msm_carve(sh, nullptr).total; }

// ------------------------------------------------------------------------------------------------ 1/3: count + scatter
template <bool SCATTER>
__global__ void __launch_bounds__(256) k_msm_digits(uint32_t n_entries, uint32_t n_seg, int c, int W, uint32_t B,
                                                   const uint32_t *__restrict__ scalars, const uint32_t *__restrict__ seg_offsets,
                                                   uint32_t *__restrict__ counters, uint32_t *__restrict__ sorted) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_entries) return;
    uint32_t s[8];
    ld8(s, scalars + 8 * (size_t)i);
    // s > l/2  ->  use (l - s, -P): the recoded scalar is < 2^251, so with c*W >= 252 the top window never overflows and
    // no half-empty carry-only window exists (which would put half of all entries into one bucket)
    uint32_t sneg;
    {
        uint32_t t[8];
        int64_t bw = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { bw += (int64_t)sc_l(k) - (int64_t)s[k]; t[k] = (uint32_t)bw; bw >>= 32; }
        bool gt = false, decided = false;      // s > t ?
#pragma unroll
        for (int k = 7; k >= 0; k--) {
            if (!decided && s[k] != t[k]) { gt = s[k] > t[k]; decided = true; }
        }
        sneg = gt ? 1u : 0u;
        if (gt) {
#pragma unroll
            for (int k = 0; k < 8; k++) s[k] = t[k];
        }
    }
    uint32_t seg = seg_of(seg_offsets, n_seg, i);
    uint32_t key_base = seg * (uint32_t)W * B;
    uint32_t carry = 0;
    for (int w = 0; w < W; w++) {
        uint32_t d = bits_at(s, w * c, c) + carry;
        uint32_t neg = sneg;
        if (d > B) { d = (2u * B) - d; neg ^= 1u; carry = 1u; } else carry = 0u;
        if (d != 0) {
            uint32_t key = key_base + (uint32_t)w * B + (d - 1u);
            uint32_t pos = atomicAdd(&counters[key], 1u);
            if (SCATTER) sorted[pos] = i | (neg << 31);
        }
    }
}

// ------------------------------------------------------------------------------------------------ 1': sort in shared memory
// Small segments (the verifier's <= 256-proof chunks: ~4.2 k entries, 256 buckets per window): ONE CTA sorts one (segment, window)
// entirely in shared memory -- digit of every entry of the segment for this window, histogram, exclusive scan, scatter -- and writes
// the window's records as one contiguous run plus (start, count) per bucket.  Replaces memset + count pass + three scan kernels +
// scatter pass (6 launches, two rounds of global atomics on n_seg * W * B counters) by one launch without global atomics.
// The signed recoding carries from window to window; the carry INTO window w is 1 iff the low c*w bits of the scalar exceed
// T_w = B * (2^(c w) - 1) / (2^c - 1)  (digits B, B, .., B in base 2^c; B = 2^(c-1)): induction over "digit + carry > B".
// Region of (segment, window) inside `sorted`: seg_lo * W + w * len (len = entries of the segment): the same n_entries * W words as the
// scan-based layout, records of empty (zero) digits simply missing at the end of each region.
template <int BMAX>
__global__ void __launch_bounds__(256) k_msm_sort_seg(uint32_t n_entries, uint32_t n_seg, int c, int W, uint32_t B, const uint32_t *__restrict__ scalars,
                                                     const uint32_t *__restrict__ seg_offsets, uint32_t *__restrict__ starts, uint32_t *__restrict__ counts,
                                                     uint32_t *__restrict__ sorted) {
    __shared__ uint32_t s_hist[BMAX], s_scan[256], s_thr[8];
    const uint32_t seg = blockIdx.x / (uint32_t)W, w = blockIdx.x % (uint32_t)W, tid = threadIdx.x;
    const uint32_t lo = n_seg > 1 ? seg_offsets[seg] : 0u, hi = n_seg > 1 ? seg_offsets[seg + 1] : n_entries;
    const uint32_t len = hi - lo;
    const uint32_t region = lo * (uint32_t)W + w * len;
    const uint32_t key_base = (seg * (uint32_t)W + w) * B;
    for (uint32_t bkt = tid; bkt < B; bkt += 256) s_hist[bkt] = 0;
    if (tid < 8) {                                   // T_w: bit (c - 1) + c k set for k < w
        uint32_t t = 0;
        for (uint32_t k = 0; k < w; k++) { const uint32_t bit = (uint32_t)(c - 1) + (uint32_t)c * k; if ((bit >> 5) == tid) t |= 1u << (bit & 31); }
        s_thr[tid] = t;
    }
    __syncthreads();
    const int lowbits = c * (int)w;                  // bits of the scalar below this window
    auto digit_of = [&](uint32_t i, uint32_t &neg) -> uint32_t {
        uint32_t s[8];
        ld8(s, scalars + 8 * (size_t)i);
        // s > l/2  ->  (l - s, -P): the recoded scalar is < 2^251 (see k_msm_digits)
        uint32_t t[8];
        int64_t bw = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { bw += (int64_t)sc_l(k) - (int64_t)s[k]; t[k] = (uint32_t)bw; bw >>= 32; }
        bool gt = false, decided = false;
#pragma unroll
        for (int k = 7; k >= 0; k--) if (!decided && s[k] != t[k]) { gt = s[k] > t[k]; decided = true; }
        neg = gt ? 1u : 0u;
        if (gt) {
#pragma unroll
            for (int k = 0; k < 8; k++) s[k] = t[k];
        }
        // carry into this window: (s mod 2^lowbits) > T_w
        uint32_t carry = 0;
        {
            bool g2 = false, dec2 = false;
#pragma unroll
            for (int k = 7; k >= 0; k--) {
                const int base = 32 * k;
                uint32_t m = lowbits >= base + 32 ? 0xffffffffu : lowbits <= base ? 0u : ((1u << (lowbits - base)) - 1u);
                const uint32_t a = s[k] & m, th = s_thr[k];
                if (!dec2 && a != th) { g2 = a > th; dec2 = true; }
            }
            carry = g2 ? 1u : 0u;
        }
        uint32_t dg = bits_at(s, lowbits, c) + carry;
        if (dg > B) { dg = 2u * B - dg; neg ^= 1u; }
        return dg;
    };
    for (uint32_t i = lo + tid; i < hi; i += 256) {
        uint32_t neg;
        const uint32_t dg = digit_of(i, neg);
        if (dg) atomicAdd(&s_hist[dg - 1u], 1u);
    }
    __syncthreads();
    // exclusive scan of the B counters: thread t owns buckets [t * per, (t + 1) * per)
    const uint32_t per = (B + 255u) / 256u;
    uint32_t acc = 0;
    for (uint32_t k = 0; k < per; k++) { const uint32_t bkt = tid * per + k; if (bkt < B) acc += s_hist[bkt]; }
    s_scan[tid] = acc;
    __syncthreads();
    for (int sft = 1; sft < 256; sft <<= 1) {
        const uint32_t t = (int)tid >= sft ? s_scan[tid - sft] : 0u;
        __syncthreads();
        s_scan[tid] += t;
        __syncthreads();
    }
    uint32_t run = s_scan[tid] - acc;
    for (uint32_t k = 0; k < per; k++) {
        const uint32_t bkt = tid * per + k;
        if (bkt < B) {
            const uint32_t cnt = s_hist[bkt];
            starts[key_base + bkt] = region + run;
            counts[key_base + bkt] = cnt;
            s_hist[bkt] = run;                       // becomes the bucket's cursor
            run += cnt;
        }
    }
    __syncthreads();
    for (uint32_t i = lo + tid; i < hi; i += 256) {
        uint32_t neg;
        const uint32_t dg = digit_of(i, neg);
        if (dg) {
            const uint32_t pos = atomicAdd(&s_hist[dg - 1u], 1u);
            sorted[region + pos] = i | (neg << 31);
        }
    }
}

// The same for a GROUP of windows per CTA, for the verifier's width (C = 9, W = 28): CTA (segment, g) handles windows
// [g * G, (g + 1) * G) -- G * B counters in shared memory -- and walks the segment's scalars twice (count, scatter).  The carry into the
// group's first window comes from the closed form, the G digits follow sequentially with compile-time bit positions (the scalar stays
// in registers), so the redundant recoding of k_msm_sort_seg (once per window and pass: 134 M warp instructions per 16-job pass, 13 %
// of it, ALU-bound) shrinks by G.  G is chosen so that the grid still covers the machine: scattered 4-byte stores are limited per SM
// (one CTA per segment, G = W, took 120 us however many segments there were; measured).  Regions as in k_msm_sort_seg.
template <int C>
__global__ void __launch_bounds__(512) k_msm_sort_segw(uint32_t n_entries, uint32_t n_seg, int G, const uint32_t *__restrict__ scalars,
                                                      const uint32_t *__restrict__ seg_offsets, uint32_t *__restrict__ starts, uint32_t *__restrict__ counts,
                                                      uint32_t *__restrict__ sorted) {
    constexpr int W = (252 + C - 1) / C;
    constexpr uint32_t B = 1u << (C - 1);
    extern __shared__ uint32_t s_cnt[];              // G * B counters, then cursors (relative to the window's region)
    __shared__ uint32_t s_part[512], s_thr[8];
    const uint32_t seg = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
    const int w0 = (int)blockIdx.y * G, w1 = min(W, w0 + G), nw = w1 - w0;
    const uint32_t lo = n_seg > 1 ? seg_offsets[seg] : 0u, hi = n_seg > 1 ? seg_offsets[seg + 1] : n_entries;
    const uint32_t len = hi - lo, nk = (uint32_t)nw * B;
    const uint32_t key_base = (seg * (uint32_t)W + (uint32_t)w0) * B, region0 = lo * (uint32_t)W + (uint32_t)w0 * len;
    for (uint32_t k = tid; k < nk; k += nthr) s_cnt[k] = 0;
    if (tid < 8) {                                   // T_w0: bit (C - 1) + C k set for k < w0
        uint32_t t = 0;
        for (int k = 0; k < w0; k++) { const uint32_t bit = (uint32_t)(C - 1) + (uint32_t)C * (uint32_t)k; if ((bit >> 5) == tid) t |= 1u << (bit & 31); }
        s_thr[tid] = t;
    }
    __syncthreads();
    const int lowbits = C * w0;
    auto walk = [&](uint32_t i, bool scatter) {
        uint32_t s[8], t[8];
        ld8(s, scalars + 8 * (size_t)i);
        int64_t bw = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { bw += (int64_t)sc_l(k) - (int64_t)s[k]; t[k] = (uint32_t)bw; bw >>= 32; }
        bool gt = false, decided = false;            // s > l - s: use (l - s, -P), |recoded| < 2^251
#pragma unroll
        for (int k = 7; k >= 0; k--) if (!decided && s[k] != t[k]) { gt = s[k] > t[k]; decided = true; }
        if (gt) {
#pragma unroll
            for (int k = 0; k < 8; k++) s[k] = t[k];
        }
        uint32_t carry;                              // into window w0: (s mod 2^lowbits) > T_w0
        {
            bool g2 = false, dec2 = false;
#pragma unroll
            for (int k = 7; k >= 0; k--) {
                const int base = 32 * k;
                const uint32_t m = lowbits >= base + 32 ? 0xffffffffu : lowbits <= base ? 0u : ((1u << (lowbits - base)) - 1u);
                const uint32_t a = s[k] & m, th = s_thr[k];
                if (!dec2 && a != th) { g2 = a > th; dec2 = true; }
            }
            carry = g2 ? 1u : 0u;
        }
#pragma unroll
        for (int w = 0; w < W; w++) {
            if (w < w0 || w >= w1) continue;
            constexpr uint32_t mask = (1u << C) - 1u;
            const int off = w * C, wi = off >> 5, sh = off & 31;          // compile-time after unrolling
            uint32_t d = s[wi] >> sh;
            if (sh + C > 32 && wi + 1 < 8) d |= s[wi + 1] << (32 - sh);
            d = (d & mask) + carry;
            uint32_t neg = gt ? 1u : 0u;
            if (d > B) { d = 2u * B - d; neg ^= 1u; carry = 1u; } else carry = 0u;
            if (d) {
                const uint32_t pos = atomicAdd(&s_cnt[(uint32_t)(w - w0) * B + d - 1u], 1u);
                if (scatter) sorted[region0 + (uint32_t)(w - w0) * len + pos] = i | (neg << 31);
            }
        }
    };
    for (uint32_t i = lo + tid; i < hi; i += nthr) walk(i, false);
    __syncthreads();
    // exclusive scan over the nw * B counters (thread t owns a contiguous run), then made relative to each window's first bucket
    const uint32_t per = (nk + nthr - 1u) / nthr;
    uint32_t acc = 0;
    for (uint32_t k = 0; k < per; k++) { const uint32_t kk = tid * per + k; if (kk < nk) acc += s_cnt[kk]; }
    s_part[tid] = acc;
    __syncthreads();
    for (uint32_t sft = 1; sft < nthr; sft <<= 1) {
        const uint32_t v = tid >= sft ? s_part[tid - sft] : 0u;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    uint32_t run = s_part[tid] - acc;
    for (uint32_t k = 0; k < per; k++) {
        const uint32_t kk = tid * per + k;
        if (kk < nk) { const uint32_t cnt = s_cnt[kk]; s_cnt[kk] = run; run += cnt; counts[key_base + kk] = cnt; }
    }
    __syncthreads();
    for (uint32_t w = tid; w < (uint32_t)nw; w += nthr) s_part[w] = s_cnt[w * B];      // scan value at each window's first bucket
    __syncthreads();
    for (uint32_t kk = tid; kk < nk; kk += nthr) {
        const uint32_t w = kk / B, rel = s_cnt[kk] - s_part[w];                          // position inside the window's region
        starts[key_base + kk] = region0 + w * len + rel;
        s_cnt[kk] = rel;                                                                 // cursor of the scatter pass
    }
    __syncthreads();
    for (uint32_t i = lo + tid; i < hi; i += nthr) walk(i, true);
}

// ------------------------------------------------------------------------------------------------ 2: scan
__global__ void __launch_bounds__(256) k_scan_tile_sums(const uint32_t *__restrict__ in, size_t n, uint32_t *__restrict__ tile_sums) {
    __shared__ uint32_t sh[256];
    size_t base = (size_t)blockIdx.x * SCAN_TILE;
    uint32_t acc = 0;
    for (uint32_t k = threadIdx.x; k < SCAN_TILE; k += 256) {
        size_t idx = base + k;
        if (idx < n) acc += in[idx];
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = sh[0];
}
__global__ void k_scan_tiles(uint32_t *tile_sums, size_t n_tiles) {
    // single thread block, serial over tiles in chunks: n_tiles <= a few thousand
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (size_t t = 0; t < n_tiles; t++) { uint32_t v = tile_sums[t]; tile_sums[t] = acc; acc += v; }
        tile_sums[n_tiles] = acc;
    }
}
// starts[i] = exclusive prefix; also cursor[i] = starts[i]; the grand total lands in starts[n]
__global__ void __launch_bounds__(256) k_scan_apply(uint32_t *__restrict__ counts_then_starts, uint32_t *__restrict__ cursor, size_t n,
                                                   const uint32_t *__restrict__ tile_sums) {
    __shared__ uint32_t sh[256];
    size_t base = (size_t)blockIdx.x * SCAN_TILE;
    constexpr uint32_t per = SCAN_TILE / 256;      // 16 consecutive elements per thread
    uint32_t v[per];
    uint32_t acc = 0;
    size_t first = base + (size_t)threadIdx.x * per;
#pragma unroll
    for (uint32_t k = 0; k < per; k++) {
        size_t idx = first + k;
        v[k] = idx < n ? counts_then_starts[idx] : 0;
        acc += v[k];
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    // Hillis-Steele inclusive scan over 256 thread sums
    for (int s = 1; s < 256; s <<= 1) {
        uint32_t t = ((int)threadIdx.x >= s) ? sh[threadIdx.x - s] : 0;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    uint32_t run = tile_sums[blockIdx.x] + sh[threadIdx.x] - acc;
#pragma unroll
    for (uint32_t k = 0; k < per; k++) {
        size_t idx = first + k;
        if (idx < n) { counts_then_starts[idx] = run; cursor[idx] = run; }
        run += v[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) counts_then_starts[n] = tile_sums[gridDim.x];
}

// ------------------------------------------------------------------------------------------------ 4: bucket sums
// one quad per bucket; the bucket is written in cached form for the running sums.  A warp walks max(count) of its 8 buckets
// (idle quads add the identity), so the 64 buckets of a CTA are first ranked by size and handed to the quads in rank order:
// the 8 buckets of a warp then have neighbouring sizes (bucket sizes of a 4226-entry segment at c = 9 spread 16.5 +- 4, which
// without the ranking costs ~35 % of the additions as padding).
#define BUCKET_CTA 256
__global__ void __launch_bounds__(BUCKET_CTA) k_msm_bucket(uint32_t n_keys, const uint32_t *__restrict__ starts, const uint32_t *__restrict__ counts, const uint32_t *__restrict__ sorted,
                                                   const uint32_t *__restrict__ pidx, const aniels *__restrict__ dyn,
                                                   const aniels *__restrict__ gens, const cached *__restrict__ dync,
                                                   cached *__restrict__ buckets, uint32_t heavy_min) {
    constexpr uint32_t QUADS = BUCKET_CTA / 4;
    __shared__ uint32_t s_cnt[QUADS], s_perm[QUADS];
    const uint32_t quad = threadIdx.x >> 2, k0 = blockIdx.x * QUADS;
    const int role = threadIdx.x & 3, base = (threadIdx.x & 31) & ~3;
    {
        const uint32_t kk = k0 + quad;
        uint32_t cc = kk < n_keys ? (counts ? counts[kk] : starts[kk + 1] - starts[kk]) : 0u;
        if (heavy_min && cc > heavy_min) cc = 0u;
        if (role == 0) s_cnt[quad] = cc;
        __syncthreads();
        uint32_t rank = 0;
#pragma unroll 8
        for (uint32_t j = 0; j < QUADS; j++) { uint32_t o = s_cnt[j]; rank += (o > cc || (o == cc && j < quad)) ? 1u : 0u; }
        if (role == 0) s_perm[rank] = quad;
        __syncthreads();
    }
    const uint32_t k = k0 + s_perm[quad];
    const bool valid = k < n_keys;
    const uint32_t lo = valid ? starts[k] : 0u;
    uint32_t cnt = valid ? (counts ? counts[k] : starts[k + 1] - lo) : 0u;
    if (heavy_min && cnt > heavy_min) cnt = 0u;          // left to k_msm_heavy_*
    const uint32_t maxcnt = __reduce_max_sync(0xffffffffu, cnt);
    fe c = quad_identity(role);
    for (uint32_t j = 0; j < maxcnt; j++) {
        fe q = quad_cached_identity(role);
        if (j < cnt) {
            uint32_t e = sorted[lo + j];
            uint32_t idx = e & 0x7fffffffu;
            uint32_t pi = pidx ? pidx[idx] : idx;
            bool neg = (e >> 31) != 0;
            // -Q swaps (y-x, y+x) and negates 2dxy (2dT)
            if ((pi & 0xc0000000u) == 0x40000000u) {          // projective point in cached form (prover's folded generators)
                const cached *src = dync + (pi & 0x3fffffffu);
                if (role == 0) q = ld_fe(neg ? &src->ypx : &src->ymx);
                else if (role == 1) q = ld_fe(neg ? &src->ymx : &src->ypx);
                else if (role == 2) q = ld_fe(&src->z2);
                else { q = ld_fe(&src->t2d); if (neg) q = fe_sub_ll(fe_zero(), q); }
            } else {
                const aniels *src = (pi & 0x80000000u) ? (gens + (pi & 0x7fffffffu)) : (dyn + pi);
                if (role == 0) q = ld_fe(neg ? &src->ypx : &src->ymx);
                else if (role == 1) q = ld_fe(neg ? &src->ymx : &src->ypx);
                else if (role == 3) { q = ld_fe(&src->t2d); if (neg) q = fe_sub_l(fe_zero(), q); }
            }
        }
        c = quad_add(c, role, base, q);
    }
    fe out = quad_to_cached(c, role, base);
    if (valid) st_fe(reinterpret_cast<fe *>(&buckets[k]) + role, out);
}

// (X : Y : Z : T) += +-Q for one sorted entry e = index | sign << 31: whole-thread mixed addition with an affine-Niels table entry
// (7 multiplications) or with a projective "cached" point (8; the prover's folded generators)
static __device__ __forceinline__ void bucket_add_entry(fe &X, fe &Y, fe &Z, fe &T, uint32_t e, const uint32_t *__restrict__ pidx,
                                                        const aniels *__restrict__ dyn, const aniels *__restrict__ gens,
                                                        const cached *__restrict__ dync) {
    const uint32_t idx = e & 0x7fffffffu;
    const uint32_t pi = pidx ? pidx[idx] : idx;
    const bool neg = (e >> 31) != 0;
    const bool proj = (pi & 0xc0000000u) == 0x40000000u;
    const fe *qm, *qp, *qt;                                        // (y-x, y+x, 2dxy); -Q swaps the first two and negates the third
    fe D;
    if (proj) {
        const cached *src = dync + (pi & 0x3fffffffu);
        qm = &src->ymx; qp = &src->ypx; qt = &src->t2d;
        D = fe_mul(Z, ld_fe(&src->z2));
    } else {
        const aniels *src = (pi & 0x80000000u) ? (gens + (pi & 0x7fffffffu)) : (dyn + pi);
        qm = &src->ymx; qp = &src->ypx; qt = &src->t2d;
        D = fe_add(Z, Z);
    }
    const fe A = fe_mul(fe_sub_l(Y, X), ld_fe(neg ? qp : qm));
    const fe B = fe_mul(fe_add_l(Y, X), ld_fe(neg ? qm : qp));
    const fe C = fe_mul(T, ld_fe(qt));
    const fe E = fe_sub_l(B, A), H = fe_add_l(B, A);
    const fe F0 = fe_sub_l(D, C), G0 = fe_add_l(D, C);
    const fe F = fe_select(F0, G0, neg), G = fe_select(G0, F0, neg);      // -Q: C changes sign
    X = fe_mul(E, F); Y = fe_mul(G, H); Z = fe_mul(F, G); T = fe_mul(E, H);
}

// (X : Y : Z : T) = +-Q for the FIRST entry of a bucket: adding to the identity needs no addition formula.  With (y+x, y-x, 2dxy)
// of an affine Q:  X = (y+x) - (y-x) = 2x, Y = 2y, Z = 2, and T with T Z = X Y is 2xy = (2dxy) / d -- ONE multiplication by the constant
// 1/d instead of the seven of a mixed addition (the verifier's buckets hold ~16.5 entries: 1/16 of the bucket kernel's additions).
// -Q swaps y+x and y-x and takes -1/d.  Results are tight (< 2^255) as the callers expect of an accumulator that may be stored as it is.
static __device__ __forceinline__ void bucket_first_entry(fe &X, fe &Y, fe &Z, fe &T, uint32_t e, const uint32_t *__restrict__ pidx,
                                                          const aniels *__restrict__ dyn, const aniels *__restrict__ gens,
                                                          const cached *__restrict__ dync) {
    const uint32_t idx = e & 0x7fffffffu;
    const uint32_t pi = pidx ? pidx[idx] : idx;
    const bool neg = (e >> 31) != 0;
    if ((pi & 0xc0000000u) == 0x40000000u) {       // projective entries (the prover's folded generators) may hold loose values: the general addition
        X = fe_zero(); Y = fe_one(); Z = fe_one(); T = fe_zero();
        bucket_add_entry(X, Y, Z, T, e, pidx, dyn, gens, dync);
        return;
    }
    const aniels *src = (pi & 0x80000000u) ? (gens + (pi & 0x7fffffffu)) : (dyn + pi);
    const fe *qm = &src->ymx, *qp = &src->ypx, *qt = &src->t2d;
    Z = fe_from_u32(2u);
    const fe a = ld_fe(neg ? qm : qp), b = ld_fe(neg ? qp : qm);
    X = fe_sub(a, b);
    Y = fe_add(a, b);
    T = fe_mul(ld_fe(qt), neg ? fe_const_neg_inv_d() : fe_const_inv_d());
}

// Global size order of the buckets.  k_msm_bucket_thread used to rank the 256 buckets of a CTA among themselves: the lanes of a warp
// then walk buckets of neighbouring sizes, but the CTA lives as long as its largest bucket while most of its warps have left (the
// verifier's buckets hold 16.5 +- 4 entries: 8..28 inside every CTA; ncu: 27 % achieved occupancy, FMA pipe active 40 % of the
// cycles).  Three small kernels rank ALL buckets by size (counting sort on min(count, 255), largest first), so that a CTA's 256
// buckets are of one size and CTAs start in longest-first order.
__global__ void __launch_bounds__(256) k_msm_size_hist(uint32_t n_keys, const uint32_t *__restrict__ starts, const uint32_t *__restrict__ counts,
                                                      uint32_t heavy_min, uint32_t *__restrict__ bins) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t k = blockIdx.x * 1024u + threadIdx.x; k < min(n_keys, (blockIdx.x + 1u) * 1024u); k += 256u) {
        uint32_t cc = counts ? counts[k] : starts[k + 1] - starts[k];
        if (heavy_min && cc > heavy_min) cc = 0u;
        atomicAdd(&h[255u - min(cc, 255u)], 1u);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&bins[threadIdx.x], h[threadIdx.x]);
}
__global__ void __launch_bounds__(256) k_msm_size_scan(uint32_t *__restrict__ bins) {       // bins -> exclusive prefix (cursors), one CTA
    __shared__ uint32_t h[256];
    const uint32_t v = bins[threadIdx.x];
    h[threadIdx.x] = v;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {
        const uint32_t t = (int)threadIdx.x >= d ? h[threadIdx.x - d] : 0u;
        __syncthreads();
        h[threadIdx.x] += t;
        __syncthreads();
    }
    bins[threadIdx.x] = h[threadIdx.x] - v;
}
__global__ void __launch_bounds__(256) k_msm_size_scatter(uint32_t n_keys, const uint32_t *__restrict__ starts, const uint32_t *__restrict__ counts,
                                                         uint32_t heavy_min, uint32_t *__restrict__ bins, uint32_t *__restrict__ perm) {
    __shared__ uint32_t h[256], base[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    uint32_t mybin[4], mypos[4];
    int n = 0;
    for (uint32_t k = blockIdx.x * 1024u + threadIdx.x; k < min(n_keys, (blockIdx.x + 1u) * 1024u); k += 256u) {
        uint32_t cc = counts ? counts[k] : starts[k + 1] - starts[k];
        if (heavy_min && cc > heavy_min) cc = 0u;
        mybin[n] = 255u - min(cc, 255u);
        mypos[n] = atomicAdd(&h[mybin[n]], 1u);
        n++;
    }
    __syncthreads();
    if (h[threadIdx.x]) base[threadIdx.x] = atomicAdd(&bins[threadIdx.x], h[threadIdx.x]);      // one global atomic per (CTA, bin)
    __syncthreads();
    n = 0;
    for (uint32_t k = blockIdx.x * 1024u + threadIdx.x; k < min(n_keys, (blockIdx.x + 1u) * 1024u); k += 256u) {
        perm[base[mybin[n]] + mypos[n]] = k;
        n++;
    }
}

// Throughput variant: one THREAD per bucket (7 sequential multiplications per mixed addition, no shuffles or role selects:
// about half the instructions of the quad kernel per addition).  The 256 buckets of a CTA are counting-sorted by size so that
// the 32 buckets of a warp have neighbouring sizes.  Used when there are enough buckets to fill the machine with whole threads.
template <int MIN_CTAS>
__global__ void __launch_bounds__(256, MIN_CTAS) k_msm_bucket_thread(uint32_t n_keys, const uint32_t *__restrict__ starts, const uint32_t *__restrict__ counts, const uint32_t *__restrict__ sorted,
                                                          const uint32_t *__restrict__ pidx, const aniels *__restrict__ dyn,
                                                          const aniels *__restrict__ gens, const cached *__restrict__ dync,
                                                          cached *__restrict__ buckets, uint32_t heavy_min, uint32_t part_size,
                                                          const uint32_t *__restrict__ heavy_n, const uint2 *__restrict__ items, uint32_t cap,
                                                          ge *__restrict__ parts, const uint32_t *__restrict__ perm) {
    __shared__ uint32_t s_hist[256], s_perm[256];
    const uint32_t tid = threadIdx.x, k0 = blockIdx.x * 256u;
    if (k0 >= n_keys) {
        // CTAs behind the buckets: one thread per PART of an over-full bucket (section 4b); parts are all about part_size long
        const uint32_t it = (blockIdx.x - (n_keys + 255u) / 256u) * 256u + tid;
        if (it >= min(*heavy_n, cap)) return;
        const uint2 item = items[it];
        const uint32_t total = counts ? counts[item.x] : starts[item.x + 1] - starts[item.x];
        const uint32_t lo = starts[item.x] + item.y * part_size, cnt = min(part_size, total - item.y * part_size);
        fe X = fe_zero(), Y = fe_one(), Z = fe_one(), T = fe_zero();
        if (cnt) bucket_first_entry(X, Y, Z, T, sorted[lo], pidx, dyn, gens, dync);
        for (uint32_t j = 1; j < cnt; j++) bucket_add_entry(X, Y, Z, T, sorted[lo + j], pidx, dyn, gens, dync);
        ge *w = parts + it;
        st_fe(&w->X, X); st_fe(&w->Y, Y); st_fe(&w->Z, Z); st_fe(&w->T, T);
        return;
    }
    if (perm) {                                        // globally ranked: slot -> bucket
        s_perm[tid] = k0 + tid < n_keys ? perm[k0 + tid] - k0 : tid;
    } else {
        const uint32_t kk = k0 + tid;
        uint32_t cc = kk < n_keys ? (counts ? counts[kk] : starts[kk + 1] - starts[kk]) : 0u;
        if (heavy_min && cc > heavy_min) cc = 0u;
        const uint32_t bin = 255u - (cc < 255u ? cc : 255u);          // descending sizes
        s_hist[tid] = 0;
        __syncthreads();
        atomicAdd(&s_hist[bin], 1u);
        __syncthreads();
        uint32_t v = s_hist[tid];                                      // inclusive Hillis-Steele scan over the 256 bins
        for (int d = 1; d < 256; d <<= 1) {
            uint32_t t = (int)tid >= d ? s_hist[tid - d] : 0u;
            __syncthreads();
            v += t;
            s_hist[tid] = v;
            __syncthreads();
        }
        const uint32_t before = bin ? s_hist[bin - 1] : 0u;           // buckets in strictly larger-size bins
        __syncthreads();
        s_hist[tid] = 0;                                               // reuse as per-bin cursors
        __syncthreads();
        s_perm[before + atomicAdd(&s_hist[bin], 1u)] = tid;
        __syncthreads();
    }
    const uint32_t k = k0 + s_perm[tid];               // (with perm: s_perm holds perm[slot] - k0, modulo 2^32)
    if (k >= n_keys) return;
    const uint32_t lo = starts[k];
    uint32_t cnt = counts ? counts[k] : starts[k + 1] - lo;
    if (heavy_min && cnt > heavy_min) cnt = 0u;          // left to the part threads above and k_msm_heavy_finish
    fe X = fe_zero(), Y = fe_one(), Z = fe_one(), T = fe_zero();
    if (cnt) bucket_first_entry(X, Y, Z, T, sorted[lo], pidx, dyn, gens, dync);
    for (uint32_t j = 1; j < cnt; j++) bucket_add_entry(X, Y, Z, T, sorted[lo + j], pidx, dyn, gens, dync);
    cached *out = buckets + k;
    st_fe(&out->ymx, fe_sub_l(Y, X));
    st_fe(&out->ypx, fe_add_l(Y, X));
    st_fe(&out->z2, fe_add_l(Z, Z));
    st_fe(&out->t2d, fe_mul(T, fe_const_2d()));
}

// ------------------------------------------------------------------------------------------------ 5: window sums
// `parts` CTAs per (segment, window); CTA (win, part) covers buckets [part*Bp, (part+1)*Bp) with Bp = nq*L, its quad t the L buckets from
// k = part*Bp + t*L: S = sum, R = sum_j (j+1)*bucket[k+j]; the quad contributes R + k*S, and the CTA adds the contributions up.
// Control flow is CTA-uniform (idle quads work on the identity).  With parts > 1 the CTA writes a partial window sum that
// k_msm_window_parts adds up: a large single MSM has only W = 16-20 windows, i.e. as many CTAs, each walking B / 256 buckets per
// quad in sequence (c = 14: 0.46 ms whatever the point count); split eight ways the chains are L = 8 long and 8 W CTAs share the SMs.
__global__ void __launch_bounds__(1024) k_msm_reduce(uint32_t B, uint32_t L, uint32_t nq, uint32_t parts, const cached *__restrict__ buckets,
                                                    ge *__restrict__ windows) {
    __shared__ fe part[32][4];
    const uint32_t t = threadIdx.x >> 2;
    const int role = threadIdx.x & 3, lane = threadIdx.x & 31, base = lane & ~3;
    const uint32_t win = blockIdx.x / parts, k0 = (blockIdx.x % parts) * nq * L;
    const cached *bk = buckets + (size_t)win * B + k0 + (size_t)t * L;
    fe S = quad_identity(role), R = quad_identity(role);
    for (int j = (int)L - 1; j >= 0; j--) {
        fe b = t < nq ? ld_fe(reinterpret_cast<const fe *>(&bk[j]) + role) : quad_cached_identity(role);
        S = quad_add(S, role, base, b);
        R = quad_add(R, role, base, quad_to_cached(S, role, base));
    }
    // R += k * S by uniform double-and-add over the bits of the largest offset in the CTA
    {
        const uint32_t k = t < nq ? k0 + t * L : 0u, kmax = k0 + (nq - 1) * L;
        const fe Sc = quad_to_cached(S, role, base);
        fe M = quad_identity(role);
        for (int bit = 31 - __clz(kmax | 1u); bit >= 0; bit--) {
            M = quad_dbl(M, role, base);
            M = quad_add(M, role, base, ((k >> bit) & 1u) ? Sc : quad_cached_identity(role));
        }
        if (kmax) R = quad_add(R, role, base, quad_to_cached(M, role, base));
    }
    // tree over the 8 quads of a warp, then over warps through shared memory
    for (int d = 16; d >= 4; d >>= 1) {
        fe o = shfl_down_fe(R, d);
        if (lane + d >= 32) o = quad_identity(role);
        R = quad_add(R, role, base, quad_to_cached(o, role, base));
    }
    const uint32_t warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31u) >> 5;
    if (lane < 4) part[warp][role] = R;
    __syncthreads();
    if (warp == 0) {
        const uint32_t q = lane >> 2;                   // 8 quads, each folds warps q, q+8, q+16, q+24
        fe V = quad_identity(role);
        for (uint32_t r = 0; r < 4; r++) {
            uint32_t src = q + 8 * r;
            fe o = src < nwarps ? part[src][role] : quad_identity(role);
            V = quad_add(V, role, base, quad_to_cached(o, role, base));
        }
        for (int d = 16; d >= 4; d >>= 1) {
            fe o = shfl_down_fe(V, d);
            if (lane + d >= 32) o = quad_identity(role);
            V = quad_add(V, role, base, quad_to_cached(o, role, base));
        }
        if (lane < 4) st_fe(reinterpret_cast<fe *>(&windows[blockIdx.x]) + role, V);
    }
}
// window sum = sum of its `parts` partial sums (one quad per window)
__global__ void __launch_bounds__(32) k_msm_window_parts(uint32_t n_win, uint32_t parts, const ge *__restrict__ wparts, ge *__restrict__ windows) {
    const uint32_t win = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const int role = threadIdx.x & 3, base = (threadIdx.x & 31) & ~3;
    const bool valid = win < n_win;
    const fe *src = reinterpret_cast<const fe *>(wparts + (size_t)(valid ? win : 0) * parts);
    fe acc = valid ? ld_fe(src + role) : quad_identity(role);
    for (uint32_t p = 1; p < parts; p++) {
        fe o = valid ? ld_fe(src + 4 * p + role) : quad_identity(role);
        acc = quad_add(acc, role, base, quad_to_cached(o, role, base));
    }
    if (valid) st_fe(reinterpret_cast<fe *>(&windows[win]) + role, acc);
}

// Throughput variant for many (segment, window) pairs of moderate size (the verifier: 4 x 28 windows of 256 buckets): one WARP per
// window, one thread per L = B / 32 consecutive buckets, everything in whole-thread extended-coordinate arithmetic (about half the
// instructions of the quad kernel above).  Lane t: S_t = sum of its buckets, R_t = sum_j (j + 1) * bucket[tL + j] (running sums);
// window sum = sum_t R_t + L * sum_t t * S_t, and sum_t t * S_t = sum_{k >= 1} P_k with P_k = sum_{u >= k} S_u (suffix scan by
// shuffles); every lane then folds T_t = R_t + L * P_t (log2 L doublings) and one shuffle tree adds the 32 T_t.
struct gex { fe X, Y, Z, T; };
static __device__ __forceinline__ void gex_add_cached(gex &p, const fe &ymx, const fe &ypx, const fe &z2, const fe &t2d) {
    const fe A = fe_mul(fe_sub_l(p.Y, p.X), ymx), B = fe_mul(fe_add_l(p.Y, p.X), ypx), C = fe_mul(p.T, t2d), D = fe_mul(p.Z, z2);
    const fe E = fe_sub_l(B, A), H = fe_add_l(B, A), F = fe_sub_l(D, C), G = fe_add_l(D, C);
    p.X = fe_mul(E, F); p.Y = fe_mul(G, H); p.Z = fe_mul(F, G); p.T = fe_mul(E, H);
}
static __device__ __forceinline__ void gex_add(gex &p, const gex &q) {
    gex_add_cached(p, fe_sub_l(q.Y, q.X), fe_add_l(q.Y, q.X), fe_add_l(q.Z, q.Z), fe_mul(q.T, fe_const_2d()));
}
static __device__ __forceinline__ void gex_dbl(gex &p) {
    const fe XX = fe_sq(p.X), YY = fe_sq(p.Y), ZZ = fe_sq(p.Z), S = fe_sq(fe_add_l(p.X, p.Y));
    const fe H = fe_add_l(YY, XX), G = fe_sub_l(YY, XX);
    const fe E = fe_sub_ll(S, H), F = fe_sub_ll(fe_add_l(ZZ, ZZ), G);
    p.X = fe_mul(E, F); p.Y = fe_mul(H, G); p.Z = fe_mul(G, F); p.T = fe_mul(E, H);
}
static __device__ __forceinline__ gex gex_identity() { gex r; r.X = fe_zero(); r.Y = fe_one(); r.Z = fe_one(); r.T = fe_zero(); return r; }
// value of lane (lane + delta), or the identity beyond the warp
static __device__ __forceinline__ gex gex_shfl_down(const gex &v, int delta, int lane) {
    gex r;
    r.X = shfl_down_fe(v.X, delta); r.Y = shfl_down_fe(v.Y, delta); r.Z = shfl_down_fe(v.Z, delta); r.T = shfl_down_fe(v.T, delta);
    const bool in = lane + delta < 32;
    r.X = fe_select(fe_zero(), r.X, in); r.Y = fe_select(fe_one(), r.Y, in); r.Z = fe_select(fe_one(), r.Z, in); r.T = fe_select(fe_zero(), r.T, in);
    return r;
}
// Mid-size bucket sums: LANES = 2, 4 or 8 consecutive lanes of a warp per bucket.  Lane r sums entries r, r + LANES, ... of the bucket
// as a whole thread (the arithmetic of k_msm_bucket_thread), then a shuffle tree adds the LANES partial sums (9 multiplications per
// level).  For MSMs whose bucket count alone cannot fill the machine with whole threads (2^14..2^18 points: 10-50 k buckets of
// 30-250 entries) this keeps the whole-thread instruction count -- about half of the quad kernel's per addition -- at the quad
// kernel's parallelism.
template <int LANES>
__global__ void __launch_bounds__(256) k_msm_bucket_split(uint32_t n_keys, const uint32_t *__restrict__ starts, const uint32_t *__restrict__ counts, const uint32_t *__restrict__ sorted,
                                                         const uint32_t *__restrict__ pidx, const aniels *__restrict__ dyn,
                                                         const aniels *__restrict__ gens, const cached *__restrict__ dync,
                                                         cached *__restrict__ buckets, uint32_t heavy_min) {
    const uint32_t k = (blockIdx.x * 256u + threadIdx.x) / LANES;
    const uint32_t r = threadIdx.x & (LANES - 1);
    const bool valid = k < n_keys;
    const uint32_t lo = valid ? starts[k] : 0u;
    uint32_t cnt = valid ? (counts ? counts[k] : starts[k + 1] - lo) : 0u;
    if (heavy_min && cnt > heavy_min) cnt = 0u;          // left to k_msm_heavy_*
    gex p = gex_identity();
    for (uint32_t j = r; j < cnt; j += LANES) bucket_add_entry(p.X, p.Y, p.Z, p.T, sorted[lo + j], pidx, dyn, gens, dync);
    // groups are LANES-aligned inside the warp: lane r < d of a group receives the sum of lane r + d of the same group; what the
    // other lanes receive is never used
#pragma unroll
    for (int d = LANES / 2; d >= 1; d >>= 1) {
        gex o;
        o.X = shfl_down_fe(p.X, d); o.Y = shfl_down_fe(p.Y, d); o.Z = shfl_down_fe(p.Z, d); o.T = shfl_down_fe(p.T, d);
        gex_add(p, o);
    }
    if (valid && r == 0) {
        cached *out = buckets + k;
        st_fe(&out->ymx, fe_sub_l(p.Y, p.X));
        st_fe(&out->ypx, fe_add_l(p.Y, p.X));
        st_fe(&out->z2, fe_add_l(p.Z, p.Z));
        st_fe(&out->t2d, fe_mul(p.T, fe_const_2d()));
    }
}

// ------------------------------------------------------------------------------------------------ 4b: over-full buckets
// A bucket far above the average would be walked by one thread of k_msm_bucket_thread while the machine waits: the top window of a
// width that does not divide 252 (c = 16 at 2^24 points: 2^11 buckets of 8192 entries next to 2^15 buckets of 512), or degenerate
// scalar sets (the prover's {0, 1, l - 1}: two thirds of all entries in ONE bucket).  Buckets with more than 2 * part entries
// (part = twice the average bucket size) are therefore cut into parts: k_msm_heavy_find lists (bucket, part) items,
// k_msm_bucket_thread sums one part per thread in extra CTAs behind the buckets (the same walk as a bucket of average size),
// k_msm_heavy_finish adds the parts of a bucket with one warp.  This is what lets the window choice ignore the top window.
__global__ void __launch_bounds__(256) k_msm_heavy_find(uint32_t n_keys, const uint32_t *__restrict__ starts, const uint32_t *__restrict__ counts,
                                                       uint32_t heavy_min, uint32_t part_size, uint32_t *__restrict__ heavy_n, uint2 *__restrict__ items,
                                                       uint32_t cap) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_keys) return;
    const uint32_t cnt = counts ? counts[k] : starts[k + 1] - starts[k];
    if (cnt <= heavy_min) return;
    const uint32_t np = (cnt + part_size - 1u) / part_size;
    const uint32_t base = atomicAdd(heavy_n, np);
    for (uint32_t q = 0; q < np; q++) if (base + q < cap) items[base + q] = make_uint2(k, q);
}
// one warp per over-full bucket (the item with part 0 stands for it): sum of its parts -> the bucket, in cached form
__global__ void __launch_bounds__(128) k_msm_heavy_finish(const uint32_t *__restrict__ heavy_n, const uint2 *__restrict__ items, uint32_t cap,
                                                         const uint32_t *__restrict__ starts, const uint32_t *__restrict__ counts, uint32_t part_size,
                                                         const ge *__restrict__ parts, cached *__restrict__ buckets) {
    const uint32_t n_items = min(*heavy_n, cap);
    const int lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t it = gw; it < n_items; it += nw) {
        const uint2 item = items[it];
        if (item.y != 0) continue;                       // warp-uniform
        const uint32_t total = counts ? counts[item.x] : starts[item.x + 1] - starts[item.x];
        const uint32_t np = (total + part_size - 1u) / part_size;
        gex p = gex_identity();
        for (uint32_t q = (uint32_t)lane; q < np && it + q < n_items; q += 32) {
            const ge *src = parts + it + q;
            gex o;
            o.X = ld_fe(&src->X); o.Y = ld_fe(&src->Y); o.Z = ld_fe(&src->Z); o.T = ld_fe(&src->T);
            gex_add(p, o);
        }
        for (int d = 16; d > 0; d >>= 1) { const gex o = gex_shfl_down(p, d, lane); gex_add(p, o); }
        if (lane == 0) {
            cached *out = buckets + item.x;
            st_fe(&out->ymx, fe_sub_l(p.Y, p.X));
            st_fe(&out->ypx, fe_add_l(p.Y, p.X));
            st_fe(&out->z2, fe_add_l(p.Z, p.Z));
            st_fe(&out->t2d, fe_mul(p.T, fe_const_2d()));
        }
    }
}

__global__ void __launch_bounds__(128) k_msm_reduce_warp(uint32_t n_win, uint32_t B, const cached *__restrict__ buckets, ge *__restrict__ windows) {
    const uint32_t win = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (win >= n_win) return;                            // warp-uniform
    const uint32_t L = B >> 5;
    const cached *bk = buckets + (size_t)win * B + (size_t)lane * L;
    gex S = gex_identity(), R = gex_identity();
    for (int j = (int)L - 1; j >= 0; j--) {
        gex_add_cached(S, ld_fe(&bk[j].ymx), ld_fe(&bk[j].ypx), ld_fe(&bk[j].z2), ld_fe(&bk[j].t2d));
        gex_add(R, S);
    }
    // P = inclusive suffix sum of S over the lanes
    gex P = S;
    for (int d = 1; d < 32; d <<= 1) {
        const gex o = gex_shfl_down(P, d, lane);
        gex_add(P, o);
    }
    if (lane == 0) P = gex_identity();                   // sum_t t * S_t = sum_{k >= 1} P_k
    for (uint32_t l = L; l > 1; l >>= 1) gex_dbl(P);      // L is a power of two
    gex_add(R, P);                                        // T_t = R_t + L * P_t
    for (int d = 16; d > 0; d >>= 1) {
        const gex o = gex_shfl_down(R, d, lane);
        gex_add(R, o);
    }
    if (lane == 0) {
        ge *w = windows + win;
        st_fe(&w->X, R.X); st_fe(&w->Y, R.Y); st_fe(&w->Z, R.Z); st_fe(&w->T, R.T);
    }
}

// small bucket counts (B <= 64: the prover's many short MSMs): one quad per (segment, window) walks all B buckets itself, eight
// (segment, window) pairs per warp -- the CTA-per-window kernel above would run one active quad per warp there
__global__ void __launch_bounds__(128) k_msm_reduce_small(uint32_t n_win, uint32_t B, const cached *__restrict__ buckets, ge *__restrict__ windows) {
    const uint32_t idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const int role = threadIdx.x & 3, base = (threadIdx.x & 31) & ~3;
    const bool valid = idx < n_win;
    const cached *bk = buckets + (size_t)(valid ? idx : 0) * B;
    fe S = quad_identity(role), R = quad_identity(role);
    for (int j = (int)B - 1; j >= 0; j--) {
        fe b = valid ? ld_fe(reinterpret_cast<const fe *>(&bk[j]) + role) : quad_cached_identity(role);
        S = quad_add(S, role, base, b);
        R = quad_add(R, role, base, quad_to_cached(S, role, base));
    }
    if (valid) st_fe(reinterpret_cast<fe *>(&windows[idx]) + role, R);
}

// ------------------------------------------------------------------------------------------------ 6: Horner
// one quad per segment
__global__ void __launch_bounds__(32) k_msm_combine(uint32_t n_seg, int c, int W, const ge *__restrict__ windows, ge *__restrict__ result) {
    const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const int role = threadIdx.x & 3, base = (threadIdx.x & 31) & ~3;
    const bool valid = s < n_seg;
    const fe *win = reinterpret_cast<const fe *>(windows + (size_t)(valid ? s : 0) * W);
    fe acc = valid ? ld_fe(win + 4 * (W - 1) + role) : quad_identity(role);
    for (int w = W - 2; w >= 0; w--) {
        for (int k = 0; k < c; k++) acc = quad_dbl(acc, role, base);
        fe o = valid ? ld_fe(win + 4 * w + role) : quad_identity(role);
        acc = quad_add(acc, role, base, quad_to_cached(o, role, base));
    }
    if (valid) st_fe(reinterpret_cast<fe *>(&result[s]) + role, acc);
}

// ------------------------------------------------------------------------------------------------ driver
void launch_msm(cudaStream_t s, const MsmShape &sh, const uint32_t *scalars, const uint32_t *seg_offsets, const uint32_t *pidx,
                const aniels *dyn, const aniels *gens, void *scratch, ge *result, uint64_t *launches, cudaEvent_t *marks,
                const cached *dync) {
    MsmScratch sc = msm_carve(sh, scratch);
    size_t n_keys = (size_t)sh.n_seg * sh.W * sh.B;
    uint32_t n_tiles = (uint32_t)((n_keys + SCAN_TILE - 1) / SCAN_TILE);
    uint32_t eg = (sh.n_entries + 255u) / 256u;
    // shared-memory sort for many small segments (the verifier), scan-based counting sort otherwise
    const uint32_t *counts = nullptr;
    const bool fused_sort = !knobs().scan_sort && sh.max_seg_entries != 0 && sh.max_seg_entries <= 16384u && sh.B <= 1024u && sh.n_seg * (uint32_t)sh.W >= 32u;
    if (fused_sort && sh.c == 9 && sh.n_seg >= 8 && !knobs().window_sort) {
        // groups of G windows per CTA, as large as leaves ~2 CTAs per SM (few segments: one window per CTA, the kernel below)
        int groups = (int)((2 * 148 + sh.n_seg - 1) / sh.n_seg);
        if (groups > sh.W) groups = sh.W;
        const int G = (sh.W + groups - 1) / groups;
        groups = (sh.W + G - 1) / G;
        k_msm_sort_segw<9><<<dim3(sh.n_seg, (unsigned)groups), 512, (size_t)G * sh.B * 4, s>>>(sh.n_entries, sh.n_seg, G, scalars, seg_offsets, sc.starts, sc.cursor,
                                                                                              sc.sorted);
        counts = sc.cursor;
        if (launches) *launches += 1;
    } else if (fused_sort) {
        k_msm_sort_seg<1024><<<sh.n_seg * (uint32_t)sh.W, 256, 0, s>>>(sh.n_entries, sh.n_seg, sh.c, sh.W, sh.B, scalars, seg_offsets, sc.starts, sc.cursor, sc.sorted);
        counts = sc.cursor;
        if (launches) *launches += 1;
    } else {
        cudaMemsetAsync(sc.starts, 0, (n_keys + 1) * 4, s);
        if (sh.n_entries) {
            k_msm_digits<false><<<eg, 256, 0, s>>>(sh.n_entries, sh.n_seg, sh.c, sh.W, sh.B, scalars, seg_offsets, sc.starts, nullptr);
        }
        k_scan_tile_sums<<<n_tiles, 256, 0, s>>>(sc.starts, n_keys, sc.tile_sums);
        k_scan_tiles<<<1, 32, 0, s>>>(sc.tile_sums, n_tiles);
        k_scan_apply<<<n_tiles, 256, 0, s>>>(sc.starts, sc.cursor, n_keys, sc.tile_sums);
        if (sh.n_entries) {
            k_msm_digits<true><<<eg, 256, 0, s>>>(sh.n_entries, sh.n_seg, sh.c, sh.W, sh.B, scalars, seg_offsets, sc.cursor, sc.sorted);
        }
        if (launches) *launches += 3 + (sh.n_entries ? 2 : 0);
    }
    if (marks) cudaEventRecord(marks[0], s);
    // whole threads per bucket once there are enough buckets to occupy the SMs that way: always from ~4 warps per SM
    // sub-partition on; from 1 warp on only while the buckets are short (a thread walks its bucket alone: 2^16 points at c = 11
    // are 64 additions per bucket, measured 82 M points/s with quads against 75 M with threads; the verifier's 4226-entry
    // segments are 16 per bucket and 8 % faster with threads)
    const int force_bucket = knobs().bucket;      // 1 = quads, 2 = threads, 3 = split threads (tests)
    const size_t full = 148u * 4u * 32u;
    const size_t adds = (size_t)sh.n_entries * sh.W;
    const bool thread_buckets = force_bucket ? force_bucket == 2
                                             : n_keys >= 4 * full || (n_keys >= full && adds <= 32 * n_keys);
    // split threads (when whole threads per bucket cannot fill the machine): LANES = the power of two nearest to 170 k threads /
    // bucket count, at most 8, halved while a lane would walk fewer than 4 entries.  Measured (scripts/msm_bucket_probe.py, bucket
    // phase, quads -> split): 2^14 c = 9: 146 -> 86 us; 2^15: 266 -> 135; 2^16 c = 11: 391 -> 231; 2^17: 764 -> 391; 2^18 c = 12:
    // 821 -> 533 (whole threads: 767)
    int split = 0;
    if (force_bucket == 3) split = knobs().split;
    else if (!force_bucket && !thread_buckets && n_keys >= 1024) {
        const double want = 170000.0 / (double)n_keys;
        split = want >= 5.66 ? 8 : want >= 2.83 ? 4 : want >= 1.42 ? 2 : 1;
        while (split > 1 && adds < 4 * (size_t)split * n_keys) split >>= 1;
    }
    // over-full buckets (section 4b): with the thread-per-bucket kernel only (large sums)
    const uint32_t part_size = heavy_part_size(sh);
    const uint32_t heavy_min = (thread_buckets && !fused_sort && !knobs().no_heavy && sh.n_entries >= (1u << 14)) ? 2u * part_size : 0u;
    if (heavy_min) {
        cudaMemsetAsync(sc.heavy_n, 0, 4, s);
        k_msm_heavy_find<<<(uint32_t)((n_keys + 255) / 256), 256, 0, s>>>((uint32_t)n_keys, sc.starts, counts, heavy_min, part_size, sc.heavy_n, sc.heavy_items,
                                                                          sc.heavy_cap);
    }
    if (split == 2 || split == 4 || split == 8) {
        const uint32_t grid = (uint32_t)((n_keys * (size_t)split + 255) / 256);
        if (split == 2) k_msm_bucket_split<2><<<grid, 256, 0, s>>>((uint32_t)n_keys, sc.starts, counts, sc.sorted, pidx, dyn, gens, dync, sc.buckets, 0u);
        else if (split == 4) k_msm_bucket_split<4><<<grid, 256, 0, s>>>((uint32_t)n_keys, sc.starts, counts, sc.sorted, pidx, dyn, gens, dync, sc.buckets, 0u);
        else k_msm_bucket_split<8><<<grid, 256, 0, s>>>((uint32_t)n_keys, sc.starts, counts, sc.sorted, pidx, dyn, gens, dync, sc.buckets, 0u);
    } else if (thread_buckets)
    {
        // 96 registers per thread leave room for two 256-thread CTAs per SM; capped at 80 a third one fits: +3 % on passes that fill
        // the machine (measured), a little slower on small ones
        const uint32_t grid = (uint32_t)((n_keys + 255) / 256) + (heavy_min ? (sc.heavy_cap + 255u) / 256u : 0u);
        // rank all buckets by size first (see k_msm_size_*) unless they are few or the switch says no
        const uint32_t *perm = nullptr;
        if (n_keys >= 4096 && !knobs().local_rank) {
            const uint32_t g4 = (uint32_t)((n_keys + 1023) / 1024);
            cudaMemsetAsync(sc.size_bins, 0, 1024, s);
            k_msm_size_hist<<<g4, 256, 0, s>>>((uint32_t)n_keys, sc.starts, counts, heavy_min, sc.size_bins);
            k_msm_size_scan<<<1, 256, 0, s>>>(sc.size_bins);
            k_msm_size_scatter<<<g4, 256, 0, s>>>((uint32_t)n_keys, sc.starts, counts, heavy_min, sc.size_bins, sc.perm);
            perm = sc.perm;
            if (launches) *launches += 3;
        }
        static const size_t occ3_min = [] { const char *e = getenv("BPP_MSM_OCC3_MIN"); return e ? (size_t)atol(e) : (size_t)200000; }();
        if (n_keys >= occ3_min)
            k_msm_bucket_thread<3><<<grid, 256, 0, s>>>((uint32_t)n_keys, sc.starts, counts, sc.sorted, pidx, dyn, gens, dync, sc.buckets, heavy_min, part_size,
                                                        sc.heavy_n, sc.heavy_items, sc.heavy_cap, sc.heavy_parts, perm);
        else
            k_msm_bucket_thread<1><<<grid, 256, 0, s>>>((uint32_t)n_keys, sc.starts, counts, sc.sorted, pidx, dyn, gens, dync, sc.buckets, heavy_min, part_size,
                                                        sc.heavy_n, sc.heavy_items, sc.heavy_cap, sc.heavy_parts, perm);
    }
    else
        k_msm_bucket<<<(uint32_t)((n_keys + BUCKET_CTA / 4 - 1) / (BUCKET_CTA / 4)), BUCKET_CTA, 0, s>>>((uint32_t)n_keys, sc.starts, counts, sc.sorted, pidx, dyn, gens, dync, sc.buckets, 0u);
    if (heavy_min) {
        k_msm_heavy_finish<<<148, 128, 0, s>>>(sc.heavy_n, sc.heavy_items, sc.heavy_cap, sc.starts, counts, part_size, sc.heavy_parts, sc.buckets);
        if (launches) *launches += 2;
    }
    if (marks) cudaEventRecord(marks[1], s);
    const int force_reduce = knobs().reduce;      // 1 = CTA of quads, 2 = warp of threads (tests)
    if (sh.B <= 64 && sh.n_seg * sh.W >= 64) {
        uint32_t n_win = sh.n_seg * (uint32_t)sh.W;
        k_msm_reduce_small<<<(n_win + 31) / 32, 128, 0, s>>>(n_win, sh.B, sc.buckets, sc.windows);
    } else if (force_reduce ? (force_reduce == 2 && sh.B >= 32) : (sh.B >= 32 && sh.B <= 512 && sh.n_seg * sh.W >= 64)) {
        // enough windows to give every SM a warp: one warp of whole threads per window is half the instructions of the quad kernel
        // (verifier: 10.9 M -> 5 M per pass) for a longer chain (88 us against 63 us alone).  Measured with 32 lanes in flight:
        // 6.2 M against 5.9 M proofs/s; one batch alone: 958 k against 977 k proofs/s
        uint32_t n_win = sh.n_seg * (uint32_t)sh.W;
        k_msm_reduce_warp<<<(n_win + 3) / 4, 128, 0, s>>>(n_win, sh.B, sc.buckets, sc.windows);
    } else {
        // quads of 8 buckets; up to REDUCE_PARTS_MAX CTAs of <= 128 quads per window while the windows alone leave SMs idle
        // (BPP_MSM_REDUCE_PARTS forces the split: tests / experiments)
        const uint32_t n_win = sh.n_seg * (uint32_t)sh.W;
        uint32_t quads = sh.B >= 8 ? sh.B / 8 : 1, parts = 1;
        while (parts < REDUCE_PARTS_MAX && quads / parts > 128 && n_win * parts * 2 <= 2 * 148u /* CTAs after the split */) parts <<= 1;
        {
            const uint32_t v = (uint32_t)knobs().reduce_parts;
            if ((v == 1 || v == 2 || v == 4 || v == 8) && quads % v == 0) parts = v;
        }
        uint32_t nq = quads / parts;                   // quads per CTA
        if (nq > 256) nq = 256;
        uint32_t L = sh.B / (parts * nq);
        k_msm_reduce<<<n_win * parts, (4 * nq + 31u) / 32u * 32u, 0, s>>>(sh.B, L, nq, parts, sc.buckets, parts > 1 ? sc.wparts : sc.windows);
        if (parts > 1) k_msm_window_parts<<<(n_win + 7) / 8, 32, 0, s>>>(n_win, parts, sc.wparts, sc.windows);
    }
    if (marks) cudaEventRecord(marks[2], s);
    k_msm_combine<<<(sh.n_seg + 7) / 8, 32, 0, s>>>(sh.n_seg, sh.c, sh.W, sc.windows, result);
    if (marks) cudaEventRecord(marks[3], s);
    if (launches) *launches += 3;      // bucket sums, window reduction, Horner (the sort phase counted above)
}

} // namespace bpp
