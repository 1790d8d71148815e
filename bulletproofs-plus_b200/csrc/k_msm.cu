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
#include "kernels.cuh"

namespace bpp {

// ------------------------------------------------------------------------------------------------ helpers
static __device__ __forceinline__ void ld8(uint32_t w[8], const uint32_t *p) {
    uint4 a = reinterpret_cast<const uint4 *>(p)[0], b = reinterpret_cast<const uint4 *>(p)[1];
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}
static __device__ __forceinline__ void st8(uint32_t *p, const uint32_t w[8]) {
    reinterpret_cast<uint4 *>(p)[0] = make_uint4(w[0], w[1], w[2], w[3]);
    reinterpret_cast<uint4 *>(p)[1] = make_uint4(w[4], w[5], w[6], w[7]);
}
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
// by-value so that operands and result travel in registers through the call (see arith.cuh fe_mul)
static __device__ __noinline__ ge ge_add_nl(ge a, ge b) { return ge_add(a, b); }
static __device__ __noinline__ ge ge_dbl_nl(ge a) { return ge_dbl(a); }
static __device__ __noinline__ ge ge_madd_nl(ge a, aniels b) { return ge_madd(a, b); }
static __device__ __noinline__ ge ge_msub_nl(ge a, aniels b) { return ge_msub(a, b); }

static __device__ __forceinline__ ge shfl_down_ge(const ge &p, int delta) {
    ge r;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r.X.v[i] = __shfl_down_sync(0xffffffffu, p.X.v[i], delta);
        r.Y.v[i] = __shfl_down_sync(0xffffffffu, p.Y.v[i], delta);
        r.Z.v[i] = __shfl_down_sync(0xffffffffu, p.Z.v[i], delta);
        r.T.v[i] = __shfl_down_sync(0xffffffffu, p.T.v[i], delta);
    }
    return r;
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

// ------------------------------------------------------------------------------------------------ shape
MsmShape msm_shape(uint32_t n_entries, uint32_t n_seg, int forced_c) {
    MsmShape sh;
    sh.n_entries = n_entries;
    sh.n_seg = n_seg ? n_seg : 1;
    int c = forced_c;
    if (c <= 0) {
        // minimise adds = n*W + 2*B*W per segment (bucket accumulation + running-sum reduction)
        double per = (double)n_entries / (double)sh.n_seg;
        double best = 1e300;
        for (int cc = 2; cc <= 16; cc++) {
            int W = (252 + cc - 1) / cc;
            double cost = per * W + 2.5 * (double)(1u << (cc - 1)) * W;
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

struct MsmScratch {
    uint32_t *starts;   // n_keys + 1
    uint32_t *cursor;   // n_keys
    uint32_t *tile_sums;
    uint32_t *sorted;   // n_entries * W
    ge *buckets;        // n_keys
    ge *windows;        // n_seg * W
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
    s.buckets = (ge *)(p + off); off = align_up(off + n_keys * sizeof(ge), 256);
    s.windows = (ge *)(p + off); off = align_up(off + (size_t)sh.n_seg * sh.W * sizeof(ge), 256);
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
__global__ void __launch_bounds__(128) k_msm_bucket(uint32_t n_keys, const uint32_t *__restrict__ starts, const uint32_t *__restrict__ sorted,
                                                   const uint32_t *__restrict__ pidx, const aniels *__restrict__ dyn,
                                                   const aniels *__restrict__ gens, ge *__restrict__ buckets) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_keys) return;
    uint32_t lo = starts[k], hi = starts[k + 1];
    ge acc = ge_identity();
    for (uint32_t j = lo; j < hi; j++) {
        uint32_t e = sorted[j];
        uint32_t idx = e & 0x7fffffffu;
        uint32_t pi = pidx ? pidx[idx] : idx;
        const aniels *src = (pi & 0x80000000u) ? (gens + (pi & 0x7fffffffu)) : (dyn + pi);
        aniels q = ld_aniels(src);
        acc = (e >> 31) ? ge_msub_nl(acc, q) : ge_madd_nl(acc, q);
    }
    st_ge(&buckets[k], acc);
}

// ------------------------------------------------------------------------------------------------ 5: window sums
// block = T threads for one (seg, window); thread t owns buckets [t*L, (t+1)*L): S = sum, R = sum (j+1)*bucket[tL+j],
// contributes R + (t*L)*S; the block sums the contributions.
__global__ void __launch_bounds__(256) k_msm_reduce(uint32_t B, uint32_t L, uint32_t nthreads, const ge *__restrict__ buckets,
                                                    ge *__restrict__ windows) {
    __shared__ ge warp_part[32];
    uint32_t t = threadIdx.x;      // blockDim.x is a multiple of 32; threads >= nthreads only take part in the shuffles
    const ge *bk = buckets + (size_t)blockIdx.x * B + (size_t)t * L;
    ge S = ge_identity(), R = ge_identity();
    if (t < nthreads) {
        for (int j = (int)L - 1; j >= 0; j--) {
            ge b = ld_ge(&bk[j]);
            S = ge_add_nl(S, b);
            R = ge_add_nl(R, S);
        }
    }
    uint32_t k = t * L;
    if (k != 0 && t < nthreads) {
        // R += k * S  (left-to-right double-and-add; k < B <= 2^15)
        ge M = S;
        int top = 31 - __clz(k);
        for (int bit = top - 1; bit >= 0; bit--) {
            M = ge_dbl_nl(M);
            if ((k >> bit) & 1u) M = ge_add_nl(M, S);
        }
        R = ge_add_nl(R, M);
    }
    // warp tree
    for (int d = 16; d > 0; d >>= 1) {
        ge o = shfl_down_ge(R, d);
        bool valid = ((t & 31u) + d < 32u) && (t + d < nthreads);
        if (valid) R = ge_add_nl(R, o);
    }
    if ((t & 31u) == 0) warp_part[t >> 5] = R;
    __syncthreads();
    if (t < 32) {
        uint32_t nwarps = (nthreads + 31u) >> 5;
        ge V = (t < nwarps) ? warp_part[t] : ge_identity();
        for (int d = 16; d > 0; d >>= 1) {
            ge o = shfl_down_ge(V, d);
            if (t + d < nwarps && (int)t + d < 32) V = ge_add_nl(V, o);
        }
        if (t == 0) st_ge(&windows[blockIdx.x], V);
    }
}

// ------------------------------------------------------------------------------------------------ 6: Horner
__global__ void k_msm_combine(uint32_t n_seg, int c, int W, const ge *__restrict__ windows, ge *__restrict__ result) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    const ge *win = windows + (size_t)s * W;
    ge acc = ld_ge(&win[W - 1]);
    for (int w = W - 2; w >= 0; w--) {
        for (int k = 0; k < c; k++) acc = ge_dbl_nl(acc);
        acc = ge_add_nl(acc, ld_ge(&win[w]));
    }
    st_ge(&result[s], acc);
}

// ------------------------------------------------------------------------------------------------ driver
void launch_msm(cudaStream_t s, const MsmShape &sh, const uint32_t *scalars, const uint32_t *seg_offsets, const uint32_t *pidx,
                const aniels *dyn, const aniels *gens, void *scratch, ge *result, uint64_t *launches, cudaEvent_t *marks) {
    MsmScratch sc = msm_carve(sh, scratch);
    size_t n_keys = (size_t)sh.n_seg * sh.W * sh.B;
    uint32_t n_tiles = (uint32_t)((n_keys + SCAN_TILE - 1) / SCAN_TILE);
    uint32_t eg = (sh.n_entries + 255u) / 256u;
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
    if (marks) cudaEventRecord(marks[0], s);
    k_msm_bucket<<<(uint32_t)((n_keys + 127) / 128), 128, 0, s>>>((uint32_t)n_keys, sc.starts, sc.sorted, pidx, dyn, gens, sc.buckets);
    if (marks) cudaEventRecord(marks[1], s);
    uint32_t T = sh.B >= 8 ? sh.B / 8 : 1;
    if (T > 256) T = 256;
    uint32_t L = sh.B / T;
    k_msm_reduce<<<sh.n_seg * sh.W, (T + 31u) / 32u * 32u, 0, s>>>(sh.B, L, T, sc.buckets, sc.windows);
    if (marks) cudaEventRecord(marks[2], s);
    k_msm_combine<<<(sh.n_seg + 31) / 32, 32, 0, s>>>(sh.n_seg, sh.c, sh.W, sc.windows, result);
    if (marks) cudaEventRecord(marks[3], s);
    if (launches) *launches += 6 + (sh.n_entries ? 2 : 0);
}

} // namespace bpp
