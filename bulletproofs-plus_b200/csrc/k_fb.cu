// K-FB: fixed-base multiscalar multiplication over a generator set (Gi | Hi | G_k | H) with precomputed window tables.
//
// The prover's vector commitments (/root/reference/src/range_proof.rs:334-345 A, :482-495 L / R, :574-584 A1 / B) are all
// multiscalar multiplications over the SAME static generators, P proofs at a time, and the reference's generator folding
// (:511-521, 2(N-1) two-point multiplications per proof, ~65 % of its proving time) only exists to keep those MSMs short.
// With every round's L / R re-expressed over the original generators (k_prove.cu keeps the folding as two scalar vectors)
// nothing is folded, nothing is doubled and nothing is sorted:
//   table[g][w][d-1] = d * 2^(c*w) * P_g        g < n_gens, w < W = ceil(252 / c), d = 1 .. B = 2^(c-1)     (affine Niels, 96 B)
//   sum_j s_j * P_g(j) = sum_j sum_w sign * table[g(j)][w][|digit_w(s_j)| - 1]                              (signed c-bit digits)
// i.e. W mixed additions per term.  For the 64-bit, aggregation-1 set (131 generators) c = 9 is 28 windows x 256 entries =
// 90 MB, L2-resident on B200; k_fb_msm runs one warp (or one CTA) per (proof, commitment), every lane walks its share of the
// (term, window) pairs into a private accumulator, and a shuffle tree adds the 32 accumulators up.
//
// Also exported on its own (bpp_gens_fixed_base_msm) as the static half of Precomputation::vartime_mixed_multiscalar_mul.
#define BPP_INLINE_MUL
#include <algorithm>
#include <cstdlib>
#include "kernels.cuh"
#include "quad.cuh"

namespace bpp {

FbShape fb_shape(uint32_t n_gens, int forced_c, size_t max_bytes) {
    FbShape sh;
    sh.n_gens = n_gens;
    int c = forced_c > 0 ? forced_c : 9;
    if (c < 4) c = 4;
    if (c > 13) c = 13;
    for (;;) {                              // largest window that fits the memory budget
        sh.c = c; sh.W = (252 + c - 1) / c; sh.B = 1u << (c - 1);
        if (forced_c > 0 || c == 4 || fb_table_bytes(sh) <= max_bytes) break;
        c--;
    }
    return sh;
}
size_t fb_table_bytes(const FbShape &sh) { return sizeof(aniels) * (size_t)sh.n_gens * sh.W * sh.B; }

// ------------------------------------------------------------------------------------------------ table construction
// bases[g * W + w] = 2^(c*w) * P_g   (one thread per generator: c*W sequential doublings)
__global__ void __launch_bounds__(64) k_fb_bases(uint32_t n_gens, int c, int W, const aniels *__restrict__ gens, ge *__restrict__ bases) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_gens) return;
    ge P = aniels_to_ge(gens[g]);
    for (int w = 0; w < W; w++) {
        bases[(size_t)g * W + w] = P;
        if (w + 1 < W)
            for (int k = 0; k < c; k++) P = ge_dbl(P);
    }
}
// one thread per (generator, window) of the slice [item0, item0 + n_items): multiples 1..B of its base by repeated addition,
// then ONE field inversion for all B of them (Montgomery's trick through a prefix-product array) and the affine Niels form.
// pts / prefix: scratch, [d][item] so that neighbouring threads touch neighbouring addresses.
__global__ void __launch_bounds__(64) k_fb_fill(uint32_t item0, uint32_t n_items, uint32_t B, const ge *__restrict__ bases, ge *__restrict__ pts,
                                               fe *__restrict__ prefix, aniels *__restrict__ tab) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_items) return;
    const ge P = bases[item0 + t];
    ge acc = P;
    fe pre = fe_one();
    for (uint32_t d = 0; d < B; d++) {
        pts[(size_t)d * n_items + t] = acc;
        pre = fe_mul(pre, acc.Z);
        prefix[(size_t)d * n_items + t] = pre;
        if (d + 1 < B) acc = ge_add(acc, P);
    }
    fe inv = fe_invert(pre);                     // Z never vanishes: the addition law is complete on this curve
    for (uint32_t d = B; d-- > 0;) {
        const ge q = pts[(size_t)d * n_items + t];
        const fe zi = d ? fe_mul(inv, prefix[(size_t)(d - 1) * n_items + t]) : inv;
        inv = fe_mul(inv, q.Z);
        const fe x = fe_mul(q.X, zi), y = fe_mul(q.Y, zi);
        tab[(size_t)(item0 + t) * B + d] = ge_to_aniels_affine(x, y, fe_mul(x, y));
    }
}

int fb_build(cudaStream_t s, const FbShape &sh, const aniels *gens, aniels *tab, uint64_t *launches) {
    const size_t n_items = (size_t)sh.n_gens * sh.W;
    ge *bases = nullptr;
    if (cudaMalloc(&bases, sizeof(ge) * n_items) != cudaSuccess) return 1;
    k_fb_bases<<<(sh.n_gens + 63) / 64, 64, 0, s>>>(sh.n_gens, sh.c, sh.W, gens, bases);
    // scratch for at most ~256 MB worth of (point, prefix) pairs per slice
    const size_t per_item = (sizeof(ge) + sizeof(fe)) * sh.B;
    size_t slice = std::max<size_t>(64, std::min<size_t>(n_items, ((size_t)256 << 20) / per_item));
    ge *pts = nullptr;
    fe *prefix = nullptr;
    if (cudaMalloc(&pts, sizeof(ge) * slice * sh.B) != cudaSuccess || cudaMalloc(&prefix, sizeof(fe) * slice * sh.B) != cudaSuccess) {
        cudaFree(bases); if (pts) cudaFree(pts);
        return 1;
    }
    for (size_t i0 = 0; i0 < n_items; i0 += slice) {
        const uint32_t cnt = (uint32_t)std::min(slice, n_items - i0);
        k_fb_fill<<<(cnt + 63) / 64, 64, 0, s>>>((uint32_t)i0, cnt, sh.B, bases, pts, prefix, tab);
        if (launches) (*launches)++;
    }
    if (launches) (*launches)++;
    cudaError_t e = cudaStreamSynchronize(s);
    cudaFree(bases); cudaFree(pts); cudaFree(prefix);
    return e == cudaSuccess ? 0 : 1;
}

// ------------------------------------------------------------------------------------------------ the sum
static __device__ __forceinline__ aniels fb_ld(const aniels *p) {
    aniels q;
    ld8(q.ypx.v, p->ypx.v); ld8(q.ymx.v, p->ymx.v); ld8(q.t2d.v, p->t2d.v);
    return q;
}
// acc += (neg ? -Q : Q), Q affine Niels; lazy additions between the multiplications (arith.cuh)
static __device__ __forceinline__ void fb_madd(fe &X, fe &Y, fe &Z, fe &T, const aniels &q, bool neg) {
    const fe A = fe_mul(fe_sub_l(Y, X), fe_select(q.ymx, q.ypx, neg));
    const fe B = fe_mul(fe_add_l(Y, X), fe_select(q.ypx, q.ymx, neg));
    const fe C = fe_mul(T, q.t2d);
    const fe D = fe_add(Z, Z);
    const fe E = fe_sub_l(B, A), H = fe_add_l(B, A);
    const fe F0 = fe_sub_l(D, C), G0 = fe_add_l(D, C);
    const fe F = fe_select(F0, G0, neg), G = fe_select(G0, F0, neg);
    X = fe_mul(E, F); Y = fe_mul(G, H); Z = fe_mul(F, G); T = fe_mul(E, H);
}
static __device__ __forceinline__ fe fb_shfl_down(const fe &v, int delta) {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_down_sync(0xffffffffu, v.v[i], delta);
    return r;
}

#define FB_MAXW 64          // c >= 4
// One CTA of NW warps per segment; warp w takes the 32-entry chunks w, w + NW, ...  Scalars: canonical (< l), seg_len per segment,
// segment-major.  gidx: generator index of every entry, `kinds` alternatives of seg_len each; segment s uses alternative s % kinds.
template <int NW, int MIN_CTAS> __global__ void __launch_bounds__(32 * NW, MIN_CTAS) k_fb_msm(uint32_t seg_len, uint32_t kinds, int c, int W, uint32_t B,
                                                                     const uint32_t *__restrict__ scalars, const uint32_t *__restrict__ gidx,
                                                                     const aniels *__restrict__ tab, ge *__restrict__ out) {
    __shared__ int16_t dig[NW][32][FB_MAXW + 2];
    __shared__ ge part[NW];
    const uint32_t seg = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t *sc_seg = scalars + 8 * (size_t)seg * seg_len;
    const uint32_t *gi = gidx + (size_t)(seg % kinds) * seg_len;
    fe X = fe_zero(), Y = fe_one(), Z = fe_one(), T = fe_zero();
    for (uint32_t e0 = warp * 32; e0 < seg_len; e0 += 32 * NW) {
        const uint32_t e = e0 + lane;
        if (e < seg_len) {           // signed digits of min(s, l - s), the sign of the choice folded into every digit
            uint32_t s[8], t[8];
            ld8(s, sc_seg + 8 * (size_t)e);
            int64_t bw = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) { bw += (int64_t)sc_l(k) - (int64_t)s[k]; t[k] = (uint32_t)bw; bw >>= 32; }
            bool gt = false, decided = false;
#pragma unroll
            for (int k = 7; k >= 0; k--)
                if (!decided && s[k] != t[k]) { gt = s[k] > t[k]; decided = true; }
            if (gt) {
#pragma unroll
                for (int k = 0; k < 8; k++) s[k] = t[k];
            }
            uint32_t carry = 0;
            for (int w = 0; w < W; w++) {
                const int off = w * c, wi = off >> 5, sh = off & 31;
                const uint64_t lo = s[wi], hi = wi + 1 <= 7 ? s[wi + 1] : 0;
                uint32_t dgt = ((uint32_t)((lo | (hi << 32)) >> sh) & ((1u << c) - 1u)) + carry;
                bool neg = gt;
                if (dgt > B) { dgt = 2u * B - dgt; neg = !neg; carry = 1u; } else carry = 0u;
                dig[warp][lane][w] = (int16_t)(neg ? -(int)dgt : (int)dgt);
            }
        }
        __syncwarp();
        const uint32_t n_e = seg_len - e0 < 32u ? seg_len - e0 : 32u, items = n_e * (uint32_t)W;
        // window-major item order: in a full chunk lane = entry and the iteration = window, so sparse scalar sets (the bit vectors of A:
        // 0 / +-1, one non-zero digit per term) cost one addition time per chunk instead of one per (entry, window) pair that a lane
        // happens to own; a short tail chunk still spreads its n_e * W items over all 32 lanes
        for (uint32_t it = lane; it < items; it += 32) {
            const uint32_t w = it / n_e, el = it - w * n_e;
            const int dg = dig[warp][el][w];
            if (dg != 0) {
                const uint32_t mag = (uint32_t)(dg < 0 ? -dg : dg);
                const aniels q = fb_ld(tab + ((size_t)gi[e0 + el] * (uint32_t)W + w) * B + (mag - 1u));
                fb_madd(X, Y, Z, T, q, dg < 0);
            }
        }
        __syncwarp();
    }
    // 32 accumulators -> lane 0
    ge acc;
    acc.X = X; acc.Y = Y; acc.Z = Z; acc.T = T;
    for (int delta = 16; delta > 0; delta >>= 1) {
        ge o;
        o.X = fb_shfl_down(acc.X, delta); o.Y = fb_shfl_down(acc.Y, delta); o.Z = fb_shfl_down(acc.Z, delta); o.T = fb_shfl_down(acc.T, delta);
        acc = ge_add(acc, o);
    }
    if (NW == 1) {
        if (lane == 0) out[seg] = acc;
    } else {
        if (lane == 0) part[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            ge r = part[0];
            for (int k = 1; k < NW; k++) r = ge_add(r, part[k]);
            out[seg] = r;
        }
    }
}

void launch_fb_msm(cudaStream_t s, const FbShape &sh, uint32_t n_seg, uint32_t seg_len, uint32_t kinds, const uint32_t *scalars, const uint32_t *gidx,
                   const aniels *tab, ge *out, uint64_t *launches) {
    if (n_seg == 0) return;
    // (register caps of 96 / 80 / 72 per thread -- 20 / 24 / 28 one-warp CTAs per SM -- were measured: 346 / 330 / 325 k proofs/s against
    // 352 k with the 116 registers ptxas picks; the spills cost more than the extra warps hide)
    if (seg_len <= 256) k_fb_msm<1, 1><<<n_seg, 32, 0, s>>>(seg_len, kinds, sh.c, sh.W, sh.B, scalars, gidx, tab, out);
    else if (seg_len <= 1024) k_fb_msm<4, 1><<<n_seg, 128, 0, s>>>(seg_len, kinds, sh.c, sh.W, sh.B, scalars, gidx, tab, out);
    else k_fb_msm<8, 1><<<n_seg, 256, 0, s>>>(seg_len, kinds, sh.c, sh.W, sh.B, scalars, gidx, tab, out);
    if (launches) (*launches)++;
}

} // namespace bpp
