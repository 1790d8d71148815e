// Batched Ristretto255 point kernels: decode (K-DECOMPRESS), encode (K-COMPRESS), one-way map (K-MAP).
// One point per thread; the ~254-squaring inverse-square-root chain dominates (~25.7 k IMAD).
// Replaces CompressedRistretto::decompress / RistrettoPoint::compress / from_uniform_bytes as called from
// /root/reference/src/range_proof.rs:859-866,1067-1109 and :289,:348,:499-504,:587,:598-605,
// /root/reference/src/range_statement.rs:62-65, /root/reference/src/ristretto.rs:48-52.
#include <stdlib.h>
#include "kernels.cuh"
#include "rawld.cuh"

namespace bpp {

static __device__ __forceinline__ void load8(uint32_t w[8], const uint32_t *p) {
    uint4 a = reinterpret_cast<const uint4 *>(p)[0], b = reinterpret_cast<const uint4 *>(p)[1];
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}
static __device__ __forceinline__ void store8(uint32_t *p, const uint32_t w[8]) {
    reinterpret_cast<uint4 *>(p)[0] = make_uint4(w[0], w[1], w[2], w[3]);
    reinterpret_cast<uint4 *>(p)[1] = make_uint4(w[4], w[5], w[6], w[7]);
}
static __device__ __forceinline__ void store_fe(fe *dst, const fe &a) { store8(dst->v, a.v); }

__global__ void __launch_bounds__(128) k_decompress(size_t n, const uint32_t *__restrict__ in, aniels *__restrict__ out_tab,
                                                   uint8_t *__restrict__ ok, uint32_t *__restrict__ out_enc,
                                                   uint32_t *__restrict__ bad_count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t w[8];
    load8(w, in + 8 * i);
    fe x, y, t;
    bool good = ristretto_decode(x, y, t, w);
    if (!good) { x = fe_zero(); y = fe_one(); t = fe_zero(); if (bad_count) atomicAdd(bad_count, 1u); }
    if (ok) ok[i] = good ? 1 : 0;
    if (out_tab) {
        aniels q = ge_to_aniels_affine(x, y, t);
        store_fe(&out_tab[i].ypx, q.ypx); store_fe(&out_tab[i].ymx, q.ymx); store_fe(&out_tab[i].t2d, q.t2d);
    }
    if (out_enc) {
        ge p; p.X = x; p.Y = y; p.Z = fe_one(); p.T = t;
        fe s = ristretto_encode(p);
        store8(out_enc + 8 * i, s.v);
    }
}

// K-DECOMPRESS over the points of a verification pass, read from the uploaded proof bytes / commitments (no host-side gather):
// thread i finds its proof by binary search over the per-proof point offsets, then its slot's 32 bytes inside the serialised proof
template <int MIN_CTAS>
__global__ void __launch_bounds__(128, MIN_CTAS) k_decompress_proofs(uint32_t n_pts, uint32_t n_proofs, uint32_t ext, const VProof *__restrict__ proofs,
                                                          const uint32_t *__restrict__ pt_offsets, const uint8_t *__restrict__ blob,
                                                          const uint8_t *__restrict__ commitments32, aniels *__restrict__ out_tab,
                                                          uint8_t *__restrict__ ok) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pts) return;
    uint32_t lo = 0, hi = n_proofs;                     // invariant: pt_offsets[lo] <= i < pt_offsets[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (pt_offsets[mid] <= i) lo = mid; else hi = mid;
    }
    const uint32_t slot = i - pt_offsets[lo];
    const uint32_t R = proofs[lo].rounds, raw = proofs[lo].raw_off;
    const uint8_t *src;
    if (slot < 3) src = blob + raw + BPP_RAW_A(ext) + 32u * slot;
    else if (slot < 3 + R) src = blob + raw + BPP_RAW_L(ext, slot - 3);
    else if (slot < 3 + 2 * R) src = blob + raw + BPP_RAW_R(ext, slot - 3 - R);
    else src = commitments32 + 32u * (size_t)(proofs[lo].commit_off + (slot - 3 - 2 * R));
    uint32_t w[8];
    ld32_unaligned(src, w);
    fe x, y, t;
    const bool good = ristretto_decode(x, y, t, w);
    if (!good) { x = fe_zero(); y = fe_one(); t = fe_zero(); }
    ok[i] = good ? 1 : 0;
    aniels q = ge_to_aniels_affine(x, y, t);
    store_fe(&out_tab[i].ypx, q.ypx); store_fe(&out_tab[i].ymx, q.ymx); store_fe(&out_tab[i].t2d, q.t2d);
}

__global__ void __launch_bounds__(128) k_encode(size_t n, const ge *__restrict__ in, uint32_t *__restrict__ out_enc,
                                               uint8_t *__restrict__ is_identity) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ge p;
    load8(p.X.v, in[i].X.v); load8(p.Y.v, in[i].Y.v); load8(p.Z.v, in[i].Z.v); load8(p.T.v, in[i].T.v);
    if (is_identity) is_identity[i] = ge_is_ristretto_identity(p) ? 1 : 0;
    if (out_enc) {
        fe s = ristretto_encode(p);
        store8(out_enc + 8 * i, s.v);
    }
}

__global__ void __launch_bounds__(128) k_from_uniform(size_t n, const uint32_t *__restrict__ in16, uint32_t *__restrict__ out_enc,
                                                     aniels *__restrict__ out_tab) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t w[16];
    load8(w, in16 + 16 * i);
    load8(w + 8, in16 + 16 * i + 8);
    ge p = ristretto_from_uniform_words(w);
    if (out_enc) {
        fe s = ristretto_encode(p);
        store8(out_enc + 8 * i, s.v);
    }
    if (out_tab) {
        fe zi = fe_invert(p.Z);
        fe x = fe_mul(p.X, zi), y = fe_mul(p.Y, zi);
        aniels q = ge_to_aniels_affine(x, y, fe_mul(x, y));
        store_fe(&out_tab[i].ypx, q.ypx); store_fe(&out_tab[i].ymx, q.ymx); store_fe(&out_tab[i].t2d, q.t2d);
    }
}

static inline unsigned grid_for(size_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

void launch_decompress(cudaStream_t s, size_t n, const uint32_t *in, aniels *out_tab, uint8_t *ok, uint32_t *out_enc,
                       uint32_t *bad_count) {
    if (n == 0) return;
    k_decompress<<<grid_for(n, 128), 128, 0, s>>>(n, in, out_tab, ok, out_enc, bad_count);
}
void launch_decompress_proofs(cudaStream_t s, uint32_t n_pts, uint32_t n_proofs, uint32_t ext, const VProof *proofs, const uint32_t *pt_offsets,
                              const uint8_t *blob, const uint8_t *commitments32, aniels *out_tab, uint8_t *ok) {
    if (n_pts == 0) return;
    // 116 registers per thread give four 128-thread CTAs per SM; capped at 96 a fifth one fits: -2.6 % on passes that fill the machine
    // (262 k points: 0.571 -> 0.555 ms), a little slower on small ones
    if (n_pts >= 100000) k_decompress_proofs<5><<<grid_for(n_pts, 128), 128, 0, s>>>(n_pts, n_proofs, ext, proofs, pt_offsets, blob, commitments32, out_tab, ok);
    else k_decompress_proofs<1><<<grid_for(n_pts, 128), 128, 0, s>>>(n_pts, n_proofs, ext, proofs, pt_offsets, blob, commitments32, out_tab, ok);
}
void launch_encode(cudaStream_t s, size_t n, const ge *in, uint32_t *out_enc, uint8_t *is_identity) {
    if (n == 0) return;
    k_encode<<<grid_for(n, 128), 128, 0, s>>>(n, in, out_enc, is_identity);
}
void launch_from_uniform(cudaStream_t s, size_t n, const uint32_t *in16, uint32_t *out_enc, aniels *out_tab) {
    if (n == 0) return;
    k_from_uniform<<<grid_for(n, 128), 128, 0, s>>>(n, in16, out_enc, out_tab);
}

} // namespace bpp
