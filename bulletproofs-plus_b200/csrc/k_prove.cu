// K-PROVE: device side of the batched, lock-step Bulletproofs+ prover (weighted-inner-product argument rounds).
//
// Restates the arithmetic of RangeProof::prove_with_rng (/root/reference/src/range_proof.rs:232-608) for P proofs of one
// shape (bit length n, aggregation m, N = n*m) advancing together; the Fiat-Shamir transcripts, nonces and RNG draws stay
// on the host (engine_prove.cu) and meet the device once per round:
//   k_prove_bits       bit decomposition -> the +-1 entries of the A commitment MSM (:300-345)
//   k_prove_init       y powers, y^-(2^k), a_L - z, a_R + d*y^(N-i) + z (:350-381)
//   k_prove_round_pre  a_lo*y^-n', a_hi*y^n', c_L / c_R, and the entry lists of the L and R MSMs (:413-495)
//   k_prove_round_inv  the fold scalars of a round from e and e^-1 (both from the host: one batch inversion per call and round)
//   k_prove_fold_pts   Gi' = e^-1*Gi_lo + e*y^-n'*Gi_hi, Hi' = e*Hi_lo + e^-1*Hi_hi (:511-521): one quad per output point runs a
//                      joint 2-bit-window double-scalar multiplication (the reference issues 2(N-1) two-point MSMs per
//                      proof here -- ~65 % of its proving time)
//   k_prove_fold_sc    a' = a_lo*e + a_hi'*e^-1, b' = b_lo*e^-1 + b_hi*e (:523-533)
// Scalars live in Montgomery form on the device; every MSM goes through K-MSM (k_msm.cu), segmented by (proof, L|R).
#include "kernels.cuh"
#include "quad.cuh"

namespace bpp {

static __device__ __forceinline__ sc p_ld_sc(const uint32_t *p) {
    sc r;
    uint4 a = reinterpret_cast<const uint4 *>(p)[0], b = reinterpret_cast<const uint4 *>(p)[1];
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
static __device__ __forceinline__ void p_st_sc(uint32_t *p, const sc &r) {
    reinterpret_cast<uint4 *>(p)[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    reinterpret_cast<uint4 *>(p)[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
static __device__ __forceinline__ sc pmm(const sc &a, const sc &b) { return sc_montmul(a, b); }

// ---------------------------------------------------------------------------------------------------------------- A
// entries of proof p: [N entries: bit ? (+1, G_i) : (-1, H_i)] [ext entries: alpha_k -> G_k (scalars written by the host)]
__global__ void __launch_bounds__(128) k_prove_bits(PDims d, PBuffers b) {
    uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= d.P * d.N) return;
    uint32_t p = gid / d.N, i = gid % d.N, j = i / d.n, bit_i = i % d.n;
    uint64_t v = b.offset_values[(size_t)p * d.m + j];                  // value - minimum_value_promise
    uint32_t bit = (uint32_t)(v >> bit_i) & 1u;
    size_t e = (size_t)p * (d.N + d.ext) + i;
    sc s = sc_one();
    if (!bit) s = sc_neg(s);                                             // a_R = a_L - 1 = -1
    p_st_sc(b.msm_scalars + 8 * e, s);
    b.msm_pidx[e] = 0x80000000u | (bit ? i : (uint32_t)(d.gens_nm + i));
    // a_L, a_R (Montgomery form) for the rounds
    const sc one_m = sc_const_R();
    p_st_sc(b.a + 8 * ((size_t)p * d.N + i), bit ? one_m : sc_zero());
    p_st_sc(b.b + 8 * ((size_t)p * d.N + i), bit ? sc_zero() : sc_neg(one_m));
    for (uint32_t k = i; k < d.ext; k += d.N) b.msm_pidx[(size_t)p * (d.N + d.ext) + d.N + k] = 0x80000000u | (uint32_t)(2 * d.gens_nm + k);
}

// ---------------------------------------------------------------------------------------------------------------- init
// per proof: ypow[0..N+1], yinv2[k] = y^-(2^k) (k < rounds)
__global__ void __launch_bounds__(64) k_prove_ypow(PDims d, PBuffers b) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= d.P) return;
    const sc y = sc_to_mont(p_ld_sc(b.yz + 16 * (size_t)p));
    uint32_t *yp = b.ypow + 8 * (size_t)p * (d.N + 2);
    sc acc = sc_const_R();
    for (uint32_t i = 0; i < d.N + 2; i++) { p_st_sc(yp + 8 * i, acc); acc = pmm(acc, y); }
    sc yi = sc_to_mont(p_ld_sc(b.yz + 16 * (size_t)d.P + 8 * (size_t)p));          // y^-1 from the host (batch inversion over the call)
    uint32_t *yv = b.yinv2 + 8 * (size_t)p * BPP_MAX_ROUNDS;
    for (uint32_t k = 0; k < d.rounds; k++) { p_st_sc(yv + 8 * k, yi); yi = pmm(yi, yi); }
}
// per (proof, i): a_L -= z; a_R += d[i]*y^(N-i) + z, d[i] = z^(2(j+1)) * 2^bit
__global__ void __launch_bounds__(128) k_prove_init(PDims d, PBuffers b) {
    uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= d.P * d.N) return;
    uint32_t p = gid / d.N, i = gid % d.N, j = i / d.n, bit_i = i % d.n;
    const sc z = sc_to_mont(p_ld_sc(b.yz + 16 * (size_t)p + 8));
    const sc z2 = pmm(z, z);
    sc zp = z2;
    for (uint32_t t = 0; t < j; t++) zp = pmm(zp, z2);                   // z^(2(j+1))
    sc two_b = sc_zero();
    two_b.v[bit_i >> 5] = 1u << (bit_i & 31);
    sc di = pmm(zp, sc_to_mont(two_b));
    const uint32_t *yp = b.ypow + 8 * (size_t)p * (d.N + 2);
    uint32_t *pa = b.a + 8 * ((size_t)p * d.N + i), *pb = b.b + 8 * ((size_t)p * d.N + i);
    p_st_sc(pa, sc_sub(p_ld_sc(pa), z));
    p_st_sc(pb, sc_add(p_ld_sc(pb), sc_add(pmm(di, p_ld_sc(yp + 8 * (d.N - i))), z)));
}

// ---------------------------------------------------------------------------------------------------------------- rounds
// one warp per proof.  nn = current half length.  Writes a_hi' in place (a_hi *= y^nn), keeps a_lo, and emits the entries
//   L: [c_L -> H, d_L[k] -> G_k, a_lo[i]*y^-nn -> Gi[nn+i], b_hi[i] -> Hi[i]]      R: [c_R -> H, d_R[k] -> G_k, a_hi'[i] -> Gi[i], b_lo[i] -> Hi[nn+i]]
// Segment 2p is L, 2p+1 is R; every segment has 1 + ext + 2*nn entries.
static __device__ __forceinline__ sc shfl_down_sc_p(const sc &a, int delta) {
    sc r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_down_sync(0xffffffffu, a.v[i], delta);
    return r;
}
__global__ void __launch_bounds__(128) k_prove_round_pre(PDims d, PBuffers b, uint32_t nn, uint32_t round) {
    const uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= d.P) return;
    const uint32_t seg_len = 1 + d.ext + 2 * nn;
    const size_t eL = (size_t)(2 * p) * seg_len, eR = eL + seg_len;
    const uint32_t *yp = b.ypow + 8 * (size_t)p * (d.N + 2);
    uint32_t log_nn = 31 - __clz(nn);
    const sc yinv_nn = p_ld_sc(b.yinv2 + 8 * ((size_t)p * BPP_MAX_ROUNDS + log_nn));
    const sc y_nn = p_ld_sc(yp + 8 * nn);
    uint32_t *a = b.a + 8 * (size_t)p * d.N, *bb = b.b + 8 * (size_t)p * d.N;
    // point sources: round 0 reads the shared generator table, later rounds the proof's folded vectors
    const uint32_t gbase = round == 0 ? 0x80000000u : (0x40000000u | (uint32_t)((size_t)p * d.N));
    const uint32_t hbase = round == 0 ? (0x80000000u | (uint32_t)d.gens_nm) : (0x40000000u | (uint32_t)((size_t)d.P * d.N + (size_t)p * d.N));
    sc cL = sc_zero(), cR = sc_zero();
    for (uint32_t i = lane; i < nn; i += 32) {
        sc a_lo = p_ld_sc(a + 8 * i), a_hi = p_ld_sc(a + 8 * (nn + i));
        sc b_lo = p_ld_sc(bb + 8 * i), b_hi = p_ld_sc(bb + 8 * (nn + i));
        cL = sc_add(cL, pmm(pmm(a_lo, p_ld_sc(yp + 8 * (i + 1))), b_hi));
        cR = sc_add(cR, pmm(pmm(a_hi, p_ld_sc(yp + 8 * (nn + 1 + i))), b_lo));
        sc a_lo_off = pmm(a_lo, yinv_nn), a_hi_off = pmm(a_hi, y_nn);
        p_st_sc(a + 8 * (nn + i), a_hi_off);                            // kept for the fold: a' = a_lo*e + a_hi'*e^-1
        uint32_t o = 1 + d.ext + i;
        p_st_sc(b.msm_scalars + 8 * (eL + o), sc_from_mont(a_lo_off));       b.msm_pidx[eL + o] = gbase + nn + i;
        p_st_sc(b.msm_scalars + 8 * (eL + o + nn), sc_from_mont(b_hi));      b.msm_pidx[eL + o + nn] = hbase + i;
        p_st_sc(b.msm_scalars + 8 * (eR + o), sc_from_mont(a_hi_off));       b.msm_pidx[eR + o] = gbase + i;
        p_st_sc(b.msm_scalars + 8 * (eR + o + nn), sc_from_mont(b_lo));      b.msm_pidx[eR + o + nn] = hbase + nn + i;
    }
    for (int delta = 16; delta > 0; delta >>= 1) { cL = sc_add(cL, shfl_down_sc_p(cL, delta)); cR = sc_add(cR, shfl_down_sc_p(cR, delta)); }
    if (lane == 0) {
        p_st_sc(b.msm_scalars + 8 * eL, sc_from_mont(cL)); b.msm_pidx[eL] = 0x80000000u | (uint32_t)(2 * d.gens_nm + d.ext);
        p_st_sc(b.msm_scalars + 8 * eR, sc_from_mont(cR)); b.msm_pidx[eR] = 0x80000000u | (uint32_t)(2 * d.gens_nm + d.ext);
    }
    if (lane < d.ext) {       // d_L[k], d_R[k] arrive from the host as canonical scalars: [p][L|R][k]
        const uint32_t *dl = b.dlr + 8 * ((size_t)p * 2 * d.ext + lane), *dr = dl + 8 * d.ext;
        p_st_sc(b.msm_scalars + 8 * (eL + 1 + lane), p_ld_sc(dl)); b.msm_pidx[eL + 1 + lane] = 0x80000000u | (uint32_t)(2 * d.gens_nm + lane);
        p_st_sc(b.msm_scalars + 8 * (eR + 1 + lane), p_ld_sc(dr)); b.msm_pidx[eR + 1 + lane] = 0x80000000u | (uint32_t)(2 * d.gens_nm + lane);
    }
}

// per proof: the four fold scalars in PLAIN canonical form (bit scanning) and Montgomery e, e^-1 for the scalar fold:
//   fsc[p] = [e^-1, e*y^-nn, e, e^-1 (plain) | e (mont), e^-1 (mont)]
__global__ void __launch_bounds__(64) k_prove_round_inv(PDims d, PBuffers b, uint32_t nn) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= d.P) return;
    const sc e = sc_to_mont(p_ld_sc(b.e + 8 * (size_t)p));
    const sc einv = sc_to_mont(p_ld_sc(b.e + 8 * ((size_t)d.P + p)));          // inverted on the host (batch inversion over the call)
    uint32_t log_nn = 31 - __clz(nn);
    const sc yinv_nn = p_ld_sc(b.yinv2 + 8 * ((size_t)p * BPP_MAX_ROUNDS + log_nn));
    uint32_t *f = b.fsc + 8 * (size_t)p * 6;
    const sc einv_p = sc_from_mont(einv);
    p_st_sc(f, einv_p);
    p_st_sc(f + 8, sc_from_mont(pmm(e, yinv_nn)));
    p_st_sc(f + 16, sc_from_mont(e));
    p_st_sc(f + 24, einv_p);
    p_st_sc(f + 32, e);
    p_st_sc(f + 40, einv);
}

// one quad per (proof, i < nn, which): which = 0 -> Gi'[i] = s0*Gi[i] + s1*Gi[nn+i] with (s0, s1) = (e^-1, e*y^-nn);
// which = 1 -> Hi'[i] = s0*Hi[i] + s1*Hi[nn+i] with (e, e^-1).  Joint 2-bit windows: table T[4a+b] = a*P + b*Q (a, b < 4) in
// shared memory (cached form, one quad = 2 KB), then 126 steps of (2 doublings + 1 table addition).  The table index is
// uniform inside a quad, each lane fetches its own field: no divergence whatever the scalars are.
#define FOLD_QUADS 16    // quads per CTA (64 threads): 32 KB of shared memory
__global__ void __launch_bounds__(4 * FOLD_QUADS) k_prove_fold_pts(PDims d, PBuffers b, uint32_t nn, uint32_t round, const aniels *__restrict__ gens) {
    __shared__ fe table[FOLD_QUADS][16][4];
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 2, ql = threadIdx.x >> 2;
    const int role = threadIdx.x & 3, base = (threadIdx.x & 31) & ~3;
    const uint32_t total = d.P * nn * 2;
    const bool valid = q < total;
    const uint32_t qq = valid ? q : 0;
    const uint32_t which = qq & 1u, pi = qq >> 1, p = pi / nn, i = pi % nn;
    cached *vec = b.folded + (which ? (size_t)d.P * d.N : 0) + (size_t)p * d.N;
    // operands P = vec[i], Q = vec[nn+i] as this lane's cached field
    fe Pc, Qc;
    if (round == 0) {
        const aniels *g = gens + (which ? d.gens_nm : 0);
        const aniels *sp = g + i, *sq = g + nn + i;
        if (role == 0) { Pc = ld_fe(&sp->ymx); Qc = ld_fe(&sq->ymx); }
        else if (role == 1) { Pc = ld_fe(&sp->ypx); Qc = ld_fe(&sq->ypx); }
        else if (role == 2) { Pc = fe_from_u32(2); Qc = Pc; }
        else { Pc = ld_fe(&sp->t2d); Qc = ld_fe(&sq->t2d); }
    } else {
        Pc = ld_fe(reinterpret_cast<const fe *>(&vec[i]) + role);
        Qc = ld_fe(reinterpret_cast<const fe *>(&vec[nn + i]) + role);
    }
    // table T[4a + b] = a*P + b*Q, built column by column in extended coordinates (col = b*Q, then + P three times); every
    // lane only ever touches its own field of the table, so no synchronisation is needed
    fe (*T)[4] = table[ql];
    T[0][role] = quad_cached_identity(role);
    fe acc;
    fe col = quad_identity(role);
    for (int bq = 0; bq < 4; bq++) {
        if (bq > 0) col = quad_add(col, role, base, Qc);
        fe cur = col;
        for (int a = 0; a < 4; a++) {
            if (a > 0) cur = quad_add(cur, role, base, Pc);
            if (a + bq > 0) T[4 * a + bq][role] = quad_to_cached(cur, role, base);
        }
    }
    const uint32_t *f = b.fsc + 8 * (size_t)p * 6 + (which ? 16 : 0);
    uint32_t s0[8], s1[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { s0[k] = f[k]; s1[k] = f[8 + k]; }
    acc = quad_identity(role);
    for (int bit = 252; bit >= 0; bit -= 2) {          // scalars < 2^253: 127 two-bit steps from bit 252/253 down
        acc = quad_dbl(acc, role, base);
        acc = quad_dbl(acc, role, base);
        uint32_t wa = (s0[bit >> 5] >> (bit & 31)) & 3u, wb = (s1[bit >> 5] >> (bit & 31)) & 3u;   // bit is even: no word straddle
        acc = quad_add(acc, role, base, T[4 * wa + wb][role]);
    }
    fe out = quad_to_cached(acc, role, base);
    if (valid) st_fe(reinterpret_cast<fe *>(&vec[i]) + role, out);
}

// per (proof, i < nn): a'[i] = a_lo[i]*e + a_hi'[i]*e^-1;  b'[i] = b_lo[i]*e^-1 + b_hi[i]*e
__global__ void __launch_bounds__(128) k_prove_fold_sc(PDims d, PBuffers b, uint32_t nn) {
    uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= d.P * nn) return;
    uint32_t p = gid / nn, i = gid % nn;
    const uint32_t *f = b.fsc + 8 * (size_t)p * 6;
    const sc e = p_ld_sc(f + 32), einv = p_ld_sc(f + 40);
    uint32_t *a = b.a + 8 * (size_t)p * d.N, *bb = b.b + 8 * (size_t)p * d.N;
    sc a_lo = p_ld_sc(a + 8 * i), a_hi = p_ld_sc(a + 8 * (nn + i)), b_lo = p_ld_sc(bb + 8 * i), b_hi = p_ld_sc(bb + 8 * (nn + i));
    p_st_sc(a + 8 * i, sc_add(pmm(a_lo, e), pmm(a_hi, einv)));
    p_st_sc(bb + 8 * i, sc_add(pmm(b_lo, einv), pmm(b_hi, e)));
}

// a[0], b[0] of every proof as canonical scalars (for r1, s1 and the A1 scalar on the host)
__global__ void __launch_bounds__(128) k_prove_final_ab(PDims d, PBuffers b, uint32_t *out) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= d.P) return;
    p_st_sc(out + 16 * (size_t)p, sc_from_mont(p_ld_sc(b.a + 8 * (size_t)p * d.N)));
    p_st_sc(out + 16 * (size_t)p + 8, sc_from_mont(p_ld_sc(b.b + 8 * (size_t)p * d.N)));
}

// ================================================================================================================ fixed-base path
// Same protocol, no generator folding: the folded vectors of round k are linear combinations of the ORIGINAL generators,
//   Gi^(k)[i] = sum_{j : j mod n_k = i} sG[j] * G_j,      Hi^(k)[i] = sum_{j : j mod n_k = i} sH[j] * H_j        (n_k = N / 2^k),
// with sG = sH = 1 before round 0 and, at the fold with challenge e and nn = n_k / 2 (:511-521),
//   (j mod n_k) <  nn :  sG[j] *= e^-1,         sH[j] *= e
//   (j mod n_k) >= nn :  sG[j] *= e * y^-nn,    sH[j] *= e^-1.
// Every L / R / A1 is therefore an N-term sum over the static generators, evaluated by K-FB (k_fb.cu) from window tables.
// Entry layouts (scalars canonical, one segment per commitment; the matching generator-index rows come from the host):
//   A   (seg_len 2N + ext)       [a_L[j] -> G_j | a_R[j] -> H_j | alpha_k -> G_k]
//   L/R (seg_len 1 + ext + N)    [c -> H | d[k] -> G_k | N/2 G terms | N/2 H terms]
//         L: G terms t -> j = (t / nn) * 2nn + nn + (t % nn),  a_lo[t % nn] * y^-nn * sG[j];   H terms -> j - nn,  b_hi[t % nn] * sH[j - nn]
//         R: G terms t -> j = (t / nn) * 2nn + (t % nn),       a_hi[t % nn] * y^nn  * sG[j];   H terms -> j + nn,  b_lo[t % nn] * sH[j + nn]
//   A1  (seg_len 2N + 1 + ext)   [r * sG[j] -> G_j | s * sH[j] -> H_j | d[k] -> G_k | (r*y*b + s*y*a) -> H]
__global__ void __launch_bounds__(128) k_prove_bits_fb(PDims d, PBuffers b) {
    uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= d.P * d.N) return;
    uint32_t p = gid / d.N, i = gid % d.N, j = i / d.n, bit_i = i % d.n;
    uint64_t v = b.offset_values[(size_t)p * d.m + j];
    uint32_t bit = (uint32_t)(v >> bit_i) & 1u;
    const size_t seg = (size_t)p * (2 * d.N + d.ext);
    p_st_sc(b.msm_scalars + 8 * (seg + i), bit ? sc_one() : sc_zero());                  // a_L
    p_st_sc(b.msm_scalars + 8 * (seg + d.N + i), bit ? sc_zero() : sc_neg(sc_one()));    // a_R = a_L - 1
    const sc one_m = sc_const_R();
    p_st_sc(b.a + 8 * ((size_t)p * d.N + i), bit ? one_m : sc_zero());
    p_st_sc(b.b + 8 * ((size_t)p * d.N + i), bit ? sc_zero() : sc_neg(one_m));
    p_st_sc(b.sg + 8 * ((size_t)p * d.N + i), one_m);
    p_st_sc(b.sh + 8 * ((size_t)p * d.N + i), one_m);
}

// one warp per proof; segment 2p is L, 2p + 1 is R
__global__ void __launch_bounds__(128) k_prove_round_pre_fb(PDims d, PBuffers b, uint32_t nn) {
    const uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= d.P) return;
    const uint32_t seg_len = 1 + d.ext + d.N, half = d.N / 2;
    const size_t eL = (size_t)(2 * p) * seg_len, eR = eL + seg_len;
    const uint32_t *yp = b.ypow + 8 * (size_t)p * (d.N + 2);
    const uint32_t log_nn = 31 - __clz(nn);
    const sc yinv_nn = p_ld_sc(b.yinv2 + 8 * ((size_t)p * BPP_MAX_ROUNDS + log_nn));
    const sc y_nn = p_ld_sc(yp + 8 * nn);
    uint32_t *a = b.a + 8 * (size_t)p * d.N, *bb = b.b + 8 * (size_t)p * d.N;
    const uint32_t *sg = b.sg + 8 * (size_t)p * d.N, *sh = b.sh + 8 * (size_t)p * d.N;
    sc cL = sc_zero(), cR = sc_zero();
    for (uint32_t i = lane; i < nn; i += 32) {
        sc a_lo = p_ld_sc(a + 8 * i), a_hi = p_ld_sc(a + 8 * (nn + i));
        sc b_lo = p_ld_sc(bb + 8 * i), b_hi = p_ld_sc(bb + 8 * (nn + i));
        cL = sc_add(cL, pmm(pmm(a_lo, p_ld_sc(yp + 8 * (i + 1))), b_hi));
        cR = sc_add(cR, pmm(pmm(a_hi, p_ld_sc(yp + 8 * (nn + 1 + i))), b_lo));
        p_st_sc(a + 8 * (nn + i), pmm(a_hi, y_nn));                     // a_hi' kept for the fold: a' = a_lo*e + a_hi'*e^-1
    }
    __syncwarp();
    for (uint32_t t = lane; t < half; t += 32) {
        const uint32_t i = t % nn, jl = (t / nn) * 2 * nn + i, ju = jl + nn;
        const sc a_lo_off = pmm(p_ld_sc(a + 8 * i), yinv_nn), a_hi_off = p_ld_sc(a + 8 * (nn + i));
        const uint32_t o = 1 + d.ext + t;
        p_st_sc(b.msm_scalars + 8 * (eL + o), sc_from_mont(pmm(a_lo_off, p_ld_sc(sg + 8 * ju))));
        p_st_sc(b.msm_scalars + 8 * (eL + o + half), sc_from_mont(pmm(p_ld_sc(bb + 8 * (nn + i)), p_ld_sc(sh + 8 * jl))));
        p_st_sc(b.msm_scalars + 8 * (eR + o), sc_from_mont(pmm(a_hi_off, p_ld_sc(sg + 8 * jl))));
        p_st_sc(b.msm_scalars + 8 * (eR + o + half), sc_from_mont(pmm(p_ld_sc(bb + 8 * i), p_ld_sc(sh + 8 * ju))));
    }
    for (int delta = 16; delta > 0; delta >>= 1) { cL = sc_add(cL, shfl_down_sc_p(cL, delta)); cR = sc_add(cR, shfl_down_sc_p(cR, delta)); }
    if (lane == 0) {
        p_st_sc(b.msm_scalars + 8 * eL, sc_from_mont(cL));
        p_st_sc(b.msm_scalars + 8 * eR, sc_from_mont(cR));
    }
    if (lane < d.ext) {
        const uint32_t *dl = b.dlr + 8 * ((size_t)p * 2 * d.ext + lane), *dr = dl + 8 * d.ext;
        p_st_sc(b.msm_scalars + 8 * (eL + 1 + lane), p_ld_sc(dl));
        p_st_sc(b.msm_scalars + 8 * (eR + 1 + lane), p_ld_sc(dr));
    }
}

// per (proof, j < N): the fold as scalar updates; j < nn also folds a and b (:523-533)
__global__ void __launch_bounds__(128) k_prove_fold_fb(PDims d, PBuffers b, uint32_t nn) {
    uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= d.P * d.N) return;
    const uint32_t p = gid / d.N, j = gid % d.N;
    const uint32_t *f = b.fsc + 8 * (size_t)p * 6;
    const sc e = p_ld_sc(f + 32), einv = p_ld_sc(f + 40);
    const uint32_t log_nn = 31 - __clz(nn);
    const bool upper = (j & (2 * nn - 1)) >= nn;
    uint32_t *sg = b.sg + 8 * ((size_t)p * d.N + j), *sh = b.sh + 8 * ((size_t)p * d.N + j);
    if (upper) {
        const sc eyinv = pmm(e, p_ld_sc(b.yinv2 + 8 * ((size_t)p * BPP_MAX_ROUNDS + log_nn)));
        p_st_sc(sg, pmm(p_ld_sc(sg), eyinv));
        p_st_sc(sh, pmm(p_ld_sc(sh), einv));
    } else {
        p_st_sc(sg, pmm(p_ld_sc(sg), einv));
        p_st_sc(sh, pmm(p_ld_sc(sh), e));
    }
    if (j < nn) {
        uint32_t *a = b.a + 8 * (size_t)p * d.N, *bb = b.b + 8 * (size_t)p * d.N;
        sc a_lo = p_ld_sc(a + 8 * j), a_hi = p_ld_sc(a + 8 * (nn + j)), b_lo = p_ld_sc(bb + 8 * j), b_hi = p_ld_sc(bb + 8 * (nn + j));
        p_st_sc(a + 8 * j, sc_add(pmm(a_lo, e), pmm(a_hi, einv)));
        p_st_sc(bb + 8 * j, sc_add(pmm(b_lo, einv), pmm(b_hi, e)));
    }
}

// per (proof, j < N): the generator terms of A1 (r * sG[j], s * sH[j]); rs: P x 2 canonical scalars from the host
__global__ void __launch_bounds__(128) k_prove_final_fb(PDims d, PBuffers b, const uint32_t *__restrict__ rs) {
    uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= d.P * d.N) return;
    const uint32_t p = gid / d.N, j = gid % d.N;
    const sc r = sc_to_mont(p_ld_sc(rs + 16 * (size_t)p)), s = sc_to_mont(p_ld_sc(rs + 16 * (size_t)p + 8));
    const size_t seg = (size_t)p * (2 * d.N + 1 + d.ext);
    p_st_sc(b.msm_scalars + 8 * (seg + j), sc_from_mont(pmm(r, p_ld_sc(b.sg + 8 * ((size_t)p * d.N + j)))));
    p_st_sc(b.msm_scalars + 8 * (seg + d.N + j), sc_from_mont(pmm(s, p_ld_sc(b.sh + 8 * ((size_t)p * d.N + j)))));
}

// ---------------------------------------------------------------------------------------------------------------- launchers
void launch_prove_bits(cudaStream_t s, const PDims &d, const PBuffers &b) {
    k_prove_bits<<<(d.P * d.N + 127) / 128, 128, 0, s>>>(d, b);
}
void launch_prove_init(cudaStream_t s, const PDims &d, const PBuffers &b) {
    k_prove_ypow<<<(d.P + 63) / 64, 64, 0, s>>>(d, b);
    k_prove_init<<<(d.P * d.N + 127) / 128, 128, 0, s>>>(d, b);
}
void launch_prove_round_pre(cudaStream_t s, const PDims &d, const PBuffers &b, uint32_t nn, uint32_t round) {
    k_prove_round_pre<<<(d.P + 3) / 4, 128, 0, s>>>(d, b, nn, round);
}
void launch_prove_fold(cudaStream_t s, const PDims &d, const PBuffers &b, uint32_t nn, uint32_t round, const aniels *gens) {
    k_prove_round_inv<<<(d.P + 63) / 64, 64, 0, s>>>(d, b, nn);
    uint32_t quads = d.P * nn * 2;
    k_prove_fold_pts<<<(quads + FOLD_QUADS - 1) / FOLD_QUADS, 4 * FOLD_QUADS, 0, s>>>(d, b, nn, round, gens);
    k_prove_fold_sc<<<(d.P * nn + 127) / 128, 128, 0, s>>>(d, b, nn);
}
void launch_prove_bits_fb(cudaStream_t s, const PDims &d, const PBuffers &b) {
    k_prove_bits_fb<<<(d.P * d.N + 127) / 128, 128, 0, s>>>(d, b);
}
void launch_prove_round_pre_fb(cudaStream_t s, const PDims &d, const PBuffers &b, uint32_t nn) {
    k_prove_round_pre_fb<<<(d.P + 3) / 4, 128, 0, s>>>(d, b, nn);
}
void launch_prove_fold_fb(cudaStream_t s, const PDims &d, const PBuffers &b, uint32_t nn) {
    k_prove_round_inv<<<(d.P + 63) / 64, 64, 0, s>>>(d, b, nn);
    k_prove_fold_fb<<<(d.P * d.N + 127) / 128, 128, 0, s>>>(d, b, nn);
}
void launch_prove_final_fb(cudaStream_t s, const PDims &d, const PBuffers &b, const uint32_t *rs) {
    k_prove_final_fb<<<(d.P * d.N + 127) / 128, 128, 0, s>>>(d, b, rs);
}
void launch_prove_final_ab(cudaStream_t s, const PDims &d, const PBuffers &b, uint32_t *out) {
    k_prove_final_ab<<<(d.P + 127) / 128, 128, 0, s>>>(d, b, out);
}

} // namespace bpp
