// K-VPREP: per-proof scalar synthesis of the Bulletproofs+ batch verifier.
//
// Restates loop 2 of RangeProof::verify (/root/reference/src/range_proof.rs:856-1033) as three data-parallel stages:
//   A  k_vprep_proof    one thread per proof: batch inversion, challenge powers, d_sum / y_sum, the dynamic MSM scalars
//                       (A1, B, A, L_j, R_j, V_j), the proof's h / g_k contributions, optional mask recovery (:941-969)
//   B  k_vprep_vector   one CTA per proof: contributions to gi_base_scalars[i] / hi_base_scalars[i] (:987-1003).  Every
//                       i-dependent factor (s[i], s[N-1-i], y^-i, 2^(i mod n)) is a product over the set bits of i, so three
//                       two-level tables (low / high half of the bits of i, built in shared memory by doubling) give each
//                       of them with ONE multiplication per element instead of the reference's serial recurrence
//   W  k_vprep_weight   one warp per proof: multiplies the proof's dynamic scalars by its batch weight
//   C  k_vprep_reduce   one warp per (chunk, static slot): weighted column sums over the chunk's proofs (the `+=` of
//                       :999-1000, :1017-1020), written as canonical scalars into the chunk's MSM entry list
// Every term of proof p carries its weight w_p exactly once (range_proof.rs:894, :999-1032), so A and B run weight-free and
// W / C apply w_p at the end: the weights come out of a sequential Merlin transcript on the host (:811-853), and this
// ordering lets A and B run on the device while the host is still hashing.
// All arithmetic is mod l in Montgomery form (arith.cuh sc_montmul); results are bit-exact canonical scalars.
#include <stdlib.h>
#include "kernels.cuh"
#include "rawld.cuh"

namespace bpp {

static __device__ __forceinline__ sc ld_sc(const uint32_t *p) {
    sc r;
    uint4 a = reinterpret_cast<const uint4 *>(p)[0], b = reinterpret_cast<const uint4 *>(p)[1];
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
static __device__ __forceinline__ void st_sc(uint32_t *p, const sc &r) {
    reinterpret_cast<uint4 *>(p)[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    reinterpret_cast<uint4 *>(p)[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
static __device__ __forceinline__ sc mm(const sc &a, const sc &b) { return sc_montmul(a, b); }
static __device__ __forceinline__ sc ld_mont(const uint32_t *p) { return sc_to_mont(ld_sc(p)); }
// a canonical scalar inside the uploaded proof bytes (any alignment) -> Montgomery form
static __device__ __forceinline__ sc ld_raw_mont(const uint8_t *p) {
    sc r;
    ld32_unaligned(p, r.v);
    return sc_to_mont(r);
}

// pervec layout (scalars, Montgomery form), per proof at pv_off (the batch weight is applied by stages W / C).
// Bit k of i corresponds to challenge e_j with j = R-1-k (range_proof.rs:987-996: s[i] = prod_j e_j^(+-1)):
//   [0] U0 = r1*e*s[0]      (s[0] = prod e_j^-1)         u_i = U0 * prod_{k in bits(i)} ru_k = r1*e*y^-i*s[i]
//   [1] V0 = s1*e*s[N-1]    (s[N-1] = prod e_j)          v_i = V0 * prod_{k in bits(i)} rv_k = s1*e*s[N-1-i]
//   [2] Q0 = e^2*y^N                                     q_i = Q0 * prod_{k in bits(i)} rq_k = e^2*y^(N-i)*2^(i mod n)
//   [3] e^2*z    [4..7] spare
//   [8 + k]            ru_k = e_j^2 * y^-(2^k)           k < rounds
//   [8 + R + k]        rv_k = e_j^-2
//   [8 + 2R + k]       rq_k = y^-(2^k) * (2^(2^k) if 2^k < n else 1)
//   [8 + 3R + j]       z^(2(j+1))     j < m
#define PV_HDR 8

__global__ void __launch_bounds__(64) k_vprep_proof(VDims d, VBuffers b) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= d.n_proofs) return;
    const VProof pr = b.proofs[p];
    if (!pr.replay || (!pr.active && pr.nonce_off == 0xffffffffu)) return;
    const uint32_t R = pr.rounds, m = pr.m, ext = d.ext;
    const sc one_m = sc_const_R();                 // 1 in Montgomery form
    const uint32_t *ch = b.challenges + 8 * (size_t)pr.ch_off;
    sc y = ld_mont(ch), z = ld_mont(ch + 8), e = ld_mont(ch + 16);
    const uint8_t *raw = b.blob + pr.raw_off;
    sc r1 = ld_raw_mont(raw + BPP_RAW_R1(ext)), s1 = ld_raw_mont(raw + BPP_RAW_S1(ext));

    sc z2 = mm(z, z), e2 = mm(e, e);
    sc yN = y;
    for (uint32_t k = 0; k < R; k++) yN = mm(yN, yN);       // N = 2^rounds (checked on the host)
    sc yN1 = mm(yN, y);
    sc ym1 = sc_sub(y, one_m);
    sc zy = mm(z2, yN1);

    // batch inversion of [e_0..e_{R-1}, y, y-1, e, z^2*y^(N+1)]  (range_proof.rs:897-905 plus the two inversions of :950,:958)
    sc ej[BPP_MAX_ROUNDS], pre[BPP_MAX_ROUNDS + 4];
    sc acc = one_m;
    for (uint32_t j = 0; j < R; j++) { ej[j] = ld_mont(ch + 24 + 8 * j); pre[j] = acc; acc = mm(acc, ej[j]); }
    pre[R] = acc; acc = mm(acc, y);
    pre[R + 1] = acc; acc = mm(acc, ym1);
    pre[R + 2] = acc; acc = mm(acc, e);
    pre[R + 3] = acc; acc = mm(acc, zy);
    sc inv = scm_invert_gcd(acc);      // binary Euclid: ~8x shorter latency chain than a^(l-2)
    sc zy_inv = mm(inv, pre[R + 3]); inv = mm(inv, zy);
    sc e_inv = mm(inv, pre[R + 2]); inv = mm(inv, e);
    sc ym1_inv = mm(inv, pre[R + 1]); inv = mm(inv, ym1);
    sc y_inv = mm(inv, pre[R]); inv = mm(inv, y);
    const sc s0 = inv;                  // (prod e_j)^-1 = s[0]
    sc ejinv[BPP_MAX_ROUNDS];
    for (int j = (int)R - 1; j >= 0; j--) { ejinv[j] = mm(inv, pre[j]); inv = mm(inv, ej[j]); }

    // mask recovery (range_proof.rs:941-969)
    if (pr.nonce_off != 0xffffffffu && b.masks) {
        const uint32_t *nn = b.nonces + 8 * (size_t)pr.nonce_off;
        sc e2inv = mm(e_inv, e_inv);
        for (uint32_t k = 0; k < ext; k++) {
            sc d1k = ld_raw_mont(raw + BPP_RAW_D1(ext, k));
            sc eta = ld_mont(nn + 8 * k), dk = ld_mont(nn + 8 * (ext + k)), alpha = ld_mont(nn + 8 * (2 * ext + k));
            sc mask = sc_sub(sc_sub(d1k, eta), mm(e, dk));
            mask = sc_sub(mm(mask, e2inv), alpha);
            for (uint32_t j = 0; j < R; j++) {
                sc dL = ld_mont(nn + 8 * (3 * ext + j * ext + k));
                sc dR = ld_mont(nn + 8 * (3 * ext + R * ext + j * ext + k));
                mask = sc_sub(mask, mm(mm(ej[j], ej[j]), dL));
                mask = sc_sub(mask, mm(mm(ejinv[j], ejinv[j]), dR));
            }
            mask = mm(mask, zy_inv);
            st_sc(b.masks + 8 * ((size_t)p * ext + k), sc_from_mont(mask));
        }
    }
    if (!pr.active) return;

    // y_sum = y*(y^N - 1)/(y - 1);  d_sum = (2^n - 1) * sum_{j=1..m} z^(2j) by log-doubling (range_proof.rs:908-938)
    sc y_sum = mm(mm(sc_sub(yN, one_m), y), ym1_inv);
    sc d_sum = z2, tz = z2;
    for (uint32_t mmv = m; mmv > 1; mmv >>= 1) { d_sum = sc_add(d_sum, mm(d_sum, tz)); tz = mm(tz, tz); }
    uint64_t two_n_m1 = d.bit_length >= 64 ? ~0ull : ((1ull << d.bit_length) - 1ull);
    d_sum = mm(d_sum, sc_to_mont(sc_from_u64(two_n_m1)));

    const sc neg_e2 = sc_neg(e2);
    uint32_t *out = b.msm_scalars + 8 * (size_t)pr.entry_off;
    // dynamic entries of this proof: [A1, B, A, L_0.., R_0.., V_0..], still WITHOUT the weight and in Montgomery form
    st_sc(out, sc_neg(e));
    st_sc(out + 8, sc_neg(one_m));
    st_sc(out + 16, neg_e2);
    for (uint32_t j = 0; j < R; j++) {
        st_sc(out + 8 * (3 + j), mm(neg_e2, mm(ej[j], ej[j])));
        st_sc(out + 8 * (3 + R + j), mm(neg_e2, mm(ejinv[j], ejinv[j])));
    }
    uint32_t *pv = b.pervec + 8 * (size_t)pr.pv_off;
    sc h = sc_zero();
    sc zpow = one_m;
    const sc neg_e2_yN1 = mm(neg_e2, yN1);
    for (uint32_t j = 0; j < m; j++) {
        zpow = mm(zpow, z2);
        st_sc(pv + 8 * (PV_HDR + 3 * R + j), zpow);
        sc weighted = mm(neg_e2_yN1, zpow);
        st_sc(out + 8 * (3 + 2 * R + j), weighted);
        if (b.min_present[pr.commit_off + j]) {
            sc mv = sc_to_mont(sc_from_u64(b.min_values[pr.commit_off + j]));
            h = sc_sub(h, mm(weighted, mv));
        }
    }
    // h += r1*y*s1 + e^2*(y^(N+1)*z*d_sum + (z^2 - z)*y_sum); g_k += d1[k]   (range_proof.rs:1017-1020, weight applied later)
    sc t = mm(mm(r1, y), s1);
    sc u = sc_add(mm(mm(yN1, z), d_sum), mm(sc_sub(z2, z), y_sum));
    h = sc_add(h, sc_add(t, mm(e2, u)));
    uint32_t *hg = b.hg_contrib + 8 * (size_t)p * (1 + ext);
    st_sc(hg, h);
    for (uint32_t k = 0; k < ext; k++) st_sc(hg + 8 * (1 + k), ld_raw_mont(raw + BPP_RAW_D1(ext, k)));

    // hand-off to stage B (layout above); pre[R] = prod e_j = s[N-1]
    st_sc(pv, mm(mm(e, r1), s0));
    st_sc(pv + 8, mm(mm(e, s1), pre[R]));
    st_sc(pv + 16, mm(e2, yN));
    st_sc(pv + 24, mm(e2, z));
    sc yp = y_inv;
    for (uint32_t k = 0; k < R; k++) {
        const uint32_t j = R - 1 - k;
        st_sc(pv + 8 * (PV_HDR + k), mm(mm(ej[j], ej[j]), yp));
        st_sc(pv + 8 * (PV_HDR + R + k), mm(ejinv[j], ejinv[j]));
        sc rq = yp;
        if ((1u << k) < d.bit_length) rq = mm(rq, sc_to_mont(sc_from_u64(1ull << (1u << k))));     // bit_length <= 64: k <= 5
        st_sc(pv + 8 * (PV_HDR + 2 * R + k), rq);
        yp = mm(yp, yp);
    }
}

// Stage A, warp form (default): one WARP per proof.  The thread form above walks ~150 Montgomery products and a binary-Euclid
// inversion one after the other (200 us for any batch size: the second longest latency chain of a pass in round 1); here lane j owns
// round j (rounds <= 24), the batch inversion is a shuffle scan (prefix and suffix products), sums over rounds / commitments are
// shuffle reductions, and what remains sequential is the y^(2^k) chain (rounds squarings, twice), two 5-step scans and ONE inversion.
// y_sum is taken as y * prod_{k < rounds} (1 + y^(2^k)) = y (y^N - 1) / (y - 1) and d_sum as (2^n - 1) * sum_j z^(2j) directly: the
// same scalars mod l as range_proof.rs:913-938 without the (y - 1) inverse (y = 1 is flagged by the replay).
static __device__ __forceinline__ sc shfl_sc(const sc &a, int src) {
    sc r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(0xffffffffu, a.v[i], src);
    return r;
}
static __device__ __forceinline__ sc shfl_up_sc(const sc &a, int delta) {
    sc r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_up_sync(0xffffffffu, a.v[i], delta);
    return r;
}
static __device__ __forceinline__ sc shfl_xor_sc(const sc &a, int mask) {
    sc r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_xor_sync(0xffffffffu, a.v[i], mask);
    return r;
}
static __device__ __forceinline__ sc shfl_down_sc2(const sc &a, int delta) {
    sc r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_down_sync(0xffffffffu, a.v[i], delta);
    return r;
}
static __device__ __forceinline__ sc warp_sum_sc(sc a) {
    for (int mk = 16; mk > 0; mk >>= 1) a = sc_add(a, shfl_xor_sc(a, mk));
    return a;
}
static __device__ __forceinline__ sc sc_pick(const sc &a, const sc &b, bool pick_b) {
    sc r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = pick_b ? b.v[i] : a.v[i];
    return r;
}

__global__ void __launch_bounds__(128) k_vprep_proof_w(VDims d, VBuffers b) {
    const uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= d.n_proofs) return;
    const VProof pr = b.proofs[p];
    if (!pr.replay || (!pr.active && pr.nonce_off == 0xffffffffu)) return;     // warp-uniform
    const int R = (int)pr.rounds;                   // <= BPP_MAX_ROUNDS = 24: rounds, y, e and z^2 y^(N+1) fit one warp
    const uint32_t m = pr.m, ext = d.ext;
    const sc one_m = sc_const_R();
    const uint32_t *ch = b.challenges + 8 * (size_t)pr.ch_off;
    const uint8_t *raw = b.blob + pr.raw_off;
    const sc y = ld_mont(ch), z = ld_mont(ch + 8), e = ld_mont(ch + 16);
    const sc z2 = mm(z, z), e2 = mm(e, e);
    // y^(2^k): every lane walks the chain, lane k keeps y^(2^k)
    sc yk = y, ypow = y;
    for (int k = 1; k <= R; k++) { yk = mm(yk, yk); if (lane == k) ypow = yk; }
    const sc yN = yk, yN1 = mm(yN, y);
    const sc zy = mm(z2, yN1);
    // items of the batch inversion: lane j < R: e_j; R: y; R + 1: e; R + 2: z^2 y^(N+1); others 1
    sc x = one_m;
    if (lane < R) x = ld_mont(ch + 24 + 8 * lane);
    else if (lane == R) x = y;
    else if (lane == R + 1) x = e;
    else if (lane == R + 2) x = zy;
    sc P = x, S = x;                                // inclusive prefix / suffix products
    for (int dd = 1; dd < 32; dd <<= 1) {
        const sc t = shfl_up_sc(P, dd), u = shfl_down_sc2(S, dd);
        const sc pn = mm(P, t), sn = mm(S, u);
        P = sc_pick(P, pn, lane >= dd);
        S = sc_pick(S, sn, lane + dd < 32);
    }
    const sc inv_total = scm_invert_gcd(shfl_sc(P, 31));           // every lane computes it: no broadcast on the critical path
    sc Pm1 = shfl_up_sc(P, 1), Sp1 = shfl_down_sc2(S, 1);
    Pm1 = sc_pick(Pm1, one_m, lane == 0);
    Sp1 = sc_pick(Sp1, one_m, lane == 31);
    const sc xinv = mm(inv_total, mm(Pm1, Sp1));                    // lane j < R: e_j^-1; R: y^-1; R + 1: e^-1; R + 2: (z^2 y^(N+1))^-1
    const sc y_inv = shfl_sc(xinv, R), e_inv = shfl_sc(xinv, R + 1), zy_inv = shfl_sc(xinv, R + 2);
    const sc s0 = mm(inv_total, shfl_sc(S, R));                     // prod e_j^-1 = s[0]
    const sc prod_e = shfl_sc(P, R - 1);                            // prod e_j = s[N-1]
    const sc ej2 = mm(x, x), ejinv2 = mm(xinv, xinv);               // meaningful on lanes < R

    // mask recovery (range_proof.rs:941-969)
    if (pr.nonce_off != 0xffffffffu && b.masks) {
        const uint32_t *nn = b.nonces + 8 * (size_t)pr.nonce_off;
        const sc e2inv = mm(e_inv, e_inv);
        for (uint32_t k = 0; k < ext; k++) {
            sc term = sc_zero();
            if (lane < R) {
                const sc dL = ld_mont(nn + 8 * (3 * ext + lane * ext + k)), dR = ld_mont(nn + 8 * (3 * ext + R * ext + lane * ext + k));
                term = sc_add(mm(ej2, dL), mm(ejinv2, dR));
            }
            term = warp_sum_sc(term);
            if (lane == 0) {
                const sc d1k = ld_raw_mont(raw + BPP_RAW_D1(ext, k));
                const sc eta = ld_mont(nn + 8 * k), dk = ld_mont(nn + 8 * (ext + k)), alpha = ld_mont(nn + 8 * (2 * ext + k));
                sc mask = sc_sub(sc_sub(d1k, eta), mm(e, dk));
                mask = sc_sub(sc_sub(mm(mask, e2inv), alpha), term);
                st_sc(b.masks + 8 * ((size_t)p * ext + k), sc_from_mont(mm(mask, zy_inv)));
            }
        }
    }
    if (!pr.active) return;

    const sc neg_e2 = sc_neg(e2);
    uint32_t *out = b.msm_scalars + 8 * (size_t)pr.entry_off;
    uint32_t *pv = b.pervec + 8 * (size_t)pr.pv_off;
    // dynamic entries [A1, B, A, L_0.., R_0.., V_0..], still WITHOUT the weight and in Montgomery form
    if (lane == 0) { st_sc(out, sc_neg(e)); st_sc(out + 8, sc_neg(one_m)); st_sc(out + 16, neg_e2); }
    if (lane < R) { st_sc(out + 8 * (3 + lane), mm(neg_e2, ej2)); st_sc(out + 8 * (3 + R + lane), mm(neg_e2, ejinv2)); }
    // y_sum = y * prod_k (1 + y^(2^k))
    sc f = lane < R ? sc_add(one_m, ypow) : one_m;
    for (int mk = 16; mk > 0; mk >>= 1) f = mm(f, shfl_xor_sc(f, mk));
    const sc y_sum = mm(f, y);
    // z^(2(j+1)) for the commitments: lane l starts at z2^(l+1) (scan of a constant), strides by z2^32
    sc zl = z2;
    for (int dd = 1; dd < 32; dd <<= 1) { const sc t = shfl_up_sc(zl, dd); const sc zn = mm(zl, t); zl = sc_pick(zl, zn, lane >= dd); }
    const sc z32 = shfl_sc(zl, 31);
    const sc neg_e2_yN1 = mm(neg_e2, yN1);
    sc h = sc_zero(), zsum = sc_zero();
    for (uint32_t j = (uint32_t)lane; j < m; j += 32) {
        st_sc(pv + 8 * (PV_HDR + 3 * R + j), zl);
        const sc weighted = mm(neg_e2_yN1, zl);
        st_sc(out + 8 * (3 + 2 * R + j), weighted);
        if (b.min_present[pr.commit_off + j]) h = sc_sub(h, mm(weighted, sc_to_mont(sc_from_u64(b.min_values[pr.commit_off + j]))));
        zsum = sc_add(zsum, zl);
        zl = mm(zl, z32);
    }
    h = warp_sum_sc(h);
    zsum = warp_sum_sc(zsum);
    const sc r1 = ld_raw_mont(raw + BPP_RAW_R1(ext)), s1 = ld_raw_mont(raw + BPP_RAW_S1(ext));
    if (lane == 0) {
        const uint64_t two_n_m1 = d.bit_length >= 64 ? ~0ull : ((1ull << d.bit_length) - 1ull);
        const sc d_sum = mm(zsum, sc_to_mont(sc_from_u64(two_n_m1)));
        // h += r1*y*s1 + e^2*(y^(N+1)*z*d_sum + (z^2 - z)*y_sum)   (range_proof.rs:1017-1020, weight applied later)
        const sc t = mm(mm(r1, y), s1);
        const sc u = sc_add(mm(mm(yN1, z), d_sum), mm(sc_sub(z2, z), y_sum));
        uint32_t *hg = b.hg_contrib + 8 * (size_t)p * (1 + ext);
        st_sc(hg, sc_add(h, sc_add(t, mm(e2, u))));
        // hand-off to stage B (layout above)
        st_sc(pv, mm(mm(e, r1), s0));
        st_sc(pv + 8, mm(mm(e, s1), prod_e));
        st_sc(pv + 16, mm(e2, yN));
        st_sc(pv + 24, mm(e2, z));
    }
    if ((uint32_t)lane < ext) st_sc(b.hg_contrib + 8 * ((size_t)p * (1 + ext) + 1 + lane), ld_raw_mont(raw + BPP_RAW_D1(ext, lane)));
    // per bit k of i (challenge e_j with j = R-1-k): ru_k, rv_k, rq_k; lane k needs y^-(2^k)
    sc yp = y_inv;
    for (int it = 0; it < R - 1; it++) { const sc sq = mm(yp, yp); yp = sc_pick(yp, sq, lane > it); }
    const int jsrc = R - 1 - lane;
    const sc ej2_r = shfl_sc(ej2, jsrc & 31), ejinv2_r = shfl_sc(ejinv2, jsrc & 31);
    if (lane < R) {
        st_sc(pv + 8 * (PV_HDR + lane), mm(ej2_r, yp));
        st_sc(pv + 8 * (PV_HDR + R + lane), ejinv2_r);
        sc rq = yp;
        if ((1u << lane) < d.bit_length) rq = mm(rq, sc_to_mont(sc_from_u64(1ull << (1u << lane))));     // bit_length <= 64: lane <= 5
        st_sc(pv + 8 * (PV_HDR + 2 * R + lane), rq);
    }
}

// Stage B.  Table path (rounds <= VEC_TABLE_MAX_ROUNDS): i = (hi << rl) | lo, tables T_lo[lo] = base * prod_{k in bits(lo)} r_k and
// T_hi[hi] = prod_{k in bits(hi)} r_(rl+k) for each of the three products (u, v, q), built by doubling (T[x + 2^k] = T[x] * r_k);
// an element then costs 4 multiplications (u, v, q, q * z^(2(party+1))).  Shared memory: 3 * (2^rl + 2^rh) scalars.
#define VEC_TABLE_MAX_ROUNDS 15
template <bool TABLES> __global__ void __launch_bounds__(256) k_vprep_vector(VDims d, VBuffers b) {
    const uint32_t p = blockIdx.x;
    const VProof pr = b.proofs[p];
    if (!pr.active) return;                               // CTA-uniform
    const uint32_t R = pr.rounds, N = 1u << R;
    const uint32_t *pv = b.pervec + 8 * (size_t)pr.pv_off;
    const sc we2z = ld_sc(pv + 24);
    uint32_t *c = b.contrib + 8 * (size_t)pr.contrib_off;
    const uint32_t lgn = 31 - __clz(d.bit_length);        // bit_length is a power of two
    if (TABLES) {
        extern __shared__ uint32_t vec_sm[];
        const uint32_t rl = R / 2, rh = R - rl, nlo = 1u << rl, nhi = 1u << rh, per = nlo + nhi;
        // table t in {u, v, q}: lo entries at [t*per, t*per + nlo), hi entries behind them
        if (threadIdx.x < 3) {
            st_sc(vec_sm + 8 * (threadIdx.x * per), ld_sc(pv + 8 * threadIdx.x));
            st_sc(vec_sm + 8 * (threadIdx.x * per + nlo), sc_const_R());
        }
        __syncthreads();
        for (uint32_t k = 0; k < rh; k++) {                // rh >= rl
            const uint32_t span = 1u << k, items = 6 * span;
            for (uint32_t it = threadIdx.x; it < items; it += blockDim.x) {
                const uint32_t t = it / (2 * span), r = it % (2 * span), half = r / span, x = r % span;
                if (half == 0 && k >= rl) continue;
                const uint32_t bitpos = half ? rl + k : k;
                uint32_t *T = vec_sm + 8 * (t * per + (half ? nlo : 0));
                st_sc(T + 8 * (x + span), mm(ld_sc(T + 8 * x), ld_sc(pv + 8 * (PV_HDR + t * R + bitpos))));
            }
            __syncthreads();
        }
        for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) {
            const uint32_t lo = i & (nlo - 1), hi = i >> rl;
            const sc u = mm(ld_sc(vec_sm + 8 * (nlo + hi)), ld_sc(vec_sm + 8 * lo));
            const sc v = mm(ld_sc(vec_sm + 8 * (per + nlo + hi)), ld_sc(vec_sm + 8 * (per + lo)));
            const sc q = mm(ld_sc(vec_sm + 8 * (2 * per + nlo + hi)), ld_sc(vec_sm + 8 * (2 * per + lo)));
            const sc hterm = mm(q, ld_sc(pv + 8 * (PV_HDR + 3 * R + (i >> lgn))));
            // gi: r1*e*y^-i*s[i] + e^2*z;  hi: s1*e*s[N-1-i] - e^2*(d[i]*y^(N-i) + z),  d[i] = z^(2(party+1)) * 2^(i mod n)
            st_sc(c + 8 * (size_t)i, sc_add(u, we2z));
            st_sc(c + 8 * ((size_t)N + i), sc_sub(sc_sub(v, hterm), we2z));
        }
    } else {
        // very long vectors (rounds > 15): walk the set bits of i
        for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) {
            sc u = ld_sc(pv), v = ld_sc(pv + 8), q = ld_sc(pv + 16);
            for (uint32_t k = 0; k < R; k++) {
                if (!((i >> k) & 1u)) continue;
                u = mm(u, ld_sc(pv + 8 * (PV_HDR + k)));
                v = mm(v, ld_sc(pv + 8 * (PV_HDR + R + k)));
                q = mm(q, ld_sc(pv + 8 * (PV_HDR + 2 * R + k)));
            }
            const sc hterm = mm(q, ld_sc(pv + 8 * (PV_HDR + 3 * R + (i >> lgn))));
            st_sc(c + 8 * (size_t)i, sc_add(u, we2z));
            st_sc(c + 8 * ((size_t)N + i), sc_sub(sc_sub(v, hterm), we2z));
        }
    }
}

// one WARP per (chunk, static slot), slot in [0, 2*max_mn + ext + 1): lane l adds up proofs l, l+32, .. of the chunk, then a
// shuffle tree of modular additions (a single thread walking 256 proofs was a 150 us latency chain)
static __device__ __forceinline__ sc shfl_down_sc(const sc &a, int delta) {
    sc r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_down_sync(0xffffffffu, a.v[i], delta);
    return r;
}
__global__ void __launch_bounds__(128) k_vprep_reduce(VDims d, VBuffers b) {
    const uint32_t c = blockIdx.y;
    const VChunk chk = b.chunks[c];
    if (!chk.active) return;
    const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const uint32_t n_static = 2 * chk.max_mn + d.ext + 1;
    if (slot >= n_static) return;                       // warp-uniform
    sc acc = sc_zero();
    for (uint32_t p = chk.proof_lo + lane; p < chk.proof_hi; p += 32) {
        const VProof pr = b.proofs[p];
        if (!pr.active) continue;
        uint32_t N = 1u << pr.rounds;
        const uint32_t *src = nullptr;
        if (slot < chk.max_mn) { if (slot < N) src = b.contrib + 8 * ((size_t)pr.contrib_off + slot); }
        else if (slot < 2 * chk.max_mn) { uint32_t i = slot - chk.max_mn; if (i < N) src = b.contrib + 8 * ((size_t)pr.contrib_off + N + i); }
        else if (slot < 2 * chk.max_mn + d.ext) src = b.hg_contrib + 8 * ((size_t)p * (1 + d.ext) + 1 + (slot - 2 * chk.max_mn));
        else src = b.hg_contrib + 8 * ((size_t)p * (1 + d.ext));
        const sc wp = ld_sc(b.weights_mont + 8 * (size_t)p);
        // a weight that reduced to zero is what Scalar::random_not_zero would have redrawn (probability 2^-252): reported per chunk, the host
        // redoes that chunk's weight transcript with the redraw and the pass is repeated (engine_verify.cu)
        if (slot == 0 && sc_is_zero(wp)) b.weight_zero[c] = 1;
        if (src) acc = sc_add(acc, mm(ld_sc(src), wp));
    }
    for (int delta = 16; delta > 0; delta >>= 1) acc = sc_add(acc, shfl_down_sc(acc, delta));
    if (lane == 0) {
        st_sc(b.msm_scalars + 8 * ((size_t)chk.entry_off + slot), sc_from_mont(acc));
        // the chunk's static entries: Gi[0, max_mn) | Hi[0, max_mn) | G[0, ext) | H in the generator table Gi(nm) | Hi(nm) | G | H
        const uint32_t gi = slot < chk.max_mn ? slot : slot < 2 * chk.max_mn ? d.gens_nm + (slot - chk.max_mn) : 2 * d.gens_nm + (slot - 2 * chk.max_mn);
        b.msm_pidx[chk.entry_off + slot] = 0x80000000u | gi;
    }
}

// 32 proofs per CTA of four warps.  Phase 1, one THREAD per proof (warp 0): the 64 bytes the weight transcript's rng delivered -> Scalar::from_bytes_mod_order_wide
// -> Montgomery form (kept for k_vprep_reduce), times the chunk's rho with the merged check.  Phase 2, one warp per proof, eight proofs in turn: dynamic
// scalars *= w_p, leave Montgomery form, write the point indices.  (Round 1 ran one warp per proof, every lane redoing the wide
// reductions: 3.8 k warp instructions per proof, 105 us per 16-job pass with the merged check.)
__global__ void __launch_bounds__(128) k_vprep_weight(VDims d, VBuffers b) {
    __shared__ uint32_t s_w[8][32], s_entry[32], s_pt[32], s_ndyn[32];
    const uint32_t t = threadIdx.x, p = blockIdx.x * 32u + t;
    uint32_t n_dyn = 0;
    if (t < 32 && p < d.n_proofs) {
        const VProof pr = b.proofs[p];
        if (pr.active) {
            uint32_t ww[16];
            const uint4 *src = reinterpret_cast<const uint4 *>(b.weights + 16 * (size_t)p);
            uint4 q0 = src[0], q1 = src[1], q2 = src[2], q3 = src[3];
            ww[0] = q0.x; ww[1] = q0.y; ww[2] = q0.z; ww[3] = q0.w; ww[4] = q1.x; ww[5] = q1.y; ww[6] = q1.z; ww[7] = q1.w;
            ww[8] = q2.x; ww[9] = q2.y; ww[10] = q2.z; ww[11] = q2.w; ww[12] = q3.x; ww[13] = q3.y; ww[14] = q3.z; ww[15] = q3.w;
            sc w = sc_to_mont(sc_from_wide_words(ww));
            if (d.merged) {          // merged check: every term of chunk c carries rho_c as well (a zero rho is flagged like a zero weight)
                src = reinterpret_cast<const uint4 *>(b.weights + 16 * ((size_t)d.n_proofs + pr.chunk));
                q0 = src[0]; q1 = src[1]; q2 = src[2]; q3 = src[3];
                ww[0] = q0.x; ww[1] = q0.y; ww[2] = q0.z; ww[3] = q0.w; ww[4] = q1.x; ww[5] = q1.y; ww[6] = q1.z; ww[7] = q1.w;
                ww[8] = q2.x; ww[9] = q2.y; ww[10] = q2.z; ww[11] = q2.w; ww[12] = q3.x; ww[13] = q3.y; ww[14] = q3.z; ww[15] = q3.w;
                const sc rho = sc_from_wide_words(ww);
                if (sc_is_zero(rho)) b.weight_zero[pr.chunk] = 1;
                else if (!sc_is_zero(w)) w = mm(w, sc_to_mont(rho));         // (a zero weight stays zero: k_vprep_reduce flags it)
            }
            st_sc(b.weights_mont + 8 * (size_t)p, w);
#pragma unroll
            for (int i = 0; i < 8; i++) s_w[i][t] = w.v[i];
            n_dyn = 3 + 2 * pr.rounds + pr.m;
            s_entry[t] = pr.entry_off; s_pt[t] = pr.pt_off;
        }
    }
    if (t < 32) s_ndyn[t] = n_dyn;
    __syncthreads();
    const uint32_t warp = t >> 5, lane = t & 31;
    for (uint32_t q = warp; q < 32; q += 4) {
        const uint32_t nd = s_ndyn[q];
        if (!nd) continue;                                   // warp-uniform
        sc w;
#pragma unroll
        for (int i = 0; i < 8; i++) w.v[i] = s_w[i][q];
        const uint32_t e0 = s_entry[q], pt = s_pt[q];
        uint32_t *out = b.msm_scalars + 8 * (size_t)e0;
        for (uint32_t k = lane; k < nd; k += 32) {
            st_sc(out + 8 * k, sc_from_mont(mm(ld_sc(out + 8 * k), w)));
            // entries [A1, B, A, L.., R.., V..] against the point table [A, A1, B, L.., R.., V..]
            b.msm_pidx[e0 + k] = pt + (k == 0 ? 1u : k == 1 ? 2u : k == 2 ? 0u : k);
        }
    }
}

void launch_verify_prep(cudaStream_t s, const VDims &d, const VBuffers &b, uint32_t total_vec, uint32_t max_rounds, uint64_t *launches,
                        cudaEvent_t *marks) {
    if (d.n_proofs == 0) return;
    // thread form: 32 proofs share every warp instruction (2.6 k warp instructions per proof, the inversion included); warp form: one
    // proof per warp, ~20x the instructions per proof but a 3x shorter dependent chain -- measured on B200 it only pays for batches that
    // leave most warp slots empty anyway.  BPP_VPREP_WARP=n: warp form up to n proofs per pass (default 0: never)
    static const uint32_t warp_upto = getenv("BPP_VPREP_WARP") ? (uint32_t)atoi(getenv("BPP_VPREP_WARP")) : 0u;
    static const bool thread_form = getenv("BPP_VPREP_THREAD") != nullptr;
    if (thread_form || d.n_proofs > warp_upto) k_vprep_proof<<<(d.n_proofs + 63) / 64, 64, 0, s>>>(d, b);
    else k_vprep_proof_w<<<(d.n_proofs + 3) / 4, 128, 0, s>>>(d, b);
    if (marks) cudaEventRecord(marks[0], s);
    if (launches) (*launches)++;
    if (d.action != 0 /* RecoverOnly */ && total_vec) {
        // one CTA per proof; max_rounds bounds the table size (all proofs of a call share one generator set, so their vector
        // lengths differ by the aggregation factor only)
        // (64-element vectors: ONE warp per proof, two elements per thread -- the table doubling steps keep whole warps busy for a handful
        // of products, so the second warp of a 64-thread CTA cost more than it saved; BPP_VPREP_VEC_THREADS overrides)
        static const uint32_t forced_threads = getenv("BPP_VPREP_VEC_THREADS") ? (uint32_t)atoi(getenv("BPP_VPREP_VEC_THREADS")) : 0u;
        const uint32_t threads = forced_threads ? forced_threads : max_rounds >= 8 ? 256u : max_rounds <= 6 ? 32u : (1u << max_rounds);
        static const bool force_direct = getenv("BPP_VPREP_DIRECT") != nullptr;      // test hook for the long-vector path
        if (max_rounds <= VEC_TABLE_MAX_ROUNDS && !force_direct) {
            const uint32_t rl = max_rounds / 2, rh = max_rounds - rl;
            k_vprep_vector<true><<<d.n_proofs, threads, 3 * ((1u << rl) + (1u << rh)) * 32, s>>>(d, b);
        } else {
            k_vprep_vector<false><<<d.n_proofs, threads, 0, s>>>(d, b);
        }
        if (launches) (*launches)++;
    }
    if (marks) cudaEventRecord(marks[1], s);
}

void launch_verify_weigh(cudaStream_t s, const VDims &d, const VBuffers &b, uint32_t max_static, uint64_t *launches) {
    if (d.n_proofs == 0 || d.action == 0 || max_static == 0) return;
    k_vprep_weight<<<(d.n_proofs + 31) / 32, 128, 0, s>>>(d, b);
    dim3 grid((max_static + 3) / 4, d.n_chunks);       // 4 warps (slots) per CTA
    k_vprep_reduce<<<grid, 128, 0, s>>>(d, b);
    if (launches) (*launches) += 2;
}

} // namespace bpp
