// Quad-cooperative Edwards point arithmetic (device only): shared by k_msm.cu and the latency probes of k_bench.cu.
#pragma once
#include "arith.cuh"

namespace bpp {

static __device__ __forceinline__ void ld8(uint32_t w[8], const uint32_t *p) {
    uint4 a = reinterpret_cast<const uint4 *>(p)[0], b = reinterpret_cast<const uint4 *>(p)[1];
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}
static __device__ __forceinline__ void st8(uint32_t *p, const uint32_t w[8]) {
    reinterpret_cast<uint4 *>(p)[0] = make_uint4(w[0], w[1], w[2], w[3]);
    reinterpret_cast<uint4 *>(p)[1] = make_uint4(w[4], w[5], w[6], w[7]);
}
// ------------------------------------------------------------------------------------------------ quad-cooperative point ops
// Four consecutive lanes own ONE accumulator: lane role r = lane & 3 holds coordinate r of (X, Y, Z, T).  Every point
// operation is two warp-wide field multiplications (the four independent products of each half of the unified
// Edwards formulas run in the four lanes) instead of 7-9 sequential ones per thread: the long dependent chains of the
// MSM (bucket sums, running sums, Horner) get ~4x shorter, and a lane carries 8 registers of state instead of 32.
// All 32 lanes of a warp must call these together (full-mask shuffles); idle quads add the identity.
//
// "cached" operand of a quad addition = (Y-X, Y+X, 2Z, 2dT), one field per role; an affine-Niels table entry is the
// cached form with 2Z = 2.

static __device__ __forceinline__ fe shfl_fe(const fe &v, int src) {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(0xffffffffu, v.v[i], src);
    return r;
}
static __device__ __forceinline__ fe shfl_down_fe(const fe &v, int delta) {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_down_sync(0xffffffffu, v.v[i], delta);
    return r;
}
// branch-free 4-way select by role (bit masks: a divergent select would serialise the four lanes of every quad)
static __device__ __forceinline__ fe sel4(int role, const fe &a0, const fe &a1, const fe &a2, const fe &a3) {
    const uint32_t m0 = role == 0 ? 0xffffffffu : 0u, m1 = role == 1 ? 0xffffffffu : 0u, m2 = role == 2 ? 0xffffffffu : 0u,
                   m3 = role == 3 ? 0xffffffffu : 0u;
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = (a0.v[i] & m0) | (a1.v[i] & m1) | (a2.v[i] & m2) | (a3.v[i] & m3);
    return r;
}
static __device__ __forceinline__ fe quad_identity(int role) { return (role == 1 || role == 2) ? fe_one() : fe_zero(); }
static __device__ __forceinline__ fe quad_cached_identity(int role) {       // (1, 1, 2, 0)
    fe r = fe_zero();
    r.v[0] = role == 3 ? 0u : role == 2 ? 2u : 1u;
    return r;
}
// acc += Q, Q given as this lane's field of its cached form
static __device__ __forceinline__ fe quad_add(const fe &c, int role, int base, const fe &q) {
    fe x = shfl_fe(c, base), y = shfl_fe(c, base + 1);
    fe u = sel4(role, fe_sub_l(y, x), fe_add_l(y, x), c, c);
    fe m = fe_mul(u, q);                                                 // A, B, D, C  (tight)
    fe A = shfl_fe(m, base), B = shfl_fe(m, base + 1), D = shfl_fe(m, base + 2), C = shfl_fe(m, base + 3);
    fe E = fe_sub_l(B, A), F = fe_sub_l(D, C), G = fe_add_l(D, C), H = fe_add_l(B, A);
    return fe_mul(sel4(role, E, G, F, E), sel4(role, F, H, G, H));       // X = EF, Y = GH, Z = FG, T = EH
}
static __device__ __forceinline__ fe quad_dbl(const fe &c, int role, int base) {
    fe x = shfl_fe(c, base), y = shfl_fe(c, base + 1);
    fe s = fe_sq(sel4(role, x, y, c, fe_add_l(x, y)));                    // XX, YY, ZZ, (X+Y)^2  (tight)
    fe XX = shfl_fe(s, base), YY = shfl_fe(s, base + 1), ZZ = shfl_fe(s, base + 2), S = shfl_fe(s, base + 3);
    fe H = fe_add_l(YY, XX), G = fe_sub_l(YY, XX);                        // loose
    fe E = fe_sub_ll(S, H), F = fe_sub_ll(fe_add_l(ZZ, ZZ), G);
    return fe_mul(sel4(role, E, H, G, E), sel4(role, F, G, F, H));       // X = EF, Y = HG, Z = GF, T = EH
}
// extended coordinates -> this lane's field of the cached form (one multiplication: 2d * T)
static __device__ __forceinline__ fe quad_to_cached(const fe &c, int role, int base) {
    fe x = shfl_fe(c, base), y = shfl_fe(c, base + 1);
    fe m = fe_mul(c, role == 3 ? fe_const_2d() : fe_one());
    return sel4(role, fe_sub_l(y, x), fe_add_l(y, x), fe_add_l(c, c), m);
}
static __device__ __forceinline__ fe ld_fe(const fe *p) { fe r; ld8(r.v, p->v); return r; }
static __device__ __forceinline__ void st_fe(fe *p, const fe &r) { st8(p->v, r.v); }

} // namespace bpp
