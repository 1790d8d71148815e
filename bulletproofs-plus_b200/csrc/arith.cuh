// bpp-b200 arithmetic core: GF(2^255-19), scalars mod l, twisted-Edwards / Ristretto255 point ops.
//
// Replaces, for the hot path, what the reference obtains from curve25519-dalek 4.1.3 through its trait seam
// (/root/reference/src/traits.rs:7-43, /root/reference/src/protocols/curve_point_protocol.rs:18-36,
//  /root/reference/src/ristretto.rs:28-64).  Written from the published specifications (RFC 9496, RFC 8032).
//
// Representation (B200-first): 8 x 32-bit saturated limbs, one element per thread, everything in registers.
//   fe  : value < 2^255 (bit 255 clear), NOT necessarily < p.  Every fe_* function returns such a value and
//         accepts any such value.
//   sc  : canonical (< l), plain form.   scm : canonical, Montgomery form (x * 2^256 mod l).
// The same source compiles for the host (portable C++ path) so the logic is unit-tested on CPU
// (tests/hostcheck); on the device the multiply cores are carry-chained mad.lo.cc / madc.hi.cc PTX, which
// ptxas lowers to IMAD.WIDE / IMAD.X / IADD3.X chains on sm_100a.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BPP_HD __host__ __device__ __forceinline__
#else
#define BPP_HD inline
#endif
// fe_mul / fe_sq / sc_montmul are ~300 SASS instructions each: out-of-line by default on the device so that
// kernels with dozens of call sites stay inside the instruction cache; define BPP_INLINE_MUL in a translation
// unit whose inner loop wants them inlined.
#if defined(__CUDA_ARCH__) && !defined(BPP_INLINE_MUL)
#define BPP_MULFN static __device__ __noinline__
#elif defined(__CUDACC__)
#define BPP_MULFN static __host__ __device__ __forceinline__
#else
#define BPP_MULFN static inline
#endif

#if defined(__CUDA_ARCH__) && !defined(BPP_PORTABLE_ARITH)
#define BPP_PTX 1
#else
#define BPP_PTX 0
#endif

namespace bpp {

struct fe { uint32_t v[8]; };
struct sc { uint32_t v[8]; };
struct ge { fe X, Y, Z, T; };            // extended coordinates, x = X/Z, y = Y/Z, T = XY/Z
struct aniels { fe ypx, ymx, t2d; };     // affine Niels: (y+x, y-x, 2d*x*y); identity = (1, 1, 0)
struct cached { fe ymx, ypx, z2, t2d; }; // projective Niels ("cached"): (Y-X, Y+X, 2Z, 2dT); fields may be loose 256-bit values

// ================================================================================================ PTX helpers
#if BPP_PTX
namespace ptx {
__device__ __forceinline__ uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

// acc[j], acc[j+1] = a[j] * b for even j < N  (disjoint 64-bit columns, no carries)
template <int N> __device__ __forceinline__ void mul_n(uint32_t *acc, const uint32_t *a, uint32_t b) {
#pragma unroll
    for (int j = 0; j < N; j += 2) { acc[j] = mul_lo(a[j], b); acc[j + 1] = mul_hi(a[j], b); }
}
// acc[0..N) += a[even j] * b as one carry chain; carry-out stays in CC.CF
template <int N> __device__ __forceinline__ void cmad_n(uint32_t *acc, const uint32_t *a, uint32_t b) {
    acc[0] = mad_lo_cc(a[0], b, acc[0]);
    acc[1] = madc_hi_cc(a[0], b, acc[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) { acc[j] = madc_lo_cc(a[j], b, acc[j]); acc[j + 1] = madc_hi_cc(a[j], b, acc[j + 1]); }
}
// one row of the even/odd schoolbook: `odd` receives a[odd j]*b (N-2 existing limbs + 2 new), `even` receives
// a[even j]*b (N existing limbs); the even chain's carry-out lands in odd[N-1], which has the same weight.
template <int N> __device__ __forceinline__ void mad_row(uint32_t *odd, uint32_t *even, const uint32_t *a, uint32_t b) {
    cmad_n<N - 2>(odd, a + 1, b);
    odd[N - 2] = madc_lo_cc(a[N - 1], b, 0);
    odd[N - 1] = madc_hi(a[N - 1], b, 0);
    cmad_n<N>(even, a, b);
    odd[N - 1] = addc(odd[N - 1], 0);
}
// t[0..16) = a * b
__device__ __forceinline__ void mul256(uint32_t t[16], const uint32_t a[8], const uint32_t b[8]) {
    uint32_t odd[14];
    mul_n<8>(t, a, b[0]);
    mul_n<8>(odd, a + 1, b[0]);
    mad_row<8>(&t[2], &odd[0], a, b[1]);
#pragma unroll
    for (int i = 2; i < 8; i += 2) {
        mad_row<8>(&odd[i], &t[i], a, b[i]);
        mad_row<8>(&t[i + 2], &odd[i], a, b[i + 1]);
    }
    t[1] = add_cc(t[1], odd[0]);
#pragma unroll
    for (int i = 1; i < 14; i++) t[i + 1] = addc_cc(t[i + 1], odd[i]);
    t[15] = addc(t[15], 0);
}
// t[0..8) = a * b for 4-limb operands: the even / odd row scheme of mul256 on half the width (16 multiplies)
__device__ __forceinline__ void mul128(uint32_t t[8], const uint32_t a[4], const uint32_t b[4]) {
    uint32_t odd[6];
    mul_n<4>(t, a, b[0]);
    mul_n<4>(odd, a + 1, b[0]);
    mad_row<4>(&t[2], &odd[0], a, b[1]);
    mad_row<4>(&odd[2], &t[2], a, b[2]);
    mad_row<4>(&t[4], &odd[2], a, b[3]);
    t[1] = add_cc(t[1], odd[0]);
#pragma unroll
    for (int i = 1; i < 6; i++) t[i + 1] = addc_cc(t[i + 1], odd[i]);
    t[7] = addc(t[7], 0);
}
// t[0..16) = a * b with one level of Karatsuba: 48 multiplies instead of 64 and ~45 more additions.  The multiplier pipe is what
// bounds the field kernels (ncu: math-pipe throttle is their top stall; a 32x32->64 product issues about once per 5.5 cycles per
// scheduler whether it is one IMAD.WIDE or a lo / hi pair, bpp_microbench 2 and 12) while the ALU pipe and the issue slots are half
// idle, so trading 16 products for additions looked like a win.  MEASURED, NOT A WIN (kept behind -DBPP_KARATSUBA, bit-exact, all parity
// tests pass with it): ptxas places the extra carry logic on the same FMA pipe (33 IMAD.MOV + 12 IMAD.X per multiplication, two issue
// cycles each, next to the 16 saved IMAD.WIDE at ~5.6) and the longer live ranges spill under the occupancy cap of the bucket kernel:
// k_msm_bucket_thread 0.544 -> 0.680 ms, k_msm_reduce_warp 0.199 -> 0.238 ms, 9.8 -> 8.7 M proofs/s.
//   a = a0 + a1 W, b = b0 + b1 W (W = 2^128):  z0 = a0 b0, z2 = a1 b1, z1 = (a0 + a1)(b0 + b1) - z0 - z2,  a b = z0 + z1 W + z2 W^2
// The half sums are 129 bits: with sa, sb their low 128 bits and ca, cb the carries, (a0 + a1)(b0 + b1) = sa sb + (ca sb + cb sa) W
// + ca cb W^2, a 258-bit value kept as m[0..7] and m8 <= 3.
__device__ __forceinline__ void mul256_k(uint32_t t[16], const uint32_t a[8], const uint32_t b[8]) {
    uint32_t m[8], sa[4], sb[4];
    mul128(t, a, b);
    mul128(t + 8, a + 4, b + 4);
    sa[0] = add_cc(a[0], a[4]); sa[1] = addc_cc(a[1], a[5]); sa[2] = addc_cc(a[2], a[6]); sa[3] = addc_cc(a[3], a[7]);
    const uint32_t ca = addc(0, 0);
    sb[0] = add_cc(b[0], b[4]); sb[1] = addc_cc(b[1], b[5]); sb[2] = addc_cc(b[2], b[6]); sb[3] = addc_cc(b[3], b[7]);
    const uint32_t cb = addc(0, 0);
    mul128(m, sa, sb);
    const uint32_t ma = 0u - ca, mb = 0u - cb;
    m[4] = add_cc(m[4], sb[0] & ma); m[5] = addc_cc(m[5], sb[1] & ma); m[6] = addc_cc(m[6], sb[2] & ma); m[7] = addc_cc(m[7], sb[3] & ma);
    uint32_t m8 = addc(ca & cb, 0);
    m[4] = add_cc(m[4], sa[0] & mb); m[5] = addc_cc(m[5], sa[1] & mb); m[6] = addc_cc(m[6], sa[2] & mb); m[7] = addc_cc(m[7], sa[3] & mb);
    m8 = addc(m8, 0);
    m[0] = sub_cc(m[0], t[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) m[i] = subc_cc(m[i], t[i]);
    m8 = subc(m8, 0);
    m[0] = sub_cc(m[0], t[8]);
#pragma unroll
    for (int i = 1; i < 8; i++) m[i] = subc_cc(m[i], t[8 + i]);
    m8 = subc(m8, 0);
    t[4] = add_cc(t[4], m[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) t[4 + i] = addc_cc(t[4 + i], m[i]);
    t[12] = addc_cc(t[12], m8);
    t[13] = addc_cc(t[13], 0);
    t[14] = addc_cc(t[14], 0);
    t[15] = addc(t[15], 0);
}
// t[0..16) = a^2 with 36 multiplies: the 28 off-diagonal products a_i * a_j (i < j) once, doubled by a funnel shift, plus the 8
// diagonal squares.  Row i (multiplier a_i, elements a_{i+1..7}, first column 2i + 1) is split like the rows of mul256 into its
// elements at even and at odd distance: each half is one pure carry chain (low half of element k + 2 lands right after the high
// half of element k).  Chains that start on an odd column accumulate into x[], chains that start on an even column into y[]; a
// chain's carry-out goes into the limb after its end, which at that point holds nothing but earlier carry-outs (checked per row in
// targets x[9], x[9], x[11], x[11], x[13], x[13], x[15] and y[8], y[10], y[10], y[12], y[12], y[14]).
__device__ __forceinline__ void sq256(uint32_t t[16], const uint32_t a[8]) {
    uint32_t x[16], y[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { x[i] = 0; y[i] = 0; }
    // row 0: elements a1 a3 a5 a7 -> x[1..8];  a2 a4 a6 -> y[2..7]
    cmad_n<8>(x + 1, a + 1, a[0]);  x[9] = addc(x[9], 0);
    cmad_n<6>(y + 2, a + 2, a[0]);  y[8] = addc(y[8], 0);
    // row 1: a2 a4 a6 -> x[3..8];  a3 a5 a7 -> y[4..9]
    cmad_n<6>(x + 3, a + 2, a[1]);  x[9] = addc(x[9], 0);
    cmad_n<6>(y + 4, a + 3, a[1]);  y[10] = addc(y[10], 0);
    // row 2: a3 a5 a7 -> x[5..10];  a4 a6 -> y[6..9]
    cmad_n<6>(x + 5, a + 3, a[2]);  x[11] = addc(x[11], 0);
    cmad_n<4>(y + 6, a + 4, a[2]);  y[10] = addc(y[10], 0);
    // row 3: a4 a6 -> x[7..10];  a5 a7 -> y[8..11]
    cmad_n<4>(x + 7, a + 4, a[3]);  x[11] = addc(x[11], 0);
    cmad_n<4>(y + 8, a + 5, a[3]);  y[12] = addc(y[12], 0);
    // row 4: a5 a7 -> x[9..12];  a6 -> y[10..11]
    cmad_n<4>(x + 9, a + 5, a[4]);  x[13] = addc(x[13], 0);
    cmad_n<2>(y + 10, a + 6, a[4]); y[12] = addc(y[12], 0);
    // row 5: a6 -> x[11..12];  a7 -> y[12..13]
    cmad_n<2>(x + 11, a + 6, a[5]); x[13] = addc(x[13], 0);
    cmad_n<2>(y + 12, a + 7, a[5]); y[14] = addc(y[14], 0);
    // row 6: a7 -> x[13..14]
    cmad_n<2>(x + 13, a + 7, a[6]); x[15] = addc(x[15], 0);
    // u = x + y (columns 1..15)
    x[1] = add_cc(x[1], y[1]);
#pragma unroll
    for (int i = 2; i < 15; i++) x[i] = addc_cc(x[i], y[i]);
    x[15] = addc(x[15], y[15]);
    // t = 2u + diagonal
    uint32_t v[16];
    v[0] = 0;
#pragma unroll
    for (int i = 1; i < 16; i++) v[i] = __funnelshift_l(x[i - 1], x[i], 1);
    t[0] = add_cc(v[0], mul_lo(a[0], a[0]));
    t[1] = addc_cc(v[1], mul_hi(a[0], a[0]));
#pragma unroll
    for (int i = 1; i < 8; i++) { t[2 * i] = addc_cc(v[2 * i], mul_lo(a[i], a[i])); t[2 * i + 1] = (i == 7) ? addc(v[15], mul_hi(a[7], a[7])) : addc_cc(v[2 * i + 1], mul_hi(a[i], a[i])); }
}
} // namespace ptx
#endif

// portable 8x8 schoolbook
BPP_HD void mul256_portable(uint32_t t[16], const uint32_t a[8], const uint32_t b[8]) {
#pragma unroll
    for (int i = 0; i < 16; i++) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            c += (uint64_t)a[i] * b[j] + t[i + j];
            t[i + j] = (uint32_t)c;
            c >>= 32;
        }
        t[i + 8] = (uint32_t)c;
    }
}

// t[0..16) = a^2: the 28 off-diagonal products once, doubled, plus the 8 diagonal squares (36 multiplies instead of 64)
BPP_HD void sq256_any(uint32_t t[16], const uint32_t a[8]) {
    uint32_t u[16];
#pragma unroll
    for (int i = 0; i < 16; i++) u[i] = 0;
#pragma unroll
    for (int i = 0; i < 7; i++) {
        uint64_t c = 0;
#pragma unroll
        for (int j = i + 1; j < 8; j++) {
            c += (uint64_t)a[i] * a[j] + u[i + j];
            u[i + j] = (uint32_t)c;
            c >>= 32;
        }
        u[i + 8] = (uint32_t)c;
    }
    uint64_t c = 0;
    uint32_t carry_bit = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)a[i] * a[i];
        uint32_t lo2 = (u[2 * i] << 1) | carry_bit;
        carry_bit = u[2 * i] >> 31;
        uint32_t hi2 = (u[2 * i + 1] << 1) | carry_bit;
        carry_bit = u[2 * i + 1] >> 31;
        c += (uint64_t)lo2 + (uint32_t)d;
        t[2 * i] = (uint32_t)c;
        c >>= 32;
        c += (uint64_t)hi2 + (uint32_t)(d >> 32);
        t[2 * i + 1] = (uint32_t)c;
        c >>= 32;
    }
}

// Measured on B200 with by-value operands (gpurun r01b): hand-chained mad.cc core 104 G mul/s, compiler-scheduled
// portable schoolbook 81 G mul/s -> the PTX core is the default; BPP_PORTABLE_MUL selects the other one.
BPP_HD void mul256_any(uint32_t t[16], const uint32_t a[8], const uint32_t b[8]) {
#if BPP_PTX && !defined(BPP_PORTABLE_MUL) && defined(BPP_KARATSUBA)
    ptx::mul256_k(t, a, b);
#elif BPP_PTX && !defined(BPP_PORTABLE_MUL)
    ptx::mul256(t, a, b);
#else
    mul256_portable(t, a, b);
#endif
}

// ================================================================================================ field
BPP_HD fe fe_zero() { fe r; for (int i = 0; i < 8; i++) r.v[i] = 0; return r; }
BPP_HD fe fe_one() { fe r = fe_zero(); r.v[0] = 1; return r; }
BPP_HD fe fe_from_u32(uint32_t x) { fe r = fe_zero(); r.v[0] = x; return r; }

#define BPP_FE(a0, a1, a2, a3, a4, a5, a6, a7) fe{{a0, a1, a2, a3, a4, a5, a6, a7}}
BPP_HD fe fe_const_d() { return BPP_FE(0x135978a3u, 0x75eb4dcau, 0x4141d8abu, 0x00700a4du, 0x7779e898u, 0x8cc74079u, 0x2b6ffe73u, 0x52036ceeu); }
BPP_HD fe fe_const_inv_d() { return BPP_FE(0xcdc9f843u, 0x25e0f276u, 0x4279542eu, 0x0b5dd698u, 0xcdb9cf66u, 0x2b162114u, 0x14d5ce43u, 0x40907ed2u); }          // 1/d
BPP_HD fe fe_const_neg_inv_d() { return BPP_FE(0x323607aau, 0xda1f0d89u, 0xbd86abd1u, 0xf4a22967u, 0x32463099u, 0xd4e9deebu, 0xeb2a31bcu, 0x3f6f812du); }     // -1/d
BPP_HD fe fe_const_2d() { return BPP_FE(0x26b2f159u, 0xebd69b94u, 0x8283b156u, 0x00e0149au, 0xeef3d130u, 0x198e80f2u, 0x56dffce7u, 0x2406d9dcu); }
BPP_HD fe fe_const_sqrtm1() { return BPP_FE(0x4a0ea0b0u, 0xc4ee1b27u, 0xad2fe478u, 0x2f431806u, 0x3dfbd7a7u, 0x2b4d0099u, 0x4fc1df0bu, 0x2b832480u); }
BPP_HD fe fe_const_invsqrt_a_minus_d() { return BPP_FE(0x805d40eau, 0x99c8fdaau, 0x5a4172beu, 0x9d2f1617u, 0xfe01d840u, 0x16c27b91u, 0xcfaffca2u, 0x786c8905u); }
BPP_HD fe fe_const_sqrt_ad_minus_one() { return BPP_FE(0x497b2e1bu, 0x7e97f6a0u, 0x1b7854bdu, 0xaf9d8e0cu, 0x31f5d1fdu, 0x0f3cfcc9u, 0x2b8348acu, 0x376931bfu); }
BPP_HD fe fe_const_one_minus_d_sq() { return BPP_FE(0x945fc176u, 0xe27c09c1u, 0xcd5e350fu, 0x2c81a138u, 0xbe70dfe4u, 0x9994abddu, 0xb2b3e0d7u, 0x029072a8u); }
BPP_HD fe fe_const_d_minus_one_sq() { return BPP_FE(0x44ed4d20u, 0x31ad5aaau, 0xb01e1999u, 0xd29e4a2cu, 0x529b4eebu, 0x4cdcd32fu, 0xf66c2241u, 0x5968b37au); }

// r (8 limbs) + 2^256 * hi  ->  < 2^255.   hi < 2^26.
BPP_HD void fe_fold(fe &r, uint32_t hi) {
    uint32_t top = (hi << 1) | (r.v[7] >> 31);
    r.v[7] &= 0x7fffffffu;
    uint32_t add = top * 19u;
#if BPP_PTX
    r.v[0] = ptx::add_cc(r.v[0], add);
#pragma unroll
    for (int i = 1; i < 7; i++) r.v[i] = ptx::addc_cc(r.v[i], 0);
    r.v[7] = ptx::addc(r.v[7], 0);
#else
    uint64_t c = add;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += r.v[i]; r.v[i] = (uint32_t)c; c >>= 32; }
#endif
    // now < 2^255 + 2^32; bit 255 can only be set if the low 255 bits are < 2^32 - so one more tiny fold
    top = r.v[7] >> 31;
    r.v[7] &= 0x7fffffffu;
    r.v[0] += 19u * top;
}

BPP_HD fe fe_add(const fe &a, const fe &b) {
    fe r;
#if BPP_PTX
    r.v[0] = ptx::add_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) r.v[i] = ptx::addc_cc(a.v[i], b.v[i]);
    r.v[7] = ptx::addc(a.v[7], b.v[7]);
#else
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (uint64_t)a.v[i] + b.v[i]; r.v[i] = (uint32_t)c; c >>= 32; }
#endif
    fe_fold(r, 0); // a + b < 2^256
    return r;
}

// a - b  computed as  a + (2^256 - 38 - b)  =  a - b + 2p, 9 limbs, then folded
BPP_HD fe fe_sub(const fe &a, const fe &b) {
    fe r;
    uint32_t hi;
#if BPP_PTX
    // r = a - b - 38 + 2^256 : two's complement arithmetic, the final borrow tells whether the 2^256 was consumed
    r.v[0] = ptx::sub_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) r.v[i] = ptx::subc_cc(a.v[i], b.v[i]);
    uint32_t bw1 = ptx::subc(0, 0);           // 0 or 0xffffffff
    r.v[0] = ptx::sub_cc(r.v[0], 38u);
#pragma unroll
    for (int i = 1; i < 8; i++) r.v[i] = ptx::subc_cc(r.v[i], 0);
    uint32_t bw2 = ptx::subc(0, 0);
    hi = 1u + bw1 + bw2;                       // 2^256 minus the borrows: 1, 0 (never -1: a - b - 38 > -2^256)
#else
    int64_t c = -38;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (int64_t)a.v[i] - (int64_t)b.v[i]; r.v[i] = (uint32_t)c; c >>= 32; }
    hi = (uint32_t)(1 + c);
#endif
    fe_fold(r, hi);
    return r;
}

BPP_HD fe fe_neg(const fe &a) { return fe_sub(fe_zero(), a); }

// ------------------------------------------------------------------------------------------------ lazy add / sub
// "loose" values: any 256-bit integer (not necessarily < 2^255).  fe_mul / fe_sq accept loose operands (the 512-bit
// product is reduced whatever its size) and return tight ones (< 2^255), so the additions between two multiplications of
// the point formulas need no reduction at all:
//   fe_add_l(tight, tight) -> loose      one 8-limb carry chain
//   fe_sub_l(loose a, TIGHT b) -> loose  a - b; if that borrowed, the wrapped value is a - b + 2^256 = a - b + 38 (mod p) and
//                                        is >= 2^256 - 2^255 > 38, so 38 is simply taken off again (two carry chains)
//   fe_sub_ll(loose a, loose b) -> loose same, but b - a may exceed 2^256 - 38, so taking 38 off can wrap once more
//                                        (three carry chains; the third cannot borrow)
BPP_HD fe fe_add_l(const fe &a, const fe &b) {
    fe r;
#if BPP_PTX
    r.v[0] = ptx::add_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) r.v[i] = ptx::addc_cc(a.v[i], b.v[i]);
    r.v[7] = ptx::addc(a.v[7], b.v[7]);
#else
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (uint64_t)a.v[i] + b.v[i]; r.v[i] = (uint32_t)c; c >>= 32; }
#endif
    return r;
}
BPP_HD fe fe_sub_l(const fe &a, const fe &b) {
    fe r;
#if BPP_PTX
    r.v[0] = ptx::sub_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) r.v[i] = ptx::subc_cc(a.v[i], b.v[i]);
    uint32_t bw = ptx::subc(0, 0);                 // 0 or 0xffffffff
    r.v[0] = ptx::sub_cc(r.v[0], 38u & bw);
#pragma unroll
    for (int i = 1; i < 7; i++) r.v[i] = ptx::subc_cc(r.v[i], 0);
    r.v[7] = ptx::subc(r.v[7], 0);
#else
    int64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (int64_t)a.v[i] - (int64_t)b.v[i]; r.v[i] = (uint32_t)c; c >>= 32; }
    int64_t d = c < 0 ? -38 : 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { d += (int64_t)r.v[i]; r.v[i] = (uint32_t)d; d >>= 32; }
#endif
    return r;
}
BPP_HD fe fe_sub_ll(const fe &a, const fe &b) {
    fe r;
#if BPP_PTX
    r.v[0] = ptx::sub_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) r.v[i] = ptx::subc_cc(a.v[i], b.v[i]);
    uint32_t bw = ptx::subc(0, 0);
    r.v[0] = ptx::sub_cc(r.v[0], 38u & bw);
#pragma unroll
    for (int i = 1; i < 8; i++) r.v[i] = ptx::subc_cc(r.v[i], 0);
    bw = ptx::subc(0, 0);
    r.v[0] = ptx::sub_cc(r.v[0], 38u & bw);
#pragma unroll
    for (int i = 1; i < 7; i++) r.v[i] = ptx::subc_cc(r.v[i], 0);
    r.v[7] = ptx::subc(r.v[7], 0);
#else
    int64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (int64_t)a.v[i] - (int64_t)b.v[i]; r.v[i] = (uint32_t)c; c >>= 32; }
    for (int pass = 0; pass < 2; pass++) {
        int64_t d = c < 0 ? -38 : 0;
#pragma unroll
        for (int i = 0; i < 8; i++) { d += (int64_t)r.v[i]; r.v[i] = (uint32_t)d; d >>= 32; }
        c = d;
    }
#endif
    return r;
}

// 16-limb product -> fe:  lo + 38 * hi, then fold
BPP_HD fe fe_reduce512(const uint32_t t[16]) {
    fe r;
    uint32_t c8;
#if BPP_PTX
    uint32_t lo[8];
#pragma unroll
    for (int i = 0; i < 8; i++) lo[i] = t[i];
    const uint32_t *hi = t + 8;
    ptx::cmad_n<8>(lo, hi, 38u);                 // lo += hi[0,2,4,6] * 38
    c8 = ptx::addc(0, 0);
    ptx::cmad_n<6>(lo + 1, hi + 1, 38u);         // lo[1..6] += hi[1,3,5] * 38
    lo[7] = ptx::madc_lo_cc(hi[7], 38u, lo[7]);
    c8 = ptx::madc_hi(hi[7], 38u, c8);
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = lo[i];
#else
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (uint64_t)t[i + 8] * 38u + t[i]; r.v[i] = (uint32_t)c; c >>= 32; }
    c8 = (uint32_t)c;
#endif
    fe_fold(r, c8);
    return r;
}

// NB by-value parameters: through a __noinline__ call the CUDA ABI then keeps both operands and the result in
// registers (a `const fe &` parameter forces the caller to spill them to local memory first).
BPP_MULFN fe fe_mul(fe a, fe b) {
    uint32_t t[16];
    mul256_any(t, a.v, b.v);
    return fe_reduce512(t);
}

BPP_MULFN fe fe_sq(fe a) {
    uint32_t t[16];
#if BPP_PTX && !defined(BPP_PORTABLE_MUL)
    ptx::sq256(t, a.v);
#else
    sq256_any(t, a.v);
#endif
    return fe_reduce512(t);
}

BPP_HD fe fe_sqn(fe a, int n) {
    for (int i = 0; i < n; i++) a = fe_sq(a);
    return a;
}

// canonical representative (< p)
BPP_HD fe fe_canon(const fe &a) {
    // a < 2^255.  a >= p  <=>  a + 19 >= 2^255
    fe t;
    uint64_t c = 19;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += a.v[i]; t.v[i] = (uint32_t)c; c >>= 32; }
    uint32_t ge_p = t.v[7] >> 31;          // 1 if a >= p; then a - p = (a + 19) - 2^255 = t with bit 255 cleared
    t.v[7] &= 0x7fffffffu;
    fe r;
    uint32_t m = 0u - ge_p;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = (t.v[i] & m) | (a.v[i] & ~m);
    return r;
}

BPP_HD bool fe_is_zero(const fe &a) {
    fe c = fe_canon(a);
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= c.v[i];
    return o == 0;
}
BPP_HD bool fe_is_negative(const fe &a) { return (fe_canon(a).v[0] & 1u) != 0; }
BPP_HD bool fe_eq(const fe &a, const fe &b) { return fe_is_zero(fe_sub(a, b)); }
BPP_HD fe fe_select(const fe &a, const fe &b, bool pick_b) {
    fe r;
    uint32_t m = pick_b ? 0xffffffffu : 0u;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = (b.v[i] & m) | (a.v[i] & ~m);
    return r;
}
BPP_HD fe fe_cneg(const fe &a, bool neg) { return fe_select(a, fe_neg(a), neg); }
BPP_HD fe fe_abs(const fe &a) { return fe_cneg(a, fe_is_negative(a)); }

// 32 little-endian bytes -> fe; bit 255 is dropped.  *canonical = bytes encode a value < p with bit 255 clear
BPP_HD fe fe_frombytes(const uint8_t *s, bool *canonical) {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++)
        r.v[i] = (uint32_t)s[4 * i] | ((uint32_t)s[4 * i + 1] << 8) | ((uint32_t)s[4 * i + 2] << 16) | ((uint32_t)s[4 * i + 3] << 24);
    bool hibit = (r.v[7] >> 31) != 0;
    r.v[7] &= 0x7fffffffu;
    if (canonical) {
        fe c = fe_canon(r);
        uint32_t diff = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) diff |= c.v[i] ^ r.v[i];
        *canonical = !hibit && diff == 0;
    }
    return r;
}
BPP_HD fe fe_fromwords(const uint32_t *w, bool *canonical) {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = w[i];
    bool hibit = (r.v[7] >> 31) != 0;
    r.v[7] &= 0x7fffffffu;
    if (canonical) {
        fe c = fe_canon(r);
        uint32_t diff = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) diff |= c.v[i] ^ r.v[i];
        *canonical = !hibit && diff == 0;
    }
    return r;
}
BPP_HD void fe_tobytes(uint8_t *s, const fe &a) {
    fe c = fe_canon(a);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        s[4 * i] = (uint8_t)c.v[i]; s[4 * i + 1] = (uint8_t)(c.v[i] >> 8);
        s[4 * i + 2] = (uint8_t)(c.v[i] >> 16); s[4 * i + 3] = (uint8_t)(c.v[i] >> 24);
    }
}

// z^(2^252 - 3) = z^((p-5)/8)
BPP_HD fe fe_pow22523(const fe &z) {
    fe t0 = fe_sq(z);                         // 2
    fe t1 = fe_mul(z, fe_sqn(t0, 2));         // 9
    fe t2 = fe_mul(t0, t1);                   // 11
    fe t4 = fe_mul(t1, fe_sq(t2));            // 31 = 2^5 - 1
    fe t5 = fe_mul(fe_sqn(t4, 5), t4);        // 2^10 - 1
    fe t6 = fe_mul(fe_sqn(t5, 10), t5);       // 2^20 - 1
    fe t7 = fe_mul(fe_sqn(t6, 20), t6);       // 2^40 - 1
    fe t9 = fe_mul(fe_sqn(t7, 10), t5);       // 2^50 - 1
    fe t11 = fe_mul(fe_sqn(t9, 50), t9);      // 2^100 - 1
    fe t13 = fe_mul(fe_sqn(t11, 100), t11);   // 2^200 - 1
    fe t15 = fe_mul(fe_sqn(t13, 50), t9);     // 2^250 - 1
    return fe_mul(fe_sqn(t15, 2), z);         // 2^252 - 3
}

BPP_HD fe fe_invert(const fe &z) {
    // z^(p-2) = z^(2^255 - 21) = (z^(2^252-3))^8 * z^3
    fe t = fe_sqn(fe_pow22523(z), 3);         // 2^255 - 24
    return fe_mul(t, fe_mul(fe_sq(z), z));
}

// RFC 9496 SQRT_RATIO_M1(u, v): returns was_square, r = |sqrt(u/v)| or |sqrt(i*u/v)|
BPP_HD bool fe_sqrt_ratio_i(fe &out, const fe &u, const fe &v) {
    fe v3 = fe_mul(fe_sq(v), v);
    fe v7 = fe_mul(fe_sq(v3), v);
    fe r = fe_mul(fe_mul(u, v3), fe_pow22523(fe_mul(u, v7)));
    fe check = fe_mul(v, fe_sq(r));
    fe neg_u = fe_neg(u);
    fe neg_u_i = fe_mul(neg_u, fe_const_sqrtm1());
    bool correct = fe_eq(check, u);
    bool flipped = fe_eq(check, neg_u);
    bool flipped_i = fe_eq(check, neg_u_i);
    fe r_i = fe_mul(r, fe_const_sqrtm1());
    r = fe_select(r, r_i, flipped || flipped_i);
    out = fe_abs(r);
    return correct || flipped;
}

// 1/sqrt(v) specialisation (u = 1): saves two multiplications
BPP_HD bool fe_invsqrt(fe &out, const fe &v) {
    fe v3 = fe_mul(fe_sq(v), v);
    fe v7 = fe_mul(fe_sq(v3), v);
    fe r = fe_mul(v3, fe_pow22523(v7));
    fe check = fe_mul(v, fe_sq(r));
    fe one = fe_one();
    fe m1 = fe_neg(one);
    fe mi = fe_neg(fe_const_sqrtm1());
    bool correct = fe_eq(check, one);
    bool flipped = fe_eq(check, m1);
    bool flipped_i = fe_eq(check, mi);
    fe r_i = fe_mul(r, fe_const_sqrtm1());
    r = fe_select(r, r_i, flipped || flipped_i);
    out = fe_abs(r);
    return correct || flipped;
}

// ================================================================================================ scalars mod l
// l = 2^252 + 27742317777372353535851937790883648493
BPP_HD uint32_t sc_l(int i) {
    const uint32_t L[8] = {0x5cf5d3edu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu, 0u, 0u, 0u, 0x10000000u};
    return L[i];
}
#define BPP_SC(a0, a1, a2, a3, a4, a5, a6, a7) sc{{a0, a1, a2, a3, a4, a5, a6, a7}}
BPP_HD sc sc_const_R() { return BPP_SC(0x8d98951du, 0xd6ec3174u, 0x737dcf70u, 0xc6ef5bf4u, 0xfffffffeu, 0xffffffffu, 0xffffffffu, 0x0fffffffu); }
BPP_HD sc sc_const_RR() { return BPP_SC(0x449c0f01u, 0xa40611e3u, 0x68859347u, 0xd00e1ba7u, 0x17f5be65u, 0xceec73d2u, 0x7c309a3du, 0x0399411bu); }
#define BPP_SC_LFACTOR 0x12547e1bu   // -l^-1 mod 2^32

BPP_HD sc sc_zero() { sc r; for (int i = 0; i < 8; i++) r.v[i] = 0; return r; }
BPP_HD sc sc_one() { sc r = sc_zero(); r.v[0] = 1; return r; }
BPP_HD sc sc_from_u64(uint64_t x) { sc r = sc_zero(); r.v[0] = (uint32_t)x; r.v[1] = (uint32_t)(x >> 32); return r; }
BPP_HD bool sc_is_zero(const sc &a) { uint32_t o = 0; for (int i = 0; i < 8; i++) o |= a.v[i]; return o == 0; }
BPP_HD bool sc_eq(const sc &a, const sc &b) { uint32_t o = 0; for (int i = 0; i < 8; i++) o |= a.v[i] ^ b.v[i]; return o == 0; }

// r = a - l if a >= l else a   (a < 2l)
BPP_HD sc sc_cond_sub_l(const sc &a, uint32_t extra_hi) {
    sc t;
    int64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (int64_t)a.v[i] - (int64_t)sc_l(i); t.v[i] = (uint32_t)c; c >>= 32; }
    c += extra_hi;
    bool ge = c >= 0;
    sc r;
    uint32_t m = ge ? 0xffffffffu : 0u;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = (t.v[i] & m) | (a.v[i] & ~m);
    return r;
}

BPP_HD sc sc_add(const sc &a, const sc &b) {
    sc r;
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (uint64_t)a.v[i] + b.v[i]; r.v[i] = (uint32_t)c; c >>= 32; }
    return sc_cond_sub_l(r, 0);
}

BPP_HD sc sc_sub(const sc &a, const sc &b) {
    sc t;
    int64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (int64_t)a.v[i] - (int64_t)b.v[i]; t.v[i] = (uint32_t)c; c >>= 32; }
    uint32_t m = c < 0 ? 0xffffffffu : 0u;
    uint64_t d = 0;
    sc r;
#pragma unroll
    for (int i = 0; i < 8; i++) { d += (uint64_t)t.v[i] + (sc_l(i) & m); r.v[i] = (uint32_t)d; d >>= 32; }
    return r;
}

BPP_HD sc sc_neg(const sc &a) { return sc_sub(sc_zero(), a); }

// Montgomery product a*b*2^-256 mod l; needs a*b < l*2^256; result canonical
// Montgomery product a * b / 2^256 mod l, result canonical.
// Device: the 512-bit product comes from the hand-chained field multiplier core (ptx::mul256, 64 products as pure mad.cc chains), then
// eight word-serial reduction steps on it: m = t[i] * (-l^-1), t += m * l * 2^(32 i).  l = l_lo (4 limbs) + 2^252, so a step is four
// products, three carry hops over l's zero limbs and the shifted m; its carry-out belongs to limb i + 8, which no later step reads
// for its m, so the eight carry-outs are parked in k[] and added once at the end.  (The interleaved word-serial form below, left to
// the compiler, was ~500 SASS instructions per product against ~170 for a field multiplication.)
BPP_MULFN sc sc_montmul(sc a, sc b) {
#if BPP_PTX && !defined(BPP_PORTABLE_MUL)
    uint32_t t[16], k[8];
    ptx::mul256(t, a.v, b.v);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t m = t[i] * BPP_SC_LFACTOR;
        uint64_t c = ((uint64_t)m * sc_l(0) + t[i]) >> 32;
#pragma unroll
        for (int j = 1; j < 4; j++) { c += (uint64_t)m * sc_l(j) + t[i + j]; t[i + j] = (uint32_t)c; c >>= 32; }
#pragma unroll
        for (int j = 4; j < 7; j++) { c += t[i + j]; t[i + j] = (uint32_t)c; c >>= 32; }       // limbs 4..6 of l are zero
        c += (uint64_t)m * 0x10000000u + t[i + 7]; t[i + 7] = (uint32_t)c; c >>= 32;            // l's top limb is 2^28
        k[i] = (uint32_t)c;                                                                     // < 2^29: weight 2^(32 (i + 8))
    }
    // r = t[8..16) + k[0..8); the total is < 2l < 2^254, so there is no carry out of limb 15
    uint32_t r9;
    sc r;
    {
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) { c += (uint64_t)t[8 + i] + k[i]; r.v[i] = (uint32_t)c; c >>= 32; }
        r9 = (uint32_t)c;
    }
    return sc_cond_sub_l(r, r9);
#else
    uint32_t t[10];
#pragma unroll
    for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) { c += (uint64_t)a.v[j] * b.v[i] + t[j]; t[j] = (uint32_t)c; c >>= 32; }
        c += t[8]; t[8] = (uint32_t)c; t[9] = (uint32_t)(c >> 32);
        uint32_t m = t[0] * BPP_SC_LFACTOR;
        c = ((uint64_t)m * sc_l(0) + t[0]) >> 32;
#pragma unroll
        for (int j = 1; j < 4; j++) { c += (uint64_t)m * sc_l(j) + t[j]; t[j - 1] = (uint32_t)c; c >>= 32; }
        // limbs 4..6 of l are zero
#pragma unroll
        for (int j = 4; j < 7; j++) { c += t[j]; t[j - 1] = (uint32_t)c; c >>= 32; }
        c += (uint64_t)m * 0x10000000u + t[7]; t[6] = (uint32_t)c; c >>= 32;
        c += t[8]; t[7] = (uint32_t)c; c >>= 32;
        t[8] = t[9] + (uint32_t)c;
        t[9] = 0;
    }
    sc r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
    return sc_cond_sub_l(r, t[8]);
#endif
}
BPP_HD sc sc_mul(const sc &a, const sc &b) { return sc_montmul(sc_montmul(a, b), sc_const_RR()); }
BPP_HD sc sc_to_mont(const sc &a) { return sc_montmul(a, sc_const_RR()); }
BPP_HD sc sc_from_mont(const sc &a) { return sc_montmul(a, sc_one()); }

// any 256-bit value -> canonical
BPP_HD sc sc_reduce256(const sc &a) { return sc_montmul(a, sc_const_R()); }
// 512-bit little-endian value (16 words) -> canonical   (Scalar::from_bytes_mod_order_wide)
BPP_HD sc sc_from_wide_words(const uint32_t w[16]) {
    sc lo, hi;
#pragma unroll
    for (int i = 0; i < 8; i++) { lo.v[i] = w[i]; hi.v[i] = w[i + 8]; }
    return sc_add(sc_montmul(lo, sc_const_R()), sc_montmul(hi, sc_const_RR()));
}
BPP_HD bool sc_is_canonical_words(const uint32_t w[8]) {
    int64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (int64_t)w[i] - (int64_t)sc_l(i); c >>= 32; }
    return c < 0;
}
BPP_HD sc sc_frombytes_raw(const uint8_t *s) {
    sc r;
#pragma unroll
    for (int i = 0; i < 8; i++)
        r.v[i] = (uint32_t)s[4 * i] | ((uint32_t)s[4 * i + 1] << 8) | ((uint32_t)s[4 * i + 2] << 16) | ((uint32_t)s[4 * i + 3] << 24);
    return r;
}
BPP_HD void sc_tobytes(uint8_t *s, const sc &a) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        s[4 * i] = (uint8_t)a.v[i]; s[4 * i + 1] = (uint8_t)(a.v[i] >> 8);
        s[4 * i + 2] = (uint8_t)(a.v[i] >> 16); s[4 * i + 3] = (uint8_t)(a.v[i] >> 24);
    }
}

// Montgomery-form inversion: a^(l-2); input/output in Montgomery form
BPP_HD sc scm_invert(const sc &a) {
    // l - 2 = 2^252 + 0x14def9dea2f79cd65812631a5cf5d3eb
    const uint32_t E[4] = {0x5cf5d3ebu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu};
    // left-to-right square and multiply over the low 125 bits, then 127 squarings for the 2^252 term:
    // a^(2^252 + e) = (a^(2^(252-125)) ... ) -- simpler: process all 253 bits MSB first.
    sc acc = a; // bit 252
    for (int i = 251; i >= 0; i--) {
        acc = sc_montmul(acc, acc);
        uint32_t bit = (i < 128) ? ((E[i >> 5] >> (i & 31)) & 1u) : 0u;
        if (bit) acc = sc_montmul(acc, a);
    }
    return acc;
}

// Inverse mod l by the binary extended Euclid with branch-free steps (plain domain; 0 -> 0).
// Invariants: x1 * a = u and x2 * a = v (mod l), gcd(u, v) = 1.  One step: make u the even one (swap if only v is even;
// if both are odd, order them and subtract), then halve u and x1.  bitlen(u) + bitlen(v) <= 506 drops by >= 1 per step.
// ~130 straight-line instructions per step against ~380 dependent Montgomery multiplications for a^(l-2): about 8x
// shorter as a single-thread latency chain, which is what the verifier's per-proof scalar prep is bound by.
BPP_HD sc sc_invert_gcd(const sc &a) {
    uint32_t u[8], v[8], x1[8], x2[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { u[i] = a.v[i]; v[i] = sc_l(i); x1[i] = 0; x2[i] = 0; }
    x1[0] = 1;
    for (int step = 0; step < 520; step++) {
        uint32_t nz = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) nz |= u[i];
        if (nz == 0) break;
        // 1. only v even -> swap
        uint32_t sw = ((u[0] & 1u) & ~(v[0] & 1u)) ? 0xffffffffu : 0u;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t t = (u[i] ^ v[i]) & sw; u[i] ^= t; v[i] ^= t;
            t = (x1[i] ^ x2[i]) & sw; x1[i] ^= t; x2[i] ^= t;
        }
        // 2. both odd -> u = |u - v| ordering by swap, x1 -= x2 (mod l)
        uint32_t odd = (u[0] & 1u) ? 0xffffffffu : 0u;
        {
            int64_t c = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) { c += (int64_t)u[i] - (int64_t)v[i]; c >>= 32; }
            uint32_t lt = (c < 0) ? odd : 0u;       // u < v (and both odd) -> swap first
#pragma unroll
            for (int i = 0; i < 8; i++) {
                uint32_t t = (u[i] ^ v[i]) & lt; u[i] ^= t; v[i] ^= t;
                t = (x1[i] ^ x2[i]) & lt; x1[i] ^= t; x2[i] ^= t;
            }
            c = 0;
            int64_t d = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                c += (int64_t)u[i] - (int64_t)(v[i] & odd); u[i] = (uint32_t)c; c >>= 32;
                d += (int64_t)x1[i] - (int64_t)(x2[i] & odd); x1[i] = (uint32_t)d; d >>= 32;
            }
            uint32_t fix = d < 0 ? 0xffffffffu : 0u;
            uint64_t e = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) { e += (uint64_t)x1[i] + (sc_l(i) & fix); x1[i] = (uint32_t)e; e >>= 32; }
        }
        // 3. u is even: if it is not zero, halve u and x1 (x1 odd -> (x1 + l) / 2)
        nz = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) nz |= u[i];
        if (nz != 0) {
            uint32_t xo = (x1[0] & 1u) ? 0xffffffffu : 0u;
            uint64_t e = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) { e += (uint64_t)x1[i] + (sc_l(i) & xo); x1[i] = (uint32_t)e; e >>= 32; }
#pragma unroll
            for (int i = 0; i < 7; i++) { u[i] = (u[i] >> 1) | (u[i + 1] << 31); x1[i] = (x1[i] >> 1) | (x1[i + 1] << 31); }
            u[7] >>= 1; x1[7] >>= 1;
        }
    }
    sc r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = x2[i];
    return r;
}
// Inverse mod l by Bernstein-Yang "safegcd" division steps in batches of 30 (variable time; plain domain; 0 -> 0).
// A division step looks at the low bits of f, g and a counter only, so 30 of them run on single 32-bit words and yield a 2x2
// transition matrix that is then applied once to the full-length (f, g) and, modulo l, to the Bezout pair (d, e).  ~19 batches of
// ~300 instructions against ~380 steps of ~110 instructions for the binary Euclid above: the single longest piece of the verifier's
// per-proof scalar prep becomes ~7x shorter.  Numbers are 9 signed limbs of 30 bits (little-endian, value = sum v[i] 2^(30 i)).
// Follows the published algorithm (Bernstein, Yang: "Fast constant-time gcd computation and modular inversion", 2019; the
// variable-time batch form used by libsecp256k1's modinv32), written for l.
struct sg30 { int32_t v[9]; };
BPP_HD int32_t sg_l(int i) {
    const int32_t LL[9] = {0x1cf5d3ed, 0x20498c69, 0x2f79cd65, 0x37be77a8, 0x14, 0, 0, 0, 0x1000};
    return LL[i];
}
#define BPP_SG_LINV30 0x2dab81e5u     // l^-1 mod 2^30
BPP_HD int sg_ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
// up to 30 division steps on the low words; returns the new eta, t = (u, v, q, r) scaled by 2^30
BPP_HD int32_t sg_divsteps_30_var(int32_t eta, uint32_t f0, uint32_t g0, int32_t t[4]) {
    uint32_t u = 1, v = 0, q = 0, r = 1, f = f0, g = g0;
    int i = 30;
    for (;;) {
        const int zeros = sg_ctz32(g | (0xffffffffu << i));
        g >>= zeros; u <<= zeros; v <<= zeros;
        eta -= zeros; i -= zeros;
        if (i == 0) break;
        if (eta < 0) {
            eta = -eta;
            uint32_t tmp = f; f = g; g = 0u - tmp;
            tmp = u; u = q; q = 0u - tmp;
            tmp = v; v = r; r = 0u - tmp;
        }
        // cancel up to min(eta + 1, i, 8) low bits of g at once: w = -g / f mod 2^limit
        int limit = (eta + 1) > i ? i : (eta + 1);
        if (limit > 8) limit = 8;
        const uint32_t m = 0xffffffffu >> (32 - limit);
        uint32_t finv = f;                         // f odd: f * f = 1 mod 8; two Newton steps reach 12 bits
        finv *= 2u - f * finv;
        finv *= 2u - f * finv;
        const uint32_t w = (0u - g * finv) & m;
        g += f * w; q += u * w; r += v * w;
    }
    t[0] = (int32_t)u; t[1] = (int32_t)v; t[2] = (int32_t)q; t[3] = (int32_t)r;
    return eta;
}
// (d, e) <- t (d, e) / 2^30 mod l   (both stay in (-2l, l))
BPP_HD void sg_update_de(sg30 &d, sg30 &e, const int32_t t[4]) {
    const int32_t M30 = 0x3fffffff;
    const int64_t u = t[0], v = t[1], q = t[2], r = t[3];
    const int32_t sd = d.v[8] >> 31, se = e.v[8] >> 31;
    int32_t md = (t[0] & sd) + (t[1] & se), me = (t[2] & sd) + (t[3] & se);
    int64_t cd = u * d.v[0] + v * e.v[0], ce = q * d.v[0] + r * e.v[0];
    md -= (int32_t)((BPP_SG_LINV30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
    me -= (int32_t)((BPP_SG_LINV30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
    cd += (int64_t)sg_l(0) * md; ce += (int64_t)sg_l(0) * me;
    cd >>= 30; ce >>= 30;
#pragma unroll
    for (int i = 1; i < 9; i++) {
        cd += u * d.v[i] + v * e.v[i]; ce += q * d.v[i] + r * e.v[i];
        cd += (int64_t)sg_l(i) * md; ce += (int64_t)sg_l(i) * me;
        d.v[i - 1] = (int32_t)cd & M30; cd >>= 30;
        e.v[i - 1] = (int32_t)ce & M30; ce >>= 30;
    }
    d.v[8] = (int32_t)cd; e.v[8] = (int32_t)ce;
}
// (f, g) <- t (f, g) / 2^30 (exact)
BPP_HD void sg_update_fg(sg30 &f, sg30 &g, const int32_t t[4]) {
    const int32_t M30 = 0x3fffffff;
    const int64_t u = t[0], v = t[1], q = t[2], r = t[3];
    int64_t cf = u * f.v[0] + v * g.v[0], cg = q * f.v[0] + r * g.v[0];
    cf >>= 30; cg >>= 30;
#pragma unroll
    for (int i = 1; i < 9; i++) {
        cf += u * f.v[i] + v * g.v[i]; cg += q * f.v[i] + r * g.v[i];
        f.v[i - 1] = (int32_t)cf & M30; cf >>= 30;
        g.v[i - 1] = (int32_t)cg & M30; cg >>= 30;
    }
    f.v[8] = (int32_t)cf; g.v[8] = (int32_t)cg;
}
// r in (-2l, l), negated when sign < 0, brought into [0, l)
BPP_HD void sg_normalize(sg30 &r, int32_t sign) {
    const int32_t M30 = 0x3fffffff;
    int32_t cond_add = r.v[8] >> 31;
    const int32_t cond_negate = sign >> 31;
#pragma unroll
    for (int i = 0; i < 9; i++) r.v[i] = ((r.v[i] + (sg_l(i) & cond_add)) ^ cond_negate) - cond_negate;
#pragma unroll
    for (int i = 0; i < 8; i++) { r.v[i + 1] += r.v[i] >> 30; r.v[i] &= M30; }
    cond_add = r.v[8] >> 31;
#pragma unroll
    for (int i = 0; i < 9; i++) r.v[i] += sg_l(i) & cond_add;
#pragma unroll
    for (int i = 0; i < 8; i++) { r.v[i + 1] += r.v[i] >> 30; r.v[i] &= M30; }
}
BPP_HD sc sc_invert_sg(const sc &a) {
    sg30 d, e, f, g;
    // a (8 x 32 bits, < l) -> 9 x 30 bits
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const int bit = 30 * i, wi = bit >> 5, sh = bit & 31;
        uint64_t w = a.v[wi];
        if (wi + 1 < 8) w |= (uint64_t)a.v[wi + 1] << 32;
        g.v[i] = (int32_t)((uint32_t)(w >> sh) & 0x3fffffffu);
        f.v[i] = sg_l(i);
        d.v[i] = 0; e.v[i] = 0;
    }
    e.v[0] = 1;
    int32_t eta = -1;
    for (int it = 0; it < 40; it++) {              // 25 batches suffice for 256-bit inputs; the loop leaves when g is zero
        int32_t t[4];
        eta = sg_divsteps_30_var(eta, (uint32_t)f.v[0], (uint32_t)g.v[0], t);
        sg_update_de(d, e, t);
        sg_update_fg(f, g, t);
        int32_t nz = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) nz |= g.v[i];
        if (nz == 0) break;
    }
    // f = +-gcd = +-1 (or +-l when a = 0: then d = 0 and the result is 0)
    sg_normalize(d, f.v[8]);
    sc r;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        // word i = bits [32 i, 32 i + 32) of sum d.v[k] 2^(30 k)
        const int bit = 32 * i, k = bit / 30, sh = bit - 30 * k;
        uint64_t w = (uint64_t)(uint32_t)d.v[k] >> sh;
        if (k + 1 < 9) w |= (uint64_t)(uint32_t)d.v[k + 1] << (30 - sh);
        if (k + 2 < 9) w |= (uint64_t)(uint32_t)d.v[k + 2] << (60 - sh);
        r.v[i] = (uint32_t)w;
    }
    return r;
}
// Montgomery-form inversion through the plain-domain Euclid: (aR)^-1 = a^-1 R^-1, times R^3 (Montgomery) = a^-1 R
BPP_HD sc sc_const_RRR() { return BPP_SC(0x7b83a2dbu, 0x2a9e4968u, 0xaef7f3ecu, 0x278324e6u, 0x04ec5b65u, 0x8065dc6cu, 0x3599cec7u, 0x0e530b77u); }
BPP_HD sc scm_invert_gcd(const sc &a) { return sc_montmul(sc_invert_sg(a), sc_const_RRR()); }
// the same through the binary Euclid (kept for comparison and as a cross-check in the tests)
BPP_HD sc scm_invert_euclid(const sc &a) { return sc_montmul(sc_invert_gcd(a), sc_const_RRR()); }

// ================================================================================================ points
BPP_HD ge ge_identity() { ge r; r.X = fe_zero(); r.Y = fe_one(); r.Z = fe_one(); r.T = fe_zero(); return r; }
BPP_HD aniels aniels_identity() { aniels r; r.ypx = fe_one(); r.ymx = fe_one(); r.t2d = fe_zero(); return r; }

// extended + affine Niels (7M)
BPP_HD ge ge_madd(const ge &p, const aniels &q) {
    fe A = fe_mul(fe_sub(p.Y, p.X), q.ymx);
    fe B = fe_mul(fe_add(p.Y, p.X), q.ypx);
    fe C = fe_mul(p.T, q.t2d);
    fe D = fe_add(p.Z, p.Z);
    fe E = fe_sub(B, A), F = fe_sub(D, C), G = fe_add(D, C), H = fe_add(B, A);
    ge r;
    r.X = fe_mul(E, F); r.Y = fe_mul(G, H); r.Z = fe_mul(F, G); r.T = fe_mul(E, H);
    return r;
}
BPP_HD ge ge_msub(const ge &p, const aniels &q) {
    fe A = fe_mul(fe_sub(p.Y, p.X), q.ypx);
    fe B = fe_mul(fe_add(p.Y, p.X), q.ymx);
    fe C = fe_mul(p.T, q.t2d);
    fe D = fe_add(p.Z, p.Z);
    fe E = fe_sub(B, A), F = fe_add(D, C), G = fe_sub(D, C), H = fe_add(B, A);
    ge r;
    r.X = fe_mul(E, F); r.Y = fe_mul(G, H); r.Z = fe_mul(F, G); r.T = fe_mul(E, H);
    return r;
}
// extended + extended (9M incl. the 2d multiply)
BPP_HD ge ge_add(const ge &p, const ge &q) {
    fe A = fe_mul(fe_sub(p.Y, p.X), fe_sub(q.Y, q.X));
    fe B = fe_mul(fe_add(p.Y, p.X), fe_add(q.Y, q.X));
    fe C = fe_mul(fe_mul(p.T, q.T), fe_const_2d());
    fe D = fe_mul(p.Z, q.Z);
    D = fe_add(D, D);
    fe E = fe_sub(B, A), F = fe_sub(D, C), G = fe_add(D, C), H = fe_add(B, A);
    ge r;
    r.X = fe_mul(E, F); r.Y = fe_mul(G, H); r.Z = fe_mul(F, G); r.T = fe_mul(E, H);
    return r;
}
BPP_HD ge ge_neg(const ge &p) { ge r; r.X = fe_neg(p.X); r.Y = p.Y; r.Z = p.Z; r.T = fe_neg(p.T); return r; }
BPP_HD ge ge_dbl(const ge &p) {
    fe XX = fe_sq(p.X), YY = fe_sq(p.Y), ZZ = fe_sq(p.Z);
    fe ZZ2 = fe_add(ZZ, ZZ);
    fe S = fe_sq(fe_add(p.X, p.Y));
    fe H = fe_add(YY, XX), G = fe_sub(YY, XX);
    fe E = fe_sub(S, H);
    fe F = fe_sub(ZZ2, G);
    ge r;
    r.X = fe_mul(E, F); r.Y = fe_mul(H, G); r.Z = fe_mul(G, F); r.T = fe_mul(E, H);
    return r;
}
BPP_HD aniels ge_to_aniels_affine(const fe &x, const fe &y, const fe &xy) {
    aniels r;
    r.ypx = fe_add(y, x); r.ymx = fe_sub(y, x); r.t2d = fe_mul(xy, fe_const_2d());
    return r;
}
BPP_HD ge aniels_to_ge(const aniels &q) { return ge_madd(ge_identity(), q); }

// Ristretto identity test: X == 0 or Y == 0  (RFC 9496 equality against (0,1,1,0))
BPP_HD bool ge_is_ristretto_identity(const ge &p) { return fe_is_zero(p.X) || fe_is_zero(p.Y); }
BPP_HD bool ge_ristretto_eq(const ge &p, const ge &q) {
    return fe_eq(fe_mul(p.X, q.Y), fe_mul(p.Y, q.X)) || fe_eq(fe_mul(p.X, q.X), fe_mul(p.Y, q.Y));
}

// RFC 9496 §4.3.1 (CompressedRistretto::decompress). words: 8 little-endian u32. Returns ok; x,y,t affine.
BPP_HD bool ristretto_decode(fe &x, fe &y, fe &t, const uint32_t words[8]) {
    bool canonical;
    fe s = fe_fromwords(words, &canonical);
    bool ok = canonical && ((words[0] & 1u) == 0);
    fe one = fe_one();
    fe ss = fe_sq(s);
    fe u1 = fe_sub(one, ss), u2 = fe_add(one, ss);
    fe u2s = fe_sq(u2);
    fe v = fe_sub(fe_neg(fe_mul(fe_const_d(), fe_sq(u1))), u2s);
    fe I;
    bool was_sq = fe_invsqrt(I, fe_mul(v, u2s));
    fe dx = fe_mul(I, u2);
    fe dy = fe_mul(fe_mul(I, dx), v);
    x = fe_abs(fe_mul(fe_add(s, s), dx));
    y = fe_mul(u1, dy);
    t = fe_mul(x, y);
    return ok && was_sq && !fe_is_negative(t) && !fe_is_zero(y);
}

// RFC 9496 §4.3.2 (RistrettoPoint::compress) -> canonical fe (write with fe_tobytes / words of fe_canon)
BPP_HD fe ristretto_encode(const ge &p) {
    fe u1 = fe_mul(fe_add(p.Z, p.Y), fe_sub(p.Z, p.Y));
    fe u2 = fe_mul(p.X, p.Y);
    fe I;
    fe_invsqrt(I, fe_mul(u1, fe_sq(u2)));
    fe d1 = fe_mul(I, u1), d2 = fe_mul(I, u2);
    fe zinv = fe_mul(fe_mul(d1, d2), p.T);
    bool rotate = fe_is_negative(fe_mul(p.T, zinv));
    fe ix = fe_mul(p.X, fe_const_sqrtm1()), iy = fe_mul(p.Y, fe_const_sqrtm1());
    fe eden = fe_mul(d1, fe_const_invsqrt_a_minus_d());
    fe x = fe_select(p.X, iy, rotate);
    fe y = fe_select(p.Y, ix, rotate);
    fe den = fe_select(d2, eden, rotate);
    y = fe_cneg(y, fe_is_negative(fe_mul(x, zinv)));
    fe s = fe_abs(fe_mul(den, fe_sub(p.Z, y)));
    return fe_canon(s);
}

// RFC 9496 §4.3.4 MAP
BPP_HD ge ristretto_elligator(const fe &t0) {
    fe one = fe_one(), minus_one = fe_neg(fe_one());
    fe r = fe_mul(fe_const_sqrtm1(), fe_sq(t0));
    fe u = fe_mul(fe_add(r, one), fe_const_one_minus_d_sq());
    fe v = fe_mul(fe_sub(minus_one, fe_mul(r, fe_const_d())), fe_add(r, fe_const_d()));
    fe s;
    bool was_sq = fe_sqrt_ratio_i(s, u, v);
    fe sp = fe_neg(fe_abs(fe_mul(s, t0)));
    s = fe_select(sp, s, was_sq);
    fe c = fe_select(r, minus_one, was_sq);
    fe N = fe_sub(fe_mul(fe_mul(c, fe_sub(r, one)), fe_const_d_minus_one_sq()), v);
    fe w0 = fe_mul(fe_add(s, s), v);
    fe w1 = fe_mul(N, fe_const_sqrt_ad_minus_one());
    fe ss = fe_sq(s);
    fe w2 = fe_sub(one, ss), w3 = fe_add(one, ss);
    ge p;
    p.X = fe_mul(w0, w3); p.Y = fe_mul(w2, w1); p.Z = fe_mul(w1, w3); p.T = fe_mul(w0, w2);
    return p;
}
// RistrettoPoint::from_uniform_bytes (16 words)
BPP_HD ge ristretto_from_uniform_words(const uint32_t w[16]) {
    fe r0 = fe_fromwords(w, nullptr), r1 = fe_fromwords(w + 8, nullptr);
    return ge_add(ristretto_elligator(r0), ristretto_elligator(r1));
}

} // namespace bpp
