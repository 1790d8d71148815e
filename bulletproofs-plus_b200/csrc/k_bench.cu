// Integer-pipe microbenchmarks: the measured denominators for the roofline (the path is bound by
// the INT32 multiply issue rate, not by HBM or tensor cores).  Every kernel runs ILP-8 dependent chains per
// thread over a grid that fills all 148 SMs at full occupancy; the caller converts elapsed time to ops/s.
#include "kernels.cuh"
#include "quad.cuh"

namespace bpp {

#define MB_BLOCKS (148 * 8)
#define MB_THREADS 256
#define MB_ILP 8

template <int WHICH> __global__ void __launch_bounds__(MB_THREADS) k_mb_instr(int iters, uint32_t seed, uint32_t *sink) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t a = seed * 2654435761u + tid, b = (seed ^ 0x9e3779b9u) + tid * 7u;
    uint32_t x[MB_ILP];
    uint64_t y[MB_ILP];
#pragma unroll
    for (int k = 0; k < MB_ILP; k++) { x[k] = a + k; y[k] = ((uint64_t)b << 32) | (a + k); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 8; rep++) {
#pragma unroll
            for (int k = 0; k < MB_ILP; k++) {
                if (WHICH == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(a), "r"(b));
                if (WHICH == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(a), "r"(b));
                if (WHICH == 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y[k]) : "r"((uint32_t)y[k]), "r"(b));
                if (WHICH == 3) asm volatile("add.u32 %0, %0, %1;\n\txor.b32 %0, %0, %2;" : "+r"(x[k]) : "r"(a), "r"(b));
                if (WHICH == 10) {
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(a), "r"(b));
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(((uint32_t *)&y[k])[0]) : "r"(a));
                }
                if (WHICH == 12) {        // one 32x32 -> 64 product issued as its two halves (two self-dependent chains, nothing to hoist): do the
                                          // lo and hi forms share a pipe?  Counted as pairs.
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(a), "r"(b));
                    asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(((uint32_t *)&y[k])[1]) : "r"(a), "r"(b));
                }
                if (WHICH == 11) {
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y[k]) : "r"(x[k]), "r"(b));
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(a));
                }
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < MB_ILP; k++) acc ^= x[k] ^ (uint32_t)y[k] ^ (uint32_t)(y[k] >> 32);
    if (acc == 0x12345678u) sink[0] = acc;
}

static __device__ __noinline__ fe fe_mul_portable(fe a, fe b) {   // which == 9: the hand-chained mad.cc PTX core
    uint32_t t[16];
#if BPP_PTX
    ptx::mul256(t, a.v, b.v);
#else
    mul256_portable(t, a.v, b.v);
#endif
    return fe_reduce512(t);
}

template <int WHICH> __global__ void __launch_bounds__(MB_THREADS) k_mb_field(int iters, uint32_t seed, uint32_t *sink) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    fe a, b;
#pragma unroll
    for (int i = 0; i < 8; i++) { a.v[i] = seed * (i + 3) + tid * 2654435761u; b.v[i] = (seed ^ 0x5bd1e995u) * (i + 7) + tid; }
    a.v[7] &= 0x7fffffffu; b.v[7] &= 0x7fffffffu;
    uint32_t acc = 0;
    if (WHICH == 4) { for (int it = 0; it < iters; it++) { a = fe_mul(a, b); b = fe_mul(b, a); } }
    if (WHICH == 5) { for (int it = 0; it < iters; it++) { a = fe_sq(a); b = fe_sq(b); } }
    if (WHICH == 9) { for (int it = 0; it < iters; it++) { a = fe_mul_portable(a, b); b = fe_mul_portable(b, a); } }
    if (WHICH == 6 || WHICH == 7) {
        ge p = ge_identity();
        aniels q; q.ypx = a; q.ymx = b; q.t2d = fe_add(a, b);
        p.X = a; p.T = b;
        if (WHICH == 6) for (int it = 0; it < iters; it++) { p = ge_madd(p, q); p = ge_msub(p, q); }
        else for (int it = 0; it < iters; it++) { p = ge_dbl(p); p = ge_dbl(p); }
        a = fe_add(p.X, p.Y); b = fe_add(p.Z, p.T);
    }
    if (WHICH == 8) {
        sc u, v;
#pragma unroll
        for (int i = 0; i < 8; i++) { u.v[i] = a.v[i]; v.v[i] = b.v[i]; }
        u.v[7] &= 0x0fffffffu; v.v[7] &= 0x0fffffffu;
        for (int it = 0; it < iters; it++) { u = sc_montmul(u, v); v = sc_montmul(v, u); }
#pragma unroll
        for (int i = 0; i < 8; i++) { a.v[i] = u.v[i]; b.v[i] = v.v[i]; }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= a.v[i] ^ b.v[i];
    if (acc == 0x12345678u) sink[0] = acc;
}

// latency probes: ONE warp per SM runs a dependent chain (20 fe_mul, 21 fe_sq, 24 sc_montmul, 25 ge_dbl, 27 ge_madd,
// 30 quad_dbl, 31 quad_add, 32 quad_add + quad_to_cached)
template <int WHICH> __global__ void __launch_bounds__(32) k_mb_lat(int iters, uint32_t seed, uint32_t *sink) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    fe a, b;
#pragma unroll
    for (int i = 0; i < 8; i++) { a.v[i] = seed * (i + 3) + tid * 2654435761u; b.v[i] = (seed ^ 0x5bd1e995u) * (i + 7) + tid; }
    a.v[7] &= 0x7fffffffu; b.v[7] &= 0x7fffffffu;
    if (WHICH == 20) for (int it = 0; it < iters; it++) { a = fe_mul(a, b); a = fe_mul(a, b); }
    if (WHICH == 21) for (int it = 0; it < iters; it++) { a = fe_sq(a); a = fe_sq(a); }
    if (WHICH >= 30 && WHICH <= 32) {
        const int role = threadIdx.x & 3, base = threadIdx.x & 28;
        if (WHICH == 30) for (int it = 0; it < iters; it++) { a = quad_dbl(a, role, base); a = quad_dbl(a, role, base); }
        if (WHICH == 31) for (int it = 0; it < iters; it++) { a = quad_add(a, role, base, b); a = quad_add(a, role, base, b); }
        if (WHICH == 32) for (int it = 0; it < iters; it++) { a = quad_add(a, role, base, quad_to_cached(b, role, base)); b = quad_add(b, role, base, quad_to_cached(a, role, base)); }
    }
    if (WHICH == 24) {
        sc u, v;
#pragma unroll
        for (int i = 0; i < 8; i++) { u.v[i] = a.v[i]; v.v[i] = b.v[i]; }
        u.v[7] &= 0x0fffffffu; v.v[7] &= 0x0fffffffu;
        for (int it = 0; it < iters; it++) { u = sc_montmul(u, v); u = sc_montmul(u, v); }
#pragma unroll
        for (int i = 0; i < 8; i++) a.v[i] = u.v[i];
    }
    if (WHICH == 25 || WHICH == 27) {
        ge p = ge_identity();
        aniels q; q.ypx = a; q.ymx = b; q.t2d = fe_add(a, b);
        p.X = a; p.T = b;
        if (WHICH == 25) for (int it = 0; it < iters; it++) { p = ge_dbl(p); p = ge_dbl(p); }
        if (WHICH == 27) for (int it = 0; it < iters; it++) { p = ge_madd(p, q); p = ge_madd(p, q); }
        a = fe_add(fe_add(p.X, p.Y), fe_add(p.Z, p.T));
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= a.v[i];
    if (acc == 0x12345678u) sink[0] = acc;
}

int microbench_run(cudaStream_t s, int which, int iters, double *ops_per_sec, double *seconds, uint64_t *launches) {
    static uint32_t *sink = nullptr;
    if (!sink && cudaMalloc(&sink, 64) != cudaSuccess) return -1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double ops_per_thread = 0;
    for (int pass = 0; pass < 2; pass++) {   // pass 0 = warm-up
        if (pass == 1) cudaEventRecord(e0, s);
        switch (which) {
            case 0: k_mb_instr<0><<<MB_BLOCKS, MB_THREADS, 0, s>>>(iters, 1u, sink); ops_per_thread = 8.0 * MB_ILP * iters; break;
            case 1: k_mb_instr<1><<<MB_BLOCKS, MB_THREADS, 0, s>>>(iters, 1u, sink); ops_per_thread = 8.0 * MB_ILP * iters; break;
            case 2: k_mb_instr<2><<<MB_BLOCKS, MB_THREADS, 0, s>>>(iters, 1u, sink); ops_per_thread = 8.0 * MB_ILP * iters; break;
            case 3: k_mb_instr<3><<<MB_BLOCKS, MB_THREADS, 0, s>>>(iters, 1u, sink); ops_per_thread = 2 * 8.0 * MB_ILP * iters; break;
            case 10: k_mb_instr<10><<<MB_BLOCKS, MB_THREADS, 0, s>>>(iters, 1u, sink); ops_per_thread = 2 * 8.0 * MB_ILP * iters; break;
            case 12: k_mb_instr<12><<<MB_BLOCKS, MB_THREADS, 0, s>>>(iters, 1u, sink); ops_per_thread = 8.0 * MB_ILP * iters; break;
            case 11: k_mb_instr<11><<<MB_BLOCKS, MB_THREADS, 0, s>>>(iters, 1u, sink); ops_per_thread = 2 * 8.0 * MB_ILP * iters; break;
            case 4: k_mb_field<4><<<MB_BLOCKS, MB_THREADS, 0, s>>>(iters, 1u, sink); ops_per_thread = 2.0 * iters; break;
            case 5: k_mb_field<5><<<MB_BLOCKS, MB_THREADS, 0, s>>>(iters, 1u, sink); ops_per_thread = 2.0 * iters; break;
            case 6: k_mb_field<6><<<MB_BLOCKS, MB_THREADS, 0, s>>>(iters, 1u, sink); ops_per_thread = 2.0 * iters; break;
            case 7: k_mb_field<7><<<MB_BLOCKS, MB_THREADS, 0, s>>>(iters, 1u, sink); ops_per_thread = 2.0 * iters; break;
            case 8: k_mb_field<8><<<MB_BLOCKS, MB_THREADS, 0, s>>>(iters, 1u, sink); ops_per_thread = 2.0 * iters; break;
            case 9: k_mb_field<9><<<MB_BLOCKS, MB_THREADS, 0, s>>>(iters, 1u, sink); ops_per_thread = 2.0 * iters; break;
#define LATCASE(W) case W: k_mb_lat<W><<<148, 32, 0, s>>>(iters, 1u, sink); ops_per_thread = 2.0 * iters * (MB_BLOCKS * (double)MB_THREADS) / (148.0 * 32.0); break;
            LATCASE(20) LATCASE(21) LATCASE(24) LATCASE(25) LATCASE(27) LATCASE(30) LATCASE(31) LATCASE(32)
            default: cudaEventDestroy(e0); cudaEventDestroy(e1); return -2;
        }
        if (launches) (*launches)++;
    }
    cudaEventRecord(e1, s);
    cudaError_t err = cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (err != cudaSuccess || cudaGetLastError() != cudaSuccess) return -1;
    *seconds = ms * 1e-3;
    *ops_per_sec = ops_per_thread * (double)MB_BLOCKS * MB_THREADS / (*seconds);
    return 0;
}

} // namespace bpp
