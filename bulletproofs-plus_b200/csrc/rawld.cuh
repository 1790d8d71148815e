// Unaligned loads from the uploaded byte stream of a verification pass.  Serialised proofs start with a one-byte extension
// degree (range_proof.rs:1120-1150), so every 32-byte element of a proof sits at an arbitrary byte offset; the device reads the
// bytes where the caller's buffers put them instead of having the host re-lay them out (engine_verify.cu).
#pragma once
#include <stdint.h>

namespace bpp {

// 32 bytes at p (any alignment) as 8 little-endian words: 8 or 9 aligned word loads + funnel shifts.  The word behind the last
// byte may be read: every section of the blob is padded by at least 4 bytes.
static __device__ __forceinline__ void ld32_unaligned(const uint8_t *p, uint32_t w[8]) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t *q = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3u) * 8u;
    uint32_t t[9];
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = q[i];
    t[8] = sh ? q[8] : 0u;
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = __funnelshift_r(t[i], t[i + 1], sh);
}

} // namespace bpp
