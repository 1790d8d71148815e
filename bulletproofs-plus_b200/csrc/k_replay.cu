// K-REPLAY: loop 1 of the batch verifier on the device -- the Fiat-Shamir replay of every proof's Merlin transcript
// (/root/reference/src/range_proof.rs:816-850 over src/transcripts.rs:59-179 and src/protocols/transcript_protocol.rs:39-79).
// Outputs the challenges (canonical scalars) straight into the buffer K-VPREP reads, the 32 bytes each proof feeds into the
// verifier-weight transcript, the advanced transcript states and per-proof failure flags.  Inputs are read where the caller's
// buffers put them: points and scalars at their byte offsets inside the serialised proofs (rawld.cuh).
//
// Three kernels, selectable (bpp_ctx_set_replay_mode), bit-identical results:
//   k_replay_sm     (default) one thread per proof, one warp per CTA.  The STROBE-128 state lives in SHARED memory
//                   ([word][lane]: conflict-free whatever the byte position, dynamic positions cost an address, not a
//                   local-memory round trip), Keccak-f[1600] runs in registers (50 loads, 24 rolled rounds, 50 stores).  The
//                   round-1 kernel kept the state in local memory behind out-of-line byte loops: 152 k warp instructions per
//                   32 proofs at 0.27 IPC (291 us for one warp's ~21 permutations, profiles/r01_ncu_summary.md).
//   k_replay<false> the round-1 thread-per-proof kernel (hash.cuh Merlin, state in local memory), kept for comparison
//   k_replay<true>  one warp per proof (wstrobe.cuh)
#include <stdlib.h>
#include "kernels.cuh"
#include "rawld.cuh"
#include "replay.cuh"
#include "wstrobe.cuh"

namespace bpp {

// ------------------------------------------------------------------------------------------------ shared-memory sponge
namespace {

constexpr int SM_RATE = 166;
constexpr int SM_WORDS = 51;          // 50 state words + {pos, pos_begin, cur_flags} packed into word 50 (the 203-byte wire form)
__constant__ uint32_t c_keccak_rc[48] = {
    0x00000001u, 0x00000000u, 0x00008082u, 0x00000000u, 0x0000808au, 0x80000000u, 0x80008000u, 0x80000000u,
    0x0000808bu, 0x00000000u, 0x80000001u, 0x00000000u, 0x80008081u, 0x80000000u, 0x00008009u, 0x80000000u,
    0x0000008au, 0x00000000u, 0x00000088u, 0x00000000u, 0x80008009u, 0x00000000u, 0x8000000au, 0x00000000u,
    0x8000808bu, 0x00000000u, 0x0000008bu, 0x80000000u, 0x00008089u, 0x80000000u, 0x00008003u, 0x80000000u,
    0x00008002u, 0x80000000u, 0x00000080u, 0x80000000u, 0x0000800au, 0x00000000u, 0x8000000au, 0x80000000u,
    0x80008081u, 0x80000000u, 0x00008080u, 0x80000000u, 0x80000001u, 0x00000000u, 0x80008008u, 0x80000000u};

struct L64 { uint32_t lo, hi; };
static __device__ __forceinline__ L64 x5(const L64 &a, const L64 &b, const L64 &c, const L64 &d, const L64 &e) {
    L64 r;
    r.lo = a.lo ^ b.lo ^ c.lo ^ d.lo ^ e.lo;
    r.hi = a.hi ^ b.hi ^ c.hi ^ d.hi ^ e.hi;
    return r;
}
template <int N> static __device__ __forceinline__ L64 rotl(const L64 &a) {      // compile-time rotation: two funnel shifts
    L64 r;
    if (N == 0) return a;
    if (N == 32) { r.lo = a.hi; r.hi = a.lo; return r; }
    if (N < 32) {
        r.lo = __funnelshift_l(a.hi, a.lo, N);
        r.hi = __funnelshift_l(a.lo, a.hi, N);
    } else {
        r.lo = __funnelshift_l(a.lo, a.hi, N - 32);
        r.hi = __funnelshift_l(a.hi, a.lo, N - 32);
    }
    return r;
}
static __device__ __forceinline__ L64 xr(const L64 &a, const L64 &b) { L64 r; r.lo = a.lo ^ b.lo; r.hi = a.hi ^ b.hi; return r; }
static __device__ __forceinline__ L64 chi(const L64 &a, const L64 &b, const L64 &c) { L64 r; r.lo = a.lo ^ (~b.lo & c.lo); r.hi = a.hi ^ (~b.hi & c.hi); return r; }

// Keccak-f[1600] on the state of this thread: S[w * 32] is word w (lanes of the warp are interleaved)
static __device__ __noinline__ void sm_keccak(uint32_t *S) {
    L64 a[25];
#pragma unroll
    for (int i = 0; i < 25; i++) { a[i].lo = S[(2 * i) * 32]; a[i].hi = S[(2 * i + 1) * 32]; }
#pragma unroll 1
    for (int round = 0; round < 24; round++) {
        const L64 c0 = x5(a[0], a[5], a[10], a[15], a[20]), c1 = x5(a[1], a[6], a[11], a[16], a[21]), c2 = x5(a[2], a[7], a[12], a[17], a[22]),
                  c3 = x5(a[3], a[8], a[13], a[18], a[23]), c4 = x5(a[4], a[9], a[14], a[19], a[24]);
        const L64 d0 = xr(c4, rotl<1>(c1)), d1 = xr(c0, rotl<1>(c2)), d2 = xr(c1, rotl<1>(c3)), d3 = xr(c2, rotl<1>(c4)), d4 = xr(c3, rotl<1>(c0));
        const L64 b0 = xr(a[0], d0);
        const L64 b10 = rotl<1>(xr(a[1], d1)), b20 = rotl<62>(xr(a[2], d2)), b5 = rotl<28>(xr(a[3], d3)), b15 = rotl<27>(xr(a[4], d4));
        const L64 b16 = rotl<36>(xr(a[5], d0)), b1 = rotl<44>(xr(a[6], d1)), b11 = rotl<6>(xr(a[7], d2)), b21 = rotl<55>(xr(a[8], d3)), b6 = rotl<20>(xr(a[9], d4));
        const L64 b7 = rotl<3>(xr(a[10], d0)), b17 = rotl<10>(xr(a[11], d1)), b2 = rotl<43>(xr(a[12], d2)), b12 = rotl<25>(xr(a[13], d3)), b22 = rotl<39>(xr(a[14], d4));
        const L64 b23 = rotl<41>(xr(a[15], d0)), b8 = rotl<45>(xr(a[16], d1)), b18 = rotl<15>(xr(a[17], d2)), b3 = rotl<21>(xr(a[18], d3)), b13 = rotl<8>(xr(a[19], d4));
        const L64 b14 = rotl<18>(xr(a[20], d0)), b24 = rotl<2>(xr(a[21], d1)), b9 = rotl<61>(xr(a[22], d2)), b19 = rotl<56>(xr(a[23], d3)), b4 = rotl<14>(xr(a[24], d4));
        a[0] = chi(b0, b1, b2); a[1] = chi(b1, b2, b3); a[2] = chi(b2, b3, b4); a[3] = chi(b3, b4, b0); a[4] = chi(b4, b0, b1);
        a[5] = chi(b5, b6, b7); a[6] = chi(b6, b7, b8); a[7] = chi(b7, b8, b9); a[8] = chi(b8, b9, b5); a[9] = chi(b9, b5, b6);
        a[10] = chi(b10, b11, b12); a[11] = chi(b11, b12, b13); a[12] = chi(b12, b13, b14); a[13] = chi(b13, b14, b10); a[14] = chi(b14, b10, b11);
        a[15] = chi(b15, b16, b17); a[16] = chi(b16, b17, b18); a[17] = chi(b17, b18, b19); a[18] = chi(b18, b19, b15); a[19] = chi(b19, b15, b16);
        a[20] = chi(b20, b21, b22); a[21] = chi(b21, b22, b23); a[22] = chi(b22, b23, b24); a[23] = chi(b23, b24, b20); a[24] = chi(b24, b20, b21);
        a[0].lo ^= c_keccak_rc[2 * round];
        a[0].hi ^= c_keccak_rc[2 * round + 1];
    }
#pragma unroll
    for (int i = 0; i < 25; i++) { S[(2 * i) * 32] = a[i].lo; S[(2 * i + 1) * 32] = a[i].hi; }
}

// STROBE-128 / Merlin over the shared-memory state (same operations as hash.cuh Strobe128 / Merlin; merlin 3.0.0 strobe.rs,
// transcript.rs).  Per thread: words 0..49 = Keccak state, word 50 = pos | pos_begin << 8 | cur_flags << 16 (the tail of the
// 203-byte wire form), words 51..66 = staging area for the bytes an operation absorbs or squeezes.  Every operation is an
// out-of-line function over that memory, so the kernel body is a short list of calls (fully inlined it was 27 k instructions with
// 1100 call sites of the permutation).
constexpr int SM_STAGE = 51;
constexpr int SM_TOTAL = 67;
enum : uint32_t { FI = 1, FA = 2, FC = 4, FT = 8, FM = 16, FK = 32 };

__constant__ char c_labels[] = "dom-sep\0H\0G\0N\0T\0M\0Ci\0vi - minimum_value\0A\0y\0z\0L\0R\0e\0A1\0B\0r1\0s1\0d1\0rng\0Bulletproofs+ Range Proof\0proof";
// (offset, length) of each label inside c_labels
enum : uint32_t { LB_DOMSEP = 0 | 7 << 8, LB_H = 8 | 1 << 8, LB_G = 10 | 1 << 8, LB_N = 12 | 1 << 8, LB_T = 14 | 1 << 8, LB_M = 16 | 1 << 8,
                  LB_CI = 18 | 2 << 8, LB_VI = 21 | 18 << 8, LB_A = 40 | 1 << 8, LB_Y = 42 | 1 << 8, LB_Z = 44 | 1 << 8, LB_L = 46 | 1 << 8,
                  LB_R = 48 | 1 << 8, LB_E = 50 | 1 << 8, LB_A1 = 52 | 2 << 8, LB_B = 55 | 1 << 8, LB_R1 = 57 | 2 << 8, LB_S1 = 60 | 2 << 8,
                  LB_D1 = 63 | 2 << 8, LB_RNG = 66 | 3 << 8, LB_PROTO = 70 | 25 << 8, LB_PROOF = 96 | 5 << 8 };

struct SmPos { uint32_t pos, pos_begin, flags; };
static __device__ __forceinline__ SmPos sm_get(const uint32_t *S) { const uint32_t t = S[50 * 32]; return SmPos{t & 0xffu, (t >> 8) & 0xffu, (t >> 16) & 0xffu}; }
static __device__ __forceinline__ void sm_put(uint32_t *S, const SmPos &p) { S[50 * 32] = p.pos | (p.pos_begin << 8) | (p.flags << 16); }
static __device__ __forceinline__ void sm_xor8(uint32_t *S, uint32_t p, uint32_t byte) { S[(p >> 2) * 32] ^= byte << (8 * (p & 3)); }
static __device__ __forceinline__ void sm_run_f(uint32_t *S, SmPos &p) {
    sm_xor8(S, p.pos, p.pos_begin);
    sm_xor8(S, p.pos + 1, 0x04u);
    sm_xor8(S, SM_RATE + 1, 0x80u);
    sm_keccak(S);
    p.pos = 0;
    p.pos_begin = 0;
}
static __device__ __forceinline__ void sm_absorb8(uint32_t *S, SmPos &p, uint32_t byte) {
    sm_xor8(S, p.pos, byte);
    if (++p.pos == SM_RATE) sm_run_f(S, p);
}
static __device__ __forceinline__ void sm_begin(uint32_t *S, SmPos &p, uint32_t flags) {
    const uint32_t old_begin = p.pos_begin;
    p.pos_begin = p.pos + 1;
    p.flags = flags;
    sm_absorb8(S, p, old_begin);
    sm_absorb8(S, p, flags);
    if ((flags & (FC | FK)) && p.pos != 0) sm_run_f(S, p);
}
// the staged bytes [0, n): whole words while the sponge position allows it
static __device__ __forceinline__ void sm_absorb_staged(uint32_t *S, SmPos &p, uint32_t n) {
    const uint32_t *T = S + SM_STAGE * 32;
    uint32_t i = 0;
    while (i < n) {
        if ((i & 3) == 0 && n - i >= 4 && p.pos + 4 <= SM_RATE) {
            const uint32_t v = T[(i >> 2) * 32], wi = p.pos >> 2, sh = 8 * (p.pos & 3);
            S[wi * 32] ^= v << sh;
            if (sh) S[(wi + 1) * 32] ^= v >> (32 - sh);
            p.pos += 4; i += 4;
        } else {
            sm_xor8(S, p.pos, (T[(i >> 2) * 32] >> (8 * (i & 3))) & 0xffu);
            p.pos++; i++;
        }
        if (p.pos == SM_RATE) sm_run_f(S, p);
    }
}
// meta_ad(label), meta_ad(LE32(len), more)
static __device__ __forceinline__ void sm_meta_label_len(uint32_t *S, SmPos &p, uint32_t label, uint32_t len) {
    sm_begin(S, p, FM | FA);
    const uint32_t off = label & 0xffu, ll = label >> 8;
    for (uint32_t i = 0; i < ll; i++) sm_absorb8(S, p, (uint32_t)(uint8_t)c_labels[off + i]);
    for (uint32_t k = 0; k < 4; k++) sm_absorb8(S, p, (len >> (8 * k)) & 0xffu);
}
// Transcript::append_message(label, staged[0..n))
static __device__ __noinline__ void sm_append(uint32_t *S, uint32_t label, uint32_t n) {
    SmPos p = sm_get(S);
    sm_meta_label_len(S, p, label, n);
    sm_begin(S, p, FA);
    sm_absorb_staged(S, p, n);
    sm_put(S, p);
}
// Transcript::challenge_bytes(label, 4 * nwords) into the staging area; with label == 0xffffffff: TranscriptRng::fill_bytes
static __device__ __noinline__ void sm_challenge(uint32_t *S, uint32_t label, uint32_t nwords) {
    SmPos p = sm_get(S);
    if (label != 0xffffffffu) sm_meta_label_len(S, p, label, 4 * nwords);
    else {
        sm_begin(S, p, FM | FA);
        for (uint32_t k = 0; k < 4; k++) sm_absorb8(S, p, ((4 * nwords) >> (8 * k)) & 0xffu);
    }
    sm_begin(S, p, FI | FA | FC);              // forces a permutation: the squeeze starts word-aligned at position 0
    uint32_t *T = S + SM_STAGE * 32;
    for (uint32_t i = 0; i < nwords; i++) {
        if ((p.pos & 3) == 0 && p.pos + 4 <= SM_RATE) {
            T[i * 32] = S[(p.pos >> 2) * 32];
            S[(p.pos >> 2) * 32] = 0;
            p.pos += 4;
            if (p.pos == SM_RATE) sm_run_f(S, p);
        } else {
            uint32_t v = 0;
            for (uint32_t k = 0; k < 4; k++) {
                const uint32_t sh = 8 * (p.pos & 3);
                v |= ((S[(p.pos >> 2) * 32] >> sh) & 0xffu) << (8 * k);
                S[(p.pos >> 2) * 32] &= ~(0xffu << sh);
                if (++p.pos == SM_RATE) sm_run_f(S, p);
            }
            T[i * 32] = v;
        }
    }
    sm_put(S, p);
}
// TranscriptRngBuilder::finalize(NullRng) on the transcript's own state: meta_ad("rng"), key(32 zero bytes)
static __device__ __noinline__ void sm_rng_finalize_null(uint32_t *S) {
    SmPos p = sm_get(S);
    sm_begin(S, p, FM | FA);
    for (uint32_t i = 0; i < 3; i++) sm_absorb8(S, p, (uint32_t)(uint8_t)c_labels[(LB_RNG & 0xffu) + i]);
    sm_begin(S, p, FA | FC);
    for (uint32_t i = 0; i < 32; i++) {
        S[(p.pos >> 2) * 32] &= ~(0xffu << (8 * (p.pos & 3)));
        if (++p.pos == SM_RATE) sm_run_f(S, p);
    }
    sm_put(S, p);
}

static __device__ __forceinline__ bool words_zero(const uint32_t (&w)[8]) {
    return (w[0] | w[1] | w[2] | w[3] | w[4] | w[5] | w[6] | w[7]) == 0;
}
static __device__ __forceinline__ void st_sc_bytes(uint8_t *dst, const sc &v) {       // dst is 32-byte aligned
    reinterpret_cast<uint4 *>(dst)[0] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
    reinterpret_cast<uint4 *>(dst)[1] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
}
// 32 bytes at src (any alignment) -> staging area; false when they are all zero (the identity encoding)
static __device__ __forceinline__ bool sm_stage32(uint32_t *S, const uint8_t *src) {
    uint32_t w[8];
    ld32_unaligned(src, w);
    uint32_t *T = S + SM_STAGE * 32;
#pragma unroll
    for (int i = 0; i < 8; i++) T[i * 32] = w[i];
    return !words_zero(w);
}
static __device__ __forceinline__ void sm_append_u64(uint32_t *S, uint32_t label, uint64_t x) {
    uint32_t *T = S + SM_STAGE * 32;
    T[0] = (uint32_t)x; T[32] = (uint32_t)(x >> 32);
    sm_append(S, label, 8);
}
// challenge_scalar (transcript_protocol.rs:67-78): 64 bytes -> Scalar::from_bytes_mod_order_wide; false when zero
static __device__ __forceinline__ bool sm_challenge_scalar(uint32_t *S, uint32_t label, sc &out) {
    sm_challenge(S, label, 16);
    const uint32_t *T = S + SM_STAGE * 32;
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = T[i * 32];
    out = sc_from_wide_words(w);
    return !sc_is_zero(out);
}

// the 203-byte wire form of this thread's sponge <-> global memory at an arbitrary byte address
static __device__ __noinline__ void sm_load_state(uint32_t *S, const uint8_t *src) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(src);
    const uint32_t *q = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3u) * 8u;
    uint32_t prev = q[0];
    for (int i = 0; i < 51; i++) {                  // 51 words cover bytes 0..203; the section is padded, the tail masked
        const uint32_t next = q[i + 1];
        S[i * 32] = __funnelshift_r(prev, next, sh);
        prev = next;
    }
    S[50 * 32] &= 0x00ffffffu;
}
static __device__ __noinline__ void sm_store_state(const uint32_t *S, uint8_t *dst) {
    // bytes of neighbouring proofs share aligned words at both ends: the unaligned head and tail go out bytewise
    const uintptr_t a = reinterpret_cast<uintptr_t>(dst);
    const int head = (int)((4u - (uint32_t)(a & 3u)) & 3u);
    int b = 0;
    for (; b < head; b++) dst[b] = (uint8_t)(S[(b >> 2) * 32] >> (8 * (b & 3)));
    const uint32_t sh = 8u * (uint32_t)(b & 3);
    for (; b + 4 <= BPP_TSTATE_BYTES; b += 4) {
        const int wi = b >> 2;
        const uint32_t w0 = S[wi * 32], w1 = wi + 1 <= 50 ? S[(wi + 1) * 32] : 0u;
        *reinterpret_cast<uint32_t *>(dst + b) = __funnelshift_r(w0, w1, sh);
    }
    for (; b < BPP_TSTATE_BYTES; b++) dst[b] = (uint8_t)(S[(b >> 2) * 32] >> (8 * (b & 3)));
}

} // namespace

__global__ void __launch_bounds__(32) k_replay_sm(VDims d, RBuffers b) {
    __shared__ uint32_t s_state[SM_TOTAL * 32];
    const uint32_t p = blockIdx.x * 32u + threadIdx.x;
    if (p >= d.n_proofs) return;
    const VProof pr = b.proofs[p];
    if (!pr.replay) { b.flags[p] = 0; return; }
    uint32_t *S = s_state + threadIdx.x;
    sm_load_state(S, b.tstates_in + BPP_TSTATE_BYTES * (size_t)pr.ts_idx);
    const uint32_t ext = d.ext, R = pr.rounds;
    const uint8_t *raw = b.blob + pr.raw_off;
    uint8_t *ch = b.challenges + 32 * (size_t)pr.ch_off;
    int rc = 1;
    sc y = sc_zero();
    // RangeProofTranscript::new (transcripts.rs:59-121)
    {
        uint32_t *T = S + SM_STAGE * 32;            // "Bulletproofs+ Range Proof": 25 bytes through the staging area
        for (int i = 0; i < 7; i++) {
            uint32_t v = 0;
            for (int k = 0; k < 4; k++) if (4 * i + k < 25) v |= (uint32_t)(uint8_t)c_labels[(LB_PROTO & 0xffu) + 4 * i + k] << (8 * k);
            T[i * 32] = v;
        }
        sm_append(S, LB_DOMSEP, 25);
    }
    do {
        if (!sm_stage32(S, b.hg32)) break;                         // validate_and_append_point: identity -> VerificationFailed
        sm_append(S, LB_H, 32);
        bool ok = true;
        for (uint32_t k = 0; k < ext && ok; k++) {
            ok = sm_stage32(S, b.hg32 + 32 * (1 + k));
            if (ok) sm_append(S, LB_G, 32);
        }
        if (!ok) break;
        sm_append_u64(S, LB_N, (uint64_t)d.bit_length);
        sm_append_u64(S, LB_T, (uint64_t)ext);
        sm_append_u64(S, LB_M, (uint64_t)pr.m);
        for (uint32_t j = 0; j < pr.m; j++) {
            sm_stage32(S, b.commitments32 + 32 * (size_t)(pr.commit_off + j));
            sm_append(S, LB_CI, 32);                               // append_point: not validated
        }
        for (uint32_t j = 0; j < pr.m; j++)
            sm_append_u64(S, LB_VI, b.min_present[pr.commit_off + j] ? b.min_values[pr.commit_off + j] : 0ull);
        // challenges_y_z (:124-136)
        if (!sm_stage32(S, raw + BPP_RAW_A(ext))) break;
        sm_append(S, LB_A, 32);
        sc z, e;
        if (!sm_challenge_scalar(S, LB_Y, y)) break;
        st_sc_bytes(ch, y);
        if (!sm_challenge_scalar(S, LB_Z, z)) break;
        st_sc_bytes(ch + 32, z);
        // challenge_round_e (:139-149)
        for (uint32_t j = 0; j < R && ok; j++) {
            ok = sm_stage32(S, raw + BPP_RAW_L(ext, j));
            if (!ok) break;
            sm_append(S, LB_L, 32);
            ok = sm_stage32(S, raw + BPP_RAW_R(ext, j));
            if (!ok) break;
            sm_append(S, LB_R, 32);
            ok = sm_challenge_scalar(S, LB_E, e);
            if (ok) st_sc_bytes(ch + 32 * (3 + j), e);
        }
        if (!ok) break;
        // challenge_final_e (:152-162)
        if (!sm_stage32(S, raw + BPP_RAW_A(ext) + 32)) break;
        sm_append(S, LB_A1, 32);
        if (!sm_stage32(S, raw + BPP_RAW_A(ext) + 64)) break;
        sm_append(S, LB_B, 32);
        if (!sm_challenge_scalar(S, LB_E, e)) break;
        st_sc_bytes(ch + 64, e);
        // to_verifier_rng (:166-179)
        sm_stage32(S, raw + BPP_RAW_R1(ext));
        sm_append(S, LB_R1, 32);
        sm_stage32(S, raw + BPP_RAW_S1(ext));
        sm_append(S, LB_S1, 32);
        for (uint32_t k = 0; k < ext; k++) {
            sm_stage32(S, raw + BPP_RAW_D1(ext, k));
            sm_append(S, LB_D1, 32);
        }
        rc = 0;
    } while (0);
    // what `&mut Transcript` holds after the call (also when loop 1 failed at this proof)
    sm_store_state(S, b.tstates_out + BPP_TSTATE_BYTES * (size_t)p);
    uint8_t flag = rc ? 1 : 0;
    if (!rc) {
        // build_rng().finalize(NullRng) on a clone of the transcript (the state above is already saved, so it is consumed in place),
        // then fill_bytes(32) (range_proof.rs:845-849)
        sm_rng_finalize_null(S);
        sm_challenge(S, 0xffffffffu, 8);
        const uint32_t *T = S + SM_STAGE * 32;
        uint8_t *wb = b.wbytes + 32 * (size_t)p;
        reinterpret_cast<uint4 *>(wb)[0] = make_uint4(T[0], T[32], T[64], T[96]);
        reinterpret_cast<uint4 *>(wb)[1] = make_uint4(T[128], T[160], T[192], T[224]);
        // y == 1 makes (y - 1) non-invertible: treated as a failed verification (see engine_verify.cu)
        if (y.v[0] == 1u && (y.v[1] | y.v[2] | y.v[3] | y.v[4] | y.v[5] | y.v[6] | y.v[7]) == 0u) flag |= 2;
    }
    b.flags[p] = flag;
}

constexpr int REPLAY_WARPS = 4;
// the round-1 kernels over hash.cuh / wstrobe.cuh (replay.cuh is the statement-level restatement they share with the host path).
// WARP = true: one warp per proof; false: one thread per proof, sponge state in local memory
template <bool WARP> __global__ void __launch_bounds__(WARP ? 32 * REPLAY_WARPS : 64) k_replay(VDims d, RBuffers b) {
    const uint32_t p = WARP ? blockIdx.x * REPLAY_WARPS + (threadIdx.x >> 5) : blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= d.n_proofs) return;
    const VProof pr = b.proofs[p];
    if (!pr.replay) { b.flags[p] = 0; return; }
    const uint8_t *raw = b.blob + pr.raw_off;
    ReplayIn in;
    in.tstate = b.tstates_in + BPP_TSTATE_BYTES * (size_t)pr.ts_idx;
    in.h32 = b.hg32; in.g32 = b.hg32 + 32;
    in.bit_length = d.bit_length; in.ext = d.ext; in.m = pr.m; in.rounds = pr.rounds;
    in.commitments32 = b.commitments32 + 32 * (size_t)pr.commit_off;
    in.min_values = b.min_values + pr.commit_off; in.min_present = b.min_present + pr.commit_off;
    in.a = raw + BPP_RAW_A(d.ext); in.a1 = in.a + 32; in.b = in.a + 64;
    in.l_base = raw + BPP_RAW_L(d.ext, 0); in.r_base = raw + BPP_RAW_R(d.ext, 0); in.lr_stride = 64;
    in.r1 = raw + BPP_RAW_R1(d.ext); in.s1 = raw + BPP_RAW_S1(d.ext); in.d1 = raw + BPP_RAW_D1(d.ext, 0);
    uint8_t *ch = b.challenges + 32 * (size_t)pr.ch_off;
    ReplayOut out;
    out.y = ch; out.z = ch + 32; out.e = ch + 64; out.ej = ch + 96;
    out.wbytes = b.wbytes + 32 * (size_t)p;
    out.tstate = b.tstates_out + BPP_TSTATE_BYTES * (size_t)p;
    int rc;
    if constexpr (WARP) rc = replay_transcript_core_t<WMerlin, WMerlinRng>(in, out);
    else rc = replay_transcript_core_t<Merlin, MerlinRng>(in, out);
    uint8_t flag = rc ? 1 : 0;
    if (!rc) {          // y == 1 makes (y - 1) non-invertible: treated as a failed verification (see engine_verify.cu)
        uint8_t acc = out.y[0] ^ 1;
        for (int i = 1; i < 32; i++) acc |= out.y[i];
        if (acc == 0) flag |= 2;
    }
    b.flags[p] = flag;
}

// K-WEIGHTS: the verifier-weight transcript of each chunk (/root/reference/src/range_proof.rs:811, :849, :853, :894) on the
// device, one WARP per chunk with the warp-cooperative sponge of wstrobe.cuh.  The transcript is inherently sequential (every
// proof appends 32 bytes, then one Keccak-f per weight drawn), so this is a ~330-permutation chain per 256-proof chunk: slower
// than a host core for one batch alone (~0.6 ms against ~0.25 ms), but it removes the only host step from the middle of a pass,
// so a pass becomes ONE graph launch and costs the host nothing -- what matters when many passes are in flight.
__global__ void __launch_bounds__(32) k_weights(VDims d, const VChunk *__restrict__ chunks, const uint8_t *__restrict__ wt_init,
                                               const uint8_t *__restrict__ wbytes, const uint8_t *__restrict__ flags, uint32_t *__restrict__ weights) {
    const VChunk chk = chunks[blockIdx.x];
    if (!chk.active) return;
    const int lane = threadIdx.x & 31;
    int bad = 0;
    for (uint32_t p = chk.proof_lo + lane; p < chk.proof_hi; p += 32) bad |= flags[p] & 1;
    if (__any_sync(0xffffffffu, bad)) return;            // loop 1 failed somewhere in the call: it ends there, no weights are drawn
    WMerlin wt;
    wt.s.load(wt_init);                                   // Transcript::new("Bulletproofs+ verifier weights"), state from the host
    const uint8_t lbl[5] = {'p', 'r', 'o', 'o', 'f'};
    for (uint32_t p = chk.proof_lo; p < chk.proof_hi; p++) wt.append_message(lbl, 5, wbytes + 32 * (size_t)p, 32);
    WMerlinRng wr;
    uint8_t zeros[32];
    for (int i = 0; i < 32; i++) zeros[i] = 0;
    wr.build(wt, nullptr, 0, false, zeros);               // NullRng
    for (uint32_t p = chk.proof_lo; p < chk.proof_hi; p++) {
        sc w;
        do {                                              // Scalar::random_not_zero (:894)
            __align__(8) uint8_t wide[64];
            wr.fill(wide, 64);
            uint32_t ww[16];
            for (int i = 0; i < 16; i++)
                ww[i] = (uint32_t)wide[4 * i] | ((uint32_t)wide[4 * i + 1] << 8) | ((uint32_t)wide[4 * i + 2] << 16) | ((uint32_t)wide[4 * i + 3] << 24);
            w = sc_from_wide_words(ww);
        } while (sc_is_zero(w));                          // warp-uniform
        if (lane < 16) weights[16 * (size_t)p + lane] = lane < 8 ? w.v[lane & 7] : 0u;
    }
    if (d.merged) {                                       // rho_c of the merged check: the next value of the same rng
        sc w;
        do {
            __align__(8) uint8_t wide[64];
            wr.fill(wide, 64);
            uint32_t ww[16];
            for (int i = 0; i < 16; i++)
                ww[i] = (uint32_t)wide[4 * i] | ((uint32_t)wide[4 * i + 1] << 8) | ((uint32_t)wide[4 * i + 2] << 16) | ((uint32_t)wide[4 * i + 3] << 24);
            w = sc_from_wide_words(ww);
        } while (sc_is_zero(w));
        if (lane < 16) weights[16 * ((size_t)d.n_proofs + blockIdx.x) + lane] = lane < 8 ? w.v[lane & 7] : 0u;
    }
}

// The same with the shared-memory sponge of k_replay_sm: one THREAD per chunk (the chain is sequential whatever is done to it; a lone
// warp runs a permutation in ~4.8 us with the state of 32 chunks in flight, the warp-cooperative form above needs ~13 us for one).
// ~331 permutations per 256-proof chunk = ~1.6 ms, next to the decompression and the weight-free scalar prep of the same pass.
__global__ void __launch_bounds__(32) k_weights_sm(VDims d, const VChunk *__restrict__ chunks, const uint8_t *__restrict__ wt_init,
                                                  const uint8_t *__restrict__ wbytes, const uint8_t *__restrict__ flags, uint32_t *__restrict__ weights) {
    __shared__ uint32_t s_state[SM_TOTAL * 32];
    const uint32_t c = blockIdx.x * 32u + threadIdx.x;
    if (c >= d.n_chunks) return;
    const VChunk chk = chunks[c];
    if (!chk.active) return;
    for (uint32_t p = chk.proof_lo; p < chk.proof_hi; p++)
        if (flags[p] & 1) return;                         // loop 1 failed somewhere in the call: it ends there, no weights are drawn
    uint32_t *S = s_state + threadIdx.x, *T = S + SM_STAGE * 32;
    sm_load_state(S, wt_init);                            // Transcript::new("Bulletproofs+ verifier weights"), state from the host
    for (uint32_t p = chk.proof_lo; p < chk.proof_hi; p++) {
        const uint4 *src = reinterpret_cast<const uint4 *>(wbytes + 32 * (size_t)p);
        const uint4 a = src[0], b = src[1];
        T[0] = a.x; T[32] = a.y; T[64] = a.z; T[96] = a.w; T[128] = b.x; T[160] = b.y; T[192] = b.z; T[224] = b.w;
        sm_append(S, LB_PROOF, 32);
    }
    sm_rng_finalize_null(S);                              // build_rng().finalize(&mut NullRng)
    for (uint32_t p = chk.proof_lo; p < chk.proof_hi; p++) {
        sc w;
        do {                                              // Scalar::random_not_zero (:894)
            sm_challenge(S, 0xffffffffu, 16);
            uint32_t ww[16];
#pragma unroll
            for (int i = 0; i < 16; i++) ww[i] = T[i * 32];
            w = sc_from_wide_words(ww);
        } while (sc_is_zero(w));
        uint4 *dst = reinterpret_cast<uint4 *>(weights + 16 * (size_t)p);
        dst[0] = make_uint4(w.v[0], w.v[1], w.v[2], w.v[3]);
        dst[1] = make_uint4(w.v[4], w.v[5], w.v[6], w.v[7]);
        dst[2] = make_uint4(0, 0, 0, 0);
        dst[3] = make_uint4(0, 0, 0, 0);
    }
    if (d.merged) {                                       // rho_c of the merged check: the next value of the same rng
        sc w;
        do {
            sm_challenge(S, 0xffffffffu, 16);
            uint32_t ww[16];
#pragma unroll
            for (int i = 0; i < 16; i++) ww[i] = T[i * 32];
            w = sc_from_wide_words(ww);
        } while (sc_is_zero(w));
        uint4 *dst = reinterpret_cast<uint4 *>(weights + 16 * ((size_t)d.n_proofs + c));
        dst[0] = make_uint4(w.v[0], w.v[1], w.v[2], w.v[3]);
        dst[1] = make_uint4(w.v[4], w.v[5], w.v[6], w.v[7]);
        dst[2] = make_uint4(0, 0, 0, 0);
        dst[3] = make_uint4(0, 0, 0, 0);
    }
}

void launch_weights(cudaStream_t s, const VDims &d, const VChunk *chunks, const uint8_t *wt_init, const uint8_t *wbytes, const uint8_t *flags,
                    uint32_t *weights, uint64_t *launches) {
    if (d.n_chunks == 0) return;
    static const bool warp_form = [] { const char *e = getenv("BPP_WEIGHTS_WARP"); return e && atoi(e) != 0; }();
    if (warp_form) k_weights<<<d.n_chunks, 32, 0, s>>>(d, chunks, wt_init, wbytes, flags, weights);
    else k_weights_sm<<<(d.n_chunks + 31) / 32, 32, 0, s>>>(d, chunks, wt_init, wbytes, flags, weights);
    if (launches) (*launches)++;
}

void launch_replay(cudaStream_t s, const VDims &d, const RBuffers &b, int kernel, uint64_t *launches) {
    if (d.n_proofs == 0) return;
    if (kernel == 2) k_replay<true><<<(d.n_proofs + REPLAY_WARPS - 1) / REPLAY_WARPS, 32 * REPLAY_WARPS, 0, s>>>(d, b);
    else if (kernel == 1) k_replay<false><<<(d.n_proofs + 63) / 64, 64, 0, s>>>(d, b);
    else k_replay_sm<<<(d.n_proofs + 31) / 32, 32, 0, s>>>(d, b);
    if (launches) (*launches)++;
}

} // namespace bpp
