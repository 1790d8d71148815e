// Loop 1 of RangeProof::verify for ONE proof (/root/reference/src/range_proof.rs:816-850): replay of its Fiat-Shamir
// transcript.  Shared by the host threads (engine_verify.cu) and the device kernel (k_replay.cu) so that both run the
// same statement-by-statement restatement of
//   RangeProofTranscript::new          /root/reference/src/transcripts.rs:59-121
//   challenges_y_z                     :124-136     challenge_round_e  :139-149     challenge_final_e  :152-162
//   to_verifier_rng                    :166-179     + 32 rng bytes for the weight transcript (range_proof.rs:845-849)
//   validate_and_append_point / challenge_scalar   /root/reference/src/protocols/transcript_protocol.rs:49-78
// The intermediate TranscriptRng rebuilds of the reference (transcripts.rs:185-194) clone the transcript and do not feed
// back into it, so only the last one (after r1, s1, d1) is materialised.
#pragma once
#include "arith.cuh"
#include "hash.cuh"

namespace bpp {

struct ReplayIn {
    const uint8_t *tstate;          // 203 B, merlin state of the caller's transcript before the call
    const uint8_t *h32;             // compressed H
    const uint8_t *g32;             // ext x 32: compressed G[k]
    uint32_t bit_length, ext, m, rounds;
    const uint8_t *commitments32;   // m x 32
    const uint64_t *min_values;     // m
    const uint8_t *min_present;     // m
    const uint8_t *a, *a1, *b;      // 32 B encodings
    const uint8_t *l_base, *r_base; // L_j = l_base + j * lr_stride, R_j = r_base + j * lr_stride
    uint32_t lr_stride;
    const uint8_t *r1, *s1, *d1;    // canonical scalars; d1: ext x 32
};
struct ReplayOut {
    uint8_t *y, *z, *e;             // 32 B each
    uint8_t *ej;                    // rounds x 32
    uint8_t *wbytes;                // 32 B for the weight transcript
    uint8_t *tstate;                // 203 B, advanced state
};

BPP_HASH_HD inline bool replay_is_zero32(const uint8_t *p) {
    uint8_t r = 0;
    for (int i = 0; i < 32; i++) r |= p[i];
    return r == 0;
}
// 64 challenge bytes -> Scalar::from_bytes_mod_order_wide; false if the scalar is zero
template <class TM> BPP_HASH_HD BPP_HASH_NOINLINE inline bool replay_challenge(TM &t, const uint8_t *label, size_t ll, uint8_t out32[32]) {
#if defined(__CUDA_ARCH__)
    __align__(8) uint8_t buf[64];
#else
    uint8_t buf[64];
#endif
    t.challenge_bytes(label, ll, buf, 64);
    uint32_t w[16];
    for (int i = 0; i < 16; i++)
        w[i] = (uint32_t)buf[4 * i] | ((uint32_t)buf[4 * i + 1] << 8) | ((uint32_t)buf[4 * i + 2] << 16) | ((uint32_t)buf[4 * i + 3] << 24);
    sc r = sc_from_wide_words(w);
    sc_tobytes(out32, r);
    return !sc_is_zero(r);
}
template <class TM> BPP_HASH_HD BPP_HASH_NOINLINE inline bool replay_point(TM &t, const uint8_t *label, size_t ll, const uint8_t pt[32]) {
    if (replay_is_zero32(pt)) return false;      // identity encoding: ProofError::VerificationFailed
    t.append_message(label, ll, pt, 32);
    return true;
}
#define BPP_LBL(s) (const uint8_t *)(s), (sizeof(s) - 1)

// returns 0 (BPP_OK) or 1 (BPP_VERIFICATION_FAILED); the transcript state is written back in both cases.
// TM / TR = Merlin / MerlinRng (one thread per transcript: host threads) or WMerlin / WMerlinRng (one warp per transcript:
// k_replay.cu; every lane runs this function with identical arguments and writes identical outputs).
template <class TM, class TR> BPP_HASH_HD inline int replay_transcript_core_t(const ReplayIn &in, const ReplayOut &out) {
    TM t;
    t.s.load(in.tstate);
    int rc = 1;
    t.append_message(BPP_LBL("dom-sep"), BPP_LBL("Bulletproofs+ Range Proof"));
    do {
        if (!replay_point(t, BPP_LBL("H"), in.h32)) break;
        bool ok = true;
        for (uint32_t k = 0; k < in.ext && ok; k++) ok = replay_point(t, BPP_LBL("G"), in.g32 + 32 * k);
        if (!ok) break;
        t.append_u64(BPP_LBL("N"), (uint64_t)in.bit_length);
        t.append_u64(BPP_LBL("T"), (uint64_t)in.ext);
        t.append_u64(BPP_LBL("M"), (uint64_t)in.m);
        for (uint32_t j = 0; j < in.m; j++) t.append_message(BPP_LBL("Ci"), in.commitments32 + 32 * j, 32);
        for (uint32_t j = 0; j < in.m; j++) t.append_u64(BPP_LBL("vi - minimum_value"), in.min_present[j] ? in.min_values[j] : 0);
        if (!replay_point(t, BPP_LBL("A"), in.a)) break;
        if (!replay_challenge(t, BPP_LBL("y"), out.y)) break;
        if (!replay_challenge(t, BPP_LBL("z"), out.z)) break;
        for (uint32_t j = 0; j < in.rounds && ok; j++) {
            ok = replay_point(t, BPP_LBL("L"), in.l_base + (size_t)j * in.lr_stride) &&
                 replay_point(t, BPP_LBL("R"), in.r_base + (size_t)j * in.lr_stride) &&
                 replay_challenge(t, BPP_LBL("e"), out.ej + 32 * j);
        }
        if (!ok) break;
        if (!replay_point(t, BPP_LBL("A1"), in.a1)) break;
        if (!replay_point(t, BPP_LBL("B"), in.b)) break;
        if (!replay_challenge(t, BPP_LBL("e"), out.e)) break;
        t.append_message(BPP_LBL("r1"), in.r1, 32);
        t.append_message(BPP_LBL("s1"), in.s1, 32);
        for (uint32_t k = 0; k < in.ext; k++) t.append_message(BPP_LBL("d1"), in.d1 + 32 * k, 32);
        TR rng;
        uint8_t zeros[32];
        for (int i = 0; i < 32; i++) zeros[i] = 0;
        rng.build(t, nullptr, 0, false, zeros);       // NullRng (/root/reference/src/utils/nullrng.rs:16-40)
        rng.fill(out.wbytes, 32);
        rc = 0;
    } while (0);
    t.s.store(out.tstate);
    return rc;
}
BPP_HASH_HD inline int replay_transcript_core(const ReplayIn &in, const ReplayOut &out) {
    return replay_transcript_core_t<Merlin, MerlinRng>(in, out);
}

} // namespace bpp
