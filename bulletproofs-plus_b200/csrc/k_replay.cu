// K-REPLAY: loop 1 of the batch verifier on the device -- one WARP replays one proof's Merlin transcript
// (/root/reference/src/range_proof.rs:816-850; statement-level restatement in replay.cuh, shared with the host path).
// STROBE-128 / Keccak-f[1600] is warp-cooperative (wstrobe.cuh: lane L holds Keccak lane L, a round is 9 64-bit shuffles +
// ~25 ALU instructions per lane); ~21 permutations per 64-bit proof.  Outputs the
// Fiat-Shamir challenges (canonical scalars) straight into the buffer K-VPREP reads, the 32 bytes each proof feeds into
// the verifier-weight transcript, the advanced transcript states and per-proof failure flags.
#include "kernels.cuh"
#include "replay.cuh"
#include "wstrobe.cuh"

namespace bpp {

constexpr int REPLAY_WARPS = 4;
// WARP = true: one warp per proof (shortest dependent chain: small batches); false: one thread per proof (14x fewer warp
// instructions per proof: large or concurrent batches)
template <bool WARP> __global__ void __launch_bounds__(WARP ? 32 * REPLAY_WARPS : 64) k_replay(VDims d, RBuffers b) {
    const uint32_t p = WARP ? blockIdx.x * REPLAY_WARPS + (threadIdx.x >> 5) : blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= d.n_proofs) return;
    const VProof pr = b.proofs[p];
    if (!pr.replay) { b.flags[p] = 0; return; }
    const uint8_t *enc = b.enc + 32 * (size_t)pr.pt_off;
    const uint8_t *ps = b.proof_scalars + 32 * (size_t)pr.sc_off;
    ReplayIn in;
    in.tstate = b.tstates_in + BPP_TSTATE_BYTES * (size_t)p;
    in.h32 = b.hg32; in.g32 = b.hg32 + 32;
    in.bit_length = d.bit_length; in.ext = d.ext; in.m = pr.m; in.rounds = pr.rounds;
    in.commitments32 = enc + 32 * (size_t)(3 + 2 * pr.rounds);
    in.min_values = b.min_values + pr.commit_off; in.min_present = b.min_present + pr.commit_off;
    in.a = enc; in.a1 = enc + 32; in.b = enc + 64;
    in.l_base = enc + 96; in.r_base = enc + 96 + 32 * (size_t)pr.rounds; in.lr_stride = 32;
    in.r1 = ps; in.s1 = ps + 32; in.d1 = ps + 64;
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
        if (lane < 8) weights[8 * (size_t)p + lane] = w.v[lane];
    }
}

void launch_weights(cudaStream_t s, const VDims &d, const VChunk *chunks, const uint8_t *wt_init, const uint8_t *wbytes, const uint8_t *flags,
                    uint32_t *weights, uint64_t *launches) {
    if (d.n_chunks == 0) return;
    k_weights<<<d.n_chunks, 32, 0, s>>>(d, chunks, wt_init, wbytes, flags, weights);
    if (launches) (*launches)++;
}

void launch_replay(cudaStream_t s, const VDims &d, const RBuffers &b, bool warp_per_proof, uint64_t *launches) {
    if (d.n_proofs == 0) return;
    if (warp_per_proof) k_replay<true><<<(d.n_proofs + REPLAY_WARPS - 1) / REPLAY_WARPS, 32 * REPLAY_WARPS, 0, s>>>(d, b);
    else k_replay<false><<<(d.n_proofs + 63) / 64, 64, 0, s>>>(d, b);
    if (launches) (*launches)++;
}

} // namespace bpp
