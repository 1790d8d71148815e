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

void launch_replay(cudaStream_t s, const VDims &d, const RBuffers &b, bool warp_per_proof, uint64_t *launches) {
    if (d.n_proofs == 0) return;
    if (warp_per_proof) k_replay<true><<<(d.n_proofs + REPLAY_WARPS - 1) / REPLAY_WARPS, 32 * REPLAY_WARPS, 0, s>>>(d, b);
    else k_replay<false><<<(d.n_proofs + 63) / 64, 64, 0, s>>>(d, b);
    if (launches) (*launches)++;
}

} // namespace bpp
