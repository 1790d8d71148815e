//! Loop 1 of `RangeProof::verify` on the host, over stock `merlin::Transcript` -- the part BASELINE.json's north_star leaves with
//! the Rust host.  This is `RangeProofTranscript` of the reference (/root/reference/src/transcripts.rs:59-179; it is `pub(crate)`
//! there, so its operation order is restated here) followed by the verifier-weight transcript of
//! /root/reference/src/range_proof.rs:811-853 and the weight draws of :894.
use curve25519_dalek::{ristretto::CompressedRistretto, scalar::Scalar};
use merlin::Transcript;
use rand_core::{CryptoRng, RngCore};
use tari_bulletproofs_plus::{errors::ProofError, ristretto::RistrettoRangeProof};

/// /root/reference/src/utils/nullrng.rs:16-40
pub(crate) struct NullRng;
impl RngCore for NullRng {
    fn next_u32(&mut self) -> u32 {
        0
    }
    fn next_u64(&mut self) -> u64 {
        0
    }
    fn fill_bytes(&mut self, dest: &mut [u8]) {
        dest.fill(0);
    }
    fn try_fill_bytes(&mut self, dest: &mut [u8]) -> Result<(), rand_core::Error> {
        self.fill_bytes(dest);
        Ok(())
    }
}
impl CryptoRng for NullRng {}

fn validate_and_append_point(t: &mut Transcript, label: &'static [u8], p: &[u8; 32]) -> Result<(), ProofError> {
    // src/protocols/transcript_protocol.rs:49-61
    if p.iter().all(|b| *b == 0) {
        Err(ProofError::VerificationFailed("Identity element cannot be added to the transcript".to_string()))
    } else {
        t.append_message(label, p);
        Ok(())
    }
}

fn challenge_scalar(t: &mut Transcript, label: &'static [u8]) -> Result<Scalar, ProofError> {
    // src/protocols/transcript_protocol.rs:67-78
    let mut buf = [0u8; 64];
    t.challenge_bytes(label, &mut buf);
    let value = Scalar::from_bytes_mod_order_wide(&buf);
    if value == Scalar::ZERO {
        Err(ProofError::VerificationFailed("Transcript challenge cannot be zero".to_string()))
    } else {
        Ok(value)
    }
}

/// Scalar::random_not_zero (src/protocols/scalar_protocol.rs:23-30)
pub(crate) fn random_not_zero<R: RngCore + CryptoRng>(rng: &mut R) -> Scalar {
    let mut value = Scalar::ZERO;
    while value == Scalar::ZERO {
        value = Scalar::random(rng);
    }
    value
}

/// the fields of a serialised proof (`to_bytes` layout, src/range_proof.rs:1120-1150): [ext] d1[ext] a a1 b r1 s1 (L R)*
pub(crate) struct ProofView<'a> {
    pub bytes: &'a [u8],
    pub ext: usize,
    pub rounds: usize,
}
impl<'a> ProofView<'a> {
    pub fn new(bytes: &'a [u8]) -> Self {
        let ext = bytes[0] as usize;
        let rounds = (bytes.len() - 1 - 32 * (ext + 5)) / 64;
        ProofView { bytes, ext, rounds }
    }
    fn el(&self, i: usize) -> &[u8; 32] {
        self.bytes[1 + 32 * i..1 + 32 * (i + 1)].try_into().unwrap()
    }
    pub fn d1(&self, k: usize) -> &[u8; 32] {
        self.el(k)
    }
    pub fn a(&self) -> &[u8; 32] {
        self.el(self.ext)
    }
    pub fn a1(&self) -> &[u8; 32] {
        self.el(self.ext + 1)
    }
    pub fn b(&self) -> &[u8; 32] {
        self.el(self.ext + 2)
    }
    pub fn r1(&self) -> &[u8; 32] {
        self.el(self.ext + 3)
    }
    pub fn s1(&self) -> &[u8; 32] {
        self.el(self.ext + 4)
    }
    pub fn l(&self, j: usize) -> &[u8; 32] {
        self.el(self.ext + 5 + 2 * j)
    }
    pub fn r(&self, j: usize) -> &[u8; 32] {
        self.el(self.ext + 6 + 2 * j)
    }
}

/// One proof's transcript replay: appends in the reference's order, returns `[y, z, e, e_0 .. e_{rounds-1}]` (the layout
/// `bpp_verify_chunks_ch` takes) and the 32 bytes the proof feeds into the weight transcript.
#[allow(clippy::too_many_arguments)]
pub(crate) fn replay_one(
    transcript: &mut Transcript,
    h_base_compressed: &CompressedRistretto,
    g_bases_compressed: &[CompressedRistretto],
    bit_length: usize,
    commitments_compressed: &[CompressedRistretto],
    minimum_value_promises: &[Option<u64>],
    proof: &ProofView<'_>,
) -> Result<(Vec<Scalar>, [u8; 32]), ProofError> {
    // RangeProofTranscript::new (transcripts.rs:59-121), verifier side: no witness, NullRng
    transcript.append_message(b"dom-sep", b"Bulletproofs+ Range Proof");
    validate_and_append_point(transcript, b"H", h_base_compressed.as_bytes())?;
    for g in g_bases_compressed {
        validate_and_append_point(transcript, b"G", g.as_bytes())?;
    }
    transcript.append_u64(b"N", bit_length as u64);
    transcript.append_u64(b"T", g_bases_compressed.len() as u64);
    transcript.append_u64(b"M", commitments_compressed.len() as u64);
    for c in commitments_compressed {
        transcript.append_message(b"Ci", c.as_bytes());
    }
    for m in minimum_value_promises {
        transcript.append_u64(b"vi - minimum_value", m.unwrap_or(0));
    }
    // challenges_y_z (:124-136); the intermediate TranscriptRng rebuilds work on clones and do not feed back
    validate_and_append_point(transcript, b"A", proof.a())?;
    let y = challenge_scalar(transcript, b"y")?;
    let z = challenge_scalar(transcript, b"z")?;
    // challenge_round_e (:139-149)
    let mut round_e = Vec::with_capacity(proof.rounds);
    for j in 0..proof.rounds {
        validate_and_append_point(transcript, b"L", proof.l(j))?;
        validate_and_append_point(transcript, b"R", proof.r(j))?;
        round_e.push(challenge_scalar(transcript, b"e")?);
    }
    // challenge_final_e (:152-162)
    validate_and_append_point(transcript, b"A1", proof.a1())?;
    validate_and_append_point(transcript, b"B", proof.b())?;
    let e = challenge_scalar(transcript, b"e")?;
    // to_verifier_rng (:166-179) and the 32 bytes for the weight transcript (range_proof.rs:845-849)
    transcript.append_message(b"r1", proof.r1());
    transcript.append_message(b"s1", proof.s1());
    for k in 0..proof.ext {
        transcript.append_message(b"d1", proof.d1(k));
    }
    let mut rng = transcript.build_rng().finalize(&mut NullRng);
    let mut wbytes = [0u8; 32];
    rng.fill_bytes(&mut wbytes);
    let mut challenges = Vec::with_capacity(3 + proof.rounds);
    challenges.push(y);
    challenges.push(z);
    challenges.push(e);
    challenges.extend(round_e);
    Ok((challenges, wbytes))
}

/// range_proof.rs:811, :849, :853, :894 -- one weight per proof of the batch, drawn from the finalised weight transcript
pub(crate) fn batch_weights(wbytes: &[[u8; 32]]) -> Vec<Scalar> {
    let mut weight_transcript = Transcript::new(b"Bulletproofs+ verifier weights");
    for w in wbytes {
        weight_transcript.append_message(b"proof", w);
    }
    let mut rng = weight_transcript.build_rng().finalize(&mut NullRng);
    wbytes.iter().map(|_| random_not_zero(&mut rng)).collect()
}

#[allow(dead_code)]
pub(crate) fn proof_bytes(p: &RistrettoRangeProof) -> Vec<u8> {
    p.to_bytes()
}
