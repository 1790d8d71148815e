//! `bpp-b200`: the reference's `RangeProof::prove_with_rng` / `RangeProof::verify_batch` on a B200, behind the reference's API.
//!
//! What runs where (BASELINE.json north_star): the MSM / inner-product arithmetic, Ristretto (de)compression and -- for the
//! exportable [`Transcript`] -- the per-proof transcript replay run in `libbpp_b200.so`; argument checks, proof (de)serialisation
//! and (for stock merlin transcripts) loop 1 stay here.  Results are bit-exact with `tari_bulletproofs_plus` 0.4.1: same proof
//! bytes for the same RNG stream, same accept / reject, same `ProofError` variant, same recovered masks.
//!
//! ```ignore
//! let params = RangeParameters::init(&engine, 64, 1, create_pedersen_gens_with_extension_degree(ExtensionDegree::DefaultPedersen))?;
//! let statement = RangeStatement::init(params.reference().clone(), vec![commitment], vec![Some(min)], Some(seed))?;
//! let proof = RangeProof::prove_with_rng(&mut Transcript::new(b"label"), &params, &statement, &witness, &mut rng)?;
//! let masks = RangeProof::verify_batch(&mut [Transcript::new(b"label")], &params, &[statement], &[proof], VerifyAction::VerifyOnly)?;
//! ```
mod merlin_host;
mod transcript;

use std::ffi::CStr;
use std::ptr;

use bpp_b200_sys as sys;
use curve25519_dalek::{ristretto::RistrettoPoint, scalar::Scalar};
use rand_core::CryptoRngCore;
pub use tari_bulletproofs_plus::{
    errors::ProofError,
    extended_mask::ExtendedMask,
    generators::pedersen_gens::ExtensionDegree,
    range_proof::VerifyAction,
    range_statement::RangeStatement,
    ristretto::{create_pedersen_gens_with_extension_degree, RistrettoRangeProof},
    PedersenGens,
};
pub use transcript::Transcript;
use zeroize::{Zeroize, ZeroizeOnDrop, Zeroizing};

type RefParameters = tari_bulletproofs_plus::range_parameters::RangeParameters<RistrettoPoint>;

/// MAX_RANGE_PROOF_BATCH_SIZE (/root/reference/src/range_proof.rs:76)
pub const MAX_RANGE_PROOF_BATCH_SIZE: usize = 256;

fn status_to_error(code: i32, what: &str) -> ProofError {
    // src/errors.rs:12-28; codes >= 100 are CUDA / runtime failures of the engine (there is no CPU fallback)
    match code {
        sys::BPP_VERIFICATION_FAILED => ProofError::VerificationFailed(what.to_string()),
        sys::BPP_INVALID_ARGUMENT => ProofError::InvalidArgument(what.to_string()),
        sys::BPP_INVALID_LENGTH => ProofError::InvalidLength(what.to_string()),
        sys::BPP_INVALID_BLAKE2B => ProofError::InvalidBlake2b,
        sys::BPP_SIZE_OVERFLOW => ProofError::SizeOverflow,
        _ => ProofError::InvalidArgument(format!("bpp-b200 engine failure {code}: {what}")),
    }
}

/// One `bpp_ctx`: a CUDA stream triple on one device.  Not `Sync`: one engine per host thread (calls on one engine are serialised).
pub struct Engine {
    ctx: *mut sys::bpp_ctx,
}
unsafe impl Send for Engine {}
impl Engine {
    pub fn new(device_ordinal: i32) -> Result<Self, ProofError> {
        let mut ctx = ptr::null_mut();
        let rc = unsafe { sys::bpp_ctx_create(device_ordinal, &mut ctx) };
        if rc != sys::BPP_OK {
            return Err(status_to_error(rc, "no usable CUDA device (bpp_ctx_create)"));
        }
        Ok(Engine { ctx })
    }
    fn last_error(&self) -> String {
        unsafe { CStr::from_ptr(sys::bpp_last_error(self.ctx)) }.to_string_lossy().into_owned()
    }
}
impl Drop for Engine {
    fn drop(&mut self) {
        unsafe { sys::bpp_ctx_destroy(self.ctx) }
    }
}

/// `RangeParameters` with the generator tables resident on the device (`bpp_gens`), next to the reference's own parameter object
/// (kept for `RangeStatement::init`, which wants it).  Must not outlive its [`Engine`].
pub struct RangeParameters<'e> {
    engine: &'e Engine,
    gens: *mut sys::bpp_gens,
    reference: RefParameters,
}
impl<'e> RangeParameters<'e> {
    /// `RangeParameters::init(bit_length, aggregation_factor, pc_gens)` (/root/reference/src/range_parameters.rs:32-58); the Pedersen
    /// bases of `pc_gens` are uploaded as they are (caller-made bases are supported), Gi / Hi are derived on the device.
    pub fn init(engine: &'e Engine, bit_length: usize, aggregation_factor: usize, pc_gens: PedersenGens<RistrettoPoint>) -> Result<Self, ProofError> {
        let reference = RefParameters::init(bit_length, aggregation_factor, pc_gens)?;
        let h = reference.h_base_compressed();
        let g: Vec<u8> = reference.g_bases_compressed().iter().flat_map(|p| p.as_bytes().iter().copied()).collect();
        let mut gens = ptr::null_mut();
        let rc = unsafe {
            sys::bpp_gens_create_with_bases(
                engine.ctx,
                bit_length as i32,
                aggregation_factor as i32,
                reference.extension_degree() as i32,
                h.as_bytes().as_ptr(),
                g.as_ptr(),
                &mut gens,
            )
        };
        if rc != sys::BPP_OK {
            return Err(status_to_error(rc, &engine.last_error()));
        }
        Ok(RangeParameters { engine, gens, reference })
    }
    /// the reference's parameter object (for `RangeStatement::init(params.reference().clone(), ..)`)
    pub fn reference(&self) -> &RefParameters {
        &self.reference
    }
    pub fn bit_length(&self) -> usize {
        self.reference.bit_length()
    }
    pub fn max_aggregation_factor(&self) -> usize {
        self.reference.max_aggregation_factor()
    }
    pub fn extension_degree(&self) -> ExtensionDegree {
        self.reference.extension_degree()
    }
    /// `PedersenGens::commit` for many openings at once (`bpp_pedersen_commit_batch`)
    pub fn commit_batch(&self, openings: &[CommitmentOpening]) -> Result<Vec<curve25519_dalek::ristretto::CompressedRistretto>, ProofError> {
        let nb = openings.first().map(|o| o.r.len()).unwrap_or(1);
        if nb == 0 || nb > self.extension_degree() as usize || openings.iter().any(|o| o.r.len() != nb) {
            return Err(ProofError::InvalidLength("blinding vector".to_string()));
        }
        let values: Vec<u64> = openings.iter().map(|o| o.v).collect();
        let blindings: Zeroizing<Vec<u8>> = Zeroizing::new(openings.iter().flat_map(|o| o.r.iter().flat_map(|s| s.to_bytes())).collect());
        let mut out = vec![0u8; 32 * openings.len()];
        let rc = unsafe { sys::bpp_pedersen_commit_batch(self.gens, openings.len(), values.as_ptr(), blindings.as_ptr(), nb as i32, out.as_mut_ptr()) };
        if rc != sys::BPP_OK {
            return Err(status_to_error(rc, &self.engine.last_error()));
        }
        Ok(out.chunks_exact(32).map(|c| curve25519_dalek::ristretto::CompressedRistretto(c.try_into().unwrap())).collect())
    }
}
impl Drop for RangeParameters<'_> {
    fn drop(&mut self) {
        unsafe { sys::bpp_gens_destroy(self.gens) }
    }
}

/// `CommitmentOpening` (/root/reference/src/commitment_opening.rs:15-38; re-declared because the reference keeps `v` and `r`
/// `pub(crate)`, and the prover has to marshal them)
#[derive(Clone, Zeroize, ZeroizeOnDrop)]
pub struct CommitmentOpening {
    v: u64,
    r: Vec<Scalar>,
}
impl CommitmentOpening {
    pub fn new(v: u64, r: Vec<Scalar>) -> Self {
        Self { v, r }
    }
    pub fn r_len(&self) -> Result<usize, ProofError> {
        if self.r.is_empty() {
            Err(ProofError::InvalidLength("Extended blinding factors cannot be empty".to_string()))
        } else {
            Ok(self.r.len())
        }
    }
}

/// `RangeWitness` (/root/reference/src/range_witness.rs:15-41)
#[derive(Clone, Zeroize, ZeroizeOnDrop)]
pub struct RangeWitness {
    openings: Vec<CommitmentOpening>,
    #[zeroize(skip)]
    extension_degree: ExtensionDegree,
}
impl RangeWitness {
    pub fn init(openings: Vec<CommitmentOpening>) -> Result<Self, ProofError> {
        if openings.is_empty() {
            return Err(ProofError::InvalidLength("Vector openings_vec length cannot be 0".to_string()));
        }
        let first = openings[0].r_len()?;
        if openings.iter().any(|o| o.r.len() != first) {
            return Err(ProofError::InvalidLength("Extended blinding factors length must be consistent".to_string()));
        }
        let extension_degree = ExtensionDegree::try_from(first)?;
        Ok(Self { openings, extension_degree })
    }
}

/// The functions of `RangeProof<RistrettoPoint>` that the engine replaces.  Proofs themselves stay the reference's
/// `RistrettoRangeProof` (byte layout /root/reference/src/range_proof.rs:1120-1150).
pub struct RangeProof;

/// `verify_statements_and_generators_consistency` (/root/reference/src/range_proof.rs:610-709) against the device-resident
/// parameter set: G, H, bit length, extension degree of every statement and `d1.len()` of every proof; promises below 2^n.
/// Gi / Hi are a function of (bit length, aggregation capacity) alone (generators/bulletproof_gens.rs:83-112) and every statement's
/// prefix of them is what the device derived, so the element-wise scan of :686-705 reduces to the parameter comparison.
fn check_consistency(params: &RangeParameters<'_>, statements: &[RangeStatement<RistrettoPoint>], proofs: &[RistrettoRangeProof]) -> Result<(), ProofError> {
    let first = statements.first().ok_or(ProofError::InvalidArgument("Empty proof statements".to_string()))?;
    if proofs.is_empty() {
        return Err(ProofError::InvalidArgument("Empty proofs".to_string()));
    }
    if statements.len() != proofs.len() {
        return Err(ProofError::InvalidArgument("Range statements and proofs length mismatch".to_string()));
    }
    let reference = params.reference();
    for (i, (statement, proof)) in statements.iter().zip(proofs.iter()).enumerate() {
        let g = &statement.generators;
        if g.g_bases_compressed() != reference.g_bases_compressed() || (i > 0 && g.g_bases() != first.generators.g_bases()) {
            return Err(ProofError::InvalidArgument("Inconsistent G generator point in batch statement".to_string()));
        }
        if g.h_base_compressed() != reference.h_base_compressed() {
            return Err(ProofError::InvalidArgument("Inconsistent H generator point in batch statement".to_string()));
        }
        if g.bit_length() != reference.bit_length() {
            return Err(ProofError::InvalidArgument("Inconsistent bit length in batch statement".to_string()));
        }
        if g.extension_degree() != reference.extension_degree() || proof.extension_degree() != reference.extension_degree() {
            return Err(ProofError::InvalidArgument("Inconsistent extension degree".to_string()));
        }
        if g.max_aggregation_factor() > reference.max_aggregation_factor() {
            return Err(ProofError::InvalidArgument("Inconsistent Gi generator point vector in batch statement".to_string()));
        }
    }
    let bit_length = reference.bit_length();
    for statement in statements {
        for value in statement.minimum_value_promises.iter().flatten() {
            if bit_length < 64 && value >> bit_length > 0 {
                return Err(ProofError::InvalidLength("Minimum value promise exceeds bit vector capacity".to_string()));
            }
        }
    }
    Ok(())
}

/// flat host buffers of one `bpp_verify_args` (one reference call = one chunk)
struct Packed {
    chunk_offsets: [u64; 2],
    proof_bytes: Vec<u8>,
    proof_offsets: Vec<u64>,
    commitments: Vec<u8>,
    commit_offsets: Vec<u64>,
    min_values: Vec<u64>,
    min_present: Vec<u8>,
    seeds: Zeroizing<Vec<u8>>,
    seed_present: Vec<u8>,
}
impl Packed {
    fn new(statements: &[RangeStatement<RistrettoPoint>], proofs: &[RistrettoRangeProof]) -> Self {
        let n = statements.len();
        let mut p = Packed {
            chunk_offsets: [0, n as u64],
            proof_bytes: Vec::new(),
            proof_offsets: vec![0],
            commitments: Vec::new(),
            commit_offsets: vec![0],
            min_values: Vec::new(),
            min_present: Vec::new(),
            seeds: Zeroizing::new(Vec::with_capacity(32 * n)),
            seed_present: Vec::with_capacity(n),
        };
        for (statement, proof) in statements.iter().zip(proofs.iter()) {
            p.proof_bytes.extend_from_slice(&proof.to_bytes());
            p.proof_offsets.push(p.proof_bytes.len() as u64);
            for c in &statement.commitments_compressed {
                p.commitments.extend_from_slice(c.as_bytes());
            }
            p.commit_offsets.push((p.commitments.len() / 32) as u64);
            for m in &statement.minimum_value_promises {
                p.min_values.push(m.unwrap_or(0));
                p.min_present.push(m.is_some() as u8);
            }
            match &statement.seed_nonce {
                Some(s) => {
                    p.seeds.extend_from_slice(s.as_bytes());
                    p.seed_present.push(1);
                },
                None => {
                    p.seeds.extend_from_slice(&[0u8; 32]);
                    p.seed_present.push(0);
                },
            }
        }
        p
    }
    fn args(&self, transcripts: *mut u8, action: VerifyAction) -> sys::bpp_verify_args {
        sys::bpp_verify_args {
            n_proofs: self.seed_present.len(),
            n_chunks: 1,
            chunk_offsets: self.chunk_offsets.as_ptr(),
            proof_bytes: self.proof_bytes.as_ptr(),
            proof_offsets: self.proof_offsets.as_ptr(),
            commitments32: self.commitments.as_ptr(),
            commit_offsets: self.commit_offsets.as_ptr(),
            min_values: self.min_values.as_ptr(),
            min_present: self.min_present.as_ptr(),
            seed_nonces32: self.seeds.as_ptr(),
            seed_present: self.seed_present.as_ptr(),
            transcripts,
            action: match action {
                VerifyAction::RecoverOnly => sys::BPP_RECOVER_ONLY,
                VerifyAction::RecoverAndVerify => sys::BPP_RECOVER_AND_VERIFY,
                VerifyAction::VerifyOnly => sys::BPP_VERIFY_ONLY,
            },
        }
    }
}

fn masks_from(ext: ExtensionDegree, n: usize, masks: &[u8], present: &[u8]) -> Result<Vec<Option<ExtendedMask>>, ProofError> {
    let e = ext as usize;
    let mut out = Vec::with_capacity(n);
    for i in 0..n {
        if present[i] != 0 {
            let mut blindings = Vec::with_capacity(e);
            for k in 0..e {
                let b: [u8; 32] = masks[32 * (i * e + k)..32 * (i * e + k + 1)].try_into().unwrap();
                blindings.push(Option::<Scalar>::from(Scalar::from_canonical_bytes(b)).ok_or(ProofError::InvalidArgument("mask".to_string()))?);
            }
            out.push(Some(ExtendedMask::assign(ext, blindings)?));
        } else {
            out.push(None);
        }
    }
    Ok(out)
}

impl RangeProof {
    /// `RangeProof::prove_with_rng` (/root/reference/src/range_proof.rs:232-608).  `rng` is drawn 32 bytes per `TranscriptRng`
    /// rebuild (`TranscriptRngBuilder::finalize`, transcripts.rs:185-194): `log2(n * m) + 3` draws, in the reference's order, so a
    /// seeded RNG yields the reference's bytes.  (On an early error return the reference has drawn fewer bytes; nothing else
    /// about the RNG's use differs.)
    pub fn prove_with_rng<R: CryptoRngCore>(
        transcript: &mut Transcript,
        params: &RangeParameters<'_>,
        statement: &RangeStatement<RistrettoPoint>,
        witness: &RangeWitness,
        rng: &mut R,
    ) -> Result<RistrettoRangeProof, ProofError> {
        let m = statement.commitments.len();
        let ext = params.extension_degree() as usize;
        // :248-260
        if witness.openings.len() != m {
            return Err(ProofError::InvalidLength("Witness openings and statement commitments do not match!".to_string()));
        }
        if witness.extension_degree != params.extension_degree() {
            return Err(ProofError::InvalidLength("Witness and statement extension degrees do not match!".to_string()));
        }
        let full_length = params.bit_length().checked_mul(m).ok_or(ProofError::SizeOverflow)?;
        let rounds = full_length.ilog2() as usize;
        let n_draws = rounds + 3;
        let mut rng_bytes = Zeroizing::new(vec![0u8; 32 * n_draws]);
        for chunk in rng_bytes.chunks_exact_mut(32) {
            rng.fill_bytes(chunk);
        }
        let commitments: Vec<u8> = statement.commitments_compressed.iter().flat_map(|c| c.as_bytes().iter().copied()).collect();
        let values: Zeroizing<Vec<u64>> = Zeroizing::new(witness.openings.iter().map(|o| o.v).collect());
        let blindings: Zeroizing<Vec<u8>> = Zeroizing::new(witness.openings.iter().flat_map(|o| o.r.iter().flat_map(|s| s.to_bytes())).collect());
        let min_values: Vec<u64> = statement.minimum_value_promises.iter().map(|v| v.unwrap_or(0)).collect();
        let min_present: Vec<u8> = statement.minimum_value_promises.iter().map(|v| v.is_some() as u8).collect();
        let seed: Zeroizing<[u8; 32]> = Zeroizing::new(statement.seed_nonce.map(|s| s.to_bytes()).unwrap_or([0u8; 32]));
        let seed_present = [statement.seed_nonce.is_some() as u8];
        let args = sys::bpp_prove_args {
            n_proofs: 1,
            aggregation: m as i32,
            commitments32: commitments.as_ptr(),
            values: values.as_ptr(),
            blindings32: blindings.as_ptr(),
            min_values: min_values.as_ptr(),
            min_present: min_present.as_ptr(),
            seed_nonces32: seed.as_ptr(),
            seed_present: seed_present.as_ptr(),
            transcripts: transcript.state.as_mut_ptr(),
            rng_bytes: rng_bytes.as_ptr(),
            rng_stride: 32 * n_draws,
        };
        let size = unsafe { sys::bpp_proof_size(ext as i32, rounds as i32) };
        let mut out = vec![0u8; size];
        let mut status = [0i32; 1];
        let rc = unsafe { sys::bpp_prove_batch(params.gens, &args, out.as_mut_ptr(), size, status.as_mut_ptr()) };
        if rc != sys::BPP_OK {
            return Err(status_to_error(rc, &params.engine.last_error()));
        }
        if status[0] != sys::BPP_OK {
            return Err(status_to_error(status[0], "prove_with_rng"));
        }
        RistrettoRangeProof::from_bytes(&out)
    }

    fn check_arguments<T>(transcripts: &[T], statements: &[RangeStatement<RistrettoPoint>], proofs: &[RistrettoRangeProof]) -> Result<usize, ProofError> {
        // /root/reference/src/range_proof.rs:719-734, then the first chunk of at most 256 (:739-751)
        if statements.is_empty() || proofs.is_empty() || transcripts.is_empty() {
            return Err(ProofError::InvalidArgument("Range statements or proofs length empty".to_string()));
        }
        if statements.len() != proofs.len() {
            return Err(ProofError::InvalidArgument("Range statements and proofs length mismatch".to_string()));
        }
        if transcripts.len() != statements.len() {
            return Err(ProofError::InvalidArgument("Range statements and transcripts length mismatch".to_string()));
        }
        Ok(statements.len().min(MAX_RANGE_PROOF_BATCH_SIZE))
    }

    /// `RangeProof::verify_batch` (/root/reference/src/range_proof.rs:712-1065) with the whole of `verify` on the device, loop 1
    /// included (the transcripts travel as their 203-byte state and come back advanced).
    pub fn verify_batch(
        transcripts: &mut [Transcript],
        params: &RangeParameters<'_>,
        statements: &[RangeStatement<RistrettoPoint>],
        proofs: &[RistrettoRangeProof],
        action: VerifyAction,
    ) -> Result<Vec<Option<ExtendedMask>>, ProofError> {
        let n = Self::check_arguments(transcripts, statements, proofs)?;
        let (statements, proofs) = (&statements[..n], &proofs[..n]);
        check_consistency(params, statements, proofs)?;
        let packed = Packed::new(statements, proofs);
        let mut states: Vec<u8> = transcripts[..n].iter().flat_map(|t| t.state.iter().copied()).collect();
        let args = packed.args(states.as_mut_ptr(), action);
        let ext = params.extension_degree();
        let (mut status, mut masks, mut present) = ([0i32; 1], Zeroizing::new(vec![0u8; 32 * n * ext as usize]), vec![0u8; n]);
        let rc = unsafe { sys::bpp_verify_chunks(params.gens, &args, status.as_mut_ptr(), masks.as_mut_ptr(), present.as_mut_ptr()) };
        if rc != sys::BPP_OK {
            return Err(status_to_error(rc, &params.engine.last_error()));
        }
        for (t, s) in transcripts[..n].iter_mut().zip(states.chunks_exact(sys::BPP_TRANSCRIPT_BYTES)) {
            t.state.copy_from_slice(s);
        }
        if status[0] != sys::BPP_OK {
            return Err(status_to_error(status[0], "Range proof batch not valid"));
        }
        masks_from(ext, n, &masks, &present)
    }

    /// The same over stock `merlin::Transcript`: loop 1 (/root/reference/src/range_proof.rs:816-850) and the weight draws (:853,
    /// :894) run here with merlin itself, everything from :856 on runs on the device (`bpp_verify_chunks_ch`).
    pub fn verify_batch_merlin(
        transcripts: &mut [merlin::Transcript],
        params: &RangeParameters<'_>,
        statements: &[RangeStatement<RistrettoPoint>],
        proofs: &[RistrettoRangeProof],
        action: VerifyAction,
    ) -> Result<Vec<Option<ExtendedMask>>, ProofError> {
        let n = Self::check_arguments(transcripts, statements, proofs)?;
        let (statements, proofs) = (&statements[..n], &proofs[..n]);
        check_consistency(params, statements, proofs)?;
        let packed = Packed::new(statements, proofs);
        let reference = params.reference();
        let h = reference.h_base_compressed();
        // loop 1: all proofs first; an identity point or a zero challenge ends the call here (VerificationFailed)
        let (mut challenges, mut offsets, mut wbytes) = (Vec::<u8>::new(), vec![0u64], Vec::with_capacity(n));
        for (i, (transcript, statement)) in transcripts[..n].iter_mut().zip(statements.iter()).enumerate() {
            let lo = packed.proof_offsets[i] as usize;
            let hi = packed.proof_offsets[i + 1] as usize;
            let view = merlin_host::ProofView::new(&packed.proof_bytes[lo..hi]);
            let (ch, wb) = merlin_host::replay_one(
                transcript,
                &h,
                reference.g_bases_compressed(),
                reference.bit_length(),
                &statement.commitments_compressed,
                &statement.minimum_value_promises,
                &view,
            )?;
            for c in &ch {
                challenges.extend_from_slice(c.as_bytes());
            }
            offsets.push((challenges.len() / 32) as u64);
            wbytes.push(wb);
        }
        let weights: Vec<u8> = merlin_host::batch_weights(&wbytes).iter().flat_map(|w| w.to_bytes()).collect();
        let ch = sys::bpp_verify_challenges { challenges32: challenges.as_ptr(), challenge_offsets: offsets.as_ptr(), weights32: weights.as_ptr() };
        let args = packed.args(ptr::null_mut(), action);
        let ext = params.extension_degree();
        let (mut status, mut masks, mut present) = ([0i32; 1], Zeroizing::new(vec![0u8; 32 * n * ext as usize]), vec![0u8; n]);
        let rc = unsafe { sys::bpp_verify_chunks_ch(params.gens, &args, &ch, status.as_mut_ptr(), masks.as_mut_ptr(), present.as_mut_ptr()) };
        if rc != sys::BPP_OK {
            return Err(status_to_error(rc, &params.engine.last_error()));
        }
        if status[0] != sys::BPP_OK {
            return Err(status_to_error(status[0], "Range proof batch not valid"));
        }
        masks_from(ext, n, &masks, &present)
    }
}
