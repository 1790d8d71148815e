//! The reference's `prove_and_verify` flow (/root/reference/tests/ristretto.rs:152-373) through the wrapper: device-made proofs must
//! be byte-identical to the reference's for the same `ChaCha12Rng` stream, and both verification entry points must agree with
//! `tari_bulletproofs_plus` on accept / reject and on the recovered masks.  Needs a B200 and `libbpp_b200.so` (BPP_B200_LIB_DIR).
use bpp_b200::*;
use curve25519_dalek::scalar::Scalar;
use rand_chacha::ChaCha12Rng;
use rand_core::{RngCore, SeedableRng};
use tari_bulletproofs_plus::{
    commitment_opening::CommitmentOpening as RefOpening,
    protocols::scalar_protocol::ScalarProtocol,
    range_witness::RangeWitness as RefWitness,
};

#[test]
fn device_proofs_equal_reference_proofs() {
    let engine = Engine::new(0).unwrap();
    for (bit_length, m, ext) in [(64usize, 1usize, ExtensionDegree::DefaultPedersen), (8, 4, ExtensionDegree::AddOneBasePoint)] {
        let params = RangeParameters::init(&engine, bit_length, m, create_pedersen_gens_with_extension_degree(ext)).unwrap();
        let mut rng = ChaCha12Rng::seed_from_u64(8675309);
        let (mut openings, mut ref_openings, mut commitments, mut mins) = (vec![], vec![], vec![], vec![]);
        for _ in 0..m {
            let value = rng.next_u64() % (1u64 << (bit_length - 1));
            let blindings = vec![Scalar::random_not_zero(&mut rng); ext as usize];
            commitments.push(params.reference().pc_gens().commit(&Scalar::from(value), &blindings).unwrap());
            mins.push(Some(value / 3));
            openings.push(CommitmentOpening::new(value, blindings.clone()));
            ref_openings.push(RefOpening::new(value, blindings));
        }
        let seed = if m == 1 { Some(Scalar::random_not_zero(&mut rng)) } else { None };
        let statement = RangeStatement::init(params.reference().clone(), commitments, mins, seed).unwrap();
        let mut rng_ref = rng.clone();
        let proof = RangeProof::prove_with_rng(&mut Transcript::new(b"test"), &params, &statement, &RangeWitness::init(openings).unwrap(), &mut rng).unwrap();
        let reference = RistrettoRangeProof::prove_with_rng(&mut merlin::Transcript::new(b"test"), &statement, &RefWitness::init(ref_openings).unwrap(), &mut rng_ref).unwrap();
        assert_eq!(proof.to_bytes(), reference.to_bytes());
        let want = RistrettoRangeProof::verify_batch(&mut [merlin::Transcript::new(b"test")], &[statement.clone()], &[reference], VerifyAction::RecoverAndVerify).unwrap();
        let got = RangeProof::verify_batch(&mut [Transcript::new(b"test")], &params, &[statement.clone()], &[proof.clone()], VerifyAction::RecoverAndVerify).unwrap();
        let got_m = RangeProof::verify_batch_merlin(&mut [merlin::Transcript::new(b"test")], &params, &[statement], &[proof], VerifyAction::RecoverAndVerify).unwrap();
        assert_eq!(want, got);
        assert_eq!(want, got_m);
    }
}
