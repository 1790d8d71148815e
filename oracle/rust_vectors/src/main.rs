//! Prints golden vectors produced by the real `tari_bulletproofs_plus` crate as one JSON document on stdout.
//!
//! The flow is the reference's own `prove_and_verify` helper (tests/ristretto.rs:152-373): one
//! `ChaCha12Rng::seed_from_u64(8675309)` per case, and per proof of the case, in this order,
//!   value = next_u64() % 2^(bit_length-1)                         (ristretto.rs:181)
//!   one `Scalar::random_not_zero` repeated `extension_degree` times as the blinding vector  (:189)
//!   seed_nonce = Some(random_not_zero) iff aggregation_size == 1   (:209-213)
//!   RangeProof::prove_with_rng(transcript("BatchedRangeProofTest"), statement, witness, rng)   (:227-228)
//! which is exactly what tests/workload.py `make_case(..., same_blinding=True)` replays through the CPU oracle's ChaCha12,
//! so `tests/test_rust_vectors.py` can demand byte-identical commitments, proofs, verdicts and recovered masks.
use curve25519_dalek::scalar::Scalar;
use merlin::Transcript;
use rand_chacha::ChaCha12Rng;
use rand_core::{RngCore, SeedableRng};
use tari_bulletproofs_plus::{
    commitment_opening::CommitmentOpening,
    generators::pedersen_gens::ExtensionDegree,
    protocols::scalar_protocol::ScalarProtocol,
    range_parameters::RangeParameters,
    range_proof::{RangeProof, VerifyAction},
    range_statement::RangeStatement,
    range_witness::RangeWitness,
    ristretto,
};

fn hex(bytes: &[u8]) -> String {
    bytes.iter().map(|b| format!("{:02x}", b)).collect()
}

fn json_list(items: &[String]) -> String {
    format!("[{}]", items.join(", "))
}

fn quoted(s: &str) -> String {
    format!("\"{}\"", s)
}

struct Case {
    bit_length: usize,
    batch: Vec<usize>, // aggregation size of every proof of the batch
    ext: ExtensionDegree,
    promise: &'static str, // "none" | "third" | "equal"  (workload.py's names)
}

fn run(case: &Case) -> String {
    let mut rng = ChaCha12Rng::seed_from_u64(8675309);
    let label = "BatchedRangeProofTest";
    let value_max = (1u128 << (case.bit_length - 1)) as u64;
    let max_aggregation = *case.batch.iter().max().unwrap();
    let mut statements = vec![];
    let mut proofs = vec![];
    let mut transcripts = vec![];
    let mut proofs_json = vec![];
    for &m in &case.batch {
        let pc_gens = ristretto::create_pedersen_gens_with_extension_degree(case.ext);
        // one parameter set per batch, as workload.py does (the generator chains of a larger aggregation factor extend the smaller)
        let generators = RangeParameters::init(case.bit_length, max_aggregation, pc_gens).unwrap();
        let (mut openings, mut commitments, mut minimum_values) = (vec![], vec![], vec![]);
        let (mut values_json, mut blind_json, mut commit_json, mut min_json) = (vec![], vec![], vec![], vec![]);
        for _ in 0..m {
            let value = rng.next_u64() % value_max;
            let minimum_value = match case.promise {
                "none" => None,
                "third" => Some(value / 3),
                _ => Some(value),
            };
            let blindings = vec![Scalar::random_not_zero(&mut rng); case.ext as usize];
            let commitment = generators.pc_gens().commit(&Scalar::from(value), blindings.as_slice()).unwrap();
            values_json.push(value.to_string());
            blind_json.push(json_list(&blindings.iter().map(|b| quoted(&hex(b.as_bytes()))).collect::<Vec<_>>()));
            commit_json.push(quoted(&hex(commitment.compress().as_bytes())));
            min_json.push(minimum_value.map_or("null".to_string(), |v| v.to_string()));
            minimum_values.push(minimum_value);
            commitments.push(commitment);
            openings.push(CommitmentOpening::new(value, blindings));
        }
        let witness = RangeWitness::init(openings).unwrap();
        let seed_nonce = if m == 1 { Some(Scalar::random_not_zero(&mut rng)) } else { None };
        let statement = RangeStatement::init(generators.clone(), commitments, minimum_values, seed_nonce).unwrap();
        let transcript = Transcript::new(label.as_bytes());
        let proof = RangeProof::prove_with_rng(&mut transcript.clone(), &statement, &witness, &mut rng).unwrap();
        proofs_json.push(format!(
            "{{\"aggregation\": {}, \"values\": {}, \"blindings\": {}, \"commitments\": {}, \"minimum_value_promises\": {}, \"seed_nonce\": {}, \"proof\": {}}}",
            m,
            json_list(&values_json),
            json_list(&blind_json),
            json_list(&commit_json),
            json_list(&min_json),
            seed_nonce.map_or("null".to_string(), |s| quoted(&hex(s.as_bytes()))),
            quoted(&hex(&proof.to_bytes()))
        ));
        statements.push(statement);
        proofs.push(proof);
        transcripts.push(transcript);
    }
    // verdict + recovered masks of the whole batch, the corrupted-proof verdict (one bit of r1 of the last proof flipped)
    let masks = RangeProof::verify_batch(&mut transcripts.clone(), &statements, &proofs, VerifyAction::RecoverAndVerify).unwrap();
    let masks_json: Vec<String> = masks
        .iter()
        .map(|m| match m {
            None => "null".to_string(),
            Some(mask) => json_list(&mask.blindings().unwrap().iter().map(|b| quoted(&hex(b.as_bytes()))).collect::<Vec<_>>()),
        })
        .collect();
    let mut bad_bytes = proofs.last().unwrap().to_bytes();
    let r1_offset = 1 + 32 * (case.ext as usize) + 96; // [ext] d1[ext] a a1 b | r1 s1 (L R)*   (range_proof.rs:1120-1150)
    bad_bytes[r1_offset] ^= 1;
    let bad_verdict = match ristretto::RistrettoRangeProof::from_bytes(&bad_bytes) {
        Err(_) => "\"parse_error\"".to_string(),
        Ok(bad) => {
            let mut all = proofs.clone();
            *all.last_mut().unwrap() = bad;
            match RangeProof::verify_batch(&mut transcripts.clone(), &statements, &all, VerifyAction::VerifyOnly) {
                Ok(_) => "\"accepted\"".to_string(),
                Err(e) => quoted(&format!("{:?}", e).split('(').next().unwrap().to_string()),
            }
        },
    };
    format!(
        "{{\"bit_length\": {}, \"max_aggregation\": {}, \"extension_degree\": {}, \"promise\": \"{}\", \"label\": \"{}\", \"rng_seed\": 8675309, \"proofs\": {}, \"recovered_masks\": {}, \"verdict_flipped_r1\": {}}}",
        case.bit_length,
        max_aggregation,
        case.ext as usize,
        case.promise,
        label,
        json_list(&proofs_json),
        json_list(&masks_json),
        bad_verdict
    )
}

fn main() {
    let cases = vec![
        Case { bit_length: 64, batch: vec![1], ext: ExtensionDegree::DefaultPedersen, promise: "third" }, // BASELINE configs[0]
        Case { bit_length: 64, batch: vec![1, 1, 1, 1], ext: ExtensionDegree::DefaultPedersen, promise: "none" },
        Case { bit_length: 8, batch: vec![1, 2, 4], ext: ExtensionDegree::AddOneBasePoint, promise: "third" },
        Case { bit_length: 64, batch: vec![1, 1], ext: ExtensionDegree::AddTwoBasePoints, promise: "equal" }, // configs[3] shape
        Case { bit_length: 64, batch: vec![32], ext: ExtensionDegree::DefaultPedersen, promise: "third" }, // configs[2]
        Case { bit_length: 32, batch: vec![4, 1, 2], ext: ExtensionDegree::DefaultPedersen, promise: "third" },
    ];
    let body: Vec<String> = cases.iter().map(run).collect();
    println!("{{\"crate\": \"tari_bulletproofs_plus 0.4.1\", \"cases\": {}}}", json_list(&body));
}
