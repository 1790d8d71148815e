#!/usr/bin/env python
"""Writes tests/golden/pyref_vectors.json: the document oracle/rust_vectors (the REAL crate) would print, produced instead by the
independent pure-Python restatement oracle/pyref.py.  Same flow as /root/reference/tests/ristretto.rs:152-373 (`prove_and_verify`)
and oracle/rust_vectors/src/main.rs: one ChaCha12Rng::seed_from_u64(8675309) per case, per proof in this drawing order
  value = next_u64() % 2^(bit_length-1); one random_not_zero repeated `ext` times as the blinding vector;
  seed_nonce = random_not_zero iff aggregation == 1; prove_with_rng(Transcript("BatchedRangeProofTest"), ..., rng)
then verify_batch(RecoverAndVerify) for the masks and the verdict of the batch with one bit of the last proof's r1 flipped.
tests/test_pyref_vectors.py demands that the C oracle reproduces every byte of the document.

  python tests/golden/make_pyref_vectors.py            # all six shapes of oracle/rust_vectors (about two minutes of pure Python)
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import pyref as R  # noqa: E402

LABEL = b"BatchedRangeProofTest"
SEED = 8675309
CASES = [  # (bit_length, batch, extension degree, promise) -- oracle/rust_vectors/src/main.rs `cases`
    (64, [1], 1, "third"),            # BASELINE.json configs[0]
    (64, [1, 1, 1, 1], 1, "none"),
    (8, [1, 2, 4], 2, "third"),
    (64, [1, 1], 3, "equal"),         # configs[3] shape
    (64, [32], 1, "third"),           # configs[2]
    (32, [4, 1, 2], 1, "third"),
]


def hexs(x):
    return R.sc_bytes(x).hex()


def run_case(bit_length, batch, ext, promise):
    rng = R.ChaCha12Rng.seed_from_u64(SEED)
    value_max = 1 << (bit_length - 1)
    max_agg = max(batch)
    params = R.Params(bit_length, max_agg, ext)
    statements, proofs, out = [], [], []
    for m in batch:
        values, blindings, commitments, mins = [], [], [], []
        for _ in range(m):
            v = rng.next_u64() % value_max
            mins.append(None if promise == "none" else v // 3 if promise == "third" else v)
            bl = [R.random_not_zero(rng)] * ext
            values.append(v)
            blindings.append(bl)
            commitments.append(params.commit(v, bl))
        seed_nonce = R.random_not_zero(rng) if m == 1 else None
        st = R.Statement(params, commitments, mins, seed_nonce)
        pr = R.prove_with_rng(R.Transcript(LABEL), st, values, blindings, rng)
        statements.append(st)
        proofs.append(pr)
        out.append({"aggregation": m, "values": values, "blindings": [[hexs(b) for b in bl] for bl in blindings],
                    "commitments": [c.hex() for c in st.commitments_c], "minimum_value_promises": mins,
                    "seed_nonce": None if seed_nonce is None else hexs(seed_nonce), "proof": pr.to_bytes().hex()})
    masks = R.verify_batch([R.Transcript(LABEL) for _ in batch], statements, proofs, R.RECOVER_AND_VERIFY)
    bad_bytes = bytearray(proofs[-1].to_bytes())
    bad_bytes[1 + 32 * ext + 96] ^= 1
    try:
        bad = R.Proof.from_bytes(bytes(bad_bytes))
        try:
            R.verify_batch([R.Transcript(LABEL) for _ in batch], statements, proofs[:-1] + [bad], R.VERIFY_ONLY)
            verdict = "accepted"
        except R.ProofError as e:
            verdict = e.variant
    except R.ProofError:
        verdict = "parse_error"
    return {"bit_length": bit_length, "max_aggregation": max_agg, "extension_degree": ext, "promise": promise, "label": LABEL.decode(),
            "rng_seed": SEED, "proofs": out, "recovered_masks": [None if mk is None else [hexs(b) for b in mk] for mk in masks],
            "verdict_flipped_r1": verdict}


def main():
    cases = []
    for c in CASES:
        t0 = time.time()
        cases.append(run_case(*c))
        print("case %r: %.1f s" % (c, time.time() - t0), file=sys.stderr)
    doc = {"crate": "pyref (independent pure-Python restatement, oracle/pyref.py); NOT the Rust crate", "cases": cases}
    with open(os.path.join(HERE, "pyref_vectors.json"), "w") as f:
        json.dump(doc, f, indent=0)
        f.write("\n")


if __name__ == "__main__":
    main()
