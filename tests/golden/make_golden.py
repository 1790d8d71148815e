#!/usr/bin/env python
"""Generates tests/golden/bpp_golden.json.

The reference (tari_bulletproofs_plus 0.4.1, Rust) cannot run in this image and holds no golden vectors of its own
(SURVEY.md §0.4, §8c), so these vectors are produced by the CPU oracle (oracle/, pinned against RFC 9496, libsodium, hashlib
and the Merlin KAT by tests/test_oracle_primitives.py) and frozen here: they pin the ORACLE against regressions and give the
device path fixed inputs with fixed expected bytes.  They are NOT outputs of the Rust crate; DESIGN.md §2 says "parity unpinned"
for that reason.  The generator constants of SURVEY.md §8c (computed there with hashlib + libsodium, independently of this
repository) are included verbatim as an external pin.

  python tests/golden/make_golden.py          # rewrites bpp_golden.json
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orc  # noqa: E402

LABEL = b"BatchedRangeProofTest"          # /root/reference/benches/range_proof.rs:49
# (bit_length, aggregation sizes of the batch, max aggregation, extension degree, promise kind, seed nonces)
CASES = [
    (64, [1], 1, 1, "third", True),        # BASELINE.json configs[0]
    (64, [1, 1, 1], 1, 1, "third", True),
    (64, [2, 4], 4, 1, "third", False),    # aggregated
    (32, [1, 2], 2, 2, "none", True),      # extension degree 2, no promises
    (64, [1, 1], 1, 3, "third", True),     # BASELINE.json configs[3] shape
    (8, [1, 8], 8, 1, "zero", True),
]
SURVEY_8C = {   # SURVEY.md §8c "Derived constants (hashlib + libsodium ...)"
    "G": ["044fad914b346d1623f0a123c90bec712c6bac717f2acbc48e12db5f6dcaef79", "429127c12411a4580d2606d2437a410a52254198b614e2d5c52ab8bb06576d55",
          "3255d8182cba353e52515411dcf0e2c28e926c5ca32689e41b722233e2bcfe70", "34f782cde81b8f949c34384912c4f978e612bd4a0609537682440e39a1864d15",
          "12a1a8238bf87962e8b102935ea88e14b2eee8d3d1719c63c44727b855205307", "aef270d5ac749567f6ffd256a692e1d4fb718565929693ca0aab94295b937b5e"],
    "Gi_0_0": "fc3b25801422672a6a8d3adb5d8457d4301fe92324b4fc56ae934c8713ddfe2d", "Gi_0_1": "ae817fdef62f713dd169dc8a26406f68be0bd3cd53652614636b0801567c4264",
    "Hi_0_0": "ba698f6dd08c501e32b55d2ee7259f6019d629fa2ba4d7039c5de157cba4df73", "Hi_0_1": "acf2d2b95428fac99b12da3bab92edf8ea3788c2fd16769e586397eede7b5052",
    "H": "e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76",
}


def make_case(ci, n, sizes, M, ext, promise, seeded):
    params = orc.Params(n, M, ext)
    rng = orc.Rng("chacha", 8675309 + ci)                # benches/range_proof.rs:47 (+ case index)
    proofs = []
    for pi, m in enumerate(sizes):
        vals = [rng.next_u64() % (1 << (n - 1)) for _ in range(m)]
        blinds = [[rng.random_not_zero() for _ in range(ext)] for _ in range(m)]
        commits = [params.commit(v, b) for v, b in zip(vals, blinds)]
        mins = [{"third": v // 3, "none": None, "zero": 0}[promise] for v in vals]
        seed = rng.random_not_zero() if (seeded and m == 1) else None
        rounds = (n * m - 1).bit_length()
        stream = hashlib.shake_256(b"golden-%d-%d" % (ci, pi)).digest(32 * (rounds + 3))      # what the external CryptoRng delivers
        st = orc.St(params, commits, mins, seed)
        rc, pr, t_after = orc.prove(orc.transcript_new(LABEL), st, orc.Wit(vals, blinds), orc.Rng("buffer", data=stream))
        assert rc == 0
        proofs.append({"values": vals, "blindings": [[hex(x) for x in b] for b in blinds], "commitments": [c.hex() for c in commits],
                       "minimum_value_promises": mins, "seed_nonce": hex(seed) if seed is not None else None, "rng_bytes": stream.hex(),
                       "proof": orc.proof_to_bytes(pr).hex(), "prover_transcript_after": t_after.hex(), "_st": st, "_pr": pr})
    sts, prs = [p["_st"] for p in proofs], [p["_pr"] for p in proofs]
    rc, masks = orc.verify_batch([orc.transcript_new(LABEL)] * len(prs), sts, prs, orc.RECOVER_AND_VERIFY)
    assert rc == 0
    verify = {"status": rc, "masks": [None if mk is None else [hex(x) for x in mk] for mk in masks]}
    # systematic corruptions of the last proof: expected reference error code of the whole call
    corrupt = []
    last = prs[-1]
    for field, idx in (("a", 3), ("a1", 0), ("b", 31), ("r1", 1), ("s1", 2), ("d1", 0), ("li", 5), ("ri", 7)):
        bad = last.copy()
        arr = getattr(bad, field)
        tgt = arr[0] if field in ("d1", "li", "ri") else arr
        tgt[idx] ^= 0x04
        rc2, _ = orc.verify_batch([orc.transcript_new(LABEL)] * len(prs), sts, prs[:-1] + [bad], orc.VERIFY_ONLY)
        corrupt.append({"field": field, "byte": idx, "xor": 4, "proof": orc.proof_to_bytes(bad).hex(), "status": rc2})
    for p in proofs:
        del p["_st"], p["_pr"]
    return {"bit_length": n, "max_aggregation": M, "extension_degree": ext, "label": LABEL.decode(), "proofs": proofs, "verify": verify,
            "corruptions_of_last_proof": corrupt}


def main():
    p = orc.Params(64, 1, 6)
    gens = {"H": p.point(0).hex(), "G": [p.point(1, k).hex() for k in range(6)], "Gi_0_0": p.point(2, 0).hex(), "Gi_0_1": p.point(2, 1).hex(),
            "Hi_0_0": p.point(3, 0).hex(), "Hi_0_1": p.point(3, 1).hex()}
    assert gens == SURVEY_8C, "oracle generators differ from the independently derived constants of SURVEY.md §8c"
    out = {"_comment": "generated by tests/golden/make_golden.py from the CPU oracle; NOT outputs of the Rust crate (see that script's docstring)",
           "generators_survey_8c": SURVEY_8C, "cases": [make_case(ci, *c) for ci, c in enumerate(CASES)]}
    with open(os.path.join(HERE, "bpp_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.join(HERE, "bpp_golden.json"), os.path.getsize(os.path.join(HERE, "bpp_golden.json")), "bytes")


if __name__ == "__main__":
    main()
