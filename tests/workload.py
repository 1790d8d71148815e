"""Seeded synthetic workloads shaped like the reference's benches/tests
(/root/reference/benches/range_proof.rs:206-292, /root/reference/tests/ristretto.rs:152-373), generated with
the CPU oracle.  Used by tests/ and bench.py (oracle = checker / input generator, never the product path)."""
import orc

LABEL = b"BatchedRangeProofTest"  # benches/range_proof.rs:49
SEED = 8675309                    # benches/range_proof.rs:47


class Case:
    def __init__(self, params, statements, witnesses, proofs, transcripts):
        self.params, self.statements, self.witnesses = params, statements, witnesses
        self.proofs, self.transcripts = proofs, transcripts

    def proof_bytes(self):
        return [orc.proof_to_bytes(p) for p in self.proofs]


def make_case(bit_length, aggregation_sizes, ext, max_aggregation=None, promise="third", seed_nonce=True,
              rng_seed=SEED, same_blinding=False, params=None, rng=None, labels=None):
    """One proof per entry of aggregation_sizes, all under one RangeParameters (like the reference tests)."""
    rng = rng or orc.Rng("chacha", rng_seed)
    M = max_aggregation or max(aggregation_sizes)
    params = params or orc.Params(bit_length, M, ext)
    sts, wits, prs, trs = [], [], [], []
    vmax = 1 << (bit_length - 1)
    for m in aggregation_sizes:
        vals, bl, coms, mins = [], [], [], []
        for _ in range(m):
            v = rng.next_u64() % vmax
            if same_blinding:  # tests/ristretto.rs:190 repeats one scalar
                b = [rng.random_not_zero()] * ext
            else:
                b = [rng.random_not_zero() for _ in range(ext)]
            vals.append(v)
            bl.append(b)
            coms.append(params.commit(v, b))
            mins.append({"third": v // 3, "none": None, "equal": v, "zero": 0}[promise])
        sn = rng.random_not_zero() if (seed_nonce and m == 1) else None
        st = orc.St(params, coms, mins, sn)
        w = orc.Wit(vals, bl)
        t0 = orc.transcript_new(labels[len(prs)] if labels else LABEL)      # per-proof labels: transcripts in different states
        rc, pr, _ = orc.prove(t0, st, w, rng)
        assert rc == 0, rc
        sts.append(st)
        wits.append(w)
        prs.append(pr)
        trs.append(t0)
    return Case(params, sts, wits, prs, trs)
