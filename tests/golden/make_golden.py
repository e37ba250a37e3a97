"""Generates tests/golden/proofs.json: SHA-256 digests (and sizes / comm) of oracle proofs for fixed (circuit, witness, seeds).

The reference holds no golden vectors for this path and cannot be built here, so these vectors are produced by the C
oracle (oracle/c) and cross-checked against the independent Python restatement (oracle/reverie_oracle.py) before being
written.  They pin the oracle against regressions and give the GPU tests a fixture that does not depend on re-running
the oracle.  Run from the repo root:  python tests/golden/make_golden.py"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np  # noqa: E402

import orc  # noqa: E402
import reverie_oracle as R  # noqa: E402
from reverie_b200 import circuits as C  # noqa: E402


def cases():
    """name -> (ops, witness, wire_counts); everything deterministic."""
    out = {}
    for n in (0, 1, 7, 8, 9, 64, 1000):
        ops, wc = C.flat_mul_circuit(n)
        out[f"flat_mul_{n}"] = (ops, np.array([1, 1], dtype=np.uint8), wc)
    from tests.test_gpu_parity import _random_circuit

    for seed in (0, 1, 2):
        rng = np.random.default_rng(1000 + seed)
        out[f"random_{seed}"] = _random_circuit(rng, int(rng.integers(1, 40)), int(rng.integers(1, 2000)))
    out["sha256_abc"] = C.sha256_abc_case()
    out["aes128_fips197"] = C.aes128_fips197_case()
    out["empty"] = (np.zeros(0, dtype=C.OP_DTYPE), np.zeros(0, dtype=np.uint8), (0, 0))
    return out


def zcases():
    """Z64 / mixed-domain cases: name -> (ops, wit_gf2, wit_z64, wire_counts)."""
    from tests._zgen import Z64_WITNESS, random_z_circuit

    out = {}
    none = np.zeros(0, dtype=np.uint8)
    for n in (0, 1, 16, 129):
        ops, wc = C.flat_mul_circuit(n, domain=C.Z64)
        out[f"z64_flat_{n}"] = (ops, none, Z64_WITNESS, wc)
    ops, nw = C.z64_mul_circuit(500)  # SURVEY.md 8(d) config 3 at test size
    out["z64_config3_500"] = (ops, none, Z64_WITNESS, (nw, 0))
    for seed in (0, 1):
        out[f"z64_random_{seed}"] = random_z_circuit(np.random.default_rng(3000 + seed), 4, 300, with_gf2=seed == 1)
    return out


def main():
    seeds = b"".join(R.default_seeds())
    seed_list = R.default_seeds()
    golden = {"seed_rule": 'seed[r] = BLAKE3("reverie-b200 seed" || LE32(r))[..16]', "cases": {}}
    for name, (ops, wit, wc) in cases().items():
        rc, pb, hashes = orc.prove(ops, wit, [], wc, seeds, want_hashes=True)
        assert rc == 0, name
        if len(ops) <= 3000:  # the literal Python restatement must produce the same bytes
            p = R.prove(orc.ops_to_tuples(ops), [int(x) for x in wit], [], wc, seed_list)
            assert R.serialize(p) == pb, name
        assert orc.verify(ops, wc, pb)[0] == 1, name
        golden["cases"][name] = {"n_ops": int(len(ops)), "proof_len": len(pb), "proof_sha256": hashlib.sha256(pb).hexdigest(),
                                 "comm": pb[:32].hex(), "rep_hashes_sha256": hashlib.sha256(hashes).hexdigest()}
        print(name, golden["cases"][name])
    golden["zcases"] = {}
    for name, (ops, gwit, zwit, wc) in zcases().items():
        rc, pb, hashes = orc.prove(ops, gwit, zwit, wc, seeds, want_hashes=True)
        assert rc == 0, name
        if len(ops) <= 400:
            p = R.prove(orc.ops_to_tuples(ops), [int(x) for x in gwit], [int(x) for x in zwit], wc, seed_list)
            assert R.serialize(p) == pb, name
        assert orc.verify(ops, wc, pb)[0] == 1, name
        golden["zcases"][name] = {"n_ops": int(len(ops)), "proof_len": len(pb), "proof_sha256": hashlib.sha256(pb).hexdigest(),
                                  "comm": pb[:32].hex(), "rep_hashes_sha256": hashlib.sha256(hashes).hexdigest()}
        print(name, golden["zcases"][name])
    with open(os.path.join(ROOT, "tests", "golden", "proofs.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
