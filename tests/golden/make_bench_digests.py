"""Generates tests/golden/bench_digests.json: SHA-256 of the CPU oracle's proof bytes for bench.py's workloads with bench.py's fixed
seeds.  bench.py compares the GPU proofs of its timed steps with these digests at every N (`parity_checked`), so the driver-run
line carries parity for BASELINE.json configs 2, 3 and 5 without importing the oracle into the GPU arm.

    python tests/golden/make_bench_digests.py [workload ...]     (default: every workload bench.py's default run touches)

The 10^8-gate circuits use the oracle's two-pass low-memory mode (orc_prove_lowmem: same functions, same bytes -- pinned against
orc_prove by tests/test_oracle_protocol.py); ~3 minutes each on 8 cores."""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import bench  # noqa: E402
import orc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "bench_digests.json")
DEFAULT = ("sha256", "flat1000000", "layered1000000", "z64mul100000", "z64mul1000000", "flat10000000", "flat100000000", "layered100000000")


def main():
    names = sys.argv[1:] or DEFAULT
    try:
        with open(OUT) as f:
            doc = json.load(f)
    except Exception:
        doc = {"what": "sha256 of the oracle's bincode proof bytes; seeds = numpy default_rng(20261017) bytes (bench.default_seeds)", "digests": {}, "proof_bytes": {}}
    seeds = bench.default_seeds()
    for name in names:
        ops, wit, wz, wc, _ = bench.make_workload(name)
        n_mul = int((ops["opcode"] == 6).sum())
        t0 = time.perf_counter()
        if n_mul > 20_000_000 or (name.startswith("z64") and n_mul > 500_000):
            # at most `n_threads` instances' transcripts (24 bytes per gate each) are alive at a time: bound them by the machine's memory
            rc, dg, n = orc.prove_digest_lowmem(ops, wit, wz, wc, seeds, n_threads=min(os.cpu_count() or 1, 8 if n_mul <= 150_000_000 else 4))
        else:
            rc, proof = orc.prove(ops, wit, wz, wc, seeds)
            dg, n = (hashlib.sha256(proof).hexdigest(), len(proof)) if rc == 0 else (None, 0)
        assert rc == 0, (name, rc)
        doc["digests"][name] = dg
        doc["proof_bytes"][name] = n
        print(f"{name}: {dg} ({n} bytes, {time.perf_counter() - t0:.1f} s)", flush=True)
        with open(OUT, "w") as f:
            json.dump(doc, f, indent=1, sort_keys=True)
            f.write("\n")


if __name__ == "__main__":
    main()
