"""Development aid: run one circuit family on the GPU and compare with the oracle (bounded; run under `timeout`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import reverie_b200 as rb
from reverie_b200 import circuits as C
import orc, reverie_oracle as R
from tests.test_gpu_parity import _random_circuit

seeds = b"".join(R.default_seeds())
which = sys.argv[1]
if which == "tiny":
    ops, wc = C.flat_mul_circuit(1); wit = [1, 1]
elif which.startswith("flat"):
    ops, wc = C.flat_mul_circuit(int(which[4:])); wit = [1, 0]
elif which.startswith("rand"):
    rng = np.random.default_rng(int(which[4:])); ops, wit, wc = _random_circuit(rng, int(rng.integers(1, 40)), int(rng.integers(1, 3000)))
elif which == "sha":
    ops, wit, wc = C.sha256_abc_case()
print(which, "ops", len(ops), flush=True)
c = rb.Circuit(ops, wc); print(c.stats(), flush=True)
got = rb.Proof.new(c, wit, (), seeds=seeds).serialize(); print("proved", len(got), flush=True)
rc, want = orc.prove(ops, wit, [], wc, seeds)
print("MATCH" if got == want else "MISMATCH", flush=True)
