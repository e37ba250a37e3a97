"""compute-sanitizer --tool racecheck target: shared-memory hazards of the kernels touched in round 2 (two-column mask VM,
4- / 8-slice mask generator, value plane with the 4-step ring) on small inputs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import orc, reverie_oracle as R, reverie_b200 as rb
from tests.test_gpu_parity import _random_circuit

seeds = b"".join(R.default_seeds())
rng = np.random.default_rng(5)
ops, wit, wc = _random_circuit(rng, 16, 400, n_cells=30)
want = orc.prove(ops, wit, [], wc, seeds)[1]
circ = rb.Circuit(ops, wc)
s = rb.Session(circ, 0, 32, n_proofs=2)  # 64 columns: the VM runs two columns per CTA
for b in range(2):
    s.upload(wit, (), seeds, slot=b)
s.prove()
assert s.fetch(0)[1] == want and s.fetch(1)[1] == want
for per in (4, 2):  # 8- and 4-slice mask generator CTAs
    x = rb.Session(circ, 0, per)
    x.upload(wit, (), seeds); x.commit(); x.hashes()
assert rb.Proof(want).verify(circ)
print("racecheck_small ok")
