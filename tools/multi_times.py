"""Development aid: per-kernel device times of one multi-proof session (P proofs side by side) on the SHA-256 workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
import reverie_b200 as rb
from reverie_b200 import circuits as C

P = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ops, wit, wc = C.sha256_abc_case()
seeds = np.random.default_rng(1).integers(0, 256, size=256 * 16, dtype=np.uint8).tobytes()
circ = rb.Circuit(ops, wc)
s = rb.Session(circ, 0, 32, n_proofs=P)
for b in range(P):
    s.upload(wit, (), seeds, slot=b)
for _ in range(3):
    s.prove()
s.sync()
s.timing(True)
reps = 10
for _ in range(reps):
    s.prove()
kt = s.kernel_times()
tot = 0
for k in kt:
    print("%-12s %9.1f us per step  (%6.1f us per proof)" % (k["name"], k["ms"] * 1e3 / reps, k["ms"] * 1e3 / reps / P))
    if k["name"] != "values":
        tot += k["ms"] * 1e3 / reps
print("main-stream sum %.1f us = %.1f us per proof" % (tot, tot / P))
