"""Development aid: per-kernel device times of one SHA-256 proof (not the bench)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import reverie_b200 as rb
from reverie_b200 import circuits as C
import reverie_oracle as R

seeds = b"".join(R.default_seeds())
which = sys.argv[1] if len(sys.argv) > 1 else "sha"
if which == "sha":
    ops, wit, wc = C.sha256_abc_case()
elif which.startswith("flat"):
    n = int(which[4:]); ops, wc = C.flat_mul_circuit(n); wit = [1, 1]
t = time.time(); circ = rb.Circuit(ops, wc); print("compile s", time.time() - t, circ.stats())
s = rb.Session(circ)
for it in range(3):
    t = time.time(); s.upload(wit, (), seeds); s.commit(); s.open(); comm, proof = s.fetch(); dt = time.time() - t
    print("e2e wall ms", dt * 1e3, len(proof))
s.timing(True)
N = 5
for it in range(N):
    s.upload(wit, (), seeds); s.commit(); s.open(); s.fetch()
for k in s.kernel_times():
    print("%-12s %8.3f us/launch-group  launches=%d" % (k["name"], k["ms"] * 1e3 / N, k["launches"] // N))
s.timing(False)
t = time.time()
for it in range(20):
    s.upload(wit, (), seeds); s.prove(); s.fetch()
dt = (time.time() - t) / 20
print("steady e2e ms/proof", dt * 1e3, "AND/s", circ.stats()["n_and"] / dt)
