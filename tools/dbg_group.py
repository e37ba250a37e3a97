import sys; sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np, reverie_b200 as rb, bench
from reverie_b200 import circuits as C
ops, wit, wz, wc, _ = bench.make_workload("sha256")
seeds = bench.default_seeds()
circ = rb.Circuit(ops, wc)
g = rb.Group.local(circ, [0, 1], n_sessions=1, slots=1)
print("created", flush=True)
import hashlib
for i in range(3):
    p = g.prove(wit, (), seeds)
    print(i, hashlib.sha256(p.data).hexdigest() == bench.golden_digest("sha256"), flush=True)
