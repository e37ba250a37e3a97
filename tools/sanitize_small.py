"""compute-sanitizer target: small instances of the paths added in round 2 (streaming segments, 1- and 2-instance shards, linked
shards on one device, a local group, the 8-byte extraction gather on short / ragged vectors).
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import orc, reverie_oracle as R, reverie_b200 as rb
from reverie_b200 import circuits as C
from tests.test_gpu_parity import _random_circuit

seeds = b"".join(R.default_seeds())
rng = np.random.default_rng(3)
ops, wit, wc = _random_circuit(rng, 16, 700, n_cells=40)
want = orc.prove(ops, wit, [], wc, seeds)[1]
assert rb.Proof.new(ops, wit, (), wc, seeds=seeds).serialize() == want
for w in (64, 200):
    assert rb.Proof.new_streaming(ops, wit, wc, seeds=seeds, window_ops=w).serialize() == want, w
circ = rb.Circuit(ops, wc)
for G in (8, 16, 32):
    per = 32 // G
    sess = [rb.Session(circ, g * per, per) for g in range(G)]
    for s in sess:
        s.upload(wit, (), seeds); s.commit()
    allh = b"".join(s.hashes() for s in sess)
    for s in sess:
        s.open(allh)
    parts = [s.fetch() for s in sess]
    assert rb.assemble(parts[0][0], [p for _, p in parts]) == want, G
    del sess
g = rb.Group.local(circ, [0, 0], n_sessions=1, slots=1)
assert g.prove(wit, (), seeds).serialize() == want
del g
for n in (1, 7, 9, 64, 1030):
    fops, fwc = C.flat_mul_circuit(n)
    fw = np.array([1, 1], dtype=np.uint8)
    fwant = orc.prove(fops, fw, [], fwc, seeds)[1]
    assert rb.Proof.new(fops, fw, (), fwc, seeds=seeds).serialize() == fwant, n
    assert rb.Proof.new_streaming(fops, fw, fwc, seeds=seeds, window_ops=64).serialize() == fwant, n
print("sanitize_small ok")
