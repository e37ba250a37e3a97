"""Development aid: streaming vs resident proving of one workload (time, digest vs the committed oracle digest)."""
import hashlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, reverie_b200 as rb
name = sys.argv[1] if len(sys.argv) > 1 else "flat100000000"
window = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ops, wit, wz, wc, _ = bench.make_workload(name)
seeds = bench.default_seeds()
n_and = int((ops["opcode"] == 6).sum())
for it in range(2):
    t0 = time.perf_counter()
    p = rb.Proof.new_streaming(ops, wit, wc, seeds=seeds, window_ops=window)
    dt = time.perf_counter() - t0
    dg = hashlib.sha256(memoryview(p._buf)).hexdigest()
    print(f"{name} window {window}: {dt:.2f} s = {n_and / dt:.3e} AND/s, {len(p)} bytes, sha256 {dg}, oracle digest match: {dg == bench.golden_digest(name)}", flush=True)
    del p
