"""ncu target: steps of ONE multi-proof session of the bench's timed configuration (default: SHA-256, 8 proofs side by side, all
32 packed instances), launched eagerly so that every kernel is its own launch.  10 launches per step: values, key_setup, mask_gen,
linear (k_mask_vm), items, chunk_cv, rep_hash, challenge, extract (+ xfinish when linked).
    ncu --set full --clock-control none --import-source on --launch-skip 20 --launch-count 9 -o gpurun_out/r2_full_sha_b8 python tools/ncu_session.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import reverie_b200 as rb  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "sha256"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 8
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
ops, wit, wz, wc, _ = bench.make_workload(name)
seeds = bench.default_seeds()
circ = rb.Circuit(ops, wc, prove_only=True)
s = rb.Session(circ, 0, 32, n_proofs=P)
for b in range(P):
    s.upload(wit, wz, seeds, slot=b)
s.timing(True)  # keeps the phases eager: one launch per kernel
for _ in range(steps):
    s.prove()
    s.sync()
print("launches per step:", s.launch_count // steps)
