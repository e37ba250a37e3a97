"""Development aid: rv_proof_new (op list with every call) on big circuits it has not seen -- streaming at first sight vs compiling for
residency -- after a warm-up call of the same size (a fresh VM's first touch of host memory is several times slower than a later one)."""
import hashlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, reverie_b200 as rb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000000
seeds = bench.default_seeds()
small = bench.make_workload("flat1000000")
rb.Proof.new_streaming(small[0], small[1], small[3], seeds=seeds, window_ops=1 << 18)  # context, kernels

def run(name, what):
    ops, wit, wz, wc, _ = bench.make_workload(name)
    t0 = time.perf_counter()
    p = rb.Proof.new(ops, wit, (), wc, seeds=seeds)
    dt = time.perf_counter() - t0
    want = bench.golden_digest(name)
    ok = (hashlib.sha256(memoryview(p._buf)).hexdigest() == want) if want else None
    print(f"{name} {what}: {dt:.2f} s = {n / dt:.3e} AND/s; oracle digest match: {ok}", flush=True)

run(f"flat{n - 1000}", "warm-up (first sight: streaming; first touch of the host memory)")
run(f"flat{n}", "first sight: streaming")
run(f"flat{n}", "second call: compile + cache + resident proof")
run(f"flat{n}", "third call: cache hit")
