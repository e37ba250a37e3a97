"""Where does the e2e time of a 32-proof step go (N = 1)?  raw C call vs the Python wrapper."""
import ctypes as C, time, sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
import bench, reverie_b200 as rb
from reverie_b200 import _native as N
from reverie_b200.proof import _batch_args
ops, wit, wz, wc, _ = bench.make_workload("sha256")
seeds = bench.default_seeds()
circ = rb.Circuit(ops, wc)
B = 32
def raw():
    keep, a_wg, n_g, a_wz, n_z, a_sd = _batch_args(B, [wit] * B, [wz] * B, [seeds] * B)
    outs, lens, sts = (C.c_void_p * B)(), (C.c_size_t * B)(), (C.c_int * B)()
    t0 = time.perf_counter()
    N.check(N.lib().rv_prove_batch(circ.handle, B, a_wg, n_g, a_wz, n_z, a_sd, outs, lens, sts))
    t1 = time.perf_counter()
    for i in range(B): N.lib().rv_free(C.c_void_p(outs[i]))
    return t1 - t0
for _ in range(5): raw()
ts = [raw() for _ in range(30)]
print("raw rv_prove_batch ms:", np.median(ts) * 1e3)
for _ in range(5): rb.Proof.new_batch(circ, [wit] * B, [wz] * B, seeds=[seeds] * B)
t0 = time.perf_counter()
for _ in range(30): p = rb.Proof.new_batch(circ, [wit] * B, [wz] * B, seeds=[seeds] * B)
print("Proof.new_batch ms:", (time.perf_counter() - t0) / 30 * 1e3)
g = rb.Group.local(circ, [0], n_sessions=4, slots=8)
for _ in range(5): g.prove_batch([wit] * B, [wz] * B, [seeds] * B)
t0 = time.perf_counter()
for _ in range(30): p = g.prove_batch([wit] * B, [wz] * B, [seeds] * B)
print("Group(1).prove_batch ms:", (time.perf_counter() - t0) / 30 * 1e3)
