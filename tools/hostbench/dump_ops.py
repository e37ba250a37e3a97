"""bench.py workload name -> file of raw rv_op records (24 bytes each) for the host benches; prints the wire counts to pass on."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
ops, wit, wz, wc, desc = bench.make_workload(sys.argv[1])
ops.tofile(sys.argv[2])
print(f"{desc}\n{ops.size} ops -> {sys.argv[2]}; z64_cells {wc[0]} gf2_cells {wc[1]}")
