"""Development aid (torchrun): where the end-to-end time of one big sharded proof goes on rank 0: upload / step / fetch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench, reverie_b200 as rb
from reverie_b200 import _native
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); _native.check(_native.lib().rv_set_device(local))
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
name = sys.argv[1] if len(sys.argv) > 1 else "flat100000000"
ops, wit, wz, wc, _ = bench.make_workload(name)
seeds = bench.default_seeds()
circ = rb.Circuit(ops, wc, prove_only=True)
g = rb.Group.rank(circ, rank, world); g.link_distributed()
x = g.sessions[0]
for it in range(6):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter(); x.upload(wit, wz, seeds)
    t1 = time.perf_counter(); x.prove()
    t2 = time.perf_counter(); x.sync()
    t3 = time.perf_counter(); c, p = x.fetch()
    t4 = time.perf_counter()
    print(f"rank {rank} it {it}: upload {1e3*(t1-t0):.1f} launch {1e3*(t2-t1):.1f} wait {1e3*(t3-t2):.1f} fetch {1e3*(t4-t3):.1f} ms ({len(p)} bytes)", flush=True)
    del p
dist.barrier(); dist.destroy_process_group()
