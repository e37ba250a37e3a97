"""torchrun worker for the multi-GPU parity test: one rank per GPU, NCCL; the sharded proof must equal the oracle's bytes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    import orc
    import reverie_oracle as R
    import reverie_b200 as rb
    from reverie_b200 import _native, circuits as C, sharding
    from tests._zgen import random_z_circuit

    torch.cuda.set_device(local)
    _native.check(_native.lib().rv_set_device(local))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    seeds = b"".join(R.default_seeds())
    cases = []
    ops, wit, wc = C.sha256_abc_case()
    cases.append(("sha256", ops, wit, np.zeros(0, dtype=np.uint64), wc))
    ops, gwit, zwit, wc = random_z_circuit(np.random.default_rng(7), 4, 500, with_gf2=True)
    cases.append(("mixed", ops, gwit, zwit, wc))
    for name, ops, gwit, zwit, wc in cases:
        circ = rb.Circuit(ops, wc)
        proof = sharding.prove_sharded(circ, gwit, zwit, seeds)
        if rank == 0:
            rc, want = orc.prove(ops, gwit, zwit, wc, seeds)
            assert rc == 0 and proof == want, f"{name}: sharded proof over {world} GPUs differs from the oracle"
            assert rb.Proof(proof).verify(circ)
            print(f"mgpu ok: {name} on {world} GPUs, {len(proof)} bytes", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
