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
    # two proofs in flight as one rv_batch: commit graph -> one NCCL group launch for both all-gathers -> open graph -> one reduce
    name, ops, gwit, zwit, wc = cases[0]
    circ = rb.Circuit(ops, wc)
    first, count = sharding.shard_of(rank, world)
    sess = [rb.Session(circ, first, count) for _ in range(2)]
    seeds2 = bytes(reversed(seeds))
    batch = rb.Batch(sess)
    lead = torch.cuda.ExternalStream(batch.stream)
    for rnd in range(3):
        sess[0].upload(gwit, zwit, seeds)
        sess[1].upload(gwit, zwit, seeds2)
        batch.commit()
        with torch.cuda.stream(lead):
            sharding.all_gather_hashes_batched([torch.as_tensor(s.all_hashes_device(), device="cuda") for s in sess],
                                               [torch.as_tensor(s.hashes_device(), device="cuda") for s in sess])
        batch.open()
        proofs = sharding.reduce_proofs(sess)
        if rank == 0:
            assert proofs[0] == orc.prove(ops, gwit, zwit, wc, seeds)[1] and proofs[1] == orc.prove(ops, gwit, zwit, wc, seeds2)[1], rnd
    if rank == 0:
        print(f"mgpu ok: batch of 2 on {world} GPUs", flush=True)
    del batch
    # device-side exchange (rv_session_peer_link, CUDA IPC between the ranks' processes): no NCCL call on the data path;
    # multi-proof sessions driven as one graph launch per rank and step
    if 32 // world >= 4:
        sops, n_wires, _ = C.sha256_compress_circuit(None)
        swc = (0, n_wires)
        wits = [C.sha256_witness(C.sha256_pad_single_block(m)) for m in (b"abc", b"", b"slot two")]
        sds = [seeds, seeds2, bytes(seeds[1:] + seeds[:1])]
        scirc = rb.Circuit(sops, swc)
        ls = [rb.Session(scirc, first, count, n_proofs=3) for _ in range(2)]
        sharding.link_sessions(ls)
        lb = rb.Batch(ls)
        for rnd in range(4):
            for i in range(2):
                for b in range(3):
                    k = (b + i + rnd) % 3
                    ls[i].upload(wits[k], (), sds[k], slot=b)
            lb.prove()
            for i in range(2):
                for b in range(3):
                    k = (b + i + rnd) % 3
                    _, got = ls[i].fetch(b)
                    if rank == 0:
                        assert got == orc.prove(sops, wits[k], [], swc, sds[k])[1], (rnd, i, b)
        del lb
    for name, ops, gwit, zwit, wc in cases:
        circ = rb.Circuit(ops, wc)
        s1 = rb.Session(circ, first, count)
        sharding.link_sessions([s1])
        for rnd in range(3):
            proof = sharding.prove_linked(circ, gwit, zwit, seeds, session=s1)
            if rank == 0:
                assert proof == orc.prove(ops, gwit, zwit, wc, seeds)[1], (name, rnd)
    # the same through the one-handle API (rv_group_create_rank + handles over the host channel): waves beyond the capacity
    if 32 // world >= 4:
        g = rb.Group.rank(scirc, rank, world, n_sessions=2, slots=2)
        g.link_distributed()
        order = [0, 1, 2, 1, 0, 2, 2]
        for rnd in range(3):
            got = g.prove_batch([wits[k] for k in order], None, [sds[k] for k in order])
            if rank == 0:
                assert [p.serialize() for p in got] == [orc.prove(sops, wits[k], [], swc, sds[k])[1] for k in order], rnd
            else:
                assert got == [None] * len(order)
        del g
    if rank == 0:
        print(f"mgpu ok: linked sessions (device-side exchange) on {world} GPUs", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
