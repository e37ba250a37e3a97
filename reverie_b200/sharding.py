"""Multi-GPU proving: the 32 packed instances of a proof (src/proof/mod.rs:127-157) sharded over the ranks of a
torch.distributed group, one process per GPU.  The only exchange on the data path is the all-gather of the 256 x 32-byte
repetition hashes that feed the Fiat-Shamir challenge (src/proof/mod.rs:160-171); every rank then derives the same
challenge and opens its own repetitions, and rank 0 assembles the shard blobs (src/proof/mod.rs:200-221).

Nothing here computes proof data on the host: commit / open run on the rank's GPU through the C ABI."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from . import _native as N


def shard_of(rank: int, world: int) -> Tuple[int, int]:
    """(first packed instance, count) owned by `rank`; `world` must divide 32."""
    if world < 1 or N.PACKED_REPS % world:
        raise ValueError("the number of ranks must divide the 32 packed instances")
    per = N.PACKED_REPS // world
    return rank * per, per


def all_gather_hashes_into(out, local, group=None):
    """Same, into a caller-provided tensor (e.g. the session's own receive buffer, Session.all_hashes_device())."""
    import torch.distributed as dist

    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


def all_gather_hashes_batched(outs, locals_, group=None):
    """The all-gathers of several proofs in flight as ONE NCCL group launch (outs[b] <- all ranks' locals_[b])."""
    import torch.distributed as dist

    pg = group or dist.group.WORLD
    if hasattr(pg, "allgather_into_tensor_coalesced"):
        work = pg.allgather_into_tensor_coalesced(list(outs), list(locals_))
        if work is not None:
            work.wait()  # stream-ordered for NCCL: makes the current stream wait, does not block the host
        return
    for o, l in zip(outs, locals_):
        dist.all_gather_into_tensor(o, l, group=group)


def all_gather_hashes(local, group=None):
    """local: uint8 tensor with this rank's n_instances * 256 bytes of repetition hashes (CUDA for NCCL, CPU for gloo).
    Returns the 256 x 32 bytes of all repetitions in repetition order (rank order = instance order)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    out = torch.empty(local.numel() * world, dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


def gather_parts(comm: bytes, part: bytes, group=None, dst: int = 0) -> Optional[bytes]:
    """Collect every rank's shard blob (equal lengths: full-length proofs, zero outside the shard's entries) on `dst` and
    assemble them into the bincode `Proof`.  Returns the proof bytes on `dst`, None elsewhere."""
    import torch
    import torch.distributed as dist

    from .proof import assemble

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda") if backend == "nccl" else torch.device("cpu")
    mine = torch.frombuffer(bytearray(part), dtype=torch.uint8).to(dev)
    if backend == "nccl":  # NCCL has no gather-to-one primitive in every torch version: all-gather is the same traffic at this size
        buf = torch.empty(mine.numel() * world, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(buf, mine, group=group)
        parts = [bytes(buf[i * mine.numel():(i + 1) * mine.numel()].cpu().numpy()) for i in range(world)] if rank == dst else None
    else:
        lst = [torch.empty_like(mine) for _ in range(world)] if rank == dst else None
        dist.gather(mine, lst, dst=dst, group=group)
        parts = [bytes(t.numpy()) for t in lst] if rank == dst else None
    return assemble(comm, parts) if rank == dst else None


def reduce_proofs(sessions, group=None, dst: int = 0):
    """Assemble the proofs of several sharded sessions on the device: the shard buffers of one proof are zero outside their
    own entries, so a sum-reduce of the bytes IS the concatenation of src/proof/mod.rs:200-221.  One NCCL reduce for all
    sessions; returns the list of proof bytes on `dst`, None elsewhere.  Raises WitnessError if any shard saw a failed assert."""
    import torch
    import torch.distributed as dist

    for s in sessions:
        s.status()
    parts = [torch.as_tensor(s.proof_device(b), device="cuda") for s in sessions for b in range(getattr(s, "n_proofs", 1))]
    buf = torch.cat(parts)
    dist.reduce(buf, dst=dst, op=dist.ReduceOp.SUM, group=group)
    if dist.get_rank(group) != dst:
        return None
    host = buf.cpu().numpy()
    out, pos = [], 0
    for p in parts:
        out.append(host[pos:pos + p.numel()].tobytes())
        pos += p.numel()
    return out


def link_sessions(sessions, group=None):
    """Link this rank's sessions (shard `rank` of the proofs they hold) to the corresponding sessions of the other ranks of
    `group`: afterwards Session.prove() / Batch.prove() run the whole sharded step -- commit, exchange of the repetition
    hashes over NVLink peer memory, challenge, extraction into rank 0's buffers -- as one CUDA graph launch per rank, with
    no collective call on the data path.  The handles travel once, through the group's host channel."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = [s.peer_handle() for s in sessions]
    everyone = [None] * world
    dist.all_gather_object(everyone, mine, group=group)
    for i, s in enumerate(sessions):
        s.peer_link(rank, world, [everyone[r][i] for r in range(world)])


def prove_linked(circuit, wit_gf2, wit_z64=(), seeds=None, group=None, session=None) -> Optional[bytes]:
    """Proof::new over all ranks of `group` with the device-side exchange (link_sessions).  Returns the proof on rank 0, None
    elsewhere (after checking this rank's status)."""
    import torch.distributed as dist

    from .proof import Session

    if seeds is None:
        raise ValueError("sharded proving needs the same seeds on every rank: draw them on rank 0 and broadcast")
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    first, count = shard_of(rank, world)
    s = session
    if s is None:
        s = Session(circuit, first, count)
        link_sessions([s], group)
    s.upload(wit_gf2, wit_z64, seeds)
    s.prove()
    _, proof = s.fetch()
    return proof if rank == 0 else None


def prove_sharded(circuit, wit_gf2, wit_z64=(), seeds=None, group=None, session=None) -> Optional[bytes]:
    """Proof::new over all ranks of `group` (NCCL, one GPU per rank).  `seeds` must be the same 256 x 16 bytes on every rank
    (rank 0 may draw them and broadcast).  Returns the proof on rank 0."""
    import torch
    import torch.distributed as dist

    from .proof import Session

    if seeds is None:
        raise ValueError("sharded proving needs the same seeds on every rank: draw them on rank 0 and broadcast")
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    first, count = shard_of(rank, world)
    s = session or Session(circuit, first, count)
    s.upload(wit_gf2, wit_z64, seeds)
    s.commit()
    recv = s.all_hashes_device()
    with torch.cuda.stream(torch.cuda.ExternalStream(s.stream)):
        all_gather_hashes_into(torch.as_tensor(recv, device="cuda"), torch.as_tensor(s.hashes_device(), device="cuda"), group)
    s.open(recv.ptr)
    out = reduce_proofs([s], group)
    return out[0] if out is not None else None
