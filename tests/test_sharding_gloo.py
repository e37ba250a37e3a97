"""Host logic of the N > 1 path on CPU: world_size-2 `gloo` process group (no GPU).  The device commit cannot run here, so
the oracle stands in for it (test infrastructure) to produce each rank's repetition hashes and shard blob; what is under
test is the product's own sharding code: shard ranges, the all-gather order, and the C-ABI assembly of the shard blobs."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _shard_blob(full: bytes, omit: np.ndarray, lo: int, hi: int, has_rep0: bool) -> bytes:
    """What a shard's extraction kernels write: the full-length proof with only the entries of repetitions [lo, hi) filled
    in (plus the vector-length headers, which repetition 0's CTA writes)."""
    import reverie_oracle as R

    d = R.deserialize(full)
    out = bytearray(len(full))
    pos = 32
    for dom in ("gf2", "z64"):
        on = d[dom]["online"]
        opened = [r for r in range(256) if omit[r] < 8]
        closed = [r for r in range(256) if omit[r] >= 8]
        if has_rep0:
            out[pos:pos + 8] = full[pos:pos + 8]
        pos += 8
        for k, o in enumerate(on):
            sz = 1 + 128 + 24 + len(o["recons"]) + len(o["corrs"]) + len(o["inputs"])
            if lo <= opened[k] < hi:
                out[pos:pos + sz] = full[pos:pos + sz]
            pos += sz
        if has_rep0:
            out[pos:pos + 8] = full[pos:pos + 8]
        pos += 8
        for k in range(216):
            if lo <= closed[k] < hi:
                out[pos:pos + 48] = full[pos:pos + 48]
            pos += 48
    assert pos == len(full)
    return bytes(out)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    import orc
    import reverie_oracle as R
    from reverie_b200 import circuits as C
    from reverie_b200 import sharding

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        seeds = b"".join(R.default_seeds())
        ops, wc = C.flat_mul_circuit(37)
        rc, full, hashes = orc.prove(ops, [1, 1], [], wc, seeds, want_hashes=True)
        assert rc == 0
        first, count = sharding.shard_of(rank, world)
        assert (first, count) == (rank * 32 // world, 32 // world)
        mine = torch.frombuffer(bytearray(hashes[first * 256:(first + count) * 256]), dtype=torch.uint8)
        allh = sharding.all_gather_hashes(mine)
        assert bytes(allh.numpy()) == hashes  # rank order = instance order = repetition order
        omit = orc.challenge(full[:32])
        part = _shard_blob(full, omit, 8 * first, 8 * (first + count), rank == 0)
        proof = sharding.gather_parts(full[:32], part)
        if rank == 0:
            assert proof == full
            assert orc.verify(ops, wc, proof)[0] == 1
        else:
            assert proof is None
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_shard_gather_assemble():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res


def test_shard_ranges():
    from reverie_b200 import sharding

    for w in (1, 2, 4, 8, 16, 32):
        cover = []
        for r in range(w):
            f, c = sharding.shard_of(r, w)
            cover += list(range(f, f + c))
        assert cover == list(range(32))
    with pytest.raises(ValueError):
        sharding.shard_of(0, 3)
