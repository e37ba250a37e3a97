"""Worker for the single-GPU emulation of multi-GPU proving (tests/test_gpu_parity.py runs it in a FRESH process).

Several "ranks" share one GPU here, and a rank's challenge kernel waits on the device for the others: that is only safe while
every stream has its own hardware work queue (CUDA_DEVICE_MAX_CONNECTIONS, set to 32 by reverie_b200._native before CUDA starts)
-- a long-lived test process that has created and dropped hundreds of streams can alias two of them onto one queue, and then a
kernel queued behind a waiting one never starts.  Ranks on their own GPUs (the real deployment, tests/_mgpu_worker.py) have their
own queues by construction."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import pytest  # noqa: E402

from tests.test_gpu_parity import _random_circuit  # noqa: E402


def linked_shards_exchange_on_device(rb, default_seeds):
    """rv_session_peer_link: the shards of a proof exchange their repetition hashes and assemble the proof over (peer) device
    memory inside the open phase -- no host all-gather, no rv_proof_assemble.  Here the G shards live on ONE GPU (same process:
    raw-pointer handles), which exercises the flag / double-buffer protocol of k_challenge / k_xfinish; the proof must equal the
    oracle's for G = 2, 4, 8, through eager run, graph capture and replays, for single- and multi-proof sessions and Z64."""
    import orc
    from reverie_b200 import circuits as C
    from tests._zgen import random_z_circuit

    rng = np.random.default_rng(21)
    ops, wit, wc = _random_circuit(rng, 20, 1500)
    circ = rb.Circuit(ops, wc)
    seeds2 = rng.integers(0, 256, size=256 * 16, dtype=np.uint8).tobytes()
    want = {sd: orc.prove(ops, wit, [], wc, sd)[1] for sd in (default_seeds, seeds2)}
    for G in (2, 4, 8):
        per = 32 // G
        sess = [rb.Session(circ, g * per, per) for g in range(G)]
        handles = [x.peer_handle() for x in sess]
        for g, x in enumerate(sess):
            x.peer_link(g, G, handles)
        for rnd in range(5):
            sd = default_seeds if rnd % 2 == 0 else seeds2
            for x in reversed(sess) if rnd & 1 else sess:  # either launch order: a rank waits on the device, never the host
                x.upload(wit, (), sd)
                x.prove()
            comm, proof = sess[0].fetch()
            assert proof == want[sd], (G, rnd)
            for x in sess[1:]:
                c2, p2 = x.fetch()
                assert c2 == comm and len(p2) == 0
    # multi-proof sessions, driven as batches (one graph per rank and step): SHA-256 without asserts, distinct witnesses per slot
    sops, n_wires, _ = C.sha256_compress_circuit(None)
    swc = (0, n_wires)
    wits = [C.sha256_witness(C.sha256_pad_single_block(m)) for m in (b"abc", b"", b"slot two")]
    sds = [default_seeds, seeds2, bytes(reversed(seeds2))]
    swant = [orc.prove(sops, w, [], swc, sd)[1] for w, sd in zip(wits, sds)]
    scirc = rb.Circuit(sops, swc)
    G = 4
    ranks = [[rb.Session(scirc, g * 8, 8, n_proofs=3) for _ in range(2)] for g in range(G)]  # two sessions per rank
    for i in range(2):
        hs = [ranks[g][i].peer_handle() for g in range(G)]
        for g in range(G):
            ranks[g][i].peer_link(g, G, hs)
    batches = [rb.Batch(ranks[g]) for g in range(G)]
    for rnd in range(4):
        for g in range(G):
            for i in range(2):
                for b in range(3):
                    k = (b + i + rnd) % 3
                    ranks[g][i].upload(wits[k], (), sds[k], slot=b)
            batches[g].prove()
        for i in range(2):
            for b in range(3):
                assert ranks[0][i].fetch(b)[1] == swant[(b + i + rnd) % 3], (rnd, i, b)
    del batches
    # a mixed GF(2) / Z64 circuit (k_zextract writes into the assembling rank's buffer too), and a failed assert seen by every rank
    zops, gwit, zwit, zwc = random_z_circuit(np.random.default_rng(7), 4, 300, with_gf2=True)
    zcirc = rb.Circuit(zops, zwc)
    zs = [rb.Session(zcirc, g * 16, 16) for g in range(2)]
    hs = [x.peer_handle() for x in zs]
    for g, x in enumerate(zs):
        x.peer_link(g, 2, hs)
    for _ in range(3):
        for x in zs:
            x.upload(gwit, zwit, default_seeds)
            x.prove()
        assert zs[0].fetch()[1] == orc.prove(zops, gwit, zwit, zwc, default_seeds)[1]
    aops, awit, awc = C.sha256_abc_case()
    acirc = rb.Circuit(aops, awc)
    xs = [rb.Session(acirc, g * 16, 16) for g in range(2)]
    hs = [x.peer_handle() for x in xs]
    for g, x in enumerate(xs):
        x.peer_link(g, 2, hs)
    bad = awit.copy()
    bad[11] ^= 1
    for w, ok in ((awit, True), (bad, False), (awit, True)):
        for x in xs:
            x.upload(w, (), default_seeds)
            x.prove()
        if ok:
            assert xs[0].fetch()[1] == orc.prove(aops, awit, [], awc, default_seeds)[1]
            xs[1].status()
        else:
            for x in xs:
                with pytest.raises(rb.WitnessError):
                    x.fetch()
    # argument checks: wrong shard order, double link
    ys = [rb.Session(circ, g * 16, 16) for g in range(2)]
    hs = [x.peer_handle() for x in ys]
    with pytest.raises(rb.ReverieError):
        ys[0].peer_link(0, 2, hs[::-1])
    ys[0].peer_link(0, 2, hs)
    with pytest.raises(rb.ReverieError):
        ys[0].peer_link(0, 2, hs)


def group_api_local(rb, default_seeds):
    """rv_group_create_local: one handle, one host thread, `world` members (here all on device 0 -- on a multi-GPU box pass its
    devices); prove / prove_batch return the oracle's bytes, including waves beyond the group's capacity, partially filled
    sessions, OS-RNG seeds (drawn once for all members) and per-proof witness errors."""
    import orc
    from reverie_b200 import circuits as C
    from reverie_b200 import _native as N

    ndev = N.lib().rv_device_count()
    sops, n_wires, _ = C.sha256_compress_circuit(None)
    swc = (0, n_wires)
    rng = np.random.default_rng(33)
    wits = [C.sha256_witness(C.sha256_pad_single_block(bytes([65 + k]) * k)) for k in range(7)]
    sds = [rng.integers(0, 256, size=256 * 16, dtype=np.uint8).tobytes() for _ in range(7)]
    want = [orc.prove(sops, w, [], swc, sd)[1] for w, sd in zip(wits, sds)]
    circ = rb.Circuit(sops, swc)
    for world in (1, 2, 4):
        devices = [r % ndev for r in range(world)]
        g = rb.Group.local(circ, devices, n_sessions=2, slots=2)  # capacity 4 proofs per step: 7 proofs = a full wave + a partial one
        for _ in range(3):
            got = g.prove_batch(wits, None, sds)
            assert [p.serialize() for p in got] == want, world
        p1 = g.prove(wits[3], (), sds[3])
        assert p1.serialize() == want[3]
        p2 = g.prove(wits[0])  # seeds from the OS RNG, the same for every member
        assert p2.verify(circ)
        del g
    # AssertZero failures are per proof; a Z64 circuit goes through a 1 x 1 group
    aops, awit, awc = C.sha256_abc_case()
    acirc = rb.Circuit(aops, awc)
    bad = awit.copy()
    bad[11] ^= 1
    g = rb.Group.local(acirc, [0, 0], n_sessions=1, slots=3)
    with pytest.raises(rb.WitnessError) as e:
        g.prove_batch([awit, bad, awit], None, [default_seeds] * 3)
    good = orc.prove(aops, awit, [], awc, default_seeds)[1]
    assert [None if p is None else p.serialize() for p in e.value.proofs] == [good, None, good]
    del g
    zops, zwc = C.flat_mul_circuit(300, domain=C.Z64)
    zw = np.array([3, 5], dtype=np.uint64)
    zc = rb.Circuit(zops, zwc)
    g = rb.Group.local(zc, [0, 0, 0, 0])
    assert g.prove((), zw, default_seeds).serialize() == orc.prove(zops, [], zw, zwc, default_seeds)[1]
    with pytest.raises(rb.ReverieError):  # multi-proof sessions do not serve Z64
        rb.Group.local(zc, [0, 0], n_sessions=1, slots=2)
    # a queue of proofs verified across the members (whole proofs per GPU); a tampered one is rejected where it sits.  Last, because
    # the verifications leave pooled sessions (streams) behind and this process emulates several ranks on one GPU (see the header).
    g = rb.Group.local(circ, [r % ndev for r in range(2)], n_sessions=1, slots=1)
    got = [rb.Proof(w) for w in want]
    bad = bytearray(want[2])
    bad[5000] ^= 1
    verdicts = g.verify_batch(got[:2] + [rb.Proof(bytes(bad))] + got[3:])
    assert verdicts == [True, True, False] + [True] * (len(got) - 3), verdicts
    del g



if __name__ == "__main__":
    import reverie_b200 as rb
    import reverie_oracle as R

    seeds = b"".join(R.default_seeds())
    {"linked": linked_shards_exchange_on_device, "group": group_api_local}[sys.argv[1]](rb, seeds)
    print("worker ok:", sys.argv[1])
