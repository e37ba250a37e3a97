"""GPU parity tests: the CUDA path through the C ABI vs. the CPU oracle, byte for byte (bit-exact bar: integer work)."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb():
    import reverie_b200 as rb
    from reverie_b200 import _native

    assert _native.lib().rv_device_count() >= 1, "no CUDA device: reverie_b200 has no CPU fallback"
    return rb


def _random_circuit(rng, n_inputs, n_gates, p_mul=0.35, n_cells=None, asserts=True):
    """Random GF(2) op list with wire-cell reuse (cells are overwritten like in the reference's bench circuit)."""
    from reverie_b200 import circuits as C

    n_cells = n_cells or max(4, n_inputs + n_gates // 3)
    recs = []
    live = []
    for i in range(n_inputs):
        recs.append((C.GF2, C.INPUT, 0, i % n_cells, 0, 0, 0))
        live.append(i % n_cells)
    for _ in range(n_gates):
        r = rng.random()
        dst = int(rng.integers(0, n_cells))
        a = int(rng.choice(live)) if live else 0
        bb = int(rng.choice(live)) if live else 0
        if r < p_mul:
            recs.append((C.GF2, C.MUL, 0, dst, a, bb, 0))
        elif r < 0.75:
            recs.append((C.GF2, C.ADD if rng.random() < 0.7 else C.SUB, 0, dst, a, bb, 0))
        elif r < 0.85:
            recs.append((C.GF2, C.ADDC if rng.random() < 0.5 else C.SUBC, 0, dst, a, 0, int(rng.integers(0, 2))))
        elif r < 0.92:
            recs.append((C.GF2, C.MULC, 0, dst, a, 0, int(rng.integers(0, 2))))
        else:
            recs.append((C.GF2, C.CONST, 0, dst, 0, 0, int(rng.integers(0, 2))))
        live.append(dst)
    ops = np.array(recs, dtype=C.OP_DTYPE)
    wit = rng.integers(0, 2, size=n_inputs).astype(np.uint8)
    if asserts:
        vals, _ = C.evaluate_gf2(ops, wit, n_cells)
        extra = []
        for w in rng.choice(n_cells, size=min(5, n_cells), replace=False):
            if vals[w] == 0:
                extra.append((C.GF2, C.ASSERT_ZERO, 0, 0, int(w), 0, 0))
        if extra:
            ops = np.concatenate([ops, np.array(extra, dtype=C.OP_DTYPE)])
    return ops, wit, (0, n_cells)


def _check(rb, ops, wit, wc, seeds):
    import orc

    rc, want = orc.prove(ops, wit, [], wc, seeds)
    assert rc == 0
    got = rb.Proof.new(ops, wit, (), wc, seeds=seeds).serialize()
    assert len(got) == len(want)
    assert got == want, "first differing byte at %d" % next(i for i in range(len(want)) if got[i] != want[i])
    assert orc.verify(ops, wc, got)[0] == 1
    return got


def test_tiny_mul(rb, default_seeds):
    from reverie_b200 import circuits as C

    ops, wc = C.flat_mul_circuit(1)
    _check(rb, ops, [1, 1], wc, default_seeds)


@pytest.mark.parametrize("n_mul", [0, 7, 8, 9, 63, 64, 65, 1023, 1024, 1025, 5000])
def test_flat_mul_lengths(rb, default_seeds, n_mul):
    """pack length quirk floor(n/8)+1 and BLAKE3 chunk boundaries (src/algebra/gf2/share.rs:131-138, src/crypto/hash.rs:5)."""
    from reverie_b200 import circuits as C

    ops, wc = C.flat_mul_circuit(n_mul)
    _check(rb, ops, [1, 0], wc, default_seeds)


@pytest.mark.parametrize("seed", range(6))
def test_random_circuits(rb, default_seeds, seed):
    rng = np.random.default_rng(seed)
    ops, wit, wc = _random_circuit(rng, int(rng.integers(1, 40)), int(rng.integers(1, 3000)))
    _check(rb, ops, wit, wc, default_seeds)


def test_sha256(rb, default_seeds):
    from reverie_b200 import circuits as C

    ops, wit, wc = C.sha256_abc_case()
    got = _check(rb, ops, wit, wc, default_seeds)
    assert len(got) == 263960


def test_other_seeds(rb):
    rng = np.random.default_rng(99)
    seeds = rng.integers(0, 256, size=256 * 16, dtype=np.uint8).tobytes()
    ops, wit, wc = _random_circuit(rng, 10, 500)
    _check(rb, ops, wit, wc, seeds)


def test_witness_errors(rb, default_seeds):
    from reverie_b200 import circuits as C

    ops, wit, wc = C.sha256_abc_case()
    bad = wit.copy()
    bad[3] ^= 1
    with pytest.raises(rb.WitnessError) as e:
        rb.Proof.new(ops, bad, (), wc, seeds=default_seeds)
    assert e.value.code == -1
    with pytest.raises(rb.WitnessError) as e:
        rb.Proof.new(ops, wit[:100], (), wc, seeds=default_seeds)
    assert e.value.code == -2


def test_sharding_invariance(rb, default_seeds):
    """The proof must not depend on how the 32 packed instances are sharded: G = 1, 2, 4, 8 shards on one GPU."""
    import orc

    rng = np.random.default_rng(5)
    ops, wit, wc = _random_circuit(rng, 20, 1500)
    rc, want = orc.prove(ops, wit, [], wc, default_seeds)
    circ = rb.Circuit(ops, wc)
    for G in (1, 2, 4, 8):
        per = 32 // G
        sess = [rb.Session(circ, g * per, per) for g in range(G)]
        for s in sess:
            s.upload(wit, (), default_seeds)
            s.commit()
        allh = b"".join(s.hashes() for s in sess)  # the all-gather
        for s in sess:
            s.open(allh)
        parts = [s.fetch() for s in sess]
        assert len({c for c, _ in parts}) == 1
        proof = rb.assemble(parts[0][0], [p for _, p in parts])
        assert proof == want, f"G={G}"


def test_os_rng_seeds_verify(rb):
    """seeds=NULL draws from the OS RNG like the reference (src/proof/mod.rs:131-134); the oracle's verifier must accept."""
    import orc

    rng = np.random.default_rng(11)
    ops, wit, wc = _random_circuit(rng, 12, 800)
    p1 = rb.Proof.new(ops, wit, (), wc).serialize()
    p2 = rb.Proof.new(ops, wit, (), wc).serialize()
    assert p1 != p2
    assert orc.verify(ops, wc, p1)[0] == 1 and orc.verify(ops, wc, p2)[0] == 1
