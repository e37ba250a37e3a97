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
    """The proof must not depend on how the 32 packed instances are sharded: G = 1 .. 32 shards on one GPU."""
    import orc

    rng = np.random.default_rng(5)
    ops, wit, wc = _random_circuit(rng, 20, 1500)
    rc, want = orc.prove(ops, wit, [], wc, default_seeds)
    circ = rb.Circuit(ops, wc)
    for G in (1, 2, 4, 8, 16, 32):  # 16 / 32 shards: the mask generator's 4-slice CTAs (spare warps take further counter blocks)
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


def _run_worker(which):
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RV_PEER_TIMEOUT_MS=os.environ.get("RV_PEER_TIMEOUT_MS", "30000"))
    res = subprocess.run([sys.executable, os.path.join(root, "tests", "_linked_worker.py"), which], capture_output=True, text=True, timeout=900, cwd=root, env=env)
    assert res.returncode == 0 and "worker ok" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]


def test_linked_shards_exchange_on_device(rb):
    """rv_session_peer_link with all the shards on ONE GPU (raw-pointer handles): the flag / double-buffer protocol of k_challenge /
    k_xfinish, G = 2, 4, 8, single- and multi-proof sessions, batches, Z64, failed asserts, argument checks; every proof equals
    the oracle's.  Runs in a fresh process (tests/_linked_worker.py says why)."""
    _run_worker("linked")


def test_group_api_local(rb):
    """rv_group_create_local through one handle: waves beyond the capacity, partial sessions, OS-RNG seeds, per-proof errors,
    Z64; on a multi-GPU box the members sit on distinct devices.  Runs in a fresh process."""
    _run_worker("group")


def test_os_rng_seeds_verify(rb):
    """seeds=NULL draws from the OS RNG like the reference (src/proof/mod.rs:131-134); the oracle's verifier must accept."""
    import orc

    rng = np.random.default_rng(11)
    ops, wit, wc = _random_circuit(rng, 12, 800)
    p1 = rb.Proof.new(ops, wit, (), wc).serialize()
    p2 = rb.Proof.new(ops, wit, (), wc).serialize()
    assert p1 != p2
    assert orc.verify(ops, wc, p1)[0] == 1 and orc.verify(ops, wc, p2)[0] == 1


# ---- Proof::verify on the GPU (src/proof/mod.rs:224-307) ------------------------------------------------------------------
def _verify_both(rb, circ, ops, wc, blob):
    """(ours, oracle) verdicts: 1 accept / 0 reject / <0 error."""
    import orc
    from reverie_b200 import _native as N

    want = orc.verify(ops, wc, blob)[0]
    try:
        got = 1 if rb.Proof(blob).verify(circ) else 0
    except rb.ReverieError as e:
        got = e.code
    if want < 0:
        want = N.E_FORMAT
    return got, want


@pytest.mark.parametrize("seed", range(4))
def test_verify_accepts_and_rejects_like_the_oracle(rb, default_seeds, seed):
    import orc

    rng = np.random.default_rng(50 + seed)
    ops, wit, wc = _random_circuit(rng, int(rng.integers(1, 40)), int(rng.integers(1, 3000)))
    circ = rb.Circuit(ops, wc)
    blob = rb.Proof.new(circ, wit, (), seeds=default_seeds).serialize()
    assert _verify_both(rb, circ, ops, wc, blob) == (1, 1)
    assert rb.Proof(orc.prove(ops, wit, [], wc, default_seeds)[1]).verify(circ)  # the oracle's proof bytes, our verifier
    for pos in [0, 31, 32, 40, 41, 170, 171, len(blob) // 2, len(blob) - 1] + [int(x) for x in rng.integers(32, len(blob), size=12)]:
        bad = bytearray(blob)
        bad[pos] ^= 1 << int(rng.integers(0, 8))
        got, want = _verify_both(rb, circ, ops, wc, bytes(bad))
        assert got == want, f"byte {pos}: ours {got}, oracle {want}"
    for cut in (1, 7, 48, 1000):
        assert _verify_both(rb, circ, ops, wc, blob[:-cut])[0] == -3  # truncated -> RV_E_FORMAT
    assert _verify_both(rb, circ, ops, wc, blob + b"\x00")[0] == -3


def test_verify_sha256_and_flat_lengths(rb, default_seeds):
    from reverie_b200 import circuits as C

    ops, wit, wc = C.sha256_abc_case()
    circ = rb.Circuit(ops, wc)
    blob = rb.Proof.new(circ, wit, (), seeds=default_seeds).serialize()
    assert rb.Proof(blob).verify(circ)
    bad = bytearray(blob)
    bad[100000] ^= 0x20
    assert not rb.Proof(bytes(bad)).verify(circ)
    for n in (0, 1, 7, 8, 9, 1025):
        ops, wc = C.flat_mul_circuit(n)
        circ = rb.Circuit(ops, wc)
        blob = rb.Proof.new(circ, [1, 0], (), seeds=default_seeds).serialize()
        assert rb.Proof(blob).verify(circ), n
        other = rb.Proof.new(circ, [1, 1], (), seeds=default_seeds).serialize()
        assert rb.Proof(other).verify(circ)
        mixed = blob[:32] + other[32:]  # commitment of one proof, openings of another
        assert not rb.Proof(mixed).verify(circ), n


def test_verify_wrong_circuit_rejects(rb, default_seeds):
    rng = np.random.default_rng(77)
    ops, wit, wc = _random_circuit(rng, 16, 600)
    blob = rb.Proof.new(ops, wit, (), wc, seeds=default_seeds).serialize()
    ops2 = ops.copy()
    k = int(np.where(ops2["opcode"] == 6)[0][3])
    ops2["opcode"][k] = 2  # one Mul becomes an Add: stream lengths change -> hashes differ
    got, want = _verify_both(rb, rb.Circuit(ops2, wc), ops2, wc, blob)
    assert got == want and got != 1


def test_golden_fixtures(rb, default_seeds):
    """GPU proofs against tests/golden/proofs.json (no oracle run involved)."""
    import hashlib
    import json
    import os

    from tests.golden.make_golden import cases

    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "proofs.json")))["cases"]
    for name, (ops, wit, wc) in cases().items():
        blob = rb.Proof.new(ops, wit, (), wc, seeds=default_seeds).serialize()
        assert len(blob) == gold[name]["proof_len"] and hashlib.sha256(blob).hexdigest() == gold[name]["proof_sha256"], name
        assert blob[:32].hex() == gold[name]["comm"], name
    from tests.golden.make_golden import zcases

    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "proofs.json")))["zcases"]
    for name, (ops, gwit, zwit, wc) in zcases().items():
        blob = rb.Proof.new(ops, gwit, zwit, wc, seeds=default_seeds).serialize()
        assert len(blob) == gold[name]["proof_len"] and hashlib.sha256(blob).hexdigest() == gold[name]["proof_sha256"], name


def test_aes128_config1(rb, default_seeds):
    """AES-128 (6400 AND), FIPS-197 C.1 witness: prove parity, verify, reject a wrong key."""
    from reverie_b200 import circuits as C

    ops, wit, wc = C.aes128_fips197_case()
    circ = rb.Circuit(ops, wc)
    assert circ.stats()["n_and"] == 6400
    blob = _check(rb, ops, wit, wc, default_seeds)
    assert _verify_both(rb, circ, ops, wc, blob) == (1, 1)
    bad = wit.copy()
    bad[17] ^= 1
    with pytest.raises(rb.WitnessError):
        rb.Proof.new(circ, bad, (), seeds=default_seeds)


# ---- Z64 domain (src/algebra/z64/*) and mixed circuits --------------------------------------------------------------------
def _check_z(rb, ops, gwit, zwit, wc, seeds):
    import orc

    rc, want = orc.prove(ops, gwit, zwit, wc, seeds)
    assert rc == 0
    got = rb.Proof.new(ops, gwit, zwit, wc, seeds=seeds).serialize()
    assert len(got) == len(want)
    assert got == want, "first differing byte at %d" % next(i for i in range(len(want)) if got[i] != want[i])
    return got


@pytest.mark.parametrize("seed", range(6))
def test_z64_random_circuits_prove_and_verify(rb, default_seeds, seed):
    import orc
    from tests._zgen import random_z_circuit

    rng = np.random.default_rng(2000 + seed)
    ops, gwit, zwit, wc = random_z_circuit(rng, int(rng.integers(1, 8)), int(rng.integers(0, 1200)), n_cells=int(rng.integers(4, 64)),
                                           with_gf2=seed % 2 == 1)
    blob = _check_z(rb, ops, gwit, zwit, wc, default_seeds)
    circ = rb.Circuit(ops, wc)
    assert _verify_both(rb, circ, ops, wc, blob) == (1, 1)
    for pos in [len(blob) // 2, len(blob) - 60, len(blob) - 20000] + [int(x) for x in rng.integers(32, len(blob), size=10)]:
        bad = bytearray(blob)
        bad[pos] ^= 1 << int(rng.integers(0, 8))
        got, want = _verify_both(rb, circ, ops, wc, bytes(bad))
        assert got == want, f"byte {pos}: ours {got}, oracle {want}"


@pytest.mark.parametrize("n_mul", [0, 1, 15, 16, 17, 127, 128, 129, 2000])
def test_z64_flat_and_config3_lengths(rb, default_seeds, n_mul):
    """BLAKE3 chunk boundaries of the 64-byte-per-Mul online stream and the 8-byte-per-Mul preprocessing stream; the
    reference-shaped flat circuit (src/proof/mod.rs:322-329 over Z64) and SURVEY.md 8(d) config 3's register-file shape."""
    from reverie_b200 import circuits as C
    from tests._zgen import Z64_WITNESS

    ops, wc = C.flat_mul_circuit(n_mul, domain=C.Z64)
    blob = _check_z(rb, ops, (), Z64_WITNESS, wc, default_seeds)
    assert rb.Proof(blob).verify(rb.Circuit(ops, wc))
    ops, nw = C.z64_mul_circuit(n_mul)
    blob = _check_z(rb, ops, (), Z64_WITNESS, (nw, 0), default_seeds)
    assert rb.Proof(blob).verify(rb.Circuit(ops, (nw, 0)))


def test_z64_witness_errors_and_sharding(rb, default_seeds):
    import orc
    from reverie_b200 import circuits as C
    from tests._zgen import random_z_circuit

    b = C.Builder(C.Z64)
    x = b.input()
    b.assert_zero(b.addc(b.mul(x, x), 5))
    with pytest.raises(rb.WitnessError) as e:
        rb.Proof.new(b.ops(), (), [3], (b.n_wires, 0), seeds=default_seeds)
    assert e.value.code == -1
    with pytest.raises(rb.WitnessError) as e:
        rb.Proof.new(b.ops(), (), [], (b.n_wires, 0), seeds=default_seeds)
    assert e.value.code == -2
    rng = np.random.default_rng(31)
    ops, gwit, zwit, wc = random_z_circuit(rng, 5, 700, with_gf2=True)
    rc, want = orc.prove(ops, gwit, zwit, wc, default_seeds)
    circ = rb.Circuit(ops, wc)
    for G in (2, 8):
        per = 32 // G
        sess = [rb.Session(circ, g * per, per) for g in range(G)]
        for s in sess:
            s.upload(gwit, zwit, default_seeds)
            s.commit()
        allh = b"".join(s.hashes() for s in sess)
        for s in sess:
            s.open(allh)
        parts = [s.fetch() for s in sess]
        assert rb.assemble(parts[0][0], [p for _, p in parts]) == want, f"G={G}"


def test_wide_layered_circuit(rb, default_seeds):
    """SURVEY.md 8(d) config 5(ii) at a size the oracle finishes in seconds: wide levels take the per-level value-plane launches."""
    from reverie_b200 import circuits as C

    width, n_and = 8192, 40000
    ops, nw = C.layered_and_circuit(width, n_and)
    wit = np.random.default_rng(0).integers(0, 2, size=width).astype(np.uint8)
    circ = rb.Circuit(ops, (0, nw))
    assert circ.stats()["n_lut_steps"] == 0 and circ.stats()["n_luts"] > 0  # the wide path is the one that runs
    blob = _check(rb, ops, wit, (0, nw), default_seeds)
    assert rb.Proof(blob).verify(circ)


def test_multi_gpu_nccl_sharded_prove(rb):
    """One process per GPU over NCCL: needs >= 2 GPUs on the box; the proof must not depend on the sharding."""
    import os
    import subprocess
    import sys

    from reverie_b200 import _native

    n = _native.lib().rv_device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else (4 if n < 8 else 8)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", os.path.join(root, "tests", "_mgpu_worker.py")], capture_output=True, text=True, timeout=600, cwd=root)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert res.stdout.count("mgpu ok") == 4


def test_cli_prove_verify_roundtrip(rb, tmp_path):
    """python -m reverie_b200 --operation prove / verify / oneshot-zk (src/main.rs:57-165); the oracle accepts the proof file."""
    import orc
    from reverie_b200 import __main__ as cli, circuits as C

    ops, wit, wc = C.aes128_fips197_case()
    prog, w, pf = tmp_path / "aes.rvops", tmp_path / "aes.wit", tmp_path / "aes.proof"
    C.save_ops(str(prog), ops, wc)
    w.write_text("".join(str(int(b)) for b in wit))
    assert cli.main(["--operation", "prove", "--program-path", str(prog), "--witness-path", str(w), "--proof-path", str(pf)]) == 0
    assert orc.verify(ops, wc, pf.read_bytes())[0] == 1
    assert cli.main(["--operation", "verify", "--program-path", str(prog), "--proof-path", str(pf)]) == 0
    blob = bytearray(pf.read_bytes())
    blob[5000] ^= 4
    pf.write_bytes(bytes(blob))
    assert cli.main(["--operation", "verify", "--program-path", str(prog), "--proof-path", str(pf)]) == 255
    assert cli.main(["--operation", "oneshot-zk", "--program-path", str(prog), "--witness-path", str(w)]) == 0
    bad = wit.copy()
    bad[3] ^= 1
    w.write_text("".join(str(int(b)) for b in bad))
    assert cli.main(["--operation", "oneshot-zk", "--program-path", str(prog), "--witness-path", str(w)]) == 255


@pytest.mark.parametrize("seed", range(4))
def test_random_and_b2a(rb, default_seeds, seed):
    """Random wires and B2A conversions (src/interpreter/single.rs:148-150, combine.rs:132-219): prove and verify parity, tampering."""
    from tests._zgen import random_mixed_circuit

    rng = np.random.default_rng(5000 + seed)
    ops, gwit, zwit, wc = random_mixed_circuit(rng, n_b2a=1 + seed % 3, n_random=seed % 4, n_gf2_ops=200, n_z_ops=60)
    blob = _check_z(rb, ops, gwit, zwit, wc, default_seeds)
    circ = rb.Circuit(ops, wc)
    assert _verify_both(rb, circ, ops, wc, blob) == (1, 1)
    for pos in [len(blob) // 2, len(blob) // 3, len(blob) - 5000] + [int(x) for x in rng.integers(32, len(blob), size=8)]:
        bad = bytearray(blob)
        bad[pos] ^= 1 << int(rng.integers(0, 8))
        got, want = _verify_both(rb, circ, ops, wc, bytes(bad))
        assert got == want, f"byte {pos}: ours {got}, oracle {want}"
    badw = gwit.copy()
    badw[3] ^= 1
    with pytest.raises(rb.WitnessError):
        rb.Proof.new(circ, badw, zwit, seeds=default_seeds)


def test_batch_of_sessions_one_graph_per_phase(rb, default_seeds):
    """rv_batch: several proofs in flight driven as one unit (one CUDA graph launch per phase); replays stay bit-exact, a
    bad witness in one session is reported for that session only."""
    import orc
    from reverie_b200 import circuits as C

    ops, wit, wc = C.aes128_fips197_case()
    rc, want = orc.prove(ops, wit, [], wc, default_seeds)
    rng = np.random.default_rng(3)
    seeds2 = rng.integers(0, 256, size=256 * 16, dtype=np.uint8).tobytes()
    rc, want2 = orc.prove(ops, wit, [], wc, seeds2)
    circ = rb.Circuit(ops, wc)
    sess = [rb.Session(circ) for _ in range(3)]
    batch = rb.Batch(sess)
    for rnd in range(4):  # eager, capture, replay, replay
        for k, s in enumerate(sess):
            s.upload(wit, (), seeds2 if (k + rnd) % 2 else default_seeds)
        batch.prove()
        for k, s in enumerate(sess):
            assert s.fetch()[1] == (want2 if (k + rnd) % 2 else want), (rnd, k)
    bad = wit.copy()
    bad[9] ^= 1
    sess[1].upload(bad, (), default_seeds)
    batch.prove()
    assert sess[0].fetch()[1] in (want, want2) and sess[2].fetch()[1] in (want, want2)
    with pytest.raises(rb.WitnessError):
        sess[1].fetch()
    del batch


def test_large_flat_circuit_paths(rb, default_seeds):
    """7x10^5 ANDs: the four-table mask generator (circuits with many waves of work), the pinned zero-copy proof return
    (proofs >= 4 MB) and multi-CTA extraction, all against the oracle's bytes."""
    import orc
    from reverie_b200 import circuits as C

    ops, wc = C.flat_mul_circuit(700000)
    rc, want = orc.prove(ops, [1, 1], [], wc, default_seeds)
    assert rc == 0 and len(want) > (4 << 20)
    circ = rb.Circuit(ops, wc)
    for _ in range(2):  # second proof: pooled pinned buffer, graph replay
        p = rb.Proof.new(circ, [1, 1], (), seeds=default_seeds)
        assert len(p) == len(want) and p.serialize() == want
    assert p.verify(circ)
    del p


def test_long_streams_wide_tree_levels_and_streaming(rb):
    """3 x 10^6 ANDs = 2930 BLAKE3 chunks per repetition and stream: the lower tree levels run grid-wide (k_cv_tree_level, odd
    counts carried up), resident, sharded over two linked sessions, and in streaming mode; a layered circuit with non-contiguous
    reconstruction positions exercises both extraction paths.  Checked against the committed oracle digests (bench_digests.json:
    flat / layered 10^6) and the oracle itself for 3 x 10^6."""
    import hashlib

    import bench
    import orc

    seeds = bench.default_seeds()
    for name in ("flat1000000", "layered1000000"):
        ops, wit, wz, wc, _ = bench.make_workload(name)
        want = bench.golden_digest(name)
        assert hashlib.sha256(rb.Proof.new(ops, wit, (), wc, seeds=seeds).serialize()).hexdigest() == want, name
        assert hashlib.sha256(rb.Proof.new_streaming(ops, wit, wc, seeds=seeds, window_ops=300000).serialize()).hexdigest() == want, name
    # Z64 at 10^5 Muls against the committed oracle digest (eager run, graph capture, replay)
    ops, wit, wz, wc, _ = bench.make_workload("z64mul100000")
    zc = rb.Circuit(ops, wc, prove_only=True)
    for _ in range(3):  # eager, graph capture, replay
        assert hashlib.sha256(memoryview(rb.Proof.new(zc, wit, wz, seeds=seeds)._buf)).hexdigest() == bench.golden_digest("z64mul100000")
    del zc
    ops, wit, wz, wc, _ = bench.make_workload("flat3000000")
    rc, proof = orc.prove(ops, wit, [], wc, seeds)
    want = hashlib.sha256(proof).hexdigest()
    del proof
    circ = rb.Circuit(ops, wc, prove_only=True)
    assert hashlib.sha256(memoryview(rb.Proof.new(circ, wit, (), seeds=seeds)._buf)).hexdigest() == want
    assert hashlib.sha256(memoryview(rb.Proof.new_streaming(ops, wit, wc, seeds=seeds, window_ops=1 << 20)._buf)).hexdigest() == want


def test_cpp_host_mirror(rb):
    """The reference's own proof tests (src/proof/mod.rs:311-428) through the C++ host mirror include/reverie_b200.hpp."""
    import subprocess

    from tests._cppbuild import build_cpp_api_test

    import os

    # (the two-member group of the test shares one GPU: one hardware queue per stream, see tests/_linked_worker.py)
    env = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS="32", RV_PEER_TIMEOUT_MS="30000")
    res = subprocess.run([build_cpp_api_test()], capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0 and "cpp host mirror ok" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


def test_multi_proof_session(rb, default_seeds):
    """rv_session_create_multi: several proofs side by side in one session, every kernel launch covering all of them; each slot's
    bytes must be the oracle's for its OWN witness and seeds (distinct per slot: the per-slot value planes, witness and seed
    addressing are what is under test), across eager run, graph capture and replays; sharded too."""
    import orc
    from reverie_b200 import circuits as C

    ops, n_wires, _ = C.sha256_compress_circuit(None)  # no output asserts: every 768-bit witness is valid
    wc = (0, n_wires)
    rng = np.random.default_rng(8)
    wits = [C.sha256_witness(C.sha256_pad_single_block(m)) for m in (b"abc", b"", b"slot two", b"3" * 55)] + [rng.integers(0, 2, size=768).astype(np.uint8)]
    seeds = [default_seeds] + [rng.integers(0, 256, size=256 * 16, dtype=np.uint8).tobytes() for _ in range(4)]
    want = [orc.prove(ops, w, [], wc, sd)[1] for w, sd in zip(wits, seeds)]
    assert len(set(want)) == 5
    circ = rb.Circuit(ops, wc)
    s = rb.Session(circ, 0, 32, n_proofs=5)
    for rnd in range(4):
        order = [(k + rnd) % 5 for k in range(5)]
        for slot, k in enumerate(order):
            s.upload(wits[k], (), seeds[k], slot=slot)
        s.prove()
        for slot, k in enumerate(order):
            assert s.fetch(slot)[1] == want[k], (rnd, slot)
    # sharded: two "ranks" of 16 instances, 3 proofs each; gathered layout [rank][proof][local hashes]
    sh = [rb.Session(circ, g * 16, 16, n_proofs=3) for g in range(2)]
    for g in range(2):
        for b in range(3):
            sh[g].upload(wits[b + 1], (), seeds[b + 1], slot=b)
        sh[g].commit()
    gathered = b"".join(x.hashes() for x in sh)
    for x in sh:
        x.open(gathered)
    for b in range(3):
        parts = [x.fetch(b) for x in sh]
        assert rb.assemble(parts[0][0], [p for _, p in parts]) == want[b + 1], b
    # a failed AssertZero is reported for its own slot only
    aops, awit, awc = C.sha256_abc_case()
    acirc = rb.Circuit(aops, awc)
    good = orc.prove(aops, awit, [], awc, seeds[1])[1]
    bad = awit.copy()
    bad[11] ^= 1
    s2 = rb.Session(acirc, 0, 32, n_proofs=3)
    for slot, w in enumerate((awit, bad, awit)):
        s2.upload(w, (), seeds[1], slot=slot)
    s2.prove()
    assert s2.fetch(0)[1] == good and s2.fetch(2)[1] == good
    with pytest.raises(rb.WitnessError):
        s2.fetch(1)
    from reverie_b200 import _native as N

    with pytest.raises(rb.ReverieError) as e:  # Z64 circuits keep one proof per session
        zops, zwc = C.flat_mul_circuit(4, domain=C.Z64)
        rb.Session(rb.Circuit(zops, zwc), 0, 32, n_proofs=2)
    assert e.value.code == N.E_UNSUPPORTED


def test_prove_batch_api(rb, default_seeds):
    """rv_prove_batch / Proof.new_batch: a queue of DISTINCT witnesses of one circuit, proved side by side; every proof = the
    oracle's bytes for its own witness and seeds; a bad witness fails alone; Z64 circuits fall back to one proof at a time."""
    import orc
    from reverie_b200 import circuits as C

    ops, n_wires, _ = C.aes128_circuit(None)  # no ciphertext asserts: every (key, plaintext) is a valid witness
    wc = (0, n_wires)
    circ = rb.Circuit(ops, wc)
    rng = np.random.default_rng(12)
    seeds = [default_seeds] + [rng.integers(0, 256, size=256 * 16, dtype=np.uint8).tobytes() for _ in range(10)]
    wits = [C.aes128_witness(rng.bytes(16), rng.bytes(16)) for _ in range(11)]
    want = [orc.prove(ops, w, [], wc, sd)[1] for w, sd in zip(wits, seeds)]
    assert len(set(want)) == 11
    for n in (1, 8, 11, 3):
        proofs = rb.Proof.new_batch(circ, wits[:n], seeds=seeds[:n])
        assert [p.serialize() for p in proofs] == want[:n], n
        assert all(p.verify(circ) for p in proofs[:2])
    aops, awit, awc = C.aes128_fips197_case()
    acirc = rb.Circuit(aops, awc)
    awant = [orc.prove(aops, awit, [], awc, sd)[1] for sd in seeds[:3]]
    bad = awit.copy()
    bad[0] ^= 1
    with pytest.raises(rb.WitnessError) as e:
        rb.Proof.new_batch(acirc, [awit, bad, awit], seeds=seeds[:3])
    got = e.value.proofs
    assert got[1] is None and got[0].serialize() == awant[0] and got[2].serialize() == awant[2]
    with pytest.raises(rb.WitnessError) as e:  # a short witness in the queue fails alone too
        rb.Proof.new_batch(acirc, [awit, awit[:10]], seeds=seeds[:2])
    assert e.value.proofs[1] is None and e.value.proofs[0].serialize() == awant[0]
    from tests._zgen import Z64_WITNESS

    zops, nw = C.z64_mul_circuit(50)
    zc = rb.Circuit(zops, (nw, 0))
    zp = rb.Proof.new_batch(zc, [()] * 3, [Z64_WITNESS] * 3, seeds=seeds[:3])
    assert [p.serialize() for p in zp] == [orc.prove(zops, [], Z64_WITNESS, (nw, 0), sd)[1] for sd in seeds[:3]]


def test_bristol_text_to_proof(rb, default_seeds):
    """SURVEY.md 8(f)1 end to end on the GPU: Bristol-Fashion text -> parse_bristol_fashion -> prove == the oracle's bytes for the
    parsed op list, verify accepts; the parsed circuit computes SHA-256 (checked against hashlib), and a wrong expected
    digest makes the prover refuse."""
    import hashlib

    import orc
    from reverie_b200 import circuits as C

    ops0, n_wires, outs = C.sha256_compress_circuit(None)
    b = C.Builder()
    b.n_wires = n_wires  # Bristol-Fashion wants the outputs as the LAST wires: copy them there with EQW gates
    tail = [b.addc(w, 0) for w in outs]
    ops0 = np.concatenate([ops0, b.ops()])
    text = C.to_bristol_fashion(ops0, [512, 256], tail)
    assert text.splitlines()[0].split()[0] == str(len(ops0) - 768) and "AND" in text and "XOR" in text and "INV" in text
    for msg in (b"abc", b"bristol fashion"):
        digest = hashlib.sha256(msg).digest()
        bits = np.unpackbits(np.frombuffer(digest, dtype=np.uint8)).tolist()
        ops, nw, _ = C.parse_bristol_fashion(text, expected_outputs=bits)
        wit = C.sha256_witness(C.sha256_pad_single_block(msg))
        blob = _check(rb, ops, wit, (0, nw), default_seeds)
        assert rb.Proof(blob).verify(rb.Circuit(ops, (0, nw)))
    bits[5] ^= 1
    ops, nw, _ = C.parse_bristol_fashion(text, expected_outputs=bits)
    with pytest.raises(rb.WitnessError):
        rb.Proof.new(ops, wit, (), (0, nw), seeds=default_seeds)


def test_strict_verify_rejects_skipped_assert(rb, default_seeds):
    """A prover that skips its own AssertZero check (src/transcript/prover.rs:221-228) produces a proof whose commitment
    verifies: the reference accepts it (`okay` is computed at verifier/online.rs:176-178 and never read).  strict=False gives
    that verdict, the default strict mode rejects."""
    import orc
    import reverie_oracle as R
    from reverie_b200 import circuits as C

    b = C.Builder()
    x, y = b.input(), b.input()
    b.assert_zero(b.addc(b.mul(x, y), 1))  # x & y == 1
    ops, wc = b.ops(), (0, b.n_wires)
    keep = R.ProverTranscript.zero_check
    R.ProverTranscript.zero_check = lambda self, recon: None  # the cheating prover
    try:
        forged = R.serialize(R.prove(orc.ops_to_tuples(ops), [1, 0], [], wc, R.default_seeds()))
    finally:
        R.ProverTranscript.zero_check = keep
    assert orc.verify(ops, wc, forged) == (1, False)
    circ = rb.Circuit(ops, wc)
    p = rb.Proof(forged)
    assert p.verify_detail(circ) == (True, False)
    assert p.verify(circ, strict=False) and not p.verify(circ)
    honest = rb.Proof.new(circ, [1, 1], (), seeds=default_seeds)
    assert honest.verify_detail(circ) == (True, True) and honest.verify(circ)


def test_streaming_prove_matches_oracle(rb, default_seeds):
    """rv_prove_streaming: the circuit is proved segment by segment with wires carried across the boundaries,
    PRG / hash streams continued and two passes (hashes, then openings); the bytes must be the oracle's whatever the window.
    Windows are forced far below the circuit sizes: random circuits with heavy cell reuse (imports, exports, slot recycling),
    flat lengths around BLAKE3 chunk boundaries, SHA-256 (LUT value plane + mask VM inside the segments), a wide layered circuit
    (per-level paths), failed asserts and the unsupported shapes."""
    import orc
    from reverie_b200 import circuits as C
    from reverie_b200 import _native as N

    def check(ops, wit, wc, windows, seeds=default_seeds):
        rc, want = orc.prove(ops, wit, [], wc, seeds)
        assert rc == 0
        for w in windows:
            got = rb.Proof.new_streaming(ops, wit, wc, seeds=seeds, window_ops=w).serialize()
            assert len(got) == len(want), (w, len(got), len(want))
            assert got == want, (w, next(i for i in range(len(want)) if got[i] != want[i]))
        return want

    for seed in range(4):
        rng = np.random.default_rng(100 + seed)
        ops, wit, wc = _random_circuit(rng, 24, 3000, n_cells=40 + 30 * seed)
        check(ops, wit, wc, (64, 257, 1000, 10 ** 6))
    for n_mul in (0, 1, 1021, 1022, 1023, 1024, 2047, 2048, 5000):
        ops, wc = C.flat_mul_circuit(n_mul)
        check(ops, np.array([1, 1], dtype=np.uint8), wc, (64, 500, 1024))
    ops, wit, wc = C.sha256_abc_case()
    want = check(ops, wit, wc, (7000, 40000))
    assert rb.Proof(want).verify(rb.Circuit(ops, wc))
    ops, nw = C.layered_and_circuit(8192, 60000)
    wit = np.random.default_rng(0).integers(0, 2, size=8192).astype(np.uint8)
    check(ops, wit, (0, nw), (20000,))
    # a failed AssertZero, a witness that is too short, unsupported domains
    aops, awit, awc = C.sha256_abc_case()
    bad = awit.copy()
    bad[11] ^= 1
    with pytest.raises(rb.WitnessError):
        rb.Proof.new_streaming(aops, bad, awc, seeds=default_seeds, window_ops=9000)
    with pytest.raises(rb.WitnessError):
        rb.Proof.new_streaming(aops, awit[:100], awc, seeds=default_seeds, window_ops=9000)
    zops, zwc = C.flat_mul_circuit(10, domain=C.Z64)
    with pytest.raises(rb.ReverieError) as e:
        rb.Proof.new_streaming(zops, (), zwc, seeds=default_seeds, window_ops=64)
    assert e.value.code == N.E_UNSUPPORTED


@pytest.mark.gpu
def test_one_shot_proof_of_a_big_circuit_streams_first_then_compiles(rb, default_seeds):
    """rv_proof_new (the reference's call shape: the op list with every call) on a circuit above rv_oneshot_streaming_min that it
    has not seen: the first call proves in streaming mode (nothing is compiled for residency, nothing enters the cache), the
    second compiles and caches; both return the oracle's bytes.  Circuits streaming does not serve (Z64) take the resident path
    at once.  The threshold is lowered for the test; the 5 x 10^6-gate flat circuit spans two default windows."""
    import ctypes as C_
    import orc
    from reverie_b200 import circuits as C
    from reverie_b200 import _native as N

    L = N.lib()

    def cache():
        h, m, e = C_.c_uint64(), C_.c_uint64(), C_.c_size_t()
        L.rv_circuit_cache_stats(C_.byref(h), C_.byref(m), C_.byref(e))
        return h.value, m.value, e.value

    L.rv_oneshot_streaming_min(10000)
    try:
        L.rv_circuit_cache_clear()
        ops, wit, wc = C.sha256_abc_case()
        rc, want = orc.prove(ops, wit, [], wc, default_seeds)
        assert rc == 0
        h0, m0, e0 = cache()
        assert rb.Proof.new(ops, wit, (), wc, seeds=default_seeds).serialize() == want  # streamed
        assert cache() == (h0, m0, e0)
        assert rb.Proof.new(ops, wit, (), wc, seeds=default_seeds).serialize() == want  # compiled + cached
        assert cache() == (h0, m0 + 1, e0 + 1)
        assert rb.Proof.new(ops, wit, (), wc, seeds=default_seeds).serialize() == want  # cache hit
        assert cache() == (h0 + 1, m0 + 1, e0 + 1)
        bad = wit.copy()
        bad[3] ^= 1
        L.rv_circuit_cache_clear()
        with pytest.raises(rb.WitnessError):  # another circuit (first sight again): the streaming path reports the failed assert like the resident one
            rb.Proof.new(np.concatenate([ops, ops[-1:]]), bad, (), wc, seeds=default_seeds)
        big, bwc = C.flat_mul_circuit(5_000_000)
        p1 = rb.Proof.new(big, np.array([1, 1], dtype=np.uint8), (), bwc, seeds=default_seeds)
        p2 = rb.Proof.new(big, np.array([1, 1], dtype=np.uint8), (), bwc, seeds=default_seeds)
        assert len(p1) == len(p2) and np.array_equal(np.frombuffer(p1._buf, dtype=np.uint8), np.frombuffer(p2._buf, dtype=np.uint8))
        rcb, digest, nbytes = orc.prove_digest_lowmem(big, [1, 1], [], bwc, default_seeds)
        assert rcb == 0 and nbytes == len(p1) and hashlib.sha256(memoryview(p1._buf)).hexdigest() == digest
        zops, zwc = C.flat_mul_circuit(20000, domain=C.Z64)  # not streamable: resident at the first call
        _, m1, _ = cache()
        zw = np.array([3, 5], dtype=np.uint64)
        rcz, wantz = orc.prove(zops, [], zw, zwc, default_seeds)
        assert rcz == 0 and rb.Proof.new(zops, (), zw, zwc, seeds=default_seeds).serialize() == wantz
        assert cache()[1] == m1 + 1
    finally:
        L.rv_oneshot_streaming_min(1 << 24)
        L.rv_circuit_cache_clear()
