"""The two CPU restatements against each other and against the reference's own test cases re-expressed
(SURVEY.md section 4): gate truth tables (src/interpreter/single.rs:231-539), the end-to-end B2A circuit
(src/proof/mod.rs:397-427), pack/unpack round trips (src/algebra/mod.rs:210-409), the omitted-player property
(src/generator/share.rs:76-141), plus tamper tests."""
import numpy as np
import pytest

import orc
import reverie_oracle as R
from reverie_b200 import circuits as C

ZERO_SEEDS = [bytes(16)] * 8


def _value(dom, circ, wit):
    """Run one packed instance with all-zero seeds like the reference's tests (single.rs:193,220) -> instance."""
    D = R.GF2 if dom == "gf2" else R.Z64
    cells = 1 + max(max(a for a in op[2:] if isinstance(a, int) and not isinstance(a, bool)) for op in circ)
    ins = R.Instance(D, R.ProverTranscript(D, wit, ZERO_SEEDS), cells)
    for op in circ:
        ins.step(*op[1:])
    return ins


@pytest.mark.parametrize("a,b", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gf2_truth_tables(a, b):
    """single.rs:231-376: mul / add / sub / addc / mulc over GF(2); value() = recon(mask) + corr."""
    base = [("gf2", "Input", 0), ("gf2", "Input", 1)]
    for name, want in (("Mul", a & b), ("Add", a ^ b), ("Sub", a ^ b)):
        ins = _value("gf2", base + [("gf2", name, 2, 0, 1)], [a, b])
        assert ins.value(2) == (R.M64 if want else 0)
    for c in (False, True):
        ins = _value("gf2", base + [("gf2", "AddConst", 2, 0, c), ("gf2", "MulConst", 3, 0, c), ("gf2", "SubConst", 4, 1, c)], [a, b])
        assert ins.value(2) == (R.M64 if a ^ c else 0)
        assert ins.value(3) == (R.M64 if a & c else 0)
        assert ins.value(4) == (R.M64 if b ^ c else 0)


@pytest.mark.parametrize("a,b", [(0, 0), (1, 2), (2**63, 2**63), (2**64 - 1, 2), (0x0123456789ABCDEF, 0xFEDCBA9876543210)])
def test_z64_truth_tables(a, b):
    """single.rs:378-539 incl. wrapping."""
    base = [("z64", "Input", 0), ("z64", "Input", 1)]
    M = R.M64
    for name, want in (("Mul", a * b & M), ("Add", (a + b) & M), ("Sub", (a - b) & M)):
        ins = _value("z64", base + [("z64", name, 2, 0, 1)], [a, b])
        assert ins.value(2) == (want,) * 8
    ins = _value("z64", base + [("z64", "AddConst", 2, 0, 7), ("z64", "MulConst", 3, 0, 7), ("z64", "SubConst", 4, 1, 7)], [a, b])
    assert ins.value(2) == ((a + 7) & M,) * 8 and ins.value(3) == ((a * 7) & M,) * 8 and ins.value(4) == ((b - 7) & M,) * 8


def test_assert_zero_on_zero_passes_and_nonzero_raises():
    _value("gf2", [("gf2", "Input", 0), ("gf2", "AssertZero", 0)], [0])  # single.rs:231-250
    with pytest.raises(R.WitnessError):
        _value("gf2", [("gf2", "Input", 0), ("gf2", "AssertZero", 0)], [1])


E2E = ([("gf2", "Input", i) for i in range(64)] + [("b2a", 0, 0), ("z64", "Input", 1), ("z64", "Input", 2), ("z64", "Mul", 3, 1, 2),
                                                    ("z64", "Add", 3, 3, 0), ("z64", "Mul", 3, 3, 2)])


def test_reference_e2e_circuit_roundtrip(default_seeds):
    """proof/mod.rs:397-427 (64 Inputs, B2A, Z64 Mul/Add/Mul; witness all-true): prove -> verify, both oracles, same bytes."""
    seeds = [default_seeds[16 * i : 16 * i + 16] for i in range(256)]
    wc = R.largest_wires(E2E)
    p = R.prove(E2E, [1] * 64, [5, 7], wc, seeds)
    blob = R.serialize(p)
    assert R.verify(R.deserialize(blob), E2E, wc)
    ops = orc.tuples_to_ops(E2E)
    rc, pb = orc.prove(ops, [1] * 64, [5, 7], wc, default_seeds)
    assert rc == 0 and pb == blob
    assert orc.verify(ops, wc, pb) == (1, True)


@pytest.mark.parametrize("seed", range(4))
def test_oracles_agree_on_random_mixed_circuits(seed, default_seeds):
    rng = np.random.default_rng(seed)
    circ = []
    ng, nz = 12, 6
    wg, wz = [], []
    for i in range(4):
        circ.append(("gf2", "Input", i)); wg.append(int(rng.integers(0, 2)))
    for i in range(3):
        circ.append(("z64", "Input", i)); wz.append(int(rng.integers(0, 2**63)))
    names = ["Add", "Sub", "Mul", "AddConst", "SubConst", "MulConst", "Const", "Random"]
    for _ in range(int(rng.integers(5, 60))):
        dom = "gf2" if rng.random() < 0.6 else "z64"
        n = ng if dom == "gf2" else nz
        name = names[int(rng.integers(0, len(names)))]
        d, a, b = (int(rng.integers(0, n)) for _ in range(3))
        cst = bool(rng.integers(0, 2)) if dom == "gf2" else int(rng.integers(0, 2**64, dtype=np.uint64))
        if name in ("Add", "Sub", "Mul"):
            circ.append((dom, name, d, a, b))
        elif name in ("AddConst", "SubConst", "MulConst"):
            circ.append((dom, name, d, a, cst))
        elif name == "Const":
            circ.append((dom, name, d, cst))
        else:
            circ.append((dom, name, d))
    wc = (nz, ng)
    seeds = [default_seeds[16 * i : 16 * i + 16] for i in range(256)]
    tap = {}
    p = R.prove(circ, wg, wz, wc, seeds, tap=tap)
    blob = R.serialize(p)
    ops = orc.tuples_to_ops(circ)
    rc, pb, hashes = orc.prove(ops, wg, wz, wc, default_seeds, want_hashes=True)
    assert rc == 0 and pb == blob and hashes == b"".join(tap["rep_hashes"])
    assert R.verify(R.deserialize(pb), circ, wc) and orc.verify(ops, wc, pb)[0] == 1


LENS = [0, 1, 2, 3, 6, 18, 32, 63, 64, 65, 127, 128]


@pytest.mark.parametrize("n", LENS)
def test_gf2_pack_roundtrip_and_length_quirk(n):
    """algebra/mod.rs:210-409 lengths; the packed length is floor(n/8)+1 (gf2/share.rs:131-138, gf2/recon.rs:224-229)."""
    rng = np.random.default_rng(n)
    shares = [int(x) for x in rng.integers(0, 2**64, size=n, dtype=np.uint64)]
    sel = [int(x) for x in rng.integers(0, 8, size=8)]
    packed = R.GF2.pack_selected(shares, sel)
    assert all(len(p) == n // 8 + 1 for p in packed)
    un = R.GF2.unpack_selected(packed, sel)
    for k in range(n):
        for r in range(8):
            bit = 63 - (8 * r + sel[r])
            assert (un[k] >> bit) & 1 == (shares[k] >> bit) & 1
            assert un[k] & (0xFF << (8 * (7 - r))) & ~(1 << bit) == 0  # every other player's bit is zero
    recons = [sum((0xFF << (8 * (7 - r))) for r in range(8) if rng.integers(0, 2)) for _ in range(n)]
    pr = R.GF2.pack_recon(recons, [True] * 8)
    assert all(len(p) == n // 8 + 1 for p in pr)
    assert R.GF2.unpack_recon(pr)[:n] == recons
    assert R.GF2.pack_recon(recons, [False] * 8) == [b""] * 8


@pytest.mark.parametrize("n", LENS)
def test_z64_pack_roundtrip(n):
    rng = np.random.default_rng(100 + n)
    shares = [tuple(tuple(int(x) for x in rng.integers(0, 2**64, size=8, dtype=np.uint64)) for _ in range(8)) for _ in range(n)]
    sel = [int(x) for x in rng.integers(0, 8, size=8)]
    packed = R.Z64.pack_selected(shares, sel)
    assert all(len(p) == 8 * n for p in packed)
    un = R.Z64.unpack_selected(packed, sel)
    for k in range(n):
        for r in range(8):
            assert un[k][r][sel[r]] == shares[k][r][sel[r]]


def test_share_generator_omitted_player_property(default_seeds):
    """generator/share.rs:76-141: with a player omitted, every other coordinate is unchanged and the omitted one is 0."""
    rng = np.random.default_rng(3)
    seeds8 = default_seeds[:128]
    full = orc.gf2_masks(seeds8, [8] * 8, 1000)
    omit = [int(x) for x in rng.integers(0, 8, size=8)]
    part = orc.gf2_masks(seeds8, omit, 1000)
    keep = 0
    for r in range(8):
        for p in range(8):
            if p != omit[r]:
                keep |= 1 << (63 - (8 * r + p))
    assert ((full & np.uint64(keep)) == part).all()


def _sha():
    return C.sha256_abc_case()


def test_sha256_circuit_is_sha256():
    import hashlib

    ops, n_wires, out = C.sha256_compress_circuit(None)
    for msg in (b"abc", b"", b"reverie-b200"):
        w, _ = C.evaluate_gf2(ops, C.sha256_witness(C.sha256_pad_single_block(msg)), n_wires)
        assert np.packbits(w[out]).tobytes() == hashlib.sha256(msg).digest()
    assert int((ops["opcode"] == C.MUL).sum()) == 22573  # the public Bristol-Fashion sha256.txt has the same AND count


def test_bristol_fashion_roundtrip():
    b = C.Builder()
    x = [b.input() for _ in range(4)]
    o0 = b.add(b.mul(x[0], x[1]), x[2])
    o1 = b.addc(b.mul(o0, x[3]), 1)
    ops = b.ops()
    text = C.to_bristol_fashion(ops, [4], [o1])
    ops2, n2, outs = C.parse_bristol_fashion(text)
    assert (ops2 == ops).all() and n2 == b.n_wires and outs == [o1]
    for wit in ([0, 0, 0, 0], [1, 1, 0, 1], [1, 1, 1, 1]):
        v1, _ = C.evaluate_gf2(ops, wit, n2)
        ops3, n3, _ = C.parse_bristol_fashion(text, expected_outputs=[int(v1[o1])])
        assert C.evaluate_gf2(ops3, wit, n3)[1]


def test_sha256_proof_verifies_and_tampering_is_rejected(default_seeds):
    ops, wit, wc = _sha()
    rc, pb = orc.prove(ops, wit, [], wc, default_seeds)
    assert rc == 0 and len(pb) == 263960
    assert orc.verify(ops, wc, pb) == (1, True)
    rng = np.random.default_rng(0)
    for pos in [0, 31, 40, 41, 200, 5000, len(pb) // 2, len(pb) - 1] + [int(x) for x in rng.integers(32, len(pb), size=6)]:
        bad = bytearray(pb)
        bad[pos] ^= 0x10
        rc, _ = orc.verify(ops, wc, bytes(bad))
        assert rc != 1, pos
    assert orc.verify(ops, wc, pb[:-1])[0] < 0  # truncated: format error
    bad_wit = wit.copy()
    bad_wit[100] ^= 1
    assert orc.prove(ops, bad_wit, [], wc, default_seeds)[0] == orc.E_WITNESS_INVALID
    assert orc.prove(ops, wit[:10], [], wc, default_seeds)[0] == orc.E_WITNESS_SHORT


def test_lowmem_two_pass_oracle_matches(default_seeds):
    """orc_prove_lowmem (hashes first, then every instance again for its openings) produces orc_prove's bytes: GF(2) with
    asserts, a mixed GF(2) / Z64 / B2A circuit, and a failing witness."""
    import hashlib

    from reverie_b200 import circuits as C
    from tests._zgen import random_z_circuit

    ops, wit, wc = C.sha256_abc_case()
    rc, want = orc.prove(ops, wit, [], wc, default_seeds)
    assert (rc, hashlib.sha256(want).hexdigest(), len(want)) == orc.prove_digest_lowmem(ops, wit, [], wc, default_seeds, n_threads=3)
    zops, gwit, zwit, zwc = random_z_circuit(np.random.default_rng(3), 4, 200, with_gf2=True)
    rc, want = orc.prove(zops, gwit, zwit, zwc, default_seeds)
    assert (rc, hashlib.sha256(want).hexdigest(), len(want)) == orc.prove_digest_lowmem(zops, gwit, zwit, zwc, default_seeds, n_threads=2)
    bad = wit.copy()
    bad[11] ^= 1
    assert orc.prove_digest_lowmem(ops, bad, [], wc, default_seeds)[0] == orc.prove(ops, bad, [], wc, default_seeds)[0] != 0
