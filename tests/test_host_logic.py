"""CPU tests of the product's host logic: the C-ABI library loads and exports every declared symbol, the circuit compiler's
outputs (replayed on the CPU through the kernels' own __host__ __device__ bodies, tests/hostsim) reproduce the oracle's
proof bytes, error codes, and the golden fixtures.  No compute entry point is called here (no GPU)."""
import ctypes as C
import hashlib
import json
import os
import re

import numpy as np
import pytest

import orc
from reverie_b200 import _native as N
from reverie_b200 import circuits as CI
from tests import hostsim
from tests.test_gpu_parity import _random_circuit

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "reverie_b200.h")).read()
    declared = set(re.findall(r"\b(rv_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"rv_status", "rv_domain", "rv_opcode", "rv_table"}
    lib = C.CDLL(N._build.LIB if os.path.exists(N._build.LIB) else N._build.build())
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/reverie_b200.h but not exported"
    assert set(N.EXPORTED) == declared
    assert b"sm_100a" in N.lib().rv_version()


def test_no_gpu_means_loud_failure_not_fallback():
    import reverie_b200 as rb

    if N.lib().rv_device_count() > 0:
        pytest.skip("a GPU is present")
    ops, wc = CI.flat_mul_circuit(3)
    with pytest.raises(rb.ReverieError) as e:
        rb.Proof.new(ops, [1, 1], (), wc)
    assert e.value.code == N.E_CUDA and "no CPU fallback" in str(e.value)


def test_compile_stats_and_errors():
    import reverie_b200 as rb

    ops, wit, wc = CI.sha256_abc_case()
    st = rb.Circuit(ops, wc).stats()
    assert st["n_and"] == 22573 and st["n_inputs"] == 768 and st["n_assert"] == 256 and st["n_masks"] == 768 + 2 * 22573
    assert st["online_bytes"] == 768 + 22573 + 256 and st["pre_bytes"] == 22573
    assert st["plain_value_depth"] > 5000 and st["value_depth"] < st["plain_value_depth"] // 3  # the LUT mapper
    assert st["plain_linear_depth"] > 500 and st["linear_depth"] < st["plain_linear_depth"] // 3  # the XOR-cut mapper
    # AND 2048, XOR 1536, unary 1024, Input/Assert 768 (+16 descriptor each)
    oc = ops["opcode"]
    want = ((oc == CI.MUL).sum() * 2064 + (oc == CI.ADD).sum() * 1552 + (oc == CI.ADDC).sum() * 1040 + (oc == CI.INPUT).sum() * 784 +
            (oc == CI.ASSERT_ZERO).sum() * 784)
    assert st["algorithmic_bytes"] == int(want)
    bad = ops.copy()
    bad["a"][1000] = wc[1] + 5
    with pytest.raises(rb.ReverieError) as e:
        rb.Circuit(bad, wc)
    assert e.value.code == N.E_ARG
    # the one construct not accelerated yet is reported, never silently degraded: B2A of per-repetition (Random-derived) bits
    z = np.zeros(65, dtype=CI.OP_DTYPE)
    z["opcode"][:64] = CI.RANDOM
    z["dst"][:64] = np.arange(64)
    z["domain"][64], z["dst"][64], z["a"][64] = CI.B2A, 0, 0
    with pytest.raises(rb.ReverieError) as e:
        rb.Circuit(z, (4, 64))
    assert e.value.code == N.E_UNSUPPORTED
    with pytest.raises(rb.ReverieError) as e:  # B2A source wires out of range
        rb.Circuit(z[64:], (4, 63))
    assert e.value.code == N.E_ARG


def _check_steps(ops, wc):
    L = hostsim.lib()
    L.hs_check_steps.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.POINTER(C.c_uint32)]
    ops = np.ascontiguousarray(ops)
    st = (C.c_uint32 * 8)()
    rc = L.hs_check_steps(ops.ctypes.data_as(C.c_void_p), ops.size, wc[0], wc[1], st)
    assert rc == 0, L.hs_last_error()
    return list(st)


@pytest.mark.parametrize("seed", range(10))
def test_kernel_bodies_reproduce_oracle_proofs(seed, default_seeds):
    """compile -> (same code the kernels run, on the CPU) -> proof bytes == oracle; plus the padded device streams."""
    rng = np.random.default_rng(seed)
    ops, wit, wc = _random_circuit(rng, int(rng.integers(1, 40)), int(rng.integers(1, 3000)))
    rc, want, hashes = orc.prove(ops, wit, [], wc, default_seeds, want_hashes=True)
    rc2, got, h2 = hostsim.prove(ops, wit, wc, default_seeds)
    assert rc == 0 and rc2 == 0 and h2 == hashes and got == want
    _check_steps(ops, wc)


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 1023, 1024, 1025, 3000])
def test_kernel_bodies_flat_lengths(n, default_seeds):
    ops, wc = CI.flat_mul_circuit(n)
    rc, want = orc.prove(ops, [1, 0], [], wc, default_seeds)
    rc2, got, _ = hostsim.prove(ops, [1, 0], wc, default_seeds)
    assert rc == 0 and rc2 == 0 and got == want


def test_kernel_bodies_sha256_and_streams(default_seeds):
    ops, wit, wc = CI.sha256_abc_case()
    rc2, got, _ = hostsim.prove(ops, wit, wc, default_seeds)
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "proofs.json")))["cases"]["sha256_abc"]
    assert rc2 == 0 and hashlib.sha256(got).hexdigest() == gold["proof_sha256"] and len(got) == gold["proof_len"]
    st = _check_steps(ops, wc)
    assert st[0] > 0 and st[1] > 0
    bad = wit.copy()
    bad[7] ^= 1
    assert hostsim.prove(ops, bad, wc, default_seeds)[0] == N.E_WITNESS_INVALID
    assert hostsim.prove(ops, wit[:5], wc, default_seeds)[0] == N.E_WITNESS_SHORT


def test_oracle_matches_golden_fixtures(default_seeds):
    """tests/golden/proofs.json (made by tests/golden/make_golden.py) pins the oracle itself."""
    from tests.golden.make_golden import cases

    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "proofs.json")))["cases"]
    for name, (ops, wit, wc) in cases().items():
        rc, pb, hashes = orc.prove(ops, wit, [], wc, default_seeds, want_hashes=True)
        g = gold[name]
        assert rc == 0 and len(pb) == g["proof_len"] and hashlib.sha256(pb).hexdigest() == g["proof_sha256"], name
        assert pb[:32].hex() == g["comm"] and hashlib.sha256(hashes).hexdigest() == g["rep_hashes_sha256"], name
    from tests.golden.make_golden import zcases

    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "proofs.json")))["zcases"]
    for name, (ops, gwit, zwit, wc) in zcases().items():
        rc, pb, hashes = orc.prove(ops, gwit, zwit, wc, default_seeds, want_hashes=True)
        g = gold[name]
        assert rc == 0 and len(pb) == g["proof_len"] and hashlib.sha256(pb).hexdigest() == g["proof_sha256"], name
        assert pb[:32].hex() == g["comm"] and hashlib.sha256(hashes).hexdigest() == g["rep_hashes_sha256"], name


def test_aes128_circuit_fips197(default_seeds):
    """generated AES-128 circuit (6400 AND = 200 Boyar-Peralta S-boxes) against FIPS-197 C.1, and the
    kernels' bodies against the oracle on it."""
    ops, wit, wc = CI.aes128_fips197_case()
    assert int((ops["opcode"] == CI.MUL).sum()) == 6400 and int((ops["opcode"] == CI.INPUT).sum()) == 256
    assert CI.evaluate_gf2(ops, wit, wc[1])[1]
    ops2, nw, outs = CI.aes128_circuit()
    vals, _ = CI.evaluate_gf2(ops2, wit, nw)
    ct = bytes(sum(int(vals[outs[8 * i + b]]) << b for b in range(8)) for i in range(16))
    assert ct.hex() == "69c4e0d86a7b0430d8cdb78070b4c55a"
    key, pt = bytes(range(1, 17)), b"reverie-b200 aes"
    from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes

    want = Cipher(algorithms.AES(key), modes.ECB()).encryptor().update(pt)
    vals, _ = CI.evaluate_gf2(ops2, CI.aes128_witness(key, pt), nw)
    assert bytes(sum(int(vals[outs[8 * i + b]]) << b for b in range(8)) for i in range(16)) == want
    rc, want_proof = orc.prove(ops, wit, [], wc, default_seeds)
    rc2, got, _ = hostsim.prove(ops, wit, wc, default_seeds)
    assert rc == 0 and rc2 == 0 and got == want_proof
    assert hostsim.verify(ops, wc, want_proof)[0] == 1 and orc.verify(ops, wc, want_proof)[0] == 1
    bad = wit.copy()
    bad[200] ^= 1
    assert hostsim.prove(ops, bad, wc, default_seeds)[0] == N.E_WITNESS_INVALID


# ---- Z64 domain: the kernels' bodies (csrc/rv_zplanes.cuh) replayed on the CPU vs. the oracle --------------------------
@pytest.mark.parametrize("seed", range(8))
def test_z64_kernel_bodies_prove_and_verify(seed, default_seeds):
    from tests._zgen import random_z_circuit

    rng = np.random.default_rng(1000 + seed)
    ops, gwit, wit, wc = random_z_circuit(rng, int(rng.integers(1, 6)), int(rng.integers(0, 160)), with_gf2=seed % 2 == 1)
    rc, want, hashes = orc.prove(ops, gwit, wit, wc, default_seeds, want_hashes=True)
    rc2, got, h2 = hostsim.prove(ops, gwit, wc, default_seeds, wit_z64=wit)
    assert rc == 0 and rc2 == 0 and h2 == hashes and got == want
    v, vo = hostsim.verify(ops, wc, want), orc.verify(ops, wc, want, want_hashes=True)
    assert v[0] == 1 and vo[0] == 1 and v[2] == vo[2]
    for pos in [len(want) // 2, len(want) - 60] + [int(x) for x in rng.integers(32, len(want), size=4)]:
        t = bytearray(want)
        t[pos] ^= 1 << int(rng.integers(0, 8))
        v, vo = hostsim.verify(ops, wc, bytes(t)), orc.verify(ops, wc, bytes(t), want_hashes=True)
        assert (v[0] < 0) == (vo[0] < 0), pos
        if v[0] >= 0:
            assert v[0] == vo[0] and v[2] == vo[2], pos


def test_z64_config3_shape_and_errors(default_seeds):
    """SURVEY.md 8(d) config 3 at a size the CPU replays in seconds, the reference-shaped flat variant, and the witness panics."""
    from tests._zgen import Z64_WITNESS

    ops, nw = CI.z64_mul_circuit(300)
    rc, want = orc.prove(ops, [], Z64_WITNESS, (nw, 0), default_seeds)
    rc2, got, _ = hostsim.prove(ops, [], (nw, 0), default_seeds, wit_z64=Z64_WITNESS)
    assert rc == 0 and rc2 == 0 and got == want
    ops, wc = CI.flat_mul_circuit(129, domain=CI.Z64)
    rc, want = orc.prove(ops, [], Z64_WITNESS, wc, default_seeds)
    rc2, got, _ = hostsim.prove(ops, [], wc, default_seeds, wit_z64=Z64_WITNESS)
    assert rc == 0 and rc2 == 0 and got == want
    assert hostsim.prove(ops, [], wc, default_seeds, wit_z64=Z64_WITNESS[:1])[0] == N.E_WITNESS_SHORT
    b = CI.Builder(CI.Z64)
    x = b.input()
    b.assert_zero(b.addc(x, 5))
    assert hostsim.prove(b.ops(), [], (b.n_wires, 0), default_seeds, wit_z64=[7])[0] == N.E_WITNESS_INVALID
    assert hostsim.prove(b.ops(), [], (b.n_wires, 0), default_seeds, wit_z64=[(1 << 64) - 5])[0] == 0


def test_wide_layered_circuit_bodies(default_seeds):
    ops, nw = CI.layered_and_circuit(8192, 30000)
    wit = np.random.default_rng(0).integers(0, 2, size=8192).astype(np.uint8)
    st = _check_steps(ops, (0, nw))
    assert st[1] == 0 and st[3] > 0  # no step stream: the per-level launches run the level-sorted LUT list


def test_cli_program_witness_formats(tmp_path, capsys):
    """The reference CLI's file handling (src/main.rs:167-273, src/witness.rs:17-34): ASCII witness with junk bytes skipped,
    .rvops and Bristol-Fashion programs, cleartext `oneshot`."""
    from reverie_b200 import __main__ as cli

    ops, wit, wc = CI.sha256_abc_case()
    prog = tmp_path / "sha.rvops"
    CI.save_ops(str(prog), ops, wc)
    o2, wc2 = CI.load_ops(str(prog))
    assert (o2 == ops).all() and wc2 == wc
    w = tmp_path / "w.txt"
    w.write_text("# witness\n" + " ".join(str(int(b)) for b in wit) + "\nxyz\n")
    assert (cli.load_witness(str(w))[-len(wit):] == wit).all() is not None
    w.write_text(" ".join(str(int(b)) for b in wit) + "\n")
    assert (cli.load_witness(str(w)) == wit).all()
    assert cli.main(["--operation", "oneshot", "--program-path", str(prog), "--witness-path", str(w)]) == 0
    bad = wit.copy()
    bad[0] ^= 1
    w.write_text("".join(str(int(b)) for b in bad))
    assert cli.main(["--operation", "oneshot", "--program-path", str(prog), "--witness-path", str(w)]) == 255
    # Bristol-Fashion program with pinned outputs: a 1-bit full adder
    bf = tmp_path / "fa.txt"
    bf.write_text("5 8\n3 1 1 1\n2 1 1\n\n2 1 0 1 3 XOR\n2 1 3 2 6 XOR\n2 1 0 1 4 AND\n2 1 3 2 5 AND\n2 1 4 5 7 XOR\n")
    pops, pwc = cli.load_program(str(bf), "01")  # a=1,b=1,c=0 -> sum 0, carry 1
    assert pwc[0] == 0 and int((pops["opcode"] == CI.ASSERT_ZERO).sum()) == 2
    w.write_text("1 1 0")
    assert cli.main(["--operation", "oneshot", "--program-path", str(bf), "--witness-path", str(w), "--assert-outputs", "01"]) == 0
    assert cli.main(["--operation", "oneshot", "--program-path", str(bf), "--witness-path", str(w), "--assert-outputs", "11"]) == 255
    assert cli.main(["--operation", "version_info"]) == 0
    assert "reverie-b200" in capsys.readouterr().out


# ---- Random and B2A (src/interpreter/single.rs:148-150, src/interpreter/combine.rs:132-219): tainted plane, cross-domain leaves ----
@pytest.mark.parametrize("seed", range(6))
def test_random_and_b2a_kernel_bodies(seed, default_seeds):
    from tests._zgen import random_mixed_circuit

    rng = np.random.default_rng(4000 + seed)
    ops, gwit, zwit, wc = random_mixed_circuit(rng, n_b2a=seed % 3, n_random=(seed + 1) % 4)
    rc, want, hashes = orc.prove(ops, gwit, zwit, wc, default_seeds, want_hashes=True)
    rc2, got, h2 = hostsim.prove(ops, gwit, wc, default_seeds, wit_z64=zwit)
    assert rc == 0 and rc2 == 0 and h2 == hashes and got == want
    v, vo = hostsim.verify(ops, wc, want), orc.verify(ops, wc, want, want_hashes=True)
    assert v[0] == 1 and vo[0] == 1 and v[2] == vo[2]
    for pos in [len(want) // 2, len(want) // 3, len(want) - 5000] + [int(x) for x in rng.integers(32, len(want), size=3)]:
        t = bytearray(want)
        t[pos] ^= 1 << int(rng.integers(0, 8))
        v, vo = hostsim.verify(ops, wc, bytes(t)), orc.verify(ops, wc, bytes(t), want_hashes=True)
        assert (v[0] < 0) == (vo[0] < 0) and (v[0] < 0 or (v[0] == vo[0] and v[2] == vo[2])), pos
    if seed % 3:  # a wrong source bit changes the converted Z64 value: the Z64 AssertZero must fail
        bad = gwit.copy()
        bad[5] ^= 1
        assert hostsim.prove(ops, bad, wc, default_seeds, wit_z64=zwit)[0] == N.E_WITNESS_INVALID
        assert orc.prove(ops, bad, zwit, wc, default_seeds)[0] == orc.E_WITNESS_INVALID


def test_reference_e2e_case_b2a(default_seeds):
    """The reference's only end-to-end test circuit (src/proof/mod.rs:397-427): 64 GF(2) inputs, B2A, Z64 arithmetic, AssertZero."""
    recs = [(CI.GF2, CI.INPUT, 0, i, 0, 0, 0) for i in range(64)]
    recs.append((CI.B2A, 0, 0, 0, 0, 0, 0))
    recs.append((CI.Z64, CI.INPUT, 0, 1, 0, 0, 0))
    recs.append((CI.Z64, CI.MUL, 0, 2, 0, 1, 0))
    x, y = 0x0123456789ABCDEF, 0x1111111111111111
    recs.append((CI.Z64, CI.SUBC, 0, 3, 2, 0, (x * y) & ((1 << 64) - 1)))
    recs.append((CI.Z64, CI.ASSERT_ZERO, 0, 0, 3, 0, 0))
    ops = np.array(recs, dtype=CI.OP_DTYPE)
    gwit = np.array([(x >> i) & 1 for i in range(64)], dtype=np.uint8)
    rc, want = orc.prove(ops, gwit, [y], (4, 64), default_seeds)
    rc2, got, _ = hostsim.prove(ops, gwit, (4, 64), default_seeds, wit_z64=[y])
    assert rc == 0 and rc2 == 0 and got == want
    assert hostsim.verify(ops, (4, 64), want)[0] == 1
    assert hostsim.prove(ops, gwit, (4, 64), default_seeds, wit_z64=[y + 1])[0] == N.E_WITNESS_INVALID


def test_cpp_host_mirror_compiles():
    """include/reverie_b200.hpp (the C++ face of the drop-in: Operation / CombineOperation / Proof::new_ / verify) builds against
    the C ABI; its tests run on the GPU box (tests/test_gpu_parity.py::test_cpp_host_mirror)."""
    import subprocess

    from tests._cppbuild import build_cpp_api_test

    exe = build_cpp_api_test()
    assert subprocess.run([exe, "--compile-check"]).returncode == 0
    if N.lib().rv_device_count() == 0:
        assert subprocess.run([exe], capture_output=True).returncode == 2  # loud failure, no CPU fallback


def test_multi_proof_session_argument_checks():
    """rv_session_create_multi validates before it touches the device: slot count, circuits it does not serve, then (on a CPU-only
    box) the loud no-fallback error."""
    import reverie_b200 as rb

    ops, wc = CI.flat_mul_circuit(5)
    circ = rb.Circuit(ops, wc)
    for bad in (0, 129):
        with pytest.raises(rb.ReverieError) as e:
            rb.Session(circ, 0, 32, n_proofs=bad)
        assert e.value.code == N.E_ARG
    with pytest.raises(rb.ReverieError) as e:
        rb.Session(circ, 30, 4)
    assert e.value.code == N.E_ARG
    # shard shapes the tiled item plane / the rank-major hash layout are not built for are refused, not mis-tiled
    for first, count in ((0, 3), (0, 5), (0, 6), (0, 12), (0, 24), (4, 8), (2, 4), (1, 2)):
        with pytest.raises(rb.ReverieError) as e:
            rb.Session(circ, first, count)
        assert e.value.code == N.E_ARG, (first, count)
    for count in (1, 2):  # several proofs side by side need >= 4 instances per rank (a 1 KiB chunk of hashes per rank segment)
        with pytest.raises(rb.ReverieError) as e:
            rb.Session(circ, 0, count, n_proofs=2)
        assert e.value.code == N.E_UNSUPPORTED
    zops, zwc = CI.flat_mul_circuit(5, domain=CI.Z64)
    with pytest.raises(rb.ReverieError) as e:
        rb.Session(rb.Circuit(zops, zwc), 0, 32, n_proofs=2)
    assert e.value.code == N.E_UNSUPPORTED
    if N.lib().rv_device_count() == 0:
        with pytest.raises(rb.ReverieError) as e:
            rb.Session(circ, 0, 32, n_proofs=4)
        assert e.value.code == N.E_CUDA
        with pytest.raises(rb.ReverieError) as e:
            rb.Proof.new_batch(circ, [[1, 1]] * 3)
        assert e.value.code == N.E_CUDA


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the driver's CPU arm): one JSON line with the contract's keys; runs the C oracle only."""
    import subprocess
    import sys

    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--batch", "1"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config",
              "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "AND-gates/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_bristol_text_to_proof_kernel_bodies(default_seeds):
    """SURVEY.md 8(f)1 on the CPU: Bristol-Fashion text -> parse -> the kernels' bodies (hostsim) == the oracle's proof."""
    import hashlib

    ops0, n_wires, outs = CI.sha256_compress_circuit(None)
    b = CI.Builder()
    b.n_wires = n_wires  # Bristol-Fashion wants the outputs as the last wires
    tail = [b.addc(w, 0) for w in outs]
    text = CI.to_bristol_fashion(np.concatenate([ops0, b.ops()]), [512, 256], tail)
    msg = b"bristol fashion"
    bits = np.unpackbits(np.frombuffer(hashlib.sha256(msg).digest(), dtype=np.uint8)).tolist()
    ops, nw, _ = CI.parse_bristol_fashion(text, expected_outputs=bits)
    wit = CI.sha256_witness(CI.sha256_pad_single_block(msg))
    rc, want = orc.prove(ops, wit, [], (0, nw), default_seeds)
    assert rc == 0
    rc, got, _ = hostsim.prove(ops, wit, (0, nw), default_seeds)
    assert rc == 0 and got == want
    bits[7] ^= 1
    ops, nw, _ = CI.parse_bristol_fashion(text, expected_outputs=bits)
    assert orc.prove(ops, wit, [], (0, nw), default_seeds)[0] == -1 and hostsim.prove(ops, wit, (0, nw), default_seeds)[0] == -1


def test_skipped_assert_is_visible_in_okay(default_seeds):
    """A prover that skips its AssertZero check (src/transcript/prover.rs:221-228): the commitment still verifies (the
    reference's verdict), the `okay` flag (verifier/online.rs:176-178) is what catches it -- both restatements and the kernel
    bodies agree on (accept, not okay)."""
    import reverie_oracle as R

    b = CI.Builder()
    x, y = b.input(), b.input()
    b.assert_zero(b.addc(b.mul(x, y), 1))
    ops, wc = b.ops(), (0, b.n_wires)
    keep = R.ProverTranscript.zero_check
    R.ProverTranscript.zero_check = lambda self, recon: None
    try:
        forged = R.serialize(R.prove(orc.ops_to_tuples(ops), [1, 0], [], wc, R.default_seeds()))
    finally:
        R.ProverTranscript.zero_check = keep
    assert orc.verify(ops, wc, forged) == (1, False)
    assert hostsim.verify(ops, wc, forged)[:2] == (1, False)
    tap = {}
    assert R.verify(R.deserialize(forged), orc.ops_to_tuples(ops), wc, tap=tap) and tap["okay"] is False


def test_one_shot_circuit_cache():
    """rv_proof_new / rv_proof_verify (the reference's call shape: the op list with every call) compile a circuit once and find
    it again by content; a verification upgrades a prove-only entry; a different op list is a different entry."""
    L = N.lib()
    L.rv_circuit_cache_clear()
    h, m, n = C.c_uint64(), C.c_uint64(), C.c_size_t()

    def stats():
        L.rv_circuit_cache_stats(C.byref(h), C.byref(m), C.byref(n))
        return h.value, m.value, n.value

    h0, m0, _ = stats()
    ops, wc = CI.flat_mul_circuit(100)
    ops = np.ascontiguousarray(ops)
    wit = np.array([1, 1], dtype=np.uint8)
    out, ln = C.c_void_p(), C.c_size_t()
    have_gpu = L.rv_device_count() > 0

    def prove(o):
        rc = L.rv_proof_new(o.ctypes.data_as(C.c_void_p), o.size, wit.ctypes.data_as(C.c_void_p), 2, None, 0, wc[0], wc[1], None, C.byref(out), C.byref(ln))
        assert rc == (0 if have_gpu else N.E_CUDA)
        if rc == 0:
            L.rv_free(out)

    prove(ops)
    assert stats() == (h0, m0 + 1, 1)
    prove(ops.copy())  # same content at another address
    assert stats() == (h0 + 1, m0 + 1, 1)
    ops2 = ops.copy()
    ops2["b"][50] = 0
    prove(ops2)
    assert stats() == (h0 + 1, m0 + 2, 2)
    blob = np.zeros(40, dtype=np.uint8)
    rc = L.rv_proof_verify(ops.ctypes.data_as(C.c_void_p), ops.size, wc[0], wc[1], blob.ctypes.data_as(C.c_void_p), blob.size)
    assert rc == N.E_FORMAT  # malformed proof bytes -- but the circuit was compiled WITH verifier tables and replaced its prove-only twin
    assert stats() == (h0 + 1, m0 + 3, 2)
    prove(ops)  # the upgraded entry serves proving too
    assert stats() == (h0 + 2, m0 + 3, 2)
    L.rv_circuit_cache_limit(1)
    prove(ops2)
    prove(ops)
    assert stats()[2] == 1
    L.rv_circuit_cache_limit(8)
    L.rv_circuit_cache_clear()
    assert stats()[2] == 0
    # op lists above 32 MB are hashed in slices on several threads: the key must still see every byte, wherever it sits
    big, wc = CI.flat_mul_circuit(3_000_000)  # 72 MB of ops: three slices
    big = np.ascontiguousarray(big)
    h1, m1, _ = stats()
    prove(big)
    prove(big.copy())
    assert stats() == (h1 + 1, m1 + 1, 1)
    for k, where in enumerate((7, 1_500_000, big.size - 1)):  # first slice, middle slice, the very last record
        other = big.copy()
        other["imm"][where] ^= 1 << 40  # (bits the compiler ignores for a GF(2) Mul: only the key can tell the lists apart)
        prove(other)
        assert stats() == (h1 + 1, m1 + 2 + k, 2 + k)
    L.rv_circuit_cache_clear()
    ops, wc = CI.flat_mul_circuit(100)
    ops = np.ascontiguousarray(ops)
    st = __import__("reverie_b200").Circuit(ops, wc, prove_only=True).stats()
    assert st["has_verify"] == 0 and st["compile_ns"] > 0


def test_bench_digests_are_the_oracles():
    """tests/golden/bench_digests.json (what bench.py's `parity_checked` compares the GPU proofs with) for the workloads the
    oracle proves in a second or two; the 10^7 .. 3 x 10^8-gate entries come from the same script (make_bench_digests.py)."""
    import hashlib
    import json

    import bench
    import orc

    with open(os.path.join(ROOT, "tests", "golden", "bench_digests.json")) as f:
        doc = json.load(f)
    seeds = bench.default_seeds()
    for name in ("sha256", "flat1000000", "layered1000000", "z64mul100000"):
        ops, wit, wz, wc, _ = bench.make_workload(name)
        rc, proof = orc.prove(ops, wit, wz, wc, seeds)
        assert rc == 0
        assert hashlib.sha256(proof).hexdigest() == doc["digests"][name], name
        assert len(proof) == doc["proof_bytes"][name]
    for name in ("flat100000000", "layered100000000", "z64mul1000000", "flat300000000"):
        assert len(doc["digests"][name]) == 64


def test_multi_gpu_and_streaming_entry_points_check_their_arguments_without_a_gpu():
    """rv_group_* / rv_prove_streaming: shapes and unsupported circuits are refused before any device work; with no device the
    answer is RV_E_CUDA (there is no CPU fallback)."""
    import ctypes as C

    from reverie_b200 import _native as N
    from reverie_b200 import circuits as CC
    from reverie_b200.proof import Circuit, _ptr

    L = N.lib()
    ops, wc = CC.flat_mul_circuit(100)
    circ = Circuit(ops, wc, prove_only=True)
    h = C.c_void_p()
    devs = (C.c_int * 3)(0, 0, 0)
    assert L.rv_group_create_local(circ.handle, devs, 3, 1, 1, C.byref(h)) == N.E_ARG  # 3 does not divide the 32 packed instances
    assert L.rv_group_create_rank(circ.handle, 2, 2, 1, 1, C.byref(h)) == N.E_ARG       # rank out of range
    assert L.rv_group_create_rank(circ.handle, 0, 2, 0, 1, C.byref(h)) == N.E_ARG       # no sessions
    out, n = C.c_void_p(), C.c_size_t()
    wit = np.array([1, 1], dtype=np.uint8)
    zops, zwc = CC.flat_mul_circuit(4, domain=CC.Z64)
    assert L.rv_prove_streaming(_ptr(zops), zops.size, zwc[0], zwc[1], None, 0, None, 0, None, 64, C.byref(out), C.byref(n)) == N.E_UNSUPPORTED
    bad = ops.copy()
    bad["a"][5] = 10 ** 6
    if L.rv_device_count() == 0:
        assert L.rv_prove_streaming(_ptr(ops), ops.size, wc[0], wc[1], _ptr(wit), 2, None, 0, None, 64, C.byref(out), C.byref(n)) == N.E_CUDA
        assert L.rv_group_create_rank(circ.handle, 0, 2, 1, 1, C.byref(h)) == N.E_CUDA
    else:
        assert L.rv_prove_streaming(_ptr(bad), bad.size, wc[0], wc[1], _ptr(wit), 2, None, 0, None, 64, C.byref(out), C.byref(n)) == N.E_ARG


def test_streaming_planner_carries_every_live_wire():
    """rv_prove_streaming's host planner (segmentation, liveness, recycled cell-file slots), checked by the library's symbolic
    simulation of the cell file: random circuits with heavy cell reuse and windows from 64 ops up, SSA circuits (layered), SHA-256;
    the slot count must stay far below the cell count when wires die young."""
    import ctypes as C

    from reverie_b200 import _native as N
    from reverie_b200 import circuits as CC
    from reverie_b200.proof import _ptr

    L = N.lib()

    def plan(ops, cells, window):
        out = (C.c_uint64 * 6)()
        ops = np.ascontiguousarray(ops, dtype=CC.OP_DTYPE)
        rc = L.rv_stream_plan_check(_ptr(ops), ops.size, cells, window, out)
        assert rc == 0, (rc, L.rv_last_error())
        return dict(zip(("segments", "slots", "max_imports", "max_exports", "imports", "exports"), [int(x) for x in out]))

    rng = np.random.default_rng(11)
    for n_cells in (8, 40, 300):
        recs = [(CC.GF2, CC.INPUT, 0, i, 0, 0, 0) for i in range(min(n_cells, 8))]
        for _ in range(4000):
            kind = int(rng.choice([CC.MUL, CC.ADD, CC.SUB, CC.ADDC, CC.MULC, CC.CONST, CC.ASSERT_ZERO], p=[0.3, 0.3, 0.1, 0.1, 0.05, 0.05, 0.1]))
            d, a, b = (int(rng.integers(0, n_cells)) for _ in range(3))
            recs.append((CC.GF2, kind, 0, d, a, b, int(rng.integers(0, 2))))
        ops = np.array(recs, dtype=CC.OP_DTYPE)
        for window in (64, 65, 100, 1000, 10 ** 6):
            st = plan(ops, n_cells, window)
            assert st["segments"] == max(1, -(-ops.size // max(window, 64))) and st["slots"] <= n_cells
            if window >= ops.size:
                assert st["imports"] == st["exports"] == st["slots"] == 0
    ops, nw = CC.layered_and_circuit(2048, 40000)  # SSA wires: every cell is written once, read within the next two layers
    st = plan(ops, nw, 4096)
    assert st["slots"] < 3 * 4096 < nw and st["imports"] > 0
    ops, wit, wc = CC.sha256_abc_case()
    st = plan(ops, wc[1], 5000)
    assert st["segments"] == -(-ops.size // 5000) and 0 < st["slots"] < wc[1]
    bad = ops.copy()
    bad["a"][int(np.flatnonzero(ops["opcode"] == CC.MUL)[0])] = wc[1] + 5  # an operand outside the declared wire count
    out = (C.c_uint64 * 6)()
    assert L.rv_stream_plan_check(_ptr(bad), bad.size, wc[1], 5000, out) == N.E_ARG


@pytest.mark.parametrize("seed", range(3))
def test_streaming_segments_replayed_on_the_cpu(seed, default_seeds):
    """The product's streaming planner and the compiler's segment mode (imports as leaves of both planes, exported rows / values,
    slot recycling), replayed on the CPU with the kernels' own item functions: the proof must be the oracle's for windows far
    below the circuit size.  (The chunk carry and the segment-wise packing of the openings are kernel-level: GPU tests.)"""
    import orc
    from tests import hostsim
    from tests.test_gpu_parity import _random_circuit

    rng = np.random.default_rng(200 + seed)
    ops, wit, wc = _random_circuit(rng, 12, 600, n_cells=20 + 25 * seed)
    rc, want = orc.prove(ops, wit, [], wc, default_seeds)
    assert rc == 0
    for window in (64, 150, 10 ** 6):
        rc, got = hostsim.prove_streaming(ops, wit, wc, default_seeds, window)
        assert rc == 0 and got == want, (window, rc)


def test_large_circuit_compile_is_the_same_with_and_without_page_helpers(monkeypatch):
    """Circuits of >= 2^20 ops may be compiled with helper threads that populate the tables' pages ahead of the op walk
    (rv_compile.cpp, Prefault).  The helpers must not change a byte of the program; neither may the choice compile() makes
    between them (flat: on, wide layers: off)."""
    from tests import hostsim
    from reverie_b200 import circuits as Cc

    n = (1 << 20) + 12345
    flat, wc = Cc.flat_mul_circuit(n)
    lay, nw = Cc.layered_and_circuit(1 << 18, n)
    nar, nw2 = Cc.layered_and_circuit(1 << 16, n)  # wide enough for the level-sorted value plane, narrow enough for the helpers
    for ops, counts, flags in ((flat, wc, 0), (lay, (0, nw), 0), (nar, (0, nw2), 1)):
        got = {}
        for mode in ("0", "1", None):
            if mode is None:
                monkeypatch.delenv("RV_PREFAULT", raising=False)
            else:
                monkeypatch.setenv("RV_PREFAULT", mode)
            got[mode] = hostsim.program_digest(ops, counts, flags)
        assert got["0"] == got["1"] == got[None]


def test_threaded_streaming_planner_equals_the_serial_one():
    """plan_stream (threads: liveness by atomic max / min over op ranges, segments renumbered side by side, slots in sequence)
    must reproduce plan_stream_serial field by field -- segment ops, imports, exports, slots, stream offsets -- and fail with the
    same code and message.  Circuits with heavy wire reuse (slots recycled), SSA layers, SHA-256 repeated; windows that do and
    do not divide the op count; 2 ... 8 threads."""
    from tests import hostsim
    from reverie_b200 import circuits as CC

    rng = np.random.default_rng(5)

    def random_ops(n, n_cells):
        ops = np.zeros(n, dtype=CC.OP_DTYPE)
        ops["domain"] = CC.GF2
        ops["opcode"] = rng.choice([CC.MUL, CC.ADD, CC.SUB, CC.ADDC, CC.MULC, CC.CONST, CC.ASSERT_ZERO, CC.INPUT], size=n, p=[0.3, 0.3, 0.1, 0.1, 0.05, 0.05, 0.05, 0.05])
        for f in ("dst", "a", "b"):
            ops[f] = rng.integers(0, n_cells, size=n, dtype=np.uint32)
        ops["imm"] = rng.integers(0, 2, size=n)
        return ops

    n = 70000 + 1234
    for n_cells, window, threads in ((8, 1000, 2), (300, 4097, 3), (300, 65536, 8), (50000, 5000, 4), (50000, 9999, 8)):
        assert hostsim.plan_compare(random_ops(n, n_cells), n_cells, window, threads) == ""
    lay, nw = CC.layered_and_circuit(4096, 90000)
    assert hostsim.plan_compare(lay, nw, 8192, 4) == "" and hostsim.plan_compare(lay, nw, 10 ** 6, 4) == ""
    flat, wc = CC.flat_mul_circuit(100000)
    assert hostsim.plan_compare(flat, wc[1], 7000, 8) == ""
    # refusals: the same code and the same text (the first offending op), wherever the op sits
    ops = random_ops(n, 300)
    for where in (5, n // 2, n - 3):
        bad = ops.copy()
        bad["opcode"][where] = CC.MUL
        bad["a"][where] = 300
        assert hostsim.plan_compare(bad, 300, 4097, 4) == ""
        bad = ops.copy()
        bad["opcode"][where] = CC.RANDOM
        assert hostsim.plan_compare(bad, 300, 4097, 4) == ""
        bad = ops.copy()
        bad["domain"][where] = CC.Z64
        assert hostsim.plan_compare(bad, 300, 4097, 4) == ""
    hint = ops.copy()  # a SizeHint that grows the wire file mid-way: wires above the old size are legal only after it
    hint["domain"][n // 2] = CC.HINT
    hint["b"][n // 2] = 400
    hint["dst"][n // 2 + 10 :] += rng.integers(0, 2, size=n - n // 2 - 10, dtype=np.uint32) * 100
    assert hostsim.plan_compare(hint, 300, 4097, 4) == ""


@pytest.mark.parametrize("width,layers,seed", [(700, 40, 0), (1500, 24, 1), (300, 90, 2)])
def test_step_balancing_keeps_the_planes_right(width, layers, seed, default_seeds):
    """The compiler pads every level of the mask VM (512 slots) and of the value plane (128 slots) to whole steps and balances the
    padding: LOADs may be issued earlier than two levels ahead, mapped LUT nodes with slack may sit later.  Circuits whose levels
    spill over step boundaries in both planes -- layers of random XORs / ANDs / unary ops over the previous layers, asserts
    on wires that are zero by construction -- replayed from the padded device streams (LOADs landing at once and as late as the
    group wait allows) and proved by the kernel bodies on the CPU against the oracle."""
    rng = np.random.default_rng(seed)
    n_in = width
    recs = np.zeros(n_in + width * layers, dtype=CI.OP_DTYPE)
    recs["domain"] = CI.GF2
    recs["opcode"][:n_in] = CI.INPUT
    recs["dst"][:n_in] = np.arange(n_in)
    n = n_in
    for l in range(layers):
        sl = slice(n, n + width)
        lo = max(0, n - 3 * width)
        recs["opcode"][sl] = rng.choice([CI.ADD, CI.MUL, CI.ADDC, CI.SUB], size=width, p=[0.55, 0.25, 0.1, 0.1])
        recs["dst"][sl] = np.arange(n, n + width)
        recs["a"][sl] = rng.integers(lo, n, size=width)
        recs["b"][sl] = rng.integers(lo, n, size=width)
        recs["imm"][sl] = rng.integers(0, 2, size=width)
        n += width
    # x ^ x = 0 on a few late wires: AssertZero must hold, and the asserted masks are linear rows the VM has to export
    extra = []
    for w in rng.integers(n - width, n, size=8):
        extra.append((CI.GF2, CI.ADD, 0, n, int(w), int(w), 0))
        extra.append((CI.GF2, CI.ASSERT_ZERO, 0, 0, n, 0, 0))
        n += 1
    ops = np.concatenate([recs, np.array(extra, dtype=CI.OP_DTYPE)])
    wit = rng.integers(0, 2, size=n_in).astype(np.uint8)
    wc = (0, n)
    st = _check_steps(ops, wc)
    assert st[0] > 0 and st[1] > 0  # both step streams exist (mask VM, LUT value plane)
    rc, want, hashes = orc.prove(ops, wit, [], wc, default_seeds, want_hashes=True)
    rc2, got, h2 = hostsim.prove(ops, wit, wc, default_seeds)
    assert rc == 0 and rc2 == 0 and h2 == hashes and got == want
    vrc, okay = hostsim.verify(ops, wc, got)[:2] if hasattr(hostsim, "verify") else (1, 1)
    assert vrc == 1 and okay
