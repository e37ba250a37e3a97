"""Random Z64 / mixed-domain circuits for the parity tests (test helper, shared by the CPU and GPU suites)."""
import numpy as np

from reverie_b200 import circuits as CI

M64 = (1 << 64) - 1


def random_z_circuit(rng, n_in, n_ops, n_cells=24, with_gf2=False, n_asserts=3):
    """-> (ops, wit_gf2, wit_z64, wire_counts).  Cells are overwritten (non-SSA) like the reference's bench circuit; every
    AssertZero checks `wire - its plaintext value`, so the witness is valid by construction."""
    recs, vals, wit = [], [0] * n_cells, []
    gvals, gwit = [0] * 16, []

    def emit(d, op, dst=0, a=0, b=0, imm=0):
        recs.append((d, op, 0, dst, a, b, imm & M64))

    for k in range(n_in):
        w = int(rng.integers(0, 1 << 63)) * 2 + int(rng.integers(0, 2))
        wit.append(w)
        vals[k % n_cells] = w
        emit(CI.Z64, CI.INPUT, k % n_cells)
    if with_gf2:
        for k in range(4):
            b = int(rng.integers(0, 2))
            gwit.append(b)
            gvals[k] = b
            emit(CI.GF2, CI.INPUT, k)
    for _ in range(n_ops):
        if with_gf2 and rng.random() < 0.3:
            d, a, b = (int(x) for x in rng.integers(0, 16, 3))
            if rng.random() < 0.5:
                emit(CI.GF2, CI.MUL, d, a, b)
                gvals[d] = gvals[a] & gvals[b]
            else:
                emit(CI.GF2, CI.ADD, d, a, b)
                gvals[d] = gvals[a] ^ gvals[b]
            continue
        k = int(rng.integers(0, 9))
        d, a, b = (int(x) for x in rng.integers(0, n_cells, 3))
        c = int(rng.integers(0, 1 << 63)) * 2 + 1 if rng.random() < 0.7 else int(rng.integers(0, 4))
        if k == 0:
            emit(CI.Z64, CI.ADD, d, a, b)
            vals[d] = (vals[a] + vals[b]) & M64
        elif k == 1:
            emit(CI.Z64, CI.SUB, d, a, b)
            vals[d] = (vals[a] - vals[b]) & M64
        elif k in (2, 3, 4):
            emit(CI.Z64, CI.MUL, d, a, b)
            vals[d] = (vals[a] * vals[b]) & M64
        elif k == 5:
            emit(CI.Z64, CI.ADDC, d, a, 0, c)
            vals[d] = (vals[a] + c) & M64
        elif k == 6:
            emit(CI.Z64, CI.SUBC, d, a, 0, c)
            vals[d] = (vals[a] - c) & M64
        elif k == 7:
            emit(CI.Z64, CI.MULC, d, a, 0, c)
            vals[d] = (vals[a] * c) & M64
        else:
            emit(CI.Z64, CI.CONST, d, 0, 0, c)
            vals[d] = c
    for _ in range(n_asserts):
        a, d = int(rng.integers(0, n_cells)), n_cells - 1
        emit(CI.Z64, CI.SUBC, d, a, 0, vals[a])
        emit(CI.Z64, CI.ASSERT_ZERO, 0, d)
        vals[d] = 0
    if with_gf2:
        for w in range(16):
            if gvals[w] == 0 and rng.random() < 0.3:
                emit(CI.GF2, CI.ASSERT_ZERO, 0, w)
    ops = np.array(recs, dtype=CI.OP_DTYPE) if recs else np.zeros(0, dtype=CI.OP_DTYPE)
    return ops, np.array(gwit, dtype=np.uint8), np.array(wit, dtype=np.uint64), (n_cells, 16)


Z64_WITNESS = np.array([0x0123456789ABCDEF, 0xFEDCBA9876543210], dtype=np.uint64)


def random_mixed_circuit(rng, n_b2a=2, n_random=3, n_gf2_ops=60, n_z_ops=20):
    """GF(2) circuit with `Random` wires and B2A conversions feeding a Z64 tail (src/interpreter/combine.rs:132-219).
    -> (ops, wit_gf2, wit_z64, wire_counts); witness valid by construction."""
    recs = []
    n_g = 64 * n_b2a + 40
    gv = [None] * n_g  # plaintext bit, or None when the wire depends on a Random (per-repetition value)
    gwit = []

    def emit(d, op, dst=0, a=0, b=0, imm=0):
        recs.append((d, op, 0, dst, a, b, imm & M64))

    n_in = 64 * n_b2a + 8
    for k in range(n_in):
        bit = int(rng.integers(0, 2))
        gwit.append(bit)
        gv[k] = bit
        emit(CI.GF2, CI.INPUT, k)
    scratch = list(range(n_in, n_g))
    for w in scratch:
        gv[w] = 0
    rnd = []
    for k in range(n_random):
        w = scratch[k]
        emit(CI.GF2, CI.RANDOM, w)
        gv[w] = None
        rnd.append(w)
    for _ in range(n_gf2_ops):  # gates over the scratch wires and inputs, some touching the Random wires
        d = int(rng.choice(scratch[n_random:]))
        a, b = int(rng.integers(0, n_g)), int(rng.integers(0, n_g))
        if rng.random() < 0.5:
            emit(CI.GF2, CI.MUL, d, a, b)
            gv[d] = None if (gv[a] is None or gv[b] is None) and not (gv[a] == 0 or gv[b] == 0) else (gv[a] & gv[b] if gv[a] is not None and gv[b] is not None else 0)
        else:
            emit(CI.GF2, CI.ADD, d, a, b)
            gv[d] = None if gv[a] is None or gv[b] is None else gv[a] ^ gv[b]
    for w in rnd:  # r ^ r = 0 in every repetition
        d = scratch[-1]
        emit(CI.GF2, CI.ADD, d, w, w)
        emit(CI.GF2, CI.ASSERT_ZERO, 0, d)
        gv[d] = 0
    for w in scratch[n_random:-1]:
        if gv[w] == 0 and rng.random() < 0.3:
            emit(CI.GF2, CI.ASSERT_ZERO, 0, w)
    # B2A of input words, then a Z64 tail
    n_z = n_b2a + 6
    zv = [0] * n_z
    for k in range(n_b2a):
        recs.append((CI.B2A, 0, 0, k, 64 * k, 0, 0))
        zv[k] = sum(gwit[64 * k + i] << i for i in range(64))
    zwit = []
    w = int(rng.integers(0, 1 << 63))
    zwit.append(w)
    zv[n_b2a] = w
    emit(CI.Z64, CI.INPUT, n_b2a)
    for _ in range(n_z_ops):
        d = int(rng.integers(n_b2a + 1, n_z))
        a, b = int(rng.integers(0, n_z)), int(rng.integers(0, n_z))
        k = int(rng.integers(0, 4))
        if k == 0:
            emit(CI.Z64, CI.ADD, d, a, b)
            zv[d] = (zv[a] + zv[b]) & M64
        elif k == 1:
            emit(CI.Z64, CI.SUB, d, a, b)
            zv[d] = (zv[a] - zv[b]) & M64
        elif k == 2:
            emit(CI.Z64, CI.MUL, d, a, b)
            zv[d] = (zv[a] * zv[b]) & M64
        else:
            c = int(rng.integers(0, 1 << 62)) * 2 + 1
            emit(CI.Z64, CI.MULC, d, a, 0, c)
            zv[d] = (zv[a] * c) & M64
    for a in range(n_z):
        d = n_z - 1
        emit(CI.Z64, CI.SUBC, d, a, 0, zv[a])
        emit(CI.Z64, CI.ASSERT_ZERO, 0, d)
        zv[d] = 0
    return np.array(recs, dtype=CI.OP_DTYPE), np.array(gwit, dtype=np.uint8), np.array(zwit, dtype=np.uint64), (n_z, n_g)
