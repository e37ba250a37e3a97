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
