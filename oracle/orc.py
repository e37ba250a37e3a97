"""ctypes loader for the C oracle (oracle/c/liborc.so).  TEST INFRASTRUCTURE ONLY -- see oracle/c/reverie_oracle.h.

Also converts between the packed 24-byte op records (numpy, shared layout with the product's rv_op) and the tuple
form the Python oracle (oracle/reverie_oracle.py) interprets."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "c", "liborc.so")

OP_DTYPE = np.dtype(
    [("domain", "u1"), ("opcode", "u1"), ("pad", "<u2"), ("dst", "<u4"), ("a", "<u4"), ("b", "<u4"), ("imm", "<u8")]
)
assert OP_DTYPE.itemsize == 24
OPCODES = ("Input", "Random", "Add", "AddConst", "Sub", "SubConst", "Mul", "MulConst", "AssertZero", "Const")
E_WITNESS_INVALID, E_WITNESS_SHORT, E_FORMAT, E_ARG = -1, -2, -3, -4


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        u8p, u64p, sz = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.c_size_t
        L.orc_prove.argtypes = [C.c_void_p, sz, C.c_void_p, sz, C.c_void_p, sz, sz, sz, C.c_void_p, C.c_int,
                                C.POINTER(C.c_void_p), C.POINTER(sz), C.c_void_p]
        L.orc_prove.restype = C.c_int
        L.orc_prove_lowmem.argtypes = [C.c_void_p, sz, C.c_void_p, sz, C.c_void_p, sz, sz, sz, C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(sz)]
        L.orc_prove_lowmem.restype = C.c_int
        L.orc_verify.argtypes = [C.c_void_p, sz, sz, sz, C.c_void_p, sz, C.c_int, C.POINTER(C.c_int), C.c_void_p]
        L.orc_verify.restype = C.c_int
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_aes128_ctr.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, sz]
        L.orc_blake3.argtypes = [C.c_void_p, sz, C.c_void_p, sz]
        L.orc_gf2_masks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, sz]
        L.orc_challenge.argtypes = [C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def prove(ops: np.ndarray, wit_gf2, wit_z64, wire_counts, seeds: bytes, n_threads: int = 0, want_hashes=False):
    """-> (rc, proof_bytes|None[, rep_hashes])   wire_counts = (z64_cells, gf2_cells) like the reference."""
    ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
    wg = np.ascontiguousarray(np.asarray(wit_gf2, dtype=np.uint8))
    wz = np.ascontiguousarray(np.asarray(wit_z64, dtype=np.uint64))
    sd = np.frombuffer(bytes(seeds), dtype=np.uint8)
    assert sd.size == 256 * 16
    out, n = C.c_void_p(), C.c_size_t()
    hashes = np.zeros(256 * 32, dtype=np.uint8)
    rc = lib().orc_prove(_ptr(ops), ops.size, _ptr(wg), wg.size, _ptr(wz), wz.size, wire_counts[0], wire_counts[1],
                         _ptr(sd), n_threads, C.byref(out), C.byref(n), _ptr(hashes))
    proof = None
    if rc == 0:
        proof = C.string_at(out, n.value)
        lib().orc_free(out)
    return (rc, proof, hashes.tobytes()) if want_hashes else (rc, proof)


def prove_digest_lowmem(ops: np.ndarray, wit_gf2, wit_z64, wire_counts, seeds: bytes, n_threads: int = 0):
    """orc_prove_lowmem: the same proof in two passes (at most n_threads instances' transcripts alive at a time), for circuits
    whose transcripts do not fit in memory all at once.  -> (rc, sha256 hex digest of the proof bytes, proof length)."""
    import hashlib

    ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
    wg = np.ascontiguousarray(np.asarray(wit_gf2, dtype=np.uint8))
    wz = np.ascontiguousarray(np.asarray(wit_z64, dtype=np.uint64))
    sd = np.frombuffer(bytes(seeds), dtype=np.uint8)
    out, n = C.c_void_p(), C.c_size_t()
    rc = lib().orc_prove_lowmem(_ptr(ops), ops.size, _ptr(wg), wg.size, _ptr(wz), wz.size, wire_counts[0], wire_counts[1], _ptr(sd), n_threads,
                                C.byref(out), C.byref(n))
    if rc != 0:
        return rc, None, 0
    view = (C.c_uint8 * n.value).from_address(out.value)
    dg = hashlib.sha256(memoryview(view)).hexdigest()
    lib().orc_free(out)
    return rc, dg, n.value


def verify(ops: np.ndarray, wire_counts, proof: bytes, n_threads: int = 0, want_hashes=False):
    """-> (rc, okay[, rep_hashes])   rc: 1 accept, 0 reject, <0 error."""
    ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
    pb = np.frombuffer(proof, dtype=np.uint8)
    okay = C.c_int(1)
    hashes = np.zeros(256 * 32, dtype=np.uint8)
    rc = lib().orc_verify(_ptr(ops), ops.size, wire_counts[0], wire_counts[1], _ptr(pb), pb.size, n_threads,
                          C.byref(okay), _ptr(hashes))
    return (rc, bool(okay.value), hashes.tobytes()) if want_hashes else (rc, bool(okay.value))


def aes128_ctr(key: bytes, first_block: int, n_blocks: int) -> bytes:
    out = np.zeros(16 * n_blocks, dtype=np.uint8)
    k = np.frombuffer(key, dtype=np.uint8)
    lib().orc_aes128_ctr(_ptr(k), first_block, _ptr(out), n_blocks)
    return out.tobytes()


def blake3(data: bytes, out_len: int = 32) -> bytes:
    d = np.frombuffer(data, dtype=np.uint8)
    out = np.zeros(out_len, dtype=np.uint8)
    lib().orc_blake3(_ptr(d), d.size, _ptr(out), out_len)
    return out.tobytes()


def gf2_masks(seeds8: bytes, omit, n: int) -> np.ndarray:
    s = np.frombuffer(seeds8, dtype=np.uint8)
    o = np.asarray(omit, dtype=np.uint8)
    out = np.zeros(n, dtype=np.uint64)
    lib().orc_gf2_masks(_ptr(s), _ptr(o), _ptr(out), n)
    return out


def challenge(comm: bytes) -> np.ndarray:
    c = np.frombuffer(comm, dtype=np.uint8)
    out = np.zeros(256, dtype=np.uint8)
    lib().orc_challenge(_ptr(c), _ptr(out))
    return out


def ops_to_tuples(ops: np.ndarray):
    """packed records -> the tuple form of oracle/reverie_oracle.py"""
    out = []
    for o in ops:
        d, c = int(o["domain"]), int(o["opcode"])
        dst, a, b, imm = int(o["dst"]), int(o["a"]), int(o["b"]), int(o["imm"])
        if d == 2:
            out.append(("b2a", dst, a))
            continue
        if d == 3:
            out.append(("hint", a, b))
            continue
        dom = "gf2" if d == 0 else "z64"
        v = bool(imm & 1) if d == 0 else imm
        name = OPCODES[c]
        if name in ("Input", "Random"):
            out.append((dom, name, dst))
        elif name in ("Add", "Sub", "Mul"):
            out.append((dom, name, dst, a, b))
        elif name in ("AddConst", "SubConst", "MulConst"):
            out.append((dom, name, dst, a, v))
        elif name == "AssertZero":
            out.append((dom, name, a))
        else:
            out.append((dom, name, dst, v))
    return out


def tuples_to_ops(circ) -> np.ndarray:
    ops = np.zeros(len(circ), dtype=OP_DTYPE)
    for i, t in enumerate(circ):
        if t[0] == "b2a":
            ops[i] = (2, 0, 0, t[1], t[2], 0, 0)
        elif t[0] == "hint":
            ops[i] = (3, 0, 0, 0, t[1], t[2], 0)
        else:
            d = 0 if t[0] == "gf2" else 1
            name, args = t[1], t[2:]
            c = OPCODES.index(name)
            if name in ("Input", "Random"):
                ops[i] = (d, c, 0, args[0], 0, 0, 0)
            elif name in ("Add", "Sub", "Mul"):
                ops[i] = (d, c, 0, args[0], args[1], args[2], 0)
            elif name in ("AddConst", "SubConst", "MulConst"):
                ops[i] = (d, c, 0, args[0], args[1], 0, int(args[2]) & ((1 << 64) - 1))
            elif name == "AssertZero":
                ops[i] = (d, c, 0, 0, args[0], 0, 0)
            else:
                ops[i] = (d, c, 0, args[0], 0, 0, int(args[1]) & ((1 << 64) - 1))
    return ops
