"""ctypes loader for tests/hostsim/hostsim.cpp (CPU replay of the kernel bodies; test infrastructure only)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_LIB = os.path.join(_HERE, "libhostsim.so")
_SRC = [os.path.join(_HERE, "hostsim.cpp"), os.path.join(_ROOT, "reverie_b200", "csrc", "rv_compile.cpp")]
_lib = None


def lib():
    global _lib
    if _lib is None:
        deps = _SRC + [os.path.join(_ROOT, "reverie_b200", "csrc", f) for f in ("rv_planes.cuh", "rv_zplanes.cuh", "rv_aes_bs.cuh", "rv_blake3.cuh", "rv_compile.h", "rv_bincode.h", "rv_stream_plan.h")]
        if not os.path.exists(_LIB) or any(os.path.getmtime(d) > os.path.getmtime(_LIB) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", _LIB] + _SRC)
        L = C.CDLL(_LIB)
        sz = C.c_size_t
        L.hs_prove.argtypes = [C.c_void_p, sz, sz, sz, C.c_void_p, sz, C.c_void_p, sz, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(sz), C.c_void_p]
        L.hs_prove.restype = C.c_int
        L.hs_prove_streaming.argtypes = [C.c_void_p, sz, sz, C.c_void_p, sz, C.c_void_p, sz, C.POINTER(C.c_void_p), C.POINTER(sz)]
        L.hs_prove_streaming.restype = C.c_int
        L.hs_free.argtypes = [C.c_void_p]
        L.hs_plan_compare.argtypes = [C.c_void_p, sz, sz, sz, C.c_uint]
        L.hs_plan_compare.restype = C.c_int
        L.hs_program_digest.argtypes = [C.c_void_p, sz, sz, sz, C.c_uint32, C.POINTER(C.c_uint64)]
        L.hs_program_digest.restype = C.c_int
        L.hs_blake3.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.hs_aes128_encrypt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.hs_gf2_masks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.hs_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def prove(ops, wit, wire_counts, seeds: bytes, wit_z64=()):
    ops = np.ascontiguousarray(ops)
    w = np.ascontiguousarray(np.asarray(wit, dtype=np.uint8))
    wz = np.ascontiguousarray(np.asarray(wit_z64, dtype=np.uint64))
    sd = np.frombuffer(seeds, dtype=np.uint8)
    out, n = C.c_void_p(), C.c_size_t()
    hashes = np.zeros(256 * 32, dtype=np.uint8)
    rc = lib().hs_prove(_p(ops), ops.size, wire_counts[0], wire_counts[1], _p(w), w.size, _p(wz), wz.size, _p(sd), C.byref(out), C.byref(n), _p(hashes))
    if rc != 0:
        return rc, None, None
    proof = C.string_at(out, n.value)
    lib().hs_free(out)
    return 0, proof, hashes.tobytes()


def prove_streaming(ops, wit, wire_counts, seeds: bytes, window_ops: int):
    """CPU replay of rv_prove_streaming's planner + segment compilation + carried cell file (GF(2) only)."""
    ops = np.ascontiguousarray(ops)
    w = np.ascontiguousarray(np.asarray(wit, dtype=np.uint8))
    sd = np.frombuffer(seeds, dtype=np.uint8)
    out, n = C.c_void_p(), C.c_size_t()
    rc = lib().hs_prove_streaming(_p(ops), ops.size, wire_counts[1], _p(w), w.size, _p(sd), window_ops, C.byref(out), C.byref(n))
    if rc != 0:
        return rc, None
    proof = C.string_at(out, n.value)
    lib().hs_free(out)
    return 0, proof


def blake3(data: bytes) -> bytes:
    d = np.frombuffer(data, dtype=np.uint8)
    out = np.zeros(32, dtype=np.uint8)
    lib().hs_blake3(_p(d), d.size, _p(out))
    return out.tobytes()


def aes128_encrypt(key: bytes, block: bytes) -> bytes:
    out = np.zeros(16, dtype=np.uint8)
    lib().hs_aes128_encrypt(_p(np.frombuffer(key, dtype=np.uint8)), _p(np.frombuffer(block, dtype=np.uint8)), _p(out))
    return out.tobytes()


def tt_aes128_encrypt(key: bytes, block: bytes) -> bytes:
    """AES-128 through the T-table rounds the GPU mask generators run (csrc/rv_aes_bs.cuh: tt_aes128_encrypt)."""
    out = np.zeros(16, dtype=np.uint8)
    lib().hs_tt_aes128_encrypt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib().hs_tt_aes128_encrypt(_p(np.frombuffer(key, dtype=np.uint8)), _p(np.frombuffer(block, dtype=np.uint8)), _p(out))
    return out.tobytes()


def gf2_masks(seeds8: bytes, omit, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.uint64)
    o = np.asarray(omit, dtype=np.uint8)
    lib().hs_gf2_masks(_p(np.frombuffer(seeds8, dtype=np.uint8)), _p(o), _p(out), n)
    return out


def verify(ops, wire_counts, proof: bytes):
    """-> (rc, okay, rep_hashes): rc 1 accept / 0 reject / <0 error, through the verifier kernels' bodies."""
    L = lib()
    L.hs_verify.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.c_void_p]
    L.hs_verify.restype = C.c_int
    ops = np.ascontiguousarray(ops)
    pb = np.frombuffer(proof, dtype=np.uint8)
    okay = C.c_int(1)
    hashes = np.zeros(256 * 32, dtype=np.uint8)
    rc = L.hs_verify(_p(ops), ops.size, wire_counts[0], wire_counts[1], _p(pb), pb.size, C.byref(okay), _p(hashes))
    return rc, bool(okay.value), hashes.tobytes()


def program_digest(ops, wire_counts, flags: int = 0) -> int:
    """FNV-1a of every device-bound table of the compiled program."""
    ops = np.ascontiguousarray(ops)
    d = C.c_uint64()
    rc = lib().hs_program_digest(_p(ops), ops.size, wire_counts[0], wire_counts[1], flags, C.byref(d))
    if rc != 0:
        raise RuntimeError(f"compile failed ({rc}): {lib().hs_last_error().decode()}")
    return d.value


def plan_compare(ops, gf2_cells: int, window_ops: int, n_threads: int) -> str:
    """'' if the threaded streaming planner and the serial one agree in every field (or fail alike), else what differs."""
    ops = np.ascontiguousarray(ops)
    return "" if lib().hs_plan_compare(_p(ops), ops.size, gf2_cells, window_ops, n_threads) == 0 else lib().hs_last_error().decode()
