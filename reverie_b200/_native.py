"""ctypes binding of libreverie_b200.so (include/reverie_b200.h).  Importing this module never touches a GPU; the first
compute call does.  If the shared library is missing it is built in-tree with nvcc (reverie_b200/_build.py); if that
fails the import error is loud -- there is no Python or CPU fallback for any compute entry point."""
from __future__ import annotations

import ctypes as C
import os

from . import _build

RV_OK, E_WITNESS_INVALID, E_WITNESS_SHORT, E_FORMAT, E_ARG, E_CUDA, E_NOMEM, E_UNSUPPORTED, E_PEER = 0, -1, -2, -3, -4, -5, -6, -7, -8
PEER_HANDLE_BYTES = 256
TOTAL_REPS, ONLINE_REPS, PACKED_REPS, PLAYERS = 256, 40, 32, 8


class CircuitStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "n_ops", "n_and", "n_inputs", "n_assert", "n_masks", "n_linear", "value_depth", "linear_depth", "plain_value_depth",
        "plain_linear_depth", "n_luts", "n_lut_steps", "n_vm_steps", "vm_cells", "online_bytes", "pre_bytes", "algorithmic_bytes",
        "device_bytes", "z64_mul", "z64_inputs", "z64_assert", "z64_masks", "z64_linear", "z64_value_depth", "z64_linear_depth",
        "z64_online_bytes", "z64_pre_bytes", "compile_ns", "has_verify", "n_vals", "n_uvals", "n_vlut_steps")]


class KernelTime(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("ms", C.c_double), ("launches", C.c_uint64), ("algorithmic_bytes", C.c_uint64)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB if os.path.exists(_build.LIB) and not _build._stale() else _build.build()
    # one hardware work queue per stream (the default is 8, shared): the sessions of a batch run on their own streams, and the
    # challenge kernel of a linked session waits on the device for its peers -- work of another stream must not queue up behind it
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    L = C.CDLL(path)
    vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
    pp, psz = C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)
    sigs = {
        "rv_last_error": ([], C.c_char_p),
        "rv_version": ([], C.c_char_p),
        "rv_device_count": ([], i32),
        "rv_set_device": ([i32], i32),
        "rv_circuit_compile": ([vp, sz, sz, sz, pp], i32),
        "rv_circuit_compile_ex": ([vp, sz, sz, sz, C.c_uint, pp], i32),
        "rv_circuit_free": ([vp], None),
        "rv_circuit_cache_clear": ([], None),
        "rv_oneshot_streaming_min": ([sz], None),
        "rv_circuit_cache_limit": ([sz], None),
        "rv_circuit_cache_stats": ([C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), psz], None),
        "rv_proof_verify_ex": ([vp, sz, sz, sz, vp, sz, C.POINTER(i32)], i32),
        "rv_circuit_get_stats": ([vp, C.POINTER(CircuitStats)], i32),
        "rv_circuit_export": ([vp, i32, vp, psz], i32),
        "rv_prove": ([vp, vp, sz, vp, sz, vp, pp, psz], i32),
        "rv_verify": ([vp, vp, sz, C.POINTER(i32)], i32),
        "rv_prove_batch": ([vp, i32, C.POINTER(vp), psz, C.POINTER(vp), psz, C.POINTER(vp), pp, psz, C.POINTER(i32)], i32),
        "rv_session_slots": ([vp], i32),
        "rv_session_proof_stride": ([vp], sz),
        "rv_proof_new": ([vp, sz, vp, sz, vp, sz, sz, sz, vp, pp, psz], i32),
        "rv_proof_verify": ([vp, sz, sz, sz, vp, sz], i32),
        "rv_free": ([vp], None),
        "rv_session_create": ([vp, i32, i32, pp], i32),
        "rv_session_free": ([vp], None),
        "rv_session_create_multi": ([vp, i32, i32, i32, pp], i32),
        "rv_session_upload_slot": ([vp, i32, vp, sz, vp, sz, vp], i32),
        "rv_session_fetch_slot": ([vp, i32, vp, pp, psz], i32),
        "rv_session_upload": ([vp, vp, sz, vp, sz, vp], i32),
        "rv_session_commit": ([vp], i32),
        "rv_session_hashes": ([vp, vp], i32),
        "rv_session_open": ([vp, vp], i32),
        "rv_session_hashes_device": ([vp], vp),
        "rv_session_all_hashes_device": ([vp], vp),
        "rv_session_prove": ([vp], i32),
        "rv_session_fetch": ([vp, vp, pp, psz], i32),
        "rv_session_sync": ([vp], i32),
        "rv_session_status": ([vp], i32),
        "rv_session_proof_device": ([vp, pp, psz], i32),
        "rv_batch_create": ([C.POINTER(vp), i32, pp], i32),
        "rv_batch_free": ([vp], None),
        "rv_batch_commit": ([vp], i32),
        "rv_batch_open": ([vp], i32),
        "rv_batch_prove": ([vp], i32),
        "rv_batch_stream": ([vp], vp),
        "rv_proof_assemble": ([vp, C.POINTER(vp), psz, i32, pp, psz], i32),
        "rv_session_stream": ([vp], vp),
        "rv_session_timing": ([vp, i32], i32),
        "rv_session_kernel_times": ([vp, C.POINTER(KernelTime), i32, i32], i32),
        "rv_session_launch_count": ([vp], C.c_uint64),
        "rv_session_peer_handle": ([vp, vp], i32),
        "rv_session_peer_link": ([vp, i32, i32, vp], i32),
        "rv_session_peer_rank": ([vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)], i32),
        "rv_stream_plan_check": ([vp, sz, sz, sz, C.POINTER(C.c_uint64)], i32),
        "rv_prove_streaming": ([vp, sz, sz, sz, vp, sz, vp, sz, vp, sz, pp, psz], i32),
        "rv_circuit_clone": ([vp, i32, pp], i32),
        "rv_group_create_local": ([vp, C.POINTER(i32), i32, i32, i32, pp], i32),
        "rv_group_create_rank": ([vp, i32, i32, i32, i32, pp], i32),
        "rv_group_handles_bytes": ([vp], sz),
        "rv_group_handles": ([vp, vp], i32),
        "rv_group_link": ([vp, vp], i32),
        "rv_group_info": ([vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)], i32),
        "rv_group_session": ([vp, i32, i32], vp),
        "rv_group_step": ([vp], i32),
        "rv_group_prove_batch": ([vp, i32, C.POINTER(vp), psz, C.POINTER(vp), psz, C.POINTER(vp), pp, psz, C.POINTER(i32)], i32),
        "rv_group_prove": ([vp, vp, sz, vp, sz, vp, pp, psz], i32),
        "rv_group_verify_batch": ([vp, i32, C.POINTER(vp), psz, C.POINTER(i32), C.POINTER(i32)], i32),
        "rv_group_free": ([vp], None),
    }
    for name, (args, res) in sigs.items():
        fn = getattr(L, name)  # AttributeError here = the library does not export what include/reverie_b200.h declares
        fn.argtypes, fn.restype = args, res
    _lib = L
    return L


EXPORTED = (
    "rv_last_error rv_version rv_device_count rv_set_device rv_circuit_compile rv_circuit_compile_ex rv_circuit_cache_clear rv_oneshot_streaming_min rv_circuit_cache_limit rv_circuit_cache_stats rv_proof_verify_ex rv_circuit_free rv_circuit_get_stats "
    "rv_circuit_export rv_prove rv_prove_batch rv_session_slots rv_session_proof_stride rv_verify rv_proof_new rv_proof_verify rv_free rv_session_create rv_session_create_multi rv_session_upload_slot rv_session_fetch_slot rv_session_free "
    "rv_session_upload rv_session_commit rv_session_hashes rv_session_hashes_device rv_session_all_hashes_device rv_session_open rv_session_prove rv_session_fetch rv_session_sync rv_session_status rv_session_proof_device "
    "rv_proof_assemble rv_batch_create rv_batch_free rv_batch_commit rv_batch_open rv_batch_prove rv_batch_stream rv_session_stream rv_session_timing rv_session_kernel_times rv_session_launch_count "
    "rv_session_peer_handle rv_session_peer_link rv_session_peer_rank rv_prove_streaming rv_stream_plan_check rv_circuit_clone rv_group_create_local rv_group_create_rank "
    "rv_group_handles_bytes rv_group_handles rv_group_link rv_group_info rv_group_session rv_group_step rv_group_prove_batch rv_group_prove rv_group_verify_batch rv_group_free"
).split()


class ReverieError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[rv_status {code}] {msg}")
        self.code = code


class WitnessError(ReverieError):
    """The reference's panics 'witness is invalid!' / 'witness is too short' (src/transcript/prover.rs:190,223)."""


class FormatError(ReverieError):
    """Malformed proof bytes (the reference panics on these, e.g. src/algebra/gf2/share.rs:158-164)."""


def check(rc: int) -> int:
    if rc >= 0:
        return rc
    msg = (lib().rv_last_error() or b"").decode()
    if rc in (E_WITNESS_INVALID, E_WITNESS_SHORT):
        raise WitnessError(rc, msg)
    if rc == E_FORMAT:
        raise FormatError(rc, msg)
    raise ReverieError(rc, msg)
