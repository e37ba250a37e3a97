"""Host-side mirror of the reference's public API for the KKW hot path.

    reverie (Rust)                                              here
    Proof::new(circuit, wit_gf2, wit_z64, wire_counts)     ->   Proof.new(circuit, wit_gf2, wit_z64, wire_counts[, seeds])
    proof.verify(circuit, wire_counts) -> bool             ->   proof.verify(circuit, wire_counts) -> bool
    bincode::serialize(&proof) / deserialize               ->   proof.serialize() / Proof.deserialize(bytes)

(src/proof/mod.rs:119-124,224; src/main.rs:74,84,103,108).  `circuit` is either a numpy array of packed op records
(reverie_b200.circuits.OP_DTYPE) or a `Circuit`, the compiled, device-resident form that plays the role of the
reference's `Arc<Vec<CombineOperation>>`: build it once, prove many times.  Everything runs on the GPU through the C ABI
(include/reverie_b200.h); nothing here computes proof data on the host.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _native as N
from .circuits import OP_DTYPE


def _ptr(a: Optional[np.ndarray]):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def _seeds_arr(seeds) -> Optional[np.ndarray]:
    if seeds is None:
        return None
    b = b"".join(seeds) if not isinstance(seeds, (bytes, bytearray, np.ndarray)) else bytes(seeds)
    if len(b) != N.TOTAL_REPS * 16:
        raise ValueError("seeds must be 256 x 16 bytes")
    return np.frombuffer(b, dtype=np.uint8)


class Circuit:
    """A compiled circuit (rv_circuit).  wire_counts = (z64_cells, gf2_cells), the reference's tuple order
    (src/proof/mod.rs:125)."""

    def __init__(self, ops: np.ndarray, wire_counts: Tuple[int, int], prove_only: bool = False):
        """prove_only: leave out the online verifier's tables (RV_COMPILE_PROVE_ONLY) -- a third of the compile time and memory."""
        self.ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        self.wire_counts = (int(wire_counts[0]), int(wire_counts[1]))
        h = C.c_void_p()
        N.check(N.lib().rv_circuit_compile_ex(_ptr(self.ops), self.ops.size, self.wire_counts[0], self.wire_counts[1], 1 if prove_only else 0, C.byref(h)))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            N.lib().rv_circuit_free(h)

    @property
    def handle(self):
        return self._h

    def stats(self) -> dict:
        st = N.CircuitStats()
        N.check(N.lib().rv_circuit_get_stats(self._h, C.byref(st)))
        return {n: int(getattr(st, n)) for n, _ in st._fields_}

    def export(self, what: int, dtype) -> np.ndarray:
        n = C.c_size_t(0)
        N.check(N.lib().rv_circuit_export(self._h, what, None, C.byref(n)))
        buf = np.zeros(n.value, dtype=np.uint8)
        N.check(N.lib().rv_circuit_export(self._h, what, _ptr(buf), C.byref(n)))
        return buf.view(dtype)


def _as_circuit(circuit, wire_counts) -> Circuit:
    if isinstance(circuit, Circuit):
        if wire_counts is not None and tuple(wire_counts) != circuit.wire_counts:
            raise ValueError("wire_counts differ from the ones the circuit was compiled with")
        return circuit
    return Circuit(circuit, wire_counts)


_ZERO_COPY_FROM = 4 << 20


def _take(out: C.c_void_p, n: C.c_size_t, zero_copy_from: int = 0):
    """Library-allocated bytes -> Python.  Small results are copied into `bytes`; big ones (proofs of Z64 / 10^8-gate circuits
    run to a gigabyte) stay in the library's pinned buffer, wrapped as a read-only numpy view that frees it when collected."""
    if n.value < (zero_copy_from or _ZERO_COPY_FROM):
        b = C.string_at(out, n.value)
        N.lib().rv_free(out)
        return b
    import weakref

    arr = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_uint8)), shape=(n.value,))
    arr.flags.writeable = False
    weakref.finalize(arr, N.lib().rv_free, C.c_void_p(out.value))
    return arr


class Session:
    """One proof shard in flight (rv_session): packed instances [first, first + count) of the 32
    (src/proof/mod.rs:127-157).  Used directly for multi-GPU sharding and for device-resident timing."""

    def __init__(self, circuit: Circuit, first_instance: int = 0, n_instances: int = N.PACKED_REPS, n_proofs: int = 1):
        """n_proofs > 1: the session holds that many independent proofs side by side (rv_session_create_multi); every phase is
        the same few kernel launches covering all of them.  Fill the slots with upload(..., slot=b), read them with fetch(slot=b)."""
        self.circuit = circuit
        h = C.c_void_p()
        N.check(N.lib().rv_session_create_multi(circuit.handle, first_instance, n_instances, n_proofs, C.byref(h)))
        self._h = h
        self.first_instance, self.n_instances, self.n_proofs = first_instance, n_instances, n_proofs

    @classmethod
    def _view(cls, circuit: Circuit, handle, first_instance: int, n_instances: int, n_proofs: int, owner) -> "Session":
        """A Session over a handle that `owner` (a Group) frees."""
        s = cls.__new__(cls)
        s.circuit, s._h, s._owner = circuit, C.c_void_p(handle), owner
        s.first_instance, s.n_instances, s.n_proofs = first_instance, n_instances, n_proofs
        return s

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and getattr(self, "_owner", None) is None:
            N.lib().rv_session_free(h)

    def upload(self, wit_gf2, wit_z64=(), seeds=None, slot: int = 0):
        self._wg = np.ascontiguousarray(np.asarray(wit_gf2, dtype=np.uint8))
        self._wz = np.ascontiguousarray(np.asarray(wit_z64, dtype=np.uint64))
        self._sd = _seeds_arr(seeds)
        N.check(N.lib().rv_session_upload_slot(self._h, slot, _ptr(self._wg), self._wg.size, _ptr(self._wz), self._wz.size, _ptr(self._sd)))

    def commit(self):
        N.check(N.lib().rv_session_commit(self._h))

    def hashes(self) -> bytes:
        out = np.zeros(self.n_proofs * self.n_instances * 8 * 32, dtype=np.uint8)
        N.check(N.lib().rv_session_hashes(self._h, _ptr(out)))
        return out.tobytes()

    def hashes_device(self):
        """The shard's repetition hashes in device memory as an object exposing __cuda_array_interface__ (uint8,
        n_instances * 256 bytes), e.g. `torch.as_tensor(s.hashes_device(), device="cuda")` for an NCCL all-gather on the
        session stream -- no host round trip."""
        ptr = int(N.lib().rv_session_hashes_device(self._h) or 0)
        if not ptr:
            raise N.ReverieError(N.E_ARG, "rv_session_commit has not run")
        n = self.n_proofs * self.n_instances * 8 * 32

        class _Dev:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 2}

        return _Dev()

    def all_hashes_device(self):
        """The session's own 256 x 32-byte receive buffer for the all-gather (__cuda_array_interface__); pass its address to
        open() afterwards: no copy, and the open phase replays as one CUDA graph."""
        ptr = int(N.lib().rv_session_all_hashes_device(self._h) or 0)
        n = self.n_proofs * N.TOTAL_REPS * 32

        class _Dev:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 2}

        d = _Dev()
        d.ptr = ptr
        return d

    def open(self, all_rep_hashes=None):
        """all_rep_hashes: 256*32 bytes (host bytes / numpy) or an int device pointer; None = own hashes (single shard)."""
        if all_rep_hashes is None:
            p = None
        elif isinstance(all_rep_hashes, int):
            p = C.c_void_p(all_rep_hashes)
        else:
            self._ah = np.frombuffer(bytes(all_rep_hashes), dtype=np.uint8)
            if self._ah.size != self.n_proofs * N.TOTAL_REPS * 32:
                raise ValueError("need 256 x 32 bytes of repetition hashes per proof of the session")
            p = _ptr(self._ah)
        N.check(N.lib().rv_session_open(self._h, p))

    def prove(self):
        """commit + open (own hashes) of a full shard, asynchronously; one CUDA graph launch after the first call."""
        N.check(N.lib().rv_session_prove(self._h))

    def fetch(self, slot: int = 0):
        comm = np.zeros(32, dtype=np.uint8)
        out, n = C.c_void_p(), C.c_size_t()
        N.check(N.lib().rv_session_fetch_slot(self._h, slot, _ptr(comm), C.byref(out), C.byref(n)))
        return comm.tobytes(), _take(out, n)

    def sync(self):
        N.check(N.lib().rv_session_sync(self._h))

    def status(self):
        """Synchronise and raise WitnessError if an AssertZero failed, without copying the proof."""
        N.check(N.lib().rv_session_status(self._h))

    def proof_device(self, slot: int = 0):
        """The shard's full-length proof buffer of `slot` in device memory (__cuda_array_interface__, uint8)."""
        ptr, n = C.c_void_p(), C.c_size_t()
        N.check(N.lib().rv_session_proof_device(self._h, C.byref(ptr), C.byref(n)))
        base = ptr.value + slot * int(N.lib().rv_session_proof_stride(self._h))

        class _Dev:
            __cuda_array_interface__ = {"shape": (n.value,), "typestr": "|u1", "data": (base, False), "version": 2}

        return _Dev()

    @property
    def stream(self) -> int:
        return int(N.lib().rv_session_stream(self._h) or 0)

    def peer_handle(self) -> bytes:
        """Opaque handle naming this session for rv_session_peer_link (raw pointers inside one process, CUDA IPC across)."""
        buf = np.zeros(N.PEER_HANDLE_BYTES, dtype=np.uint8)
        N.check(N.lib().rv_session_peer_handle(self._h, _ptr(buf)))
        return buf.tobytes()

    def peer_link(self, rank: int, world: int, handles: Sequence[bytes]):
        """Link this shard to the sessions holding the other shards of the same proofs (handles in rank order).  Afterwards
        prove() runs commit, the exchange of repetition hashes over peer memory, the challenge and the extraction into rank 0's
        proof buffer as one CUDA graph launch; fetch() on rank 0 returns the whole proof."""
        blob = np.frombuffer(b"".join(handles), dtype=np.uint8)
        if blob.size != world * N.PEER_HANDLE_BYTES:
            raise ValueError("need one handle per rank")
        N.check(N.lib().rv_session_peer_link(self._h, rank, world, _ptr(blob)))
        self.linked_rank, self.linked_world = rank, world

    def timing(self, enable: bool):
        N.check(N.lib().rv_session_timing(self._h, int(enable)))

    def kernel_times(self, reset: bool = True) -> list:
        arr = (N.KernelTime * 32)()
        n = N.check(N.lib().rv_session_kernel_times(self._h, arr, 32, int(reset)))
        return [dict(name=arr[i].name.decode(), ms=arr[i].ms, launches=int(arr[i].launches), algorithmic_bytes=int(arr[i].algorithmic_bytes))
                for i in range(min(n, 32))]

    @property
    def launch_count(self) -> int:
        return int(N.lib().rv_session_launch_count(self._h))


class Batch:
    """Several sessions of one circuit driven as a unit (rv_batch): each phase of all of them is one CUDA graph launch on the
    leader's stream.  Keep the sessions alive for as long as the batch."""

    def __init__(self, sessions: Sequence["Session"]):
        self.sessions = list(sessions)
        arr = (C.c_void_p * len(self.sessions))(*[s._h for s in self.sessions])
        h = C.c_void_p()
        N.check(N.lib().rv_batch_create(arr, len(self.sessions), C.byref(h)))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            N.lib().rv_batch_free(h)

    def commit(self):
        N.check(N.lib().rv_batch_commit(self._h))

    def open(self):
        N.check(N.lib().rv_batch_open(self._h))

    def prove(self):
        N.check(N.lib().rv_batch_prove(self._h))

    @property
    def stream(self) -> int:
        return int(N.lib().rv_batch_stream(self._h) or 0)


_ARR_TYPES = {}


def _arr(ctype, n):
    """ctypes array TYPE of n elements (creating `ctype * n` anew costs ~10 us each time)."""
    t = _ARR_TYPES.get((ctype, n))
    if t is None:
        t = _ARR_TYPES[(ctype, n)] = ctype * n
    return t


def _addr(a: np.ndarray):
    return a.__array_interface__["data"][0] if a.size else None  # (a.ctypes.data builds a ctypes helper object per call: ~3 us)


def _batch_args(n, wits_gf2, wits_z64, seeds):
    """Pointer / size arrays for n (witness, Z64 witness, seeds) triples.  A queue often repeats objects (one seed set, one Z64
    witness for all): each distinct object is converted once."""
    memo = {}

    def conv(x, dtype):
        k = (id(x), dtype)
        a = memo.get(k)
        if a is None:
            a = memo[k] = np.ascontiguousarray(np.asarray(x, dtype=dtype))
        return a

    def conv_seed(x):
        if x is None:
            return None
        k = (id(x), "s")
        a = memo.get(k)
        if a is None:
            a = memo[k] = _seeds_arr(x)
        return a

    wg = [conv(w, np.uint8) for w in wits_gf2]
    wz = [conv(w, np.uint64) for w in (wits_z64 if wits_z64 is not None else [()] * n)]
    sd = [conv_seed(x) for x in (seeds if seeds is not None else [None] * n)]
    vps, szs = _arr(C.c_void_p, n), _arr(C.c_size_t, n)
    a_wg = vps(*[_addr(w) for w in wg])
    a_wz = vps(*[_addr(w) for w in wz])
    a_sd = vps(*[_addr(x) if x is not None else None for x in sd])
    n_g = szs(*[w.size for w in wg])
    n_z = szs(*[w.size for w in wz])
    return (wg, wz, sd, memo), a_wg, n_g, a_wz, n_z, a_sd


def _batch_results(n, outs, lens, sts, want_proofs=True):
    vp = C.c_void_p
    proofs, first_err = [], None
    for i in range(n):
        if sts[i] == 0:
            proofs.append(Proof._from_library(outs[i], lens[i]) if (want_proofs and outs[i]) else None)  # wraps the library's buffer, no copy
        else:
            proofs.append(None)
            first_err = first_err if first_err is not None else sts[i]
    if first_err is not None:
        msg = {N.E_WITNESS_INVALID: "witness is invalid!", N.E_WITNESS_SHORT: "witness is too short", N.E_PEER: "a linked rank did not arrive"}.get(first_err, "proof failed")
        cls = N.WitnessError if first_err in (N.E_WITNESS_INVALID, N.E_WITNESS_SHORT) else N.ReverieError
        err = cls(first_err, msg)
        err.proofs = proofs
        raise err
    return proofs


class Group:
    """Proof::new on several GPUs behind one handle (rv_group).

    Group.local(circuit, devices, ...)        one process drives all the GPUs (peer access between them)
    Group.rank(circuit, rank, world, ...)     one process per GPU: exchange `handles()` over any host channel and `link()` them
                                              (`link_distributed()` does it through torch.distributed)
    A step proves n_sessions x slots proofs: every GPU holds its 32 / world packed instances of each of them; the exchange of
    repetition hashes and the assembly of the proofs happen on the devices, over peer memory (no collective library)."""

    def __init__(self, circuit: Circuit, handle, rank: int):
        self.circuit, self._h, self.rank_id = circuit, handle, rank
        w, m, ns, sl = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        N.check(N.lib().rv_group_info(handle, C.byref(w), C.byref(m), C.byref(ns), C.byref(sl)))
        self.world, self.n_members, self.n_sessions, self.slots = w.value, m.value, ns.value, sl.value
        per = N.PACKED_REPS // self.world
        self.members = []
        for mem in range(self.n_members):
            r = rank if self.n_members == 1 else mem
            ss = [Session._view(circuit, N.lib().rv_group_session(handle, mem, i), r * per, per, self.slots, self) for i in range(self.n_sessions)]
            self.members.append(ss)

    @staticmethod
    def local(circuit: Circuit, devices: Sequence[int], n_sessions: int = 1, slots: int = 1) -> "Group":
        arr = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        N.check(N.lib().rv_group_create_local(circuit.handle, arr, len(devices), n_sessions, slots, C.byref(h)))
        return Group(circuit, h, 0)

    @staticmethod
    def rank(circuit: Circuit, rank: int, world: int, n_sessions: int = 1, slots: int = 1) -> "Group":
        h = C.c_void_p()
        N.check(N.lib().rv_group_create_rank(circuit.handle, rank, world, n_sessions, slots, C.byref(h)))
        return Group(circuit, h, rank)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            for ss in getattr(self, "members", []):
                for s in ss:
                    s._h = None
            N.lib().rv_group_free(h)

    @property
    def sessions(self):
        """This process's sessions of member 0 (the only member of a rank group); each has its own stream."""
        return self.members[0]

    def handles(self) -> bytes:
        buf = np.zeros(int(N.lib().rv_group_handles_bytes(self._h)), dtype=np.uint8)
        N.check(N.lib().rv_group_handles(self._h, _ptr(buf)))
        return buf.tobytes()

    def link(self, all_handles: Sequence[bytes]):
        blob = np.frombuffer(b"".join(all_handles), dtype=np.uint8)
        N.check(N.lib().rv_group_link(self._h, _ptr(blob)))

    def link_distributed(self, group=None):
        """Exchange the handles through torch.distributed's host channel (once) and link."""
        import torch.distributed as dist

        if self.world == 1:
            return
        everyone = [None] * dist.get_world_size(group)
        dist.all_gather_object(everyone, self.handles(), group=group)
        self.link(everyone)

    def step(self):
        """Relaunch one step on the inputs already uploaded (asynchronous)."""
        N.check(N.lib().rv_group_step(self._h))

    def prove_batch(self, wits_gf2, wits_z64=None, seeds=None):
        """rv_group_prove_batch: returns the list of Proof on the assembling rank (rank 0 / a local group), a list of None on the
        other ranks of a rank group; raises the first per-proof error after all proofs have run."""
        n = len(wits_gf2)
        keep, a_wg, n_g, a_wz, n_z, a_sd = _batch_args(n, wits_gf2, wits_z64, seeds)
        outs, lens, sts = _arr(C.c_void_p, n)(), _arr(C.c_size_t, n)(), _arr(C.c_int, n)()
        N.check(N.lib().rv_group_prove_batch(self._h, n, a_wg, n_g, a_wz, n_z, a_sd, outs, lens, sts))
        return _batch_results(n, outs, lens, sts, want_proofs=self.rank_id == 0)

    def prove(self, wit_gf2, wit_z64=(), seeds=None):
        return self.prove_batch([wit_gf2], [wit_z64], [seeds] if seeds is not None else None)[0]

    def verify_batch(self, proofs, strict: bool = True):
        """Proof.verify for a list of proofs spread over the group's GPUs (rv_group_verify_batch: whole proofs per GPU).  Returns a
        list of bool; a rank group fills the entries i with i % world == rank and leaves None elsewhere."""
        n = len(proofs)
        bufs = [p._buf if isinstance(p, Proof) else p for p in proofs]
        arrs = [b if isinstance(b, np.ndarray) else np.frombuffer(b, dtype=np.uint8) for b in bufs]
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        lens = (C.c_size_t * n)(*[a.size for a in arrs])
        res, okay = (C.c_int * n)(*([-100] * n)), (C.c_int * n)(*([1] * n))
        N.check(N.lib().rv_group_verify_batch(self._h, n, ptrs, lens, res, okay))
        out = []
        for i in range(n):
            if res[i] == -100:
                out.append(None)
            else:
                out.append(N.check(res[i]) == 1 and (bool(okay[i]) or not strict))
        return out


def assemble(comm: bytes, parts: Sequence[bytes]) -> bytes:
    """src/proof/mod.rs:200-221 for shard blobs produced by Session.fetch()."""
    bufs = [np.frombuffer(p, dtype=np.uint8) for p in parts]
    ptrs = (C.c_void_p * len(bufs))(*[b.ctypes.data for b in bufs])
    lens = (C.c_size_t * len(bufs))(*[b.size for b in bufs])
    cm = np.frombuffer(comm, dtype=np.uint8)
    out, n = C.c_void_p(), C.c_size_t()
    N.check(N.lib().rv_proof_assemble(_ptr(cm), ptrs, lens, len(bufs), C.byref(out), C.byref(n)))
    return _take(out, n)


class Proof:
    """The bincode bytes of the reference's `Proof` struct (src/proof/mod.rs:40-66)."""

    __slots__ = ("_data", "_lib_ptr", "_lib_len", "__weakref__")

    def __init__(self, data):
        self._data = data if isinstance(data, (bytes, np.ndarray)) else bytes(data)
        self._lib_ptr = None

    @classmethod
    def _from_library(cls, ptr: int, n: int) -> "Proof":
        """Wraps a buffer the library allocated (rv_free'd when the Proof is collected) without touching its bytes: a batch of
        proofs costs a few microseconds of Python each, the bytes are only materialised when somebody reads them."""
        p = cls.__new__(cls)
        p._data, p._lib_ptr, p._lib_len = None, ptr, n
        return p

    def __del__(self):
        ptr, self._lib_ptr = getattr(self, "_lib_ptr", None), None
        if ptr:
            N.lib().rv_free(C.c_void_p(ptr))

    @property
    def _buf(self):
        """bytes, or a read-only numpy view of the library's buffer (kept alive by this object)."""
        if self._data is None:
            arr = np.ctypeslib.as_array(C.cast(C.c_void_p(self._lib_ptr), C.POINTER(C.c_uint8)), shape=(self._lib_len,))
            arr.flags.writeable = False
            self._data = arr
        return self._data

    @property
    def data(self) -> bytes:
        b = self._buf
        return b if isinstance(b, bytes) else b.tobytes()

    @staticmethod
    def new(circuit, wit_gf2, wit_z64=(), wire_counts=None, seeds=None) -> "Proof":
        """Proof::new (src/proof/mod.rs:119-222).  `seeds` (256 x 16 bytes) replaces the OsRng draw at :131-134 so a
        proof can be reproduced; None draws from the OS RNG like the reference."""
        wg = np.ascontiguousarray(np.asarray(wit_gf2, dtype=np.uint8))
        wz = np.ascontiguousarray(np.asarray(wit_z64, dtype=np.uint64))
        sd = _seeds_arr(seeds)
        out, n = C.c_void_p(), C.c_size_t()
        if isinstance(circuit, Circuit):
            c = _as_circuit(circuit, wire_counts)
            N.check(N.lib().rv_prove(c.handle, _ptr(wg), wg.size, _ptr(wz), wz.size, _ptr(sd), C.byref(out), C.byref(n)))
        else:  # the reference's call shape: the op list with every call (rv_proof_new compiles it once, then finds it in its cache)
            ops = np.ascontiguousarray(circuit, dtype=OP_DTYPE)
            N.check(N.lib().rv_proof_new(_ptr(ops), ops.size, _ptr(wg), wg.size, _ptr(wz), wz.size, int(wire_counts[0]), int(wire_counts[1]),
                                         _ptr(sd), C.byref(out), C.byref(n)))
        return Proof(_take(out, n))

    @staticmethod
    def new_streaming(ops, wit_gf2, wire_counts, seeds=None, window_ops: int = 0) -> "Proof":
        """Proof::new in streaming mode (rv_prove_streaming): the circuit is proved segment by segment with O(window_ops) device
        buffers, for circuits whose share tensor and transcripts do not fit in HBM.  Same bytes as Proof.new."""
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        wg = np.ascontiguousarray(np.asarray(wit_gf2, dtype=np.uint8))
        sd = _seeds_arr(seeds)
        out, n = C.c_void_p(), C.c_size_t()
        N.check(N.lib().rv_prove_streaming(_ptr(ops), ops.size, int(wire_counts[0]), int(wire_counts[1]), _ptr(wg), wg.size, None, 0, _ptr(sd), int(window_ops),
                                           C.byref(out), C.byref(n)))
        return Proof(_take(out, n))

    @staticmethod
    def new_batch(circuit, wits_gf2, wits_z64=None, wire_counts=None, seeds=None):
        """Proof::new for a list of independent witnesses of one circuit (rv_prove_batch): small GF(2) circuits are proved side
        by side, every kernel launch covering a group of them.  `seeds`: None or one 256 x 16-byte value (or None) per witness.
        Returns a list of Proof; raises the first per-proof error (WitnessError ...) after all proofs have run."""
        c = _as_circuit(circuit, wire_counts)
        n = len(wits_gf2)
        keep, a_wg, n_g, a_wz, n_z, a_sd = _batch_args(n, wits_gf2, wits_z64, seeds)
        outs, lens, sts = _arr(C.c_void_p, n)(), _arr(C.c_size_t, n)(), _arr(C.c_int, n)()
        N.check(N.lib().rv_prove_batch(c.handle, n, a_wg, n_g, a_wz, n_z, a_sd, outs, lens, sts))
        return _batch_results(n, outs, lens, sts)

    def verify_detail(self, circuit, wire_counts=None) -> Tuple[bool, bool]:
        """(accept, okay): `accept` is the reference's verdict -- the recomputed commitment equals the proof's
        (src/proof/mod.rs:305-306); `okay` is the AND of the online verifiers' AssertZero checks, which the reference computes
        (src/transcript/verifier/online.rs:176-178) and never reads."""
        buf = self._buf if isinstance(self._buf, np.ndarray) else np.frombuffer(self._buf, dtype=np.uint8)
        okay = C.c_int(1)
        if isinstance(circuit, Circuit):
            c = _as_circuit(circuit, wire_counts)
            accept = N.check(N.lib().rv_verify(c.handle, _ptr(buf), buf.size, C.byref(okay))) == 1
        else:
            ops = np.ascontiguousarray(circuit, dtype=OP_DTYPE)
            accept = N.check(N.lib().rv_proof_verify_ex(_ptr(ops), ops.size, int(wire_counts[0]), int(wire_counts[1]), _ptr(buf), buf.size, C.byref(okay))) == 1
        return accept, bool(okay.value)

    def verify(self, circuit, wire_counts=None, strict: bool = True) -> bool:
        """Proof::verify (src/proof/mod.rs:224-307).  strict (default) also requires every AssertZero of the opened repetitions to
        hold: the reference relies on the prover's own assert (src/transcript/prover.rs:221-228) for that, so a prover that skips
        it is accepted by `strict=False`, which is the reference's exact verdict."""
        accept, okay = self.verify_detail(circuit, wire_counts)
        return accept and (okay or not strict)

    def serialize(self) -> bytes:
        return self.data

    @staticmethod
    def deserialize(data: bytes) -> "Proof":
        return Proof(data)

    @property
    def comm(self) -> bytes:
        return bytes(self._buf[:32])

    def __len__(self):
        return self._lib_len if self._data is None else len(self._data)

    def __eq__(self, other):
        return isinstance(other, Proof) and self.data == other.data
