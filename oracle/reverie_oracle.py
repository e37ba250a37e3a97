"""CPU oracle #1 -- a literal Python restatement of trailofbits/reverie 0.3.2 (KKW MPC-in-the-head).

TEST INFRASTRUCTURE ONLY.  Nothing under ``reverie_b200/`` may import this file; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs use ``oracle/``.

PARITY PINNING STATUS
---------------------
* The reference holds NO golden vectors / KATs for this path (every test is a round trip with OsRng seeds,
  SURVEY.md section 4) and cannot be compiled here (no cargo/rustc, needs nightly + network crates).
* Primitives ARE pinned: AES-128-CTR against OpenSSL (``cryptography``) + FIPS-197 / SP 800-38A vectors,
  BLAKE3 against the ``blake3`` wheel, which wraps the very Rust crate the reference links (Cargo.toml:31).
* The protocol glue (interpreter, transcripts, packing, bincode layout) is restated from the source and checked
  against the reference's own test *cases* (truth tables single.rs:231-539, e2e circuit proof/mod.rs:397-427,
  pack/unpack length set algebra/mod.rs:304,333, omitted-player property generator/share.rs:76-141) and against
  a second independent restatement in C (``oracle/c``).  For those bytes: **parity unpinned** by reference vectors.

Every function cites the reference file:line it follows (paths relative to /root/reference).
This file deliberately mirrors the reference's *structure* (one packed instance = 8 reps x 8 players in a
u64, gate-by-gate sequential interpreter) so it can be read side by side with the Rust.
"""
from __future__ import annotations

import struct
from typing import Iterable, List, Optional, Sequence, Tuple

from blake3 import blake3 as _blake3
from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes

# --- src/lib.rs:17-38 -------------------------------------------------------------------------------
PLAYERS = 8
PACKED = 8
BATCH_SIZE = 128
ONLINE_REPS = 40
TOTAL_REPS = 256
PREPROCESSING_REPS = TOTAL_REPS - ONLINE_REPS
PACKED_REPS = TOTAL_REPS // PACKED
KEY_SIZE = 16  # src/crypto/prg.rs:9
HASH_SIZE = 32  # src/crypto/hash.rs:8
CTX_CHALLENGE = b"random-oracle challenge"  # src/proof/mod.rs:18

M64 = (1 << 64) - 1
LSB8 = 0x0101_0101_0101_0101

# --- circuit IR (mcircuit 0.1.7 Operation / CombineOperation; semantics fixed by single.rs:106-156) --
# An op is a tuple: (domain, name, *args) with domain in {"gf2","z64"}, or ("b2a", dst, src) / ("hint", z64, gf2)
OPS = ("Input", "Random", "Add", "AddConst", "Sub", "SubConst", "Mul", "MulConst", "AssertZero", "Const")


def b3(*parts: bytes) -> bytes:
    """HASH! macro, src/crypto/hash.rs:118-127 -- one-shot BLAKE3 of the concatenation."""
    h = _blake3()
    for p in parts:
        h.update(p)
    return h.digest()


class PRG:
    """src/crypto/prg.rs:13-37 -- AES-128-CTR, all-zero 128-bit big-endian counter block, keystream only."""

    def __init__(self, key: bytes):
        assert len(key) == KEY_SIZE
        self._enc = Cipher(algorithms.AES(key), modes.CTR(b"\x00" * 16)).encryptor()

    def gen(self, n: int) -> bytes:
        return self._enc.update(b"\x00" * n)


def expand_seed(seed: bytes) -> List[bytes]:
    """src/transcript/mod.rs:99-106."""
    prg = PRG(seed)
    return [prg.gen(KEY_SIZE) for _ in range(PLAYERS)]


class BufferedHasher:
    """src/crypto/hash.rs:17-58.  The 64 KiB buffering is transparent: plain BLAKE3 of the concatenation."""

    def __init__(self):
        self._h = _blake3()

    def update(self, b: bytes):
        self._h.update(b)

    def finalize(self) -> bytes:
        return self._h.copy().digest() if hasattr(self._h, "copy") else self._h.digest()


def new_packed_hasher() -> List[BufferedHasher]:
    """src/crypto/hash.rs:61-104."""
    return [BufferedHasher() for _ in range(PACKED)]


# =====================================================================================================
#  Domains
# =====================================================================================================
class GF2:
    """src/algebra/gf2/*.  Share = u64 (rep r, player p at bit 63-(8r+p)); Recon = u64 of 0x00/0xFF bytes,
    rep r at BE byte r."""

    name = "gf2"
    BATCH_BYTES = 16  # gf2/batch.rs:10

    # -- gf2/share.rs:220-238, gf2/recon.rs:333-362, gf2/domain.rs:10-18
    @staticmethod
    def share_zero():
        return 0

    @staticmethod
    def recon_zero():
        return 0

    @staticmethod
    def share_add(a, b):
        return a ^ b

    share_sub = share_add

    @staticmethod
    def share_mul_recon(s, r):
        return s & r

    @staticmethod
    def recon_add(a, b):
        return a ^ b

    recon_sub = recon_add

    @staticmethod
    def recon_mul(a, b):
        return a & b

    @staticmethod
    def recon_is_zero(a):
        return a == 0

    @staticmethod
    def const_to_recon(v) -> int:
        """gf2/recon.rs:274-287."""
        return M64 if v else 0

    @staticmethod
    def reconstruct(t: int) -> int:
        """gf2/domain.rs:47-63."""
        t ^= t >> 4
        t ^= t >> 2
        t ^= t >> 1
        t &= LSB8
        t |= (t << 1) & M64
        t |= (t << 2) & M64
        t |= (t << 4) & M64
        return t

    @staticmethod
    def batches_to_shares(batches: Sequence[Sequence[bytes]]) -> List[int]:
        """gf2/domain.rs:66-173 + byte_to_shares_avx2 :293-378.
        Byte i of every (rep,player) batch -> shares 8i..8i+7, MSB of the byte first; (rep,player) -> bit 63-(8r+p)."""
        out = []
        for i in range(GF2.BATCH_BYTES):
            src = [batches[r][p][i] for r in range(PACKED) for p in range(PLAYERS)]  # src[8r+p]
            for j in range(8):
                v = 0
                for k in range(64):
                    v |= ((src[k] >> (7 - j)) & 1) << (63 - k)
                out.append(v)
        return out

    @staticmethod
    def random_batch(prg: PRG) -> bytes:
        """gf2/batch.rs:17-21."""
        return prg.gen(16)

    ZERO_BATCH = b"\x00" * 16

    # -- hashing: gf2/share.rs:211-218, gf2/recon.rs:314-321
    @staticmethod
    def hash_share(s: int, hashers):
        bs = s.to_bytes(8, "big")
        for i in range(PACKED):
            hashers[i].update(bs[i : i + 1])

    hash_recon = hash_share

    # -- packing: gf2/share.rs:87-149
    @staticmethod
    def pack_selected(src: Sequence[int], selected: Sequence[int]) -> List[bytes]:
        dst = [bytearray() for _ in range(PACKED)]
        ext = [(idx, (PACKED - 1 - idx) * PLAYERS + (PLAYERS - 1 - pl)) for idx, pl in enumerate(selected) if pl < PLAYERS]
        if not ext:
            return [bytes(d) for d in dst]
        n = len(src)
        for c in range(0, n - n % 8, 8):
            for rep, shift in ext:
                dst[rep].append(GF2._pack8(src[c : c + 8], shift))
        arr = list(src[n - n % 8 :]) + [0] * (8 - n % 8)  # residue group is ALWAYS flushed (:131-138)
        for rep, shift in ext:
            dst[rep].append(GF2._pack8(arr, shift))
        return [bytes(d) for d in dst]

    @staticmethod
    def _pack8(arr, shift) -> int:
        """gf2/share.rs:66-85 -- first element -> MSB."""
        res = 0
        for k in range(8):
            res = (res << 1) | ((arr[k] >> shift) & 1)
        return res

    # -- gf2/share.rs:151-208
    @staticmethod
    def unpack_selected(src: Sequence[bytes], selected: Sequence[int]) -> List[int]:
        length = len(src[0])
        for s in src:
            if len(s) != length:
                raise FormatError("gf2 unpack_selected: ragged lengths")  # assert_eq! :158-164
        out = []
        for i in range(length):
            tmp = [0] * 64
            for j in range(PACKED):
                tmp[selected[j] + PLAYERS * j] = src[j][i]
            for b in range(8):
                v = 0
                for k in range(64):
                    v |= ((tmp[k] >> (7 - b)) & 1) << (63 - k)
                out.append(v)
        return out

    # -- gf2/recon.rs:190-239 (pack), :127-148 (bit order)
    @staticmethod
    def pack_recon(src: Sequence[int], selected: Sequence[bool]) -> List[bytes]:
        dst = [bytearray() for _ in range(PACKED)]
        if not any(selected):
            return [bytes(d) for d in dst]
        shifts = [(i, 64 - (i + 1) * 8) for i in range(PACKED) if selected[i]]
        n = len(src)

        def pack8(arr, shift):
            res = (arr[0] >> shift) & 2  # first element contributes bit 1, then 7 shifts -> MSB
            for k in range(1, 8):
                res |= (arr[k] >> shift) & 1
                if k < 7:
                    res = (res << 1) & 0xFF
            return res

        for c in range(0, n - n % 8, 8):
            for rep, shift in shifts:
                dst[rep].append(pack8(src[c : c + 8], shift))
        arr = list(src[n - n % 8 :]) + [0] * (8 - n % 8)  # residue ALWAYS flushed (:224-229)
        for rep, shift in shifts:
            dst[rep].append(pack8(arr, shift))
        return [bytes(d) for d in dst]

    # -- gf2/recon.rs:151-167, :241-259
    @staticmethod
    def unpack_recon(src: Sequence[bytes]) -> List[int]:
        n = len(src[0])
        for s in src:
            if len(s) < n:
                raise FormatError("gf2 unpack_recon: short slice")  # index panic in the reference
        out = []
        for i in range(n):
            for bit in range(8):
                v = 0
                for rep in range(PACKED):
                    if (src[rep][i] >> (7 - bit)) & 1:
                        v |= 0xFF << (8 * (7 - rep))
                out.append(v)
        return out


class Z64:
    """src/algebra/z64/*.  Share = [[u64; 8 players]; 8 reps]; Recon = [u64; 8 reps]; all wrapping."""

    name = "z64"
    BATCH_BYTES = 1024  # z64/batch.rs:10 (128 u64)

    @staticmethod
    def share_zero():
        return tuple((0,) * PLAYERS for _ in range(PACKED))

    @staticmethod
    def recon_zero():
        return (0,) * PACKED

    @staticmethod
    def share_add(a, b):
        return tuple(tuple((x + y) & M64 for x, y in zip(ra, rb)) for ra, rb in zip(a, b))

    @staticmethod
    def share_sub(a, b):
        return tuple(tuple((x - y) & M64 for x, y in zip(ra, rb)) for ra, rb in zip(a, b))

    @staticmethod
    def share_mul_recon(s, r):
        """z64/domain.rs:4-16."""
        return tuple(tuple((x * r[i]) & M64 for x in s[i]) for i in range(PACKED))

    @staticmethod
    def recon_add(a, b):
        return tuple((x + y) & M64 for x, y in zip(a, b))

    @staticmethod
    def recon_sub(a, b):
        return tuple((x - y) & M64 for x, y in zip(a, b))

    @staticmethod
    def recon_mul(a, b):
        return tuple((x * y) & M64 for x, y in zip(a, b))

    @staticmethod
    def recon_is_zero(a):
        return all(x == 0 for x in a)

    @staticmethod
    def const_to_recon(v):
        """z64/recon.rs:123-129."""
        return (int(v) & M64,) * PACKED

    @staticmethod
    def reconstruct(s):
        """z64/domain.rs:53-61."""
        return tuple(sum(row) & M64 for row in s)

    @staticmethod
    def batches_to_shares(batches):
        """z64/domain.rs:64-83; batch = 128 native-endian (LE) u64 (z64/batch.rs:25-30)."""
        words = [[struct.unpack("<128Q", batches[r][p]) for p in range(PLAYERS)] for r in range(PACKED)]
        return [tuple(tuple(words[r][p][i] for p in range(PLAYERS)) for r in range(PACKED)) for i in range(128)]

    @staticmethod
    def random_batch(prg: PRG) -> bytes:
        return prg.gen(1024)

    ZERO_BATCH = b"\x00" * 1024

    @staticmethod
    def hash_share(s, hashers):
        """z64/share.rs:100-108."""
        for i in range(PACKED):
            hashers[i].update(struct.pack("<8Q", *s[i]))

    @staticmethod
    def hash_recon(r, hashers):
        """z64/recon.rs:131-137."""
        for i in range(PACKED):
            hashers[i].update(struct.pack("<Q", r[i]))

    @staticmethod
    def pack_selected(src, selected):
        """z64/share.rs:37-49."""
        dst = [bytearray() for _ in range(PACKED)]
        for e in src:
            for i in range(PACKED):
                if selected[i] < PLAYERS:
                    dst[i] += struct.pack("<Q", e[i][selected[i]])
        return [bytes(d) for d in dst]

    @staticmethod
    def unpack_selected(src, selected):
        """z64/share.rs:51-91 -- count from src[0]; missing chunks in other reps read as zero."""
        n = len(src[0]) // 8
        out = []
        for k in range(n):
            val = [[0] * PLAYERS for _ in range(PACKED)]
            for j in range(PACKED):
                c = src[j][8 * k : 8 * k + 8]
                val[j][selected[j]] = struct.unpack("<Q", c)[0] if len(c) == 8 else 0
            out.append(tuple(tuple(r) for r in val))
        return out

    @staticmethod
    def pack_recon(src, selected):
        """z64/recon.rs:46-66."""
        dst = [bytearray() for _ in range(PACKED)]
        if not any(selected):
            return [bytes(d) for d in dst]
        for e in src:
            for i in range(PACKED):
                if selected[i]:
                    dst[i] += struct.pack("<Q", e[i])
        return [bytes(d) for d in dst]

    @staticmethod
    def unpack_recon(src):
        """z64/recon.rs:68-107."""
        n = len(src[0]) // 8
        out = []
        for k in range(n):
            v = []
            for j in range(PACKED):
                c = src[j][8 * k : 8 * k + 8]
                v.append(struct.unpack("<Q", c)[0] if len(c) == 8 else 0)
            out.append(tuple(v))
        return out


class FormatError(Exception):
    """Stands for the reference's panics on malformed proofs (assert_eq!/index out of range)."""


class WitnessError(Exception):
    """prover.rs:190 ('witness is too short') and :223 ('witness is invalid!')."""


# =====================================================================================================
#  Generator  (src/generator/batch.rs, share.rs)
# =====================================================================================================
class ShareGen:
    def __init__(self, D, keys: Sequence[Sequence[bytes]], omit: Sequence[int]):
        """generator/share.rs:16-52; BatchGen::new generator/batch.rs:13-28."""
        self.D = D
        self.omit = list(omit)
        self.prgs = [[PRG(keys[r][p]) for p in range(PLAYERS)] for r in range(PACKED)]
        self.batches = [[D.ZERO_BATCH for _ in range(PLAYERS)] for _ in range(PACKED)]
        self.shares: list = []
        self.next_idx = BATCH_SIZE

    def next(self):
        """generator/share.rs:54-65; BatchGen::gen generator/batch.rs:30-40 (omitted player's batch stays zero)."""
        if self.next_idx >= BATCH_SIZE:
            for r in range(PACKED):
                for p in range(PLAYERS):
                    if p != self.omit[r]:
                        self.batches[r][p] = self.D.random_batch(self.prgs[r][p])
            self.shares = self.D.batches_to_shares(self.batches)
            self.next_idx = 0
        s = self.shares[self.next_idx]
        self.next_idx += 1
        return s


def share_gen_from_rep_seeds(D, seeds: Sequence[bytes]) -> ShareGen:
    """src/transcript/mod.rs:108-122."""
    return ShareGen(D, [expand_seed(s) for s in seeds], [PLAYERS] * PACKED)


# =====================================================================================================
#  Transcripts  (src/transcript/{prover.rs, verifier/online.rs, verifier/preprocess.rs})
# =====================================================================================================
class Wire:
    __slots__ = ("mask", "corr")

    def __init__(self, mask, corr):
        self.mask = mask
        self.corr = corr


class _TranscriptBase:
    IS_PROVER = False

    def hash(self) -> List[bytes]:
        """src/transcript/mod.rs:77-96 -- per rep H(preprocess_hash || online_hash)."""
        on = self.online_hash()
        pre = self.preprocess_hash()
        return [b3(pre[i], on[i]) for i in range(PACKED)]


class ProverTranscript(_TranscriptBase):
    IS_PROVER = True

    def __init__(self, D, witness: Iterable, seeds: Sequence[bytes]):
        """prover.rs:36-53."""
        self.D = D
        self.seeds = list(seeds)
        self.witness = iter(witness)
        self.share_gen = share_gen_from_rep_seeds(D, seeds)
        self.hash_online = new_packed_hasher()
        self.hash_preprocess = new_packed_hasher()
        self.reconstructions: list = []
        self.corrections: list = []
        self.inputs: list = []

    def input(self) -> Wire:
        """prover.rs:181-199."""
        D = self.D
        mask = self.share_gen.next()
        lam = D.reconstruct(mask)
        try:
            w = next(self.witness)
        except StopIteration:
            raise WitnessError("witness is too short")
        corr = D.recon_sub(D.const_to_recon(w), lam)
        D.hash_recon(corr, self.hash_online)
        self.inputs.append(corr)
        return Wire(mask, corr)

    def online_hash(self):
        return [h.finalize() for h in self.hash_online]

    def preprocess_hash(self):
        return [h.finalize() for h in self.hash_preprocess]

    def reconstruct(self, mask):
        """prover.rs:209-213."""
        self.D.hash_share(mask, self.hash_online)
        self.reconstructions.append(mask)
        return self.D.reconstruct(mask)

    def correction(self, corr):
        """prover.rs:215-219."""
        self.D.hash_recon(corr, self.hash_preprocess)
        self.corrections.append(corr)
        return corr

    def zero_check(self, recon):
        """prover.rs:221-228."""
        if not self.D.recon_is_zero(recon):
            raise WitnessError("witness is invalid!")

    def new_mask(self):
        return self.share_gen.next()

    def extract(self, players: Sequence[int]):
        """prover.rs:57-175 -> (Vec<OpenOnline>, Vec<OpenPreprocessing>)."""
        D = self.D
        selected = [p < PLAYERS for p in players]
        dst_recon = D.pack_selected(self.reconstructions, players)
        dst_corr = D.pack_recon(self.corrections, selected)
        dst_input = D.pack_recon(self.inputs, selected)
        online, pre = [], []
        for rep in range(PACKED):
            omit = players[rep]
            if omit < PLAYERS:
                seeds = expand_seed(self.seeds[rep])
                seeds[omit] = b"\x00" * KEY_SIZE
                online.append(dict(omit=omit, seeds=seeds, recons=dst_recon[rep], corrs=dst_corr[rep], inputs=dst_input[rep]))
            else:
                pre.append(dict(seed=self.seeds[rep], comm_online=self.hash_online[rep].finalize()))
        return online, pre


class VerifierTranscriptOnline(_TranscriptBase):
    def __init__(self, D, opens: Sequence[dict]):
        """verifier/online.rs:25-121."""
        self.D = D
        self.corrs = D.unpack_recon([o["corrs"] for o in opens])
        self.inputs = D.unpack_recon([o["inputs"] for o in opens])
        omit = [o["omit"] for o in opens]
        for o in omit:
            if o >= PLAYERS:
                raise FormatError("omit out of range")
        self.recons = D.unpack_selected([o["recons"] for o in opens], omit)
        self.share_gen = ShareGen(D, [o["seeds"] for o in opens], omit)
        self.hash_online = new_packed_hasher()
        self.hash_preprocess = new_packed_hasher()
        self._ci = self._ii = self._ri = 0
        self.okay = True

    def _next(self, lst, attr, default):
        i = getattr(self, attr)
        setattr(self, attr, i + 1)
        return lst[i] if i < len(lst) else default  # unwrap_or_default(), online.rs:124,163,171

    def input(self) -> Wire:
        """online.rs:123-130 (corr drawn BEFORE the mask; both counters are independent)."""
        corr = self._next(self.inputs, "_ii", self.D.recon_zero())
        self.D.hash_recon(corr, self.hash_online)
        return Wire(self.share_gen.next(), corr)

    def online_hash(self):
        return [h.finalize() for h in self.hash_online]

    def preprocess_hash(self):
        return [h.finalize() for h in self.hash_preprocess]

    def reconstruct(self, mask):
        """online.rs:140-167."""
        msg = self._next(self.recons, "_ri", self.D.share_zero())
        mask = self.D.share_add(mask, msg)
        self.D.hash_share(mask, self.hash_online)
        return self.D.reconstruct(mask)

    def correction(self, _corr):
        """online.rs:169-174."""
        corr = self._next(self.corrs, "_ci", self.D.recon_zero())
        self.D.hash_recon(corr, self.hash_preprocess)
        return corr

    def zero_check(self, recon):
        """online.rs:176-178 (never read by Proof::verify)."""
        self.okay &= self.D.recon_is_zero(recon)

    def new_mask(self):
        return self.share_gen.next()


class VerifierTranscriptPreprocess(_TranscriptBase):
    def __init__(self, D, opens: Sequence[dict]):
        """verifier/preprocess.rs:17-43."""
        self.D = D
        self.comms_online = [o["comm_online"] for o in opens]
        self.share_gen = share_gen_from_rep_seeds(D, [o["seed"] for o in opens])
        self.hash_preprocess = new_packed_hasher()

    def input(self) -> Wire:
        return Wire(self.share_gen.next(), self.D.recon_zero())

    def online_hash(self):
        return list(self.comms_online)

    def preprocess_hash(self):
        return [h.finalize() for h in self.hash_preprocess]

    def reconstruct(self, _mask):
        return self.D.recon_zero()

    def correction(self, corr):
        self.D.hash_recon(corr, self.hash_preprocess)
        return corr

    def zero_check(self, _recon):
        pass

    def new_mask(self):
        return self.share_gen.next()


# =====================================================================================================
#  Interpreter  (src/interpreter/single.rs, combine.rs)
# =====================================================================================================
class Instance:
    def __init__(self, D, transcript, cells: int):
        """single.rs:13-19."""
        self.D = D
        self.transcript = transcript
        self.wires = [Wire(D.share_zero(), D.recon_zero()) for _ in range(cells)]

    @staticmethod
    def op_mul(D, t, w1: Wire, w2: Wire) -> Wire:
        """single.rs:25-69."""
        mask_ab = t.new_mask()
        mask_new = t.new_mask()
        a = D.reconstruct(w1.mask)
        b = D.reconstruct(w2.mask)
        c = D.reconstruct(mask_ab)
        delta = t.correction(D.recon_sub(D.recon_mul(a, b), c))
        s = D.share_sub(
            D.share_add(D.share_add(D.share_mul_recon(w2.mask, w1.corr), D.share_mul_recon(w1.mask, w2.corr)), mask_ab),
            mask_new,
        )
        recon = D.recon_add(t.reconstruct(s), delta)
        return Wire(mask_new, D.recon_add(recon, D.recon_mul(w1.corr, w2.corr)))

    @staticmethod
    def op_add(D, w1, w2):
        return Wire(D.share_add(w1.mask, w2.mask), D.recon_add(w1.corr, w2.corr))

    def step(self, name: str, *args):
        """single.rs:106-157."""
        D, t, W = self.D, self.transcript, self.wires
        if name == "Input":
            W[args[0]] = t.input()
        elif name == "Add":
            W[args[0]] = self.op_add(D, W[args[1]], W[args[2]])
        elif name == "Sub":
            a, b = W[args[1]], W[args[2]]
            W[args[0]] = Wire(D.share_sub(a.mask, b.mask), D.recon_sub(a.corr, b.corr))
        elif name == "Mul":
            W[args[0]] = self.op_mul(D, t, W[args[1]], W[args[2]])
        elif name == "AddConst":
            w = W[args[1]]
            W[args[0]] = Wire(w.mask, D.recon_add(w.corr, D.const_to_recon(args[2])))
        elif name == "SubConst":
            w = W[args[1]]
            W[args[0]] = Wire(w.mask, D.recon_sub(w.corr, D.const_to_recon(args[2])))
        elif name == "MulConst":
            w = W[args[1]]
            v = D.const_to_recon(args[2])
            W[args[0]] = Wire(D.share_mul_recon(w.mask, v), D.recon_mul(w.corr, v))
        elif name == "AssertZero":
            w = W[args[0]]
            m = t.reconstruct(w.mask)
            t.zero_check(D.recon_add(w.corr, m))
        elif name == "Random":
            W[args[0]] = Wire(t.new_mask(), D.recon_zero())
        elif name == "Const":
            W[args[0]] = Wire(D.share_zero(), D.const_to_recon(args[1]))
        else:
            raise ValueError(name)

    def value(self, idx):
        """interpreter/mod.rs:17-19 (test helper)."""
        w = self.wires[idx]
        return self.D.recon_add(self.D.reconstruct(w.mask), w.corr)


def recon_gf2_to_z64(recon, bits: Sequence[Wire]):
    """combine.rs:19-36 -- wire i becomes bit i (LSB first) of each rep's u64."""
    z = [0] * PACKED
    for w in bits:
        v = (GF2.recon_add(recon(w.mask), w.corr) & LSB8).to_bytes(8, "big")
        for j in range(PACKED):
            z[j] = ((z[j] << 1) | v[j]) & M64
    return tuple(int(f"{x:064b}"[::-1], 2) for x in z)


class CombineInstance:
    def __init__(self, gf2: Instance, z64: Instance):
        self.gf2 = gf2
        self.z64 = z64

    def add_64(self, t, a: Sequence[Wire], b: Sequence[Wire]) -> List[Wire]:
        """combine.rs:39-93 -- 64-bit ripple adder, 63 ANDs, no carry out."""
        AND = lambda x, y: Instance.op_mul(GF2, t, x, y)
        XOR = lambda x, y: Instance.op_add(GF2, x, y)
        res: List[Optional[Wire]] = [None] * 64
        carry = AND(a[0], b[0])
        res[0] = XOR(a[0], b[0])
        for i in range(1, 63):
            ac = XOR(a[i], carry)
            bc = XOR(b[i], carry)
            ac_bc = AND(ac, bc)
            res[i] = XOR(ac, b[i])
            carry = XOR(ac_bc, carry)
        res[63] = XOR(carry, XOR(a[63], b[63]))
        return res  # type: ignore

    def hash(self) -> List[bytes]:
        """combine.rs:104-118."""
        g = self.gf2.transcript.hash()
        z = self.z64.transcript.hash()
        return [b3(g[i], z[i]) for i in range(PACKED)]

    def step(self, op):
        """combine.rs:120-221."""
        kind = op[0]
        if kind == "hint":
            _, z64n, gf2n = op
            while len(self.z64.wires) < z64n:
                self.z64.wires.append(Wire(Z64.share_zero(), Z64.recon_zero()))
            while len(self.gf2.wires) < gf2n:
                self.gf2.wires.append(Wire(0, 0))
        elif kind == "gf2":
            self.gf2.step(*op[1:])
        elif kind == "z64":
            self.z64.step(*op[1:])
        elif kind == "b2a":
            _, dst, src = op
            tg, tz = self.gf2.transcript, self.z64.transcript
            gf2_wires = [Wire(tg.new_mask(), 0) for _ in range(64)]
            z64_value = recon_gf2_to_z64(GF2.reconstruct, gf2_wires)
            z64_mask = tz.new_mask()
            z64_corr = tz.correction(Z64.recon_sub(z64_value, Z64.reconstruct(z64_mask)))
            res = self.add_64(tg, gf2_wires, self.gf2.wires[src : src + 64])
            z64_recon = recon_gf2_to_z64(lambda v: tg.reconstruct(v), res)
            self.z64.wires[dst] = Wire(Z64.share_sub(Z64.share_zero(), z64_mask), Z64.recon_sub(z64_recon, z64_corr))
        else:
            raise ValueError(kind)


# =====================================================================================================
#  Proof  (src/proof/mod.rs)
# =====================================================================================================
def random_int(reader, bound: int) -> int:
    """proof/mod.rs:68-72."""
    return int.from_bytes(reader(16), "little") % bound


def challenge_to_opening(challenge: bytes) -> dict:
    """proof/mod.rs:74-83; RandomOracle src/crypto/ro.rs:7-20. Later draws overwrite earlier ones."""
    h = _blake3()
    h.update(CTX_CHALLENGE)
    h.update(b"\x00")
    h.update(challenge)
    pos = [0]

    def reader(n):
        out = h.digest(length=pos[0] + n)[pos[0] :]
        pos[0] += n
        return out

    online: dict = {}
    while len(online) < ONLINE_REPS:
        rep = random_int(reader, TOTAL_REPS)
        omit = random_int(reader, PLAYERS)
        online[rep] = omit
    return online


def opening_to_packed(open_: dict) -> List[List[int]]:
    """proof/mod.rs:85-100."""
    return [[open_.get(i * PACKED + j, PLAYERS) for j in range(PACKED)] for i in range(PACKED_REPS)]


def combine_hashes(hashes: Iterable[bytes]) -> bytes:
    """proof/mod.rs:102-108."""
    return b3(*hashes)


def default_seeds() -> List[bytes]:
    """seed[r] = BLAKE3("reverie-b200 seed" || LE32(r))[..16] (deterministic stand-in for OsRng)."""
    return [b3(b"reverie-b200 seed", struct.pack("<I", r))[:16] for r in range(TOTAL_REPS)]


def prove(circuit, wit_gf2, wit_z64, wire_counts, seeds: Sequence[bytes], instances=range(PACKED_REPS), tap=None):
    """Proof::new, proof/mod.rs:119-222, with the 256 rep seeds injected instead of OsRng (:131-134).
    Returns the Proof as a dict {comm, gf2:{online,preprocessing}, z64:{...}}."""
    z64_count, gf2_count = wire_counts  # NOTE tuple order, proof/mod.rs:125
    comms: List[bytes] = []
    transcripts = []
    for i in instances:
        keys = [seeds[i * PACKED + j] for j in range(PACKED)]
        ig = Instance(GF2, ProverTranscript(GF2, wit_gf2, keys), gf2_count)
        iz = Instance(Z64, ProverTranscript(Z64, wit_z64, keys), z64_count)  # SAME seeds, :144
        ins = CombineInstance(ig, iz)
        for op in circuit:
            ins.step(op)
        comms.extend(ins.hash())
        transcripts.append((ig.transcript, iz.transcript))
    if tap is not None:
        tap["rep_hashes"] = list(comms)
        tap["transcripts"] = transcripts
    if len(comms) != TOTAL_REPS:
        return None  # partial run (sharding tests use tap)
    comm = combine_hashes(comms)
    open_ = challenge_to_opening(comm)
    packed_open = opening_to_packed(open_)
    gf2 = dict(online=[], preprocessing=[])
    z64 = dict(online=[], preprocessing=[])
    for (tg, tz), players in zip(transcripts, packed_open):
        o, p = tg.extract(players)
        gf2["online"] += o
        gf2["preprocessing"] += p
        o, p = tz.extract(players)
        z64["online"] += o
        z64["preprocessing"] += p
    return dict(comm=comm, gf2=gf2, z64=z64)


def verify(proof: dict, circuit, wire_counts, tap=None) -> bool:
    """Proof::verify, proof/mod.rs:224-307.  Malformed inner lengths (reference: panic) -> FormatError."""
    for d in ("gf2", "z64"):
        if len(proof[d]["online"]) != ONLINE_REPS or len(proof[d]["preprocessing"]) != PREPROCESSING_REPS:
            return False
    z64_count, gf2_count = wire_counts
    hashes: List[bytes] = []
    okay = True
    for c in range(0, ONLINE_REPS, PACKED):
        ig = Instance(GF2, VerifierTranscriptOnline(GF2, proof["gf2"]["online"][c : c + PACKED]), gf2_count)
        iz = Instance(Z64, VerifierTranscriptOnline(Z64, proof["z64"]["online"][c : c + PACKED]), z64_count)
        ins = CombineInstance(ig, iz)
        for op in circuit:
            ins.step(op)
        hashes += ins.hash()
        okay &= ig.transcript.okay and iz.transcript.okay
    for c in range(0, PREPROCESSING_REPS, PACKED):
        ig = Instance(GF2, VerifierTranscriptPreprocess(GF2, proof["gf2"]["preprocessing"][c : c + PACKED]), gf2_count)
        iz = Instance(Z64, VerifierTranscriptPreprocess(Z64, proof["z64"]["preprocessing"][c : c + PACKED]), z64_count)
        ins = CombineInstance(ig, iz)
        for op in circuit:
            ins.step(op)
        hashes += ins.hash()
    open_ = challenge_to_opening(proof["comm"])
    on = iter(hashes[:ONLINE_REPS])
    pre = iter(hashes[ONLINE_REPS:])
    ordered = [next(on) if i in open_ else next(pre) for i in range(TOTAL_REPS)]
    if tap is not None:
        tap["okay"] = okay
        tap["rep_hashes"] = ordered
    return combine_hashes(ordered) == proof["comm"]


# =====================================================================================================
#  bincode 1.3 default config for `Proof` (proof/mod.rs:40-66): LE, u64 lengths, fixed arrays inline.
# =====================================================================================================
def serialize(proof: dict) -> bytes:
    out = bytearray(proof["comm"])
    for d in ("gf2", "z64"):
        ps = proof[d]
        out += struct.pack("<Q", len(ps["online"]))
        for o in ps["online"]:
            out.append(o["omit"])
            for k in o["seeds"]:
                out += k
            for f in ("recons", "corrs", "inputs"):
                out += struct.pack("<Q", len(o[f])) + o[f]
        out += struct.pack("<Q", len(ps["preprocessing"]))
        for p in ps["preprocessing"]:
            out += p["seed"] + p["comm_online"]
    return bytes(out)


def deserialize(buf: bytes) -> dict:
    pos = [0]

    def take(n):
        if pos[0] + n > len(buf):
            raise FormatError("truncated proof")
        b = buf[pos[0] : pos[0] + n]
        pos[0] += n
        return b

    def u64():
        return struct.unpack("<Q", take(8))[0]

    proof = dict(comm=take(32))
    for d in ("gf2", "z64"):
        online = []
        for _ in range(u64()):
            o = dict(omit=take(1)[0], seeds=[take(16) for _ in range(PLAYERS)])
            for f in ("recons", "corrs", "inputs"):
                o[f] = take(u64())
            online.append(o)
        pre = []
        for _ in range(u64()):
            pre.append(dict(seed=take(16), comm_online=take(32)))
        proof[d] = dict(online=online, preprocessing=pre)
    if pos[0] != len(buf):
        raise FormatError("trailing bytes")
    return proof


def largest_wires(circuit) -> Tuple[int, int]:
    """mcircuit::largest_wires as used at proof/mod.rs:125 -- returns (z64_cells, gf2_cells): 1 + the largest wire
    index touched per domain (B2A touches z64 dst and gf2 src..src+63)."""
    g = z = 0
    for op in circuit:
        k = op[0]
        if k == "hint":
            z, g = max(z, op[1]), max(g, op[2])
        elif k == "b2a":
            z, g = max(z, op[1] + 1), max(g, op[2] + 64)
        else:
            name, args = op[1], op[2:]
            if name in ("Input", "Random", "AssertZero"):
                idx = [args[0]]
            elif name in ("Add", "Sub", "Mul"):
                idx = list(args[:3])
            elif name in ("AddConst", "SubConst", "MulConst"):
                idx = list(args[:2])
            else:  # Const
                idx = [args[0]]
            m = max(idx) + 1
            if k == "gf2":
                g = max(g, m)
            else:
                z = max(z, m)
    return z, g
