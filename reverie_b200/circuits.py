"""Circuit front end: the packed op-record format, a Bristol-Fashion parser and circuit generators.

The reference's CLI reads a bincode `Vec<mcircuit::CombineOperation>` (src/main.rs:66); the op *semantics* are fixed by
the interpreter's match arms (src/interpreter/single.rs:106-156, src/interpreter/combine.rs:120-132).  This module
produces the same op set as packed 24-byte records (`rv_op`, include/reverie_b200.h) which is what crosses the C ABI.

The public Bristol-Fashion files (sha256.txt, aes_128.txt) cannot be fetched in this environment, so the SHA-256
compression circuit is generated here with the textbook construction (ripple-carry adders with one AND per bit,
`Ch = ((f^g)&e)^g`, `Maj = ((a^b)&(a^c))^a`): 22 696 AND gates vs. the public file's 22 573.
"""
from __future__ import annotations

import hashlib
import struct
from typing import Iterable, List, Sequence, Tuple

import numpy as np

OP_DTYPE = np.dtype(
    [("domain", "u1"), ("opcode", "u1"), ("pad", "<u2"), ("dst", "<u4"), ("a", "<u4"), ("b", "<u4"), ("imm", "<u8")]
)
assert OP_DTYPE.itemsize == 24

# domain
GF2, Z64, B2A, HINT = 0, 1, 2, 3
# opcode (mcircuit::Operation variants as matched at src/interpreter/single.rs:106-156)
INPUT, RANDOM, ADD, ADDC, SUB, SUBC, MUL, MULC, ASSERT_ZERO, CONST = range(10)


class Builder:
    """Append-only op list with SSA wire allocation for one domain (default GF(2))."""

    def __init__(self, domain: int = GF2):
        self.domain = domain
        self.recs: List[Tuple[int, int, int, int, int, int, int]] = []
        self.n_wires = 0

    def _new(self) -> int:
        w = self.n_wires
        self.n_wires += 1
        return w

    def _emit(self, opcode, dst=0, a=0, b=0, imm=0):
        self.recs.append((self.domain, opcode, 0, dst, a, b, imm & 0xFFFFFFFFFFFFFFFF))

    def input(self) -> int:
        w = self._new()
        self._emit(INPUT, w)
        return w

    def add(self, a, b) -> int:  # XOR in GF(2)
        w = self._new()
        self._emit(ADD, w, a, b)
        return w

    def sub(self, a, b) -> int:
        w = self._new()
        self._emit(SUB, w, a, b)
        return w

    def mul(self, a, b) -> int:  # AND in GF(2)
        w = self._new()
        self._emit(MUL, w, a, b)
        return w

    def addc(self, a, c) -> int:  # INV when c == 1 in GF(2)
        w = self._new()
        self._emit(ADDC, w, a, 0, int(c))
        return w

    def mulc(self, a, c) -> int:
        w = self._new()
        self._emit(MULC, w, a, 0, int(c))
        return w

    def const(self, c) -> int:
        w = self._new()
        self._emit(CONST, w, 0, 0, int(c))
        return w

    def assert_zero(self, a):
        self._emit(ASSERT_ZERO, 0, a)

    def ops(self) -> np.ndarray:
        return np.array(self.recs, dtype=OP_DTYPE) if self.recs else np.zeros(0, dtype=OP_DTYPE)


# ---------------------------------------------------------------------------------------------------------------------
#  SHA-256 compression function as a GF(2) circuit
# ---------------------------------------------------------------------------------------------------------------------
_K = [
    0x428A2F98, 0x71374491, 0xB5C0FBCF, 0xE9B5DBA5, 0x3956C25B, 0x59F111F1, 0x923F82A4, 0xAB1C5ED5, 0xD807AA98, 0x12835B01,
    0x243185BE, 0x550C7DC3, 0x72BE5D74, 0x80DEB1FE, 0x9BDC06A7, 0xC19BF174, 0xE49B69C1, 0xEFBE4786, 0x0FC19DC6, 0x240CA1CC,
    0x2DE92C6F, 0x4A7484AA, 0x5CB0A9DC, 0x76F988DA, 0x983E5152, 0xA831C66D, 0xB00327C8, 0xBF597FC7, 0xC6E00BF3, 0xD5A79147,
    0x06CA6351, 0x14292967, 0x27B70A85, 0x2E1B2138, 0x4D2C6DFC, 0x53380D13, 0x650A7354, 0x766A0ABB, 0x81C2C92E, 0x92722C85,
    0xA2BFE8A1, 0xA81A664B, 0xC24B8B70, 0xC76C51A3, 0xD192E819, 0xD6990624, 0xF40E3585, 0x106AA070, 0x19A4C116, 0x1E376C08,
    0x2748774C, 0x34B0BCB5, 0x391C0CB3, 0x4ED8AA4A, 0x5B9CCA4F, 0x682E6FF3, 0x748F82EE, 0x78A5636F, 0x84C87814, 0x8CC70208,
    0x90BEFFFA, 0xA4506CEB, 0xBEF9A3F7, 0xC67178F2,
]
SHA256_IV = [0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19]

# a 32-bit word is a list of 32 wire ids, index i = bit i (LSB first); None = constant 0 (used by shifts)


def _xor3(b: Builder, x, y, z):
    ws = [w for w in (x, y, z) if w is not None]
    if not ws:
        return None
    r = ws[0]
    for w in ws[1:]:
        r = b.add(r, w)
    return r


def _rotr(x, n):
    return [x[(i + n) % 32] for i in range(32)]


def _shr(x, n):
    return [x[i + n] if i + n < 32 else None for i in range(32)]


def _xor_words(b, x, y, z):
    return [_xor3(b, x[i], y[i], z[i]) for i in range(32)]


def _adder(b: Builder, x, y):
    """32-bit ripple-carry add mod 2^32: one AND per bit, carry' = ((x^c)&(y^c))^c (same cell as combine.rs:39-93)."""
    out = [None] * 32
    carry = b.mul(x[0], y[0])
    out[0] = b.add(x[0], y[0])
    for i in range(1, 31):
        xc = b.add(x[i], carry)
        yc = b.add(y[i], carry)
        t = b.mul(xc, yc)
        out[i] = b.add(xc, y[i])
        carry = b.add(t, carry)
    out[31] = b.add(carry, b.add(x[31], y[31]))
    return out


def _adder_const(b: Builder, x, k: int):
    """x + k mod 2^32 for a public constant k: still one AND per bit past the lowest set bit of k."""
    out = [None] * 32
    carry = None  # constant 0 so far
    for i in range(32):
        ki = (k >> i) & 1
        if carry is None:
            out[i] = b.addc(x[i], 1) if ki else x[i]
            if ki and i < 31:
                carry = x[i]  # carry = x & 1
            continue
        s = b.add(x[i], carry)
        out[i] = b.addc(s, 1) if ki else s
        if i < 31:
            kc = b.addc(carry, 1) if ki else carry  # k ^ c
            t = b.mul(s, kc)
            carry = b.add(t, carry)
    return out


def sha256_compress_circuit(expected_digest: bytes | None = None):
    """One SHA-256 compression: 512 message bits then 256 chaining bits as `Input`s (768 inputs, like the public
    Bristol-Fashion sha256.txt), 256 output wires.  With `expected_digest` the outputs are constrained the way
    SURVEY.md 8(f)1 prescribes (AddConst + AssertZero per output bit).

    Input order: word j of the block is inputs 32j..32j+31, most significant bit first (so the witness is just the
    block's bits in big-endian byte / MSB-first bit order, followed by the chaining words the same way).
    Returns (ops, n_wires, out_wires) with out_wires[32j + i] = bit (31-i) of output word j."""
    b = Builder(GF2)
    msg = []
    for _ in range(16):
        bits = [b.input() for _ in range(32)]  # MSB first
        msg.append(bits[::-1])  # -> LSB-first word
    st = []
    for _ in range(8):
        bits = [b.input() for _ in range(32)]
        st.append(bits[::-1])
    w = list(msg)
    for t in range(16, 64):
        x = w[t - 15]
        s0 = _xor_words(b, _rotr(x, 7), _rotr(x, 18), _shr(x, 3))
        x = w[t - 2]
        s1 = _xor_words(b, _rotr(x, 17), _rotr(x, 19), _shr(x, 10))
        w.append(_adder(b, _adder(b, _adder(b, w[t - 16], s0), w[t - 7]), s1))
    a, bb, c, d, e, f, g, h = st
    for t in range(64):
        S1 = _xor_words(b, _rotr(e, 6), _rotr(e, 11), _rotr(e, 25))
        ch = [b.add(b.mul(b.add(f[i], g[i]), e[i]), g[i]) for i in range(32)]
        t1 = _adder(b, _adder(b, _adder(b, h, S1), ch), _adder_const(b, w[t], _K[t]))
        S0 = _xor_words(b, _rotr(a, 2), _rotr(a, 13), _rotr(a, 22))
        mj = [b.add(b.mul(b.add(a[i], bb[i]), b.add(a[i], c[i])), a[i]) for i in range(32)]
        t2 = _adder(b, S0, mj)
        h, g, f, e, d, c, bb, a = g, f, e, _adder(b, d, t1), c, bb, a, _adder(b, t1, t2)
    fin = [_adder(b, x, y) for x, y in zip(st, (a, bb, c, d, e, f, g, h))]
    out_wires = [fin[j][31 - i] for j in range(8) for i in range(32)]
    if expected_digest is not None:
        assert len(expected_digest) == 32
        for k, wv in enumerate(out_wires):
            bit = (expected_digest[k // 8] >> (7 - k % 8)) & 1
            b.assert_zero(b.addc(wv, bit))
    return b.ops(), b.n_wires, out_wires


def sha256_witness(block: bytes, chaining: Sequence[int] = SHA256_IV) -> np.ndarray:
    """Witness bits for `sha256_compress_circuit`: the 64-byte block then the 8 chaining words, MSB first."""
    assert len(block) == 64
    raw = block + b"".join(struct.pack(">I", x) for x in chaining)
    return np.unpackbits(np.frombuffer(raw, dtype=np.uint8)).astype(np.uint8)


def sha256_pad_single_block(msg: bytes) -> bytes:
    assert len(msg) <= 55
    return msg + b"\x80" + b"\x00" * (55 - len(msg)) + struct.pack(">Q", 8 * len(msg))


def sha256_abc_case():
    """witness = padded block of "abc" || IV; outputs pinned to SHA-256("abc")."""
    digest = hashlib.sha256(b"abc").digest()
    ops, n_wires, _ = sha256_compress_circuit(digest)
    return ops, sha256_witness(sha256_pad_single_block(b"abc")), (0, n_wires)


# ---------------------------------------------------------------------------------------------------------------------
#  AES-128: generated Bristol-style circuit, 6400 AND gates = 200 S-boxes x 32 (Boyar-Peralta)
# ---------------------------------------------------------------------------------------------------------------------
# The 113-gate Boyar-Peralta S-box netlist (depth 27, 32 AND): the same netlist the kernels use for their key schedules
# (csrc/rv_aes_bs.cuh: bs_sbox).  ('in', k) = input bit k (LSB = 0); outputs s0..s7 = output bits 7..0.
_SBOX_NETLIST = (
    ('x0', 'in', 7), ('x1', 'in', 6), ('x2', 'in', 5), ('x3', 'in', 4), ('x4', 'in', 3), ('x5', 'in', 2),
    ('x6', 'in', 1), ('x7', 'in', 0), ('y14', 'xor', 'x3', 'x5'), ('y13', 'xor', 'x0', 'x6'),
    ('y9', 'xor', 'x0', 'x3'), ('y8', 'xor', 'x0', 'x5'), ('t0', 'xor', 'x1', 'x2'), ('y1', 'xor', 't0', 'x7'),
    ('y4', 'xor', 'y1', 'x3'), ('y12', 'xor', 'y13', 'y14'), ('y2', 'xor', 'y1', 'x0'), ('y5', 'xor', 'y1', 'x6'),
    ('y3', 'xor', 'y5', 'y8'), ('t1', 'xor', 'x4', 'y12'), ('y15', 'xor', 't1', 'x5'), ('y20', 'xor', 't1', 'x1'),
    ('y6', 'xor', 'y15', 'x7'), ('y10', 'xor', 'y15', 't0'), ('y11', 'xor', 'y20', 'y9'), ('y7', 'xor', 'x7', 'y11'),
    ('y17', 'xor', 'y10', 'y11'), ('y19', 'xor', 'y10', 'y8'), ('y16', 'xor', 't0', 'y11'),
    ('y21', 'xor', 'y13', 'y16'), ('y18', 'xor', 'x0', 'y16'), ('t2', 'and', 'y12', 'y15'),
    ('t3', 'and', 'y3', 'y6'), ('t4', 'xor', 't3', 't2'), ('t5', 'and', 'y4', 'x7'), ('t6', 'xor', 't5', 't2'),
    ('t7', 'and', 'y13', 'y16'), ('t8', 'and', 'y5', 'y1'), ('t9', 'xor', 't8', 't7'), ('t10', 'and', 'y2', 'y7'),
    ('t11', 'xor', 't10', 't7'), ('t12', 'and', 'y9', 'y11'), ('t13', 'and', 'y14', 'y17'),
    ('t14', 'xor', 't13', 't12'), ('t15', 'and', 'y8', 'y10'), ('t16', 'xor', 't15', 't12'),
    ('t17', 'xor', 't4', 't14'), ('t18', 'xor', 't6', 't16'), ('t19', 'xor', 't9', 't14'),
    ('t20', 'xor', 't11', 't16'), ('t21', 'xor', 't17', 'y20'), ('t22', 'xor', 't18', 'y19'),
    ('t23', 'xor', 't19', 'y21'), ('t24', 'xor', 't20', 'y18'), ('t25', 'xor', 't21', 't22'),
    ('t26', 'and', 't21', 't23'), ('t27', 'xor', 't24', 't26'), ('t28', 'and', 't25', 't27'),
    ('t29', 'xor', 't28', 't22'), ('t30', 'xor', 't23', 't24'), ('t31', 'xor', 't22', 't26'),
    ('t32', 'and', 't31', 't30'), ('t33', 'xor', 't32', 't24'), ('t34', 'xor', 't23', 't33'),
    ('t35', 'xor', 't27', 't33'), ('t36', 'and', 't24', 't35'), ('t37', 'xor', 't36', 't34'),
    ('t38', 'xor', 't27', 't36'), ('t39', 'and', 't29', 't38'), ('t40', 'xor', 't25', 't39'),
    ('t41', 'xor', 't40', 't37'), ('t42', 'xor', 't29', 't33'), ('t43', 'xor', 't29', 't40'),
    ('t44', 'xor', 't33', 't37'), ('t45', 'xor', 't42', 't41'), ('z0', 'and', 't44', 'y15'),
    ('z1', 'and', 't37', 'y6'), ('z2', 'and', 't33', 'x7'), ('z3', 'and', 't43', 'y16'), ('z4', 'and', 't40', 'y1'),
    ('z5', 'and', 't29', 'y7'), ('z6', 'and', 't42', 'y11'), ('z7', 'and', 't45', 'y17'),
    ('z8', 'and', 't41', 'y10'), ('z9', 'and', 't44', 'y12'), ('z10', 'and', 't37', 'y3'),
    ('z11', 'and', 't33', 'y4'), ('z12', 'and', 't43', 'y13'), ('z13', 'and', 't40', 'y5'),
    ('z14', 'and', 't29', 'y2'), ('z15', 'and', 't42', 'y9'), ('z16', 'and', 't45', 'y14'),
    ('z17', 'and', 't41', 'y8'), ('t46', 'xor', 'z15', 'z16'), ('t47', 'xor', 'z10', 'z11'),
    ('t48', 'xor', 'z5', 'z13'), ('t49', 'xor', 'z9', 'z10'), ('t50', 'xor', 'z2', 'z12'),
    ('t51', 'xor', 'z2', 'z5'), ('t52', 'xor', 'z7', 'z8'), ('t53', 'xor', 'z0', 'z3'), ('t54', 'xor', 'z6', 'z7'),
    ('t55', 'xor', 'z16', 'z17'), ('t56', 'xor', 'z12', 't48'), ('t57', 'xor', 't50', 't53'),
    ('t58', 'xor', 'z4', 't46'), ('t59', 'xor', 'z3', 't54'), ('t60', 'xor', 't46', 't57'),
    ('t61', 'xor', 'z14', 't57'), ('t62', 'xor', 't52', 't58'), ('t63', 'xor', 't49', 't58'),
    ('t64', 'xor', 'z4', 't59'), ('t65', 'xor', 't61', 't62'), ('t66', 'xor', 'z1', 't63'),
    ('s0', 'xor', 't59', 't63'), ('s6', 'xnor', 't56', 't62'), ('s7', 'xnor', 't48', 't60'),
    ('t67', 'xor', 't64', 't65'), ('s3', 'xor', 't53', 't66'), ('s4', 'xor', 't51', 't66'),
    ('s5', 'xor', 't47', 't65'), ('s1', 'xnor', 't64', 's3'), ('s2', 'xnor', 't55', 't67'),
)
_SBOX_OUT = ("s7", "s6", "s5", "s4", "s3", "s2", "s1", "s0")  # output bit 0 (LSB) .. bit 7


def _sbox(b: Builder, x: Sequence[int]) -> List[int]:
    """x: 8 wires, LSB first -> 8 wires of S[x]."""
    v = {}
    for name, kind, *args in _SBOX_NETLIST:
        if kind == "in":
            v[name] = x[args[0]]
        elif kind == "and":
            v[name] = b.mul(v[args[0]], v[args[1]])
        else:
            w = v[args[0]]
            for a in args[1:]:
                w = b.add(w, v[a])
            v[name] = b.addc(w, 1) if kind == "xnor" else w
    return [v[n] for n in _SBOX_OUT]


def _xtime(b: Builder, x: Sequence[int]) -> List[int]:
    """multiplication by 2 in GF(2^8) mod x^8+x^4+x^3+x+1, on 8 wires (LSB first): linear."""
    return [x[7], b.add(x[0], x[7]), x[1], b.add(x[2], x[7]), b.add(x[3], x[7]), x[4], x[5], x[6]]


def _xor_bytes(b: Builder, x, y):
    return [b.add(p, q) for p, q in zip(x, y)]


def aes128_circuit(expected_ciphertext: bytes | None = None):
    """AES-128 encryption of one block.  Inputs (256 `Input` ops): key bytes 0..15 then plaintext bytes 0..15, each byte
    LSB first.  With `expected_ciphertext` every output bit gets AddConst + AssertZero (the statement "I know a key that
    maps this plaintext to this ciphertext" when the plaintext is public ... here both are witness, as in SURVEY config 1).
    Returns (ops, n_wires, out_wires) with out_wires = ciphertext bytes 0..15, LSB first."""
    b = Builder()
    key = [[b.input() for _ in range(8)] for _ in range(16)]
    st = [[b.input() for _ in range(8)] for _ in range(16)]
    rcon = [0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80, 0x1B, 0x36]
    rk = [list(key)]
    for r in range(10):  # FIPS-197 5.2: w[i] = w[i-4] ^ (SubWord(RotWord(w[i-1])) ^ Rcon) for i % 4 == 0
        prev = rk[-1]
        t = [_sbox(b, prev[12 + ((k + 1) & 3)]) for k in range(4)]
        t[0] = [b.addc(w, 1) if (rcon[r] >> i) & 1 else w for i, w in enumerate(t[0])]
        nxt = [None] * 16
        for k in range(4):
            nxt[k] = _xor_bytes(b, prev[k], t[k])
        for c in range(1, 4):
            for k in range(4):
                nxt[4 * c + k] = _xor_bytes(b, prev[4 * c + k], nxt[4 * (c - 1) + k])
        rk.append(nxt)
    st = [_xor_bytes(b, st[i], rk[0][i]) for i in range(16)]
    for r in range(1, 11):
        sb = [_sbox(b, st[i]) for i in range(16)]
        sh = [sb[4 * ((c + rr) & 3) + rr] for c in range(4) for rr in range(4)]  # ShiftRows: state[row + 4 col]
        if r < 10:
            mc = [None] * 16
            for c in range(4):
                a = sh[4 * c : 4 * c + 4]
                for rr in range(4):  # out_r = xtime(a_r ^ a_{r+1}) ^ a_{r+1} ^ a_{r+2} ^ a_{r+3}
                    x = _xtime(b, _xor_bytes(b, a[rr], a[(rr + 1) & 3]))
                    mc[4 * c + rr] = _xor_bytes(b, _xor_bytes(b, x, a[(rr + 1) & 3]), _xor_bytes(b, a[(rr + 2) & 3], a[(rr + 3) & 3]))
            sh = mc
        st = [_xor_bytes(b, sh[i], rk[r][i]) for i in range(16)]
    out_wires = [w for byte in st for w in byte]
    if expected_ciphertext is not None:
        for i, w in enumerate(out_wires):
            b.assert_zero(b.addc(w, (expected_ciphertext[i // 8] >> (i % 8)) & 1))
    return b.ops(), b.n_wires, out_wires


def aes128_witness(key: bytes, pt: bytes) -> np.ndarray:
    return np.array([(byte >> i) & 1 for byte in key + pt for i in range(8)], dtype=np.uint8)


def aes128_fips197_case():
    """FIPS-197 Appendix C.1: key 00..0f, plaintext 00 11 .. ff, ciphertext 69c4e0d86a7b0430d8cdb78070b4c55a."""
    key, pt = bytes(range(16)), bytes(0x11 * i for i in range(16))
    ct = bytes.fromhex("69c4e0d86a7b0430d8cdb78070b4c55a")
    ops, nw, _ = aes128_circuit(ct)
    return ops, aes128_witness(key, pt), (0, nw)


# ---------------------------------------------------------------------------------------------------------------------
#  Bristol-Fashion text format  ->  ops     
# ---------------------------------------------------------------------------------------------------------------------
def parse_bristol_fashion(text: str, expected_outputs: Sequence[int] | None = None):
    """Bristol-Fashion: line 1 `ngates nwires`, line 2 `niv n_1..n_niv`, line 3 `nov m_1..m_nov`, then gates
    `2 1 a b o XOR|AND`, `1 1 a o INV|EQW`, `1 1 c o EQ` (constant).  Inputs are wires 0.. in order and become `Input`
    ops; outputs are the last sum(m) wires.  XOR->Add, AND->Mul, INV->AddConst(1), EQW->AddConst(0), EQ->Const.
    `expected_outputs` (bits, in output-wire order) appends AddConst+AssertZero per output.
    Returns (ops, n_wires, out_wires)."""
    toks = [ln.split() for ln in text.splitlines() if ln.strip()]
    ngates, nwires = int(toks[0][0]), int(toks[0][1])
    n_in = sum(int(x) for x in toks[1][1 : 1 + int(toks[1][0])])
    n_out = sum(int(x) for x in toks[2][1 : 1 + int(toks[2][0])])
    recs = [(GF2, INPUT, 0, w, 0, 0, 0) for w in range(n_in)]
    gates = toks[3 : 3 + ngates]
    if len(gates) != ngates:
        raise ValueError("bristol: gate count mismatch")
    for g in gates:
        kind = g[-1]
        if kind in ("XOR", "AND"):
            a, bb, o = int(g[2]), int(g[3]), int(g[4])
            recs.append((GF2, ADD if kind == "XOR" else MUL, 0, o, a, bb, 0))
        elif kind in ("INV", "NOT"):
            recs.append((GF2, ADDC, 0, int(g[3]), int(g[2]), 0, 1))
        elif kind == "EQW":
            recs.append((GF2, ADDC, 0, int(g[3]), int(g[2]), 0, 0))
        elif kind == "EQ":
            recs.append((GF2, CONST, 0, int(g[3]), 0, 0, int(g[2])))
        else:
            raise ValueError(f"bristol: unsupported gate {kind}")
    out_wires = list(range(nwires - n_out, nwires))
    n_total = nwires
    if expected_outputs is not None:
        if len(expected_outputs) != n_out:
            raise ValueError("bristol: expected_outputs length mismatch")
        for wv, bit in zip(out_wires, expected_outputs):
            recs.append((GF2, ADDC, 0, n_total, wv, 0, int(bit) & 1))
            recs.append((GF2, ASSERT_ZERO, 0, 0, n_total, 0, 0))
            n_total += 1
    return np.array(recs, dtype=OP_DTYPE), n_total, out_wires


def to_bristol_fashion(ops: np.ndarray, n_inputs_groups: Sequence[int], out_wires: Sequence[int]) -> str:
    """Inverse of the parser for pure XOR/AND/INV op lists whose outputs are the last wires (round-trip tests)."""
    lines = []
    nw = 0
    for o in ops:
        oc = int(o["opcode"])
        nw = max(nw, int(o["dst"]) + 1, int(o["a"]) + 1, int(o["b"]) + 1)
        if oc == INPUT:
            continue
        if oc == ADD:
            lines.append(f"2 1 {o['a']} {o['b']} {o['dst']} XOR")
        elif oc == MUL:
            lines.append(f"2 1 {o['a']} {o['b']} {o['dst']} AND")
        elif oc == ADDC:
            lines.append(f"1 1 {o['a']} {o['dst']} {'INV' if int(o['imm']) & 1 else 'EQW'}")
        else:
            raise ValueError("not expressible in Bristol-Fashion")
    head = [f"{len(lines)} {nw}", f"{len(n_inputs_groups)} " + " ".join(map(str, n_inputs_groups)), f"1 {len(out_wires)}", ""]
    return "\n".join(head + lines) + "\n"


# ---------------------------------------------------------------------------------------------------------------------
#  Synthetic circuits (src/proof/mod.rs:322-329)
# ---------------------------------------------------------------------------------------------------------------------
def flat_mul_circuit(n_mul: int, domain: int = GF2) -> Tuple[np.ndarray, Tuple[int, int]]:
    """`bench_prover`'s circuit scaled: Input(0), Input(1), n x Mul(2,0,1)   (src/proof/mod.rs:322-329)."""
    ops = np.zeros(2 + n_mul, dtype=OP_DTYPE)
    ops["domain"] = domain
    ops["opcode"][:2] = INPUT
    ops["dst"][0], ops["dst"][1] = 0, 1
    ops["opcode"][2:] = MUL
    ops["dst"][2:], ops["a"][2:], ops["b"][2:] = 2, 0, 1
    return ops, ((128, 128))


def layered_and_circuit(width: int, n_and: int, seed: int = 1) -> Tuple[np.ndarray, int]:
    """`width` inputs, then layers of `width` ANDs (last layer partial) whose operands are
    drawn from the previous two layers; SSA wires.  Returns (ops, n_wires)."""
    rng = np.random.default_rng(seed)
    ops = np.zeros(width + n_and, dtype=OP_DTYPE)
    ops["domain"] = GF2
    ops["opcode"][:width] = INPUT
    ops["dst"][:width] = np.arange(width, dtype=np.uint32)
    ops["opcode"][width:] = MUL
    ops["dst"][width:] = np.arange(width, width + n_and, dtype=np.uint32)
    done = 0
    layer = 0
    while done < n_and:
        n = min(width, n_and - done)
        lo = max(0, layer * width - width)  # previous two layers (layer 0 = the inputs)
        hi = (layer + 1) * width
        sl = slice(width + done, width + done + n)
        ops["a"][sl] = rng.integers(lo, hi, size=n, dtype=np.uint32)
        ops["b"][sl] = rng.integers(lo, hi, size=n, dtype=np.uint32)
        done += n
        layer += 1
    return ops, width + n_and


def z64_mul_circuit(n_mul: int) -> Tuple[np.ndarray, int]:
    """Input(0), Input(1), then n x Mul(2+(i mod 1022), (7i) mod m, (13i+1) mod m),
    m = 2+min(i,1022)."""
    i = np.arange(n_mul, dtype=np.uint64)
    m = 2 + np.minimum(i, 1022)
    ops = np.zeros(2 + n_mul, dtype=OP_DTYPE)
    ops["domain"] = Z64
    ops["opcode"][:2] = INPUT
    ops["dst"][1] = 1
    ops["opcode"][2:] = MUL
    ops["dst"][2:] = (2 + (i % 1022)).astype(np.uint32)
    ops["a"][2:] = ((i * 7) % m).astype(np.uint32)
    ops["b"][2:] = ((i * 13 + 1) % m).astype(np.uint32)
    return ops, 1024


# ---------------------------------------------------------------------------------------------------------------------
#  .rvops program files: magic, (z64_cells, gf2_cells), then the packed 24-byte rv_op records
# ---------------------------------------------------------------------------------------------------------------------
RVOPS_MAGIC = b"RVOPS\x00\x01\x00"


def save_ops(path: str, ops: np.ndarray, wire_counts: Tuple[int, int]) -> None:
    with open(path, "wb") as f:
        f.write(RVOPS_MAGIC)
        f.write(struct.pack("<QQQ", int(wire_counts[0]), int(wire_counts[1]), len(ops)))
        f.write(np.ascontiguousarray(ops, dtype=OP_DTYPE).tobytes())


def load_ops(path: str) -> Tuple[np.ndarray, Tuple[int, int]]:
    with open(path, "rb") as f:
        if f.read(8) != RVOPS_MAGIC:
            raise ValueError("not an .rvops file")
        z, g, n = struct.unpack("<QQQ", f.read(24))
        ops = np.frombuffer(f.read(), dtype=OP_DTYPE)
    if len(ops) != n:
        raise ValueError("truncated .rvops file")
    return ops.copy(), (int(z), int(g))


# ---------------------------------------------------------------------------------------------------------------------
#  helpers
# ---------------------------------------------------------------------------------------------------------------------
def largest_wires(ops: np.ndarray) -> Tuple[int, int]:
    """mcircuit::largest_wires as consumed at src/proof/mod.rs:125 -> (z64_cells, gf2_cells)."""
    g = z = 0
    for dom, cells in ((GF2, "g"), (Z64, "z")):
        sel = ops[ops["domain"] == dom]
        if sel.size == 0:
            continue
        oc = sel["opcode"]
        m = 0
        uses_dst = oc != ASSERT_ZERO
        if uses_dst.any():
            m = max(m, int(sel["dst"][uses_dst].max()) + 1)
        uses_a = np.isin(oc, (ADD, SUB, MUL, ADDC, SUBC, MULC, ASSERT_ZERO))
        if uses_a.any():
            m = max(m, int(sel["a"][uses_a].max()) + 1)
        uses_b = np.isin(oc, (ADD, SUB, MUL))
        if uses_b.any():
            m = max(m, int(sel["b"][uses_b].max()) + 1)
        if dom == GF2:
            g = m
        else:
            z = m
    b2a = ops[ops["domain"] == B2A]
    if b2a.size:
        z = max(z, int(b2a["dst"].max()) + 1)
        g = max(g, int(b2a["a"].max()) + 64)
    hint = ops[ops["domain"] == HINT]
    if hint.size:
        z = max(z, int(hint["a"].max()))
        g = max(g, int(hint["b"].max()))
    return z, g


def evaluate_gf2(ops: np.ndarray, witness: Iterable[int], n_wires: int) -> Tuple[np.ndarray, bool]:
    """Plaintext evaluation of the GF(2) ops (what mcircuit::evaluate_composite_program does for this domain).
    Returns (wire values, all AssertZero satisfied)."""
    w = np.zeros(n_wires, dtype=np.uint8)
    wit = iter(witness)
    ok = True
    for dom, oc, _, dst, a, bb, imm in ops.tolist():
        if dom != GF2:
            continue
        if oc == INPUT:
            w[dst] = next(wit) & 1
        elif oc in (ADD, SUB):
            w[dst] = w[a] ^ w[bb]
        elif oc == MUL:
            w[dst] = w[a] & w[bb]
        elif oc in (ADDC, SUBC):
            w[dst] = w[a] ^ (imm & 1)
        elif oc == MULC:
            w[dst] = w[a] & (imm & 1)
        elif oc == ASSERT_ZERO:
            ok &= w[a] == 0
        elif oc == CONST:
            w[dst] = imm & 1
        elif oc == RANDOM:
            w[dst] = 0
    return w, bool(ok)
