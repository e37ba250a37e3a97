// rv_aes_bs.cuh -- AES-128 building blocks of the mask generators.
//
// Replaces (reference, paths relative to its root):
//   PRG::new / PRG::gen                 src/crypto/prg.rs:13-37   (ctr::Ctr128BE<Aes128>, IV = 0, keystream only)
//   expand_seed                         src/transcript/mod.rs:99-106
//   BatchGen::gen + batches_to_shares   src/generator/batch.rs:30-40, src/algebra/gf2/domain.rs:66-173,293-378
//
// Three forms of the same cipher live here, all __host__ __device__ so tests can run them on the CPU against FIPS-197:
//   * tt_aes128_encrypt  -- T-table rounds, one block per thread: what the GPU generators run (k_mask_gen_tt, k_zmask_gen_tt),
//                           with the tables replicated per shared-memory bank;
//   * aes128_expand_key / aes128_encrypt_block -- scalar, S-box from the 113-gate Boyar-Peralta netlist on 4 packed bytes:
//                           seed expansion and key schedules (k_key_setup);
//   * bs_aes128_ctr_block -- bitsliced ACROSS PRG STREAMS (one u32 = the same state bit of 32 streams, so the reference's
//                           64 x 128 bit transpose costs nothing).  This was the first GPU generator (23 G blocks/s, ALU-pipe
//                           bound); tests/hostsim keeps it as an independent replay of the share-tensor layout.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RV_HD __host__ __device__ __forceinline__
#else
#define RV_HD inline
#endif

namespace rv {

// Boyar-Peralta 113-gate S-box (depth 27), inputs/outputs as bit-planes.  x[0] = bit 0 (LSB) ... x[7] = bit 7.
// `ONE` is the all-ones word of the plane type.  Checked exhaustively by tools/check_sbox.py and tests.
template <typename W>
RV_HD void bs_sbox(W *x, const W ONE) {
    const W x0 = x[7], x1 = x[6], x2 = x[5], x3 = x[4], x4 = x[3], x5 = x[2], x6 = x[1], x7 = x[0];
    const W y14 = x3 ^ x5, y13 = x0 ^ x6, y9 = x0 ^ x3, y8 = x0 ^ x5, t0 = x1 ^ x2, y1 = t0 ^ x7, y4 = y1 ^ x3;
    const W y12 = y13 ^ y14, y2 = y1 ^ x0, y5 = y1 ^ x6, y3 = y5 ^ y8, t1 = x4 ^ y12, y15 = t1 ^ x5, y20 = t1 ^ x1;
    const W y6 = y15 ^ x7, y10 = y15 ^ t0, y11 = y20 ^ y9, y7 = x7 ^ y11, y17 = y10 ^ y11, y19 = y10 ^ y8;
    const W y16 = t0 ^ y11, y21 = y13 ^ y16, y18 = x0 ^ y16;
    const W t2 = y12 & y15, t3 = y3 & y6, t4 = t3 ^ t2, t5 = y4 & x7, t6 = t5 ^ t2, t7 = y13 & y16, t8 = y5 & y1;
    const W t9 = t8 ^ t7, t10 = y2 & y7, t11 = t10 ^ t7, t12 = y9 & y11, t13 = y14 & y17, t14 = t13 ^ t12;
    const W t15 = y8 & y10, t16 = t15 ^ t12, t17 = t4 ^ t14, t18 = t6 ^ t16, t19 = t9 ^ t14, t20 = t11 ^ t16;
    const W t21 = t17 ^ y20, t22 = t18 ^ y19, t23 = t19 ^ y21, t24 = t20 ^ y18;
    const W t25 = t21 ^ t22, t26 = t21 & t23, t27 = t24 ^ t26, t28 = t25 & t27, t29 = t28 ^ t22, t30 = t23 ^ t24;
    const W t31 = t22 ^ t26, t32 = t31 & t30, t33 = t32 ^ t24, t34 = t23 ^ t33, t35 = t27 ^ t33, t36 = t24 & t35;
    const W t37 = t36 ^ t34, t38 = t27 ^ t36, t39 = t29 & t38, t40 = t25 ^ t39;
    const W t41 = t40 ^ t37, t42 = t29 ^ t33, t43 = t29 ^ t40, t44 = t33 ^ t37, t45 = t42 ^ t41;
    const W z0 = t44 & y15, z1 = t37 & y6, z2 = t33 & x7, z3 = t43 & y16, z4 = t40 & y1, z5 = t29 & y7;
    const W z6 = t42 & y11, z7 = t45 & y17, z8 = t41 & y10, z9 = t44 & y12, z10 = t37 & y3, z11 = t33 & y4;
    const W z12 = t43 & y13, z13 = t40 & y5, z14 = t29 & y2, z15 = t42 & y9, z16 = t45 & y14, z17 = t41 & y8;
    const W t46 = z15 ^ z16, t47 = z10 ^ z11, t48 = z5 ^ z13, t49 = z9 ^ z10, t50 = z2 ^ z12, t51 = z2 ^ z5;
    const W t52 = z7 ^ z8, t53 = z0 ^ z3, t54 = z6 ^ z7, t55 = z16 ^ z17, t56 = z12 ^ t48, t57 = t50 ^ t53;
    const W t58 = z4 ^ t46, t59 = z3 ^ t54, t60 = t46 ^ t57, t61 = z14 ^ t57, t62 = t52 ^ t58, t63 = t49 ^ t58;
    const W t64 = z4 ^ t59, t65 = t61 ^ t62, t66 = z1 ^ t63;
    const W s0 = t59 ^ t63, s6 = t56 ^ t62 ^ ONE, s7 = t48 ^ t60 ^ ONE, t67 = t64 ^ t65, s3 = t53 ^ t66;
    const W s4 = t51 ^ t66, s5 = t47 ^ t65, s1 = t64 ^ s3 ^ ONE, s2 = t55 ^ t67 ^ ONE;
    x[7] = s0; x[6] = s1; x[5] = s2; x[4] = s3; x[3] = s4; x[2] = s5; x[1] = s6; x[0] = s7;
}

// MixColumns on one column given as four byte-plane groups (8 planes each, LSB first); writes 32 planes to out.
template <typename W>
RV_HD void bs_mix_column(const W *a0, const W *a1, const W *a2, const W *a3, W *out) {
    const W *a[4] = {a0, a1, a2, a3};
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const W *p = a[r], *q = a[(r + 1) & 3], *u = a[(r + 2) & 3], *v = a[(r + 3) & 3];
        W t[8];
#pragma unroll
        for (int b = 0; b < 8; b++) t[b] = p[b] ^ q[b];
        // xtime(t): shift left, reduce by 0x1b on the carry plane t[7]
        W *o = out + 8 * r;
        o[0] = t[7] ^ q[0] ^ u[0] ^ v[0];
        o[1] = t[0] ^ t[7] ^ q[1] ^ u[1] ^ v[1];
        o[2] = t[1] ^ q[2] ^ u[2] ^ v[2];
        o[3] = t[2] ^ t[7] ^ q[3] ^ u[3] ^ v[3];
        o[4] = t[3] ^ t[7] ^ q[4] ^ u[4] ^ v[4];
        o[5] = t[4] ^ q[5] ^ u[5] ^ v[5];
        o[6] = t[5] ^ q[6] ^ u[6] ^ v[6];
        o[7] = t[6] ^ q[7] ^ u[7] ^ v[7];
    }
}

// Bitsliced AES-128 encryption of CTR block `ctr` (128-bit big-endian counter whose high 64 bits are zero: the
// reference's streams never reach 2^64 blocks) for 32 streams at once.
//   rk(round, k): bit-plane k (= 8*byte + bit, bit 0 = LSB) of round key `round` for this thread's 32 streams.
//   s[128]: output planes, same indexing.  State byte B = 4*column + row (FIPS-197 3.4) = keystream byte B.
template <typename RK>
RV_HD void bs_aes128_ctr_block(uint64_t ctr, RK rk, uint32_t *s) {
    // AddRoundKey(0) on the constant counter block: plane = key plane, complemented where the counter bit is set.
#pragma unroll
    for (int B = 0; B < 16; B++) {
#pragma unroll
        for (int b = 0; b < 8; b++) {
            uint32_t in = 0;
            if (B >= 8) in = (uint32_t)0 - (uint32_t)((ctr >> (8 * (15 - B) + b)) & 1);
            s[8 * B + b] = in ^ rk(0, 8 * B + b);
        }
    }
#pragma unroll 1
    for (int round = 1; round <= 10; round++) {
#pragma unroll
        for (int B = 0; B < 16; B++) bs_sbox<uint32_t>(s + 8 * B, 0xFFFFFFFFu);
        // ShiftRows: new byte (row r, col c) = old byte (row r, col c + r)
        uint32_t t[128];
#pragma unroll
        for (int c = 0; c < 4; c++)
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int b = 0; b < 8; b++) t[8 * (4 * c + r) + b] = s[8 * (4 * ((c + r) & 3) + r) + b];
        if (round < 10) {
#pragma unroll
            for (int c = 0; c < 4; c++) bs_mix_column<uint32_t>(t + 32 * c, t + 32 * c + 8, t + 32 * c + 16, t + 32 * c + 24, s + 32 * c);
        } else {
#pragma unroll
            for (int k = 0; k < 128; k++) s[k] = t[k];
        }
#pragma unroll
        for (int k = 0; k < 128; k++) s[k] ^= rk(round, k);
    }
}

// ---- scalar AES for the (tiny) seed expansion and key schedules: S-box via the same netlist, 4 bytes per u32 -------
RV_HD uint32_t sub_word(uint32_t w) {
    uint32_t x[8];
#pragma unroll
    for (int b = 0; b < 8; b++) x[b] = (w >> b) & 0x01010101u;
    bs_sbox<uint32_t>(x, 0x01010101u);
    uint32_t r = 0;
#pragma unroll
    for (int b = 0; b < 8; b++) r |= x[b] << b;
    return r;
}

struct NetlistSubWord {  // S-box from the netlist: no table, no memory
    RV_HD uint32_t operator()(uint32_t w) const { return sub_word(w); }
};
struct TableSubWord {  // S-box from a 256-byte table (shared memory on the device)
    const uint8_t *t;
    RV_HD uint32_t operator()(uint32_t w) const {
        return (uint32_t)t[w & 0xff] | ((uint32_t)t[(w >> 8) & 0xff] << 8) | ((uint32_t)t[(w >> 16) & 0xff] << 16) | ((uint32_t)t[w >> 24] << 24);
    }
};

// Key expansion (FIPS-197 5.2).  Words are little-endian loads of the key bytes: byte 0 of a word is its low byte.
template <typename SW = NetlistSubWord>
RV_HD void aes128_expand_key(const uint32_t key[4], uint32_t rk[44], SW sw = SW()) {
    const uint32_t rcon[10] = {0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80, 0x1b, 0x36};
    for (int i = 0; i < 4; i++) rk[i] = key[i];
    for (int i = 4; i < 44; i += 4) {
        uint32_t t = rk[i - 1];
        t = (t >> 8) | (t << 24);  // RotWord on a little-endian word
        t = sw(t) ^ rcon[i / 4 - 1];
        rk[i] = rk[i - 4] ^ t;
        rk[i + 1] = rk[i - 3] ^ rk[i];
        rk[i + 2] = rk[i - 2] ^ rk[i + 1];
        rk[i + 3] = rk[i - 1] ^ rk[i + 2];
    }
}

RV_HD uint32_t xtime4(uint32_t w) {  // xtime on 4 packed bytes
    const uint32_t hi = w & 0x80808080u;
    return ((w & 0x7f7f7f7fu) << 1) ^ ((hi >> 7) * 0x1bu);
}

// One AES-128 block, scalar.  Column c of the state is word c (row r in byte r).
template <typename SW = NetlistSubWord>
RV_HD void aes128_encrypt_block(const uint32_t rk[44], const uint32_t in[4], uint32_t out[4], SW sw = SW()) {
    uint32_t s[4] = {in[0] ^ rk[0], in[1] ^ rk[1], in[2] ^ rk[2], in[3] ^ rk[3]};
    for (int round = 1; round <= 10; round++) {
        uint32_t t[4];
        for (int c = 0; c < 4; c++) t[c] = sw(s[c]);
        uint32_t u[4];
        for (int c = 0; c < 4; c++)  // ShiftRows: row r of column c comes from column c + r
            u[c] = (t[c] & 0x000000ffu) | (t[(c + 1) & 3] & 0x0000ff00u) | (t[(c + 2) & 3] & 0x00ff0000u) | (t[(c + 3) & 3] & 0xff000000u);
        if (round < 10) {
            for (int c = 0; c < 4; c++) {  // MixColumns: out_r = xtime(a_r ^ a_{r+1}) ^ a_{r+1} ^ a_{r+2} ^ a_{r+3}
                const uint32_t a = u[c];
                const uint32_t r1 = (a >> 8) | (a << 24), r2 = (a >> 16) | (a << 16), r3 = (a >> 24) | (a << 8);
                u[c] = xtime4(a ^ r1) ^ r1 ^ r2 ^ r3;
            }
        }
        for (int c = 0; c < 4; c++) s[c] = u[c] ^ rk[4 * round + c];
    }
    for (int c = 0; c < 4; c++) out[c] = s[c];
}

// ---- T-table AES (one block per thread, natural byte order) for the Z64 mask generator -------------------------------
// Te0[x] = (2 s, s, s, 3 s) with s = S[x], row r of the column in byte r; Te_r = rotl(Te0, 8 r).  `tab(r, x)` returns Te_r[x]
// (on the device: a bank-private copy per lane, so the 32 lookups of a warp never conflict).
RV_HD uint32_t te0_entry(uint32_t sbox_byte) {
    const uint32_t s1 = sbox_byte & 0xff, s2 = ((s1 << 1) ^ ((s1 >> 7) * 0x1bu)) & 0xff, s3 = s2 ^ s1;
    return s2 | (s1 << 8) | (s1 << 16) | (s3 << 24);
}
template <typename TAB>  // tab(t, w, b) = Te_t[byte b of w]
RV_HD void tt_aes128_encrypt(const uint32_t rk[44], uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3, TAB tab, uint32_t out[4]) {
    s0 ^= rk[0];
    s1 ^= rk[1];
    s2 ^= rk[2];
    s3 ^= rk[3];
#pragma unroll
    for (int round = 1; round < 10; round++) {  // new column c = Te0[row 0 of col c] ^ Te1[row 1 of col c+1] ^ Te2[row 2 of col c+2] ^ Te3[row 3 of col c+3]
        const uint32_t t0 = tab(0, s0, 0) ^ tab(1, s1, 1) ^ tab(2, s2, 2) ^ tab(3, s3, 3) ^ rk[4 * round + 0];
        const uint32_t t1 = tab(0, s1, 0) ^ tab(1, s2, 1) ^ tab(2, s3, 2) ^ tab(3, s0, 3) ^ rk[4 * round + 1];
        const uint32_t t2 = tab(0, s2, 0) ^ tab(1, s3, 1) ^ tab(2, s0, 2) ^ tab(3, s1, 3) ^ rk[4 * round + 2];
        const uint32_t t3 = tab(0, s3, 0) ^ tab(1, s0, 1) ^ tab(2, s1, 2) ^ tab(3, s2, 3) ^ rk[4 * round + 3];
        s0 = t0;
        s1 = t1;
        s2 = t2;
        s3 = t3;
    }
    // last round: SubBytes + ShiftRows only; S[x] sits in byte 0 of Te2, byte 1 of Te3, byte 2 of Te0, byte 3 of Te1
    out[0] = ((tab(2, s0, 0) & 0x000000ffu) | (tab(3, s1, 1) & 0x0000ff00u) | (tab(0, s2, 2) & 0x00ff0000u) | (tab(1, s3, 3) & 0xff000000u)) ^ rk[40];
    out[1] = ((tab(2, s1, 0) & 0x000000ffu) | (tab(3, s2, 1) & 0x0000ff00u) | (tab(0, s3, 2) & 0x00ff0000u) | (tab(1, s0, 3) & 0xff000000u)) ^ rk[41];
    out[2] = ((tab(2, s2, 0) & 0x000000ffu) | (tab(3, s3, 1) & 0x0000ff00u) | (tab(0, s0, 2) & 0x00ff0000u) | (tab(1, s1, 3) & 0xff000000u)) ^ rk[42];
    out[3] = ((tab(2, s3, 0) & 0x000000ffu) | (tab(3, s0, 1) & 0x0000ff00u) | (tab(0, s1, 2) & 0x00ff0000u) | (tab(1, s2, 3) & 0xff000000u)) ^ rk[43];
}

// CTR block j of the reference's PRG as little-endian state words: bytes 8..15 = BE64(j).
RV_HD void ctr_block_words(uint64_t j, uint32_t in[4]) {
    const uint32_t hi = (uint32_t)(j >> 32), lo = (uint32_t)j;
    in[0] = 0;
    in[1] = 0;
    in[2] = ((hi >> 24) & 0xff) | ((hi >> 8) & 0xff00) | ((hi << 8) & 0xff0000) | (hi << 24);
    in[3] = ((lo >> 24) & 0xff) | ((lo >> 8) & 0xff00) | ((lo << 8) & 0xff0000) | (lo << 24);
}

}  // namespace rv
