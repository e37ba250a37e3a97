// rv_zplanes.cuh -- per-thread bodies of the Z64 kernels (__host__ __device__: tests/hostsim replays them on the CPU; the
// product only runs them inside the CUDA kernels of rv_z64.cu).
//
// Data layout in HBM for a shard of `npi` packed instances (nreps = 8 npi repetitions):
//   zrows   u64 [n_zrows][64 * npi]   Z64 share tensor: row = one mask (fresh PRG mask, linear node, or the zero row);
//                                     element 64*pi + 8*r + p = the reference's ShareZ64.pack[r][p] of instance pi
//                                     (src/algebra/z64/share.rs:10-13).  One (row, repetition) = 64 contiguous bytes.
//   zvals   u64 [n_zvals]             value plane (one per opened repetition in the verifier)
//   zon     u8  [nreps][pitch]        online hash stream of every repetition: 8 bytes per Input (LE corr,
//                                     src/algebra/z64/recon.rs:131-137), 64 bytes per Mul / AssertZero (8 players x LE u64,
//                                     src/algebra/z64/share.rs:100-108)
//   zpre    u8  [nreps][pitch]        preprocessing hash stream: 8 bytes per Mul
#pragma once
#include <stdint.h>

#include "rv_planes.cuh"

namespace rv {

// 32x32 bit-matrix transpose in registers (Hacker's Delight 7-3): new a[i] bit b = old a[31-b] bit 31-i.
RV_HD void transpose32(uint32_t a[32]) {
    uint32_t m = 0x0000FFFFu;
#pragma unroll
    for (int j = 16; j != 0; j >>= 1, m ^= (m << j)) {
#pragma unroll
        for (int k = 0; k < 32; k = (k + j + 1) & ~j) {
            const uint32_t t = (a[k] ^ (a[k | j] >> j)) & m;
            a[k] ^= t;
            a[k | j] ^= (t << j);
        }
    }
}

// Bitsliced AES planes of counter block j for one slice (plane 8B+b = keystream byte B, bit b; lane bit q = stream 31-q)
// -> Z64 mask 2j+h of each of the 32 streams as (lo[sigma], hi[sigma]), sigma = 8*(rep within slice) + player: the little-endian
// u64 at keystream byte 8h of the block (src/algebra/z64/batch.rs:25-30).  Streams whose bit is clear in lane_mask (the
// verifier's unopened player, src/generator/batch.rs:31-34) read as zero.
RV_HD void planes_to_mask_words(const uint32_t s[128], uint32_t lane_mask, int h, uint32_t lo[32], uint32_t hi[32]) {
#pragma unroll
    for (int k = 0; k < 32; k++) {
        lo[k] = s[64 * h + 31 - k] & lane_mask;
        hi[k] = s[64 * h + 32 + 31 - k] & lane_mask;
    }
    transpose32(lo);
    transpose32(hi);
}

// element index of (slice w, stream sigma) inside a zrow
RV_HD uint32_t zrow_index(uint32_t w, uint32_t sigma) { return 64 * (w >> 1) + ((w & 1) ? 0u : 32u) + sigma; }

// one (row, repetition) segment = 8 players x u64 = 64 bytes, 64-byte aligned: four 16-byte loads
struct alignas(16) ZPair {
    uint64_t x, y;
};
RV_HD void zload8(const uint64_t *p, uint64_t v[8]) {
    const ZPair *q = reinterpret_cast<const ZPair *>(p);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const ZPair t = q[i];
        v[2 * i] = t.x;
        v[2 * i + 1] = t.y;
    }
}
RV_HD uint64_t zsum8v(const uint64_t v[8]) { return v[0] + v[1] + v[2] + v[3] + v[4] + v[5] + v[6] + v[7]; }
RV_HD uint64_t zsum8(const uint64_t *p) {
    uint64_t v[8];
    zload8(p, v);
    return zsum8v(v);
}
RV_HD void put64(uint8_t *p, uint64_t v) { *reinterpret_cast<uint64_t *>(p) = v; }  // all stream offsets are multiples of 8
RV_HD uint64_t get64_unaligned(const uint8_t *p) {
    uint64_t v = 0;
    for (int i = 0; i < 8; i++) v |= (uint64_t)p[i] << (8 * i);
    return v;
}

// ---- prover ---------------------------------------------------------------------------------------------------------
// One online item of repetition `rep` (index inside the shard): src/interpreter/single.rs:25-69,140-147,
// src/transcript/prover.rs:181-232.  `stream` = this repetition's online stream.
RV_HD uint64_t z_pre_word(const ZItem &it, const uint64_t *zrows, size_t rowlen, uint32_t rep, const uint64_t *grows, uint32_t npi);
// `pre` (optional) = this repetition's preprocessing stream: the prover fills both streams in one pass over the operands.
RV_HD void z_prover_online(const ZItem &it, const uint64_t *zrows, size_t rowlen, uint32_t rep, const uint64_t *vals, uint8_t *stream, int *bad,
                           uint8_t *pre = nullptr, const uint64_t *grows = nullptr, uint32_t npi = 0) {
    if (it.kind == ITEM_B2A) {  // a conversion only appends to the preprocessing stream
        if (pre) put64(pre + 8ull * it.j, z_pre_word(it, zrows, rowlen, rep, grows, npi));
        return;
    }
    const uint64_t *A = zrows + (size_t)it.ra * rowlen + 8 * rep;
    uint8_t *dst = stream + it.off;
    if (it.kind == ITEM_INPUT) {
        put64(dst, vals[it.va] - zsum8(A));  // corr = w - reconstruct(mask), prover.rs:186-195
        return;
    }
    uint64_t a[8];
    zload8(A, a);
#pragma unroll
    for (int p = 0; p < 8; p++) a[p] *= it.ca;
    if (it.kind == ITEM_ASSERT) {
        if (vals[it.va] != 0) *bad |= 1;
#pragma unroll
        for (int p = 0; p < 8; p++) put64(dst + 8 * p, a[p]);
        return;
    }
    const uint64_t *B = zrows + (size_t)it.rb * rowlen + 8 * rep;
    const uint64_t *AB = zrows + (size_t)it.k * rowlen + 8 * rep, *NW = zrows + (size_t)(it.k + 1) * rowlen + 8 * rep;
    uint64_t b[8], ab[8], nw[8];
    zload8(B, b);
    zload8(AB, ab);
    zload8(NW, nw);
#pragma unroll
    for (int p = 0; p < 8; p++) b[p] *= it.cb;
    const uint64_t sa = zsum8v(a), sb = zsum8v(b);
    const uint64_t c1 = vals[it.va] - sa, c2 = vals[it.vb] - sb;  // corr = value - reconstruct(mask)
#pragma unroll
    for (int p = 0; p < 8; p++) put64(dst + 8 * p, b[p] * c1 + a[p] * c2 + ab[p] - nw[p]);  // single.rs:41-45
    if (pre) put64(pre + 8ull * it.j, sa * sb - zsum8v(ab));                                   // delta = a * b - c, single.rs:35-39
}

// B2A: the per-repetition plaintext of the 64 fresh GF(2) wires g0 .. g0+63 (corr = 0: value = parity of the mask shares),
// wire i = bit i (src/interpreter/combine.rs:19-36).  grows = the GF(2) share tensor [row][npi].
RV_HD uint64_t b2a_random_value(const uint64_t *grows, uint32_t npi, uint32_t g0, uint32_t rep) {
    const uint32_t pi = rep >> 3, sh = 8 * (7 - (rep & 7));
    uint64_t r = 0;
    for (uint32_t i = 0; i < 64; i++) r |= ((gf2_reconstruct(grows[(size_t)(g0 + i) * npi + pi]) >> sh) & 1ull) << i;
    return r;
}
// Mul: delta = a * b - c on reconstructed masks (single.rs:35-39).  B2A: r - reconstruct(z64_mask) (combine.rs:150-158).
RV_HD uint64_t z_pre_word(const ZItem &it, const uint64_t *zrows, size_t rowlen, uint32_t rep, const uint64_t *grows, uint32_t npi) {
    if (it.kind == ITEM_B2A) return b2a_random_value(grows, npi, it.ra, rep) - zsum8(zrows + (size_t)it.k * rowlen + 8 * rep);
    const uint64_t a = it.ca * zsum8(zrows + (size_t)it.ra * rowlen + 8 * rep), b = it.cb * zsum8(zrows + (size_t)it.rb * rowlen + 8 * rep);
    return a * b - zsum8(zrows + (size_t)it.k * rowlen + 8 * rep);
}

// ---- online verifier (src/transcript/verifier/online.rs; unpacking src/algebra/z64/share.rs:51-91, recon.rs:68-107) ----
// n_* = element counts taken from the FIRST repetition of the pack of 8; len_* = this repetition's own byte lengths: an
// element past n reads as the default (zero), a missing 8-byte chunk inside n reads as zero.
struct ZOpen {
    uint64_t off_recons, off_corrs, off_inputs;
    uint32_t n_recons, n_corrs, n_inputs;
    uint32_t len_recons, len_corrs, len_inputs;
    uint32_t omit, pad;
};
RV_HD uint64_t z_packed(const uint8_t *proof, uint64_t off, uint32_t n, uint32_t len, uint32_t e) {
    return (e < n && 8ull * e + 8 <= len) ? get64_unaligned(proof + off + 8ull * e) : 0ull;
}

// u-plane leaves of opened repetition `slot`: u = corr + rho, rho = sum of the opened players' mask shares
RV_HD uint64_t z_verify_leaf_input(const ZItem &it, uint32_t k, const ZOpen &o, const uint8_t *proof, const uint64_t *zrows, size_t rowlen, uint32_t slot) {
    return z_packed(proof, o.off_inputs, o.n_inputs, o.len_inputs, k) + zsum8(zrows + (size_t)it.ra * rowlen + 8 * slot);
}
// kappa = rho_ab - rho_a * rho_b + msg + delta
RV_HD uint64_t z_verify_leaf_kappa(const ZItem &it, uint32_t recon_idx, const ZOpen &o, const uint8_t *proof, const uint64_t *zrows, size_t rowlen,
                                   uint32_t slot) {
    const uint64_t ra = it.ca * zsum8(zrows + (size_t)it.ra * rowlen + 8 * slot), rb = it.cb * zsum8(zrows + (size_t)it.rb * rowlen + 8 * slot);
    const uint64_t rab = zsum8(zrows + (size_t)it.k * rowlen + 8 * slot);
    return rab - ra * rb + z_packed(proof, o.off_recons, o.n_recons, o.len_recons, recon_idx) + z_packed(proof, o.off_corrs, o.n_corrs, o.len_corrs, it.j);
}

// B2A output wire of an opened repetition: mask = -z64_mask, corr = recon - corr_j with recon = the 64 reconstructed result bits
// (u_i ^ msg_i, as in AssertZero) and corr_j from the proof, so u = corr + rho(mask) (src/interpreter/combine.rs:204-218).
//   guv: the repetition's GF(2) u-plane; go: its GF(2) opening; urefs: the 64 result wires' u-plane refs
RV_HD uint64_t z_verify_leaf_b2a(const ZItem &it, const ZOpen &o, const VOpen &go, const uint8_t *proof, const uint64_t *zrows, size_t rowlen,
                                 uint32_t slot, const uint8_t *guv, const uint32_t *urefs) {
    uint64_t recon = 0;
    for (uint32_t i = 0; i < 64; i++)
        recon |= (uint64_t)(val_of(guv, urefs[64 * it.va + i]) ^ packed_bit(proof + go.off_recons, go.eff_recons, it.vb + i)) << i;
    return recon - z_packed(proof, o.off_corrs, o.n_corrs, o.len_corrs, it.j) - zsum8(zrows + (size_t)it.k * rowlen + 8 * slot);
}

RV_HD void z_verify_online(const ZItem &it, uint32_t recon_idx, const ZOpen &o, const uint8_t *proof, const uint64_t *zrows, size_t rowlen, uint32_t slot,
                           const uint64_t *uvals, uint8_t *stream, int *not_okay) {
    if (it.kind == ITEM_B2A) return;
    uint8_t *dst = stream + it.off;
    if (it.kind == ITEM_INPUT) {  // the masked input from the proof is hashed as is (online.rs:123-130)
        put64(dst, z_packed(proof, o.off_inputs, o.n_inputs, o.len_inputs, it.j));
        return;
    }
    const uint64_t msg = z_packed(proof, o.off_recons, o.n_recons, o.len_recons, recon_idx);  // the unopened player's broadcast (online.rs:140-160)
    const uint64_t *A = zrows + (size_t)it.ra * rowlen + 8 * slot;
    uint64_t a[8], s[8];
    zload8(A, a);
#pragma unroll
    for (int p = 0; p < 8; p++) a[p] *= it.ca;
    if (it.kind == ITEM_ASSERT) {
#pragma unroll
        for (int p = 0; p < 8; p++) s[p] = a[p];
        if (uvals[it.va] + msg != 0) *not_okay |= 1;  // corr + reconstruct(mask + msg) = u + msg
    } else {
        const uint64_t *B = zrows + (size_t)it.rb * rowlen + 8 * slot;
        const uint64_t *AB = zrows + (size_t)it.k * rowlen + 8 * slot, *NW = zrows + (size_t)(it.k + 1) * rowlen + 8 * slot;
        uint64_t b[8], ab[8], nw[8];
        zload8(B, b);
        zload8(AB, ab);
        zload8(NW, nw);
#pragma unroll
        for (int p = 0; p < 8; p++) b[p] *= it.cb;
        const uint64_t c1 = uvals[it.va] - zsum8v(a), c2 = uvals[it.vb] - zsum8v(b);  // corr = u - rho
#pragma unroll
        for (int p = 0; p < 8; p++) s[p] = b[p] * c1 + a[p] * c2 + ab[p] - nw[p];
    }
#pragma unroll
    for (int p = 0; p < 8; p++) put64(dst + 8 * p, s[p] + (p == (int)o.omit ? msg : 0ull));
}

// value-plane instruction.  gvals / b2a_vrefs: the GF(2) value plane and the conversions' source refs (prover; nullptr in the
// verifier, whose B2A outputs are leaves computed by z_verify_leaf_b2a).
RV_HD uint64_t z_exec(const ZInstr &in, const uint64_t *v, const uint8_t *gvals, const uint32_t *b2a_vrefs) {
    switch (in.op) {
        case ZV_B2A: {
            uint64_t x = 0;
            if (gvals != nullptr)
                for (uint32_t i = 0; i < 64; i++) x |= (uint64_t)val_of(gvals, b2a_vrefs[64 * in.a + i]) << i;
            return v[in.c] + x;
        }
        case ZV_ADD: return v[in.a] + v[in.b];
        case ZV_SUB: return v[in.a] - v[in.b];
        case ZV_MUL: return v[in.a] * v[in.b] + v[in.c];
        case ZV_ADDC: return v[in.a] + in.imm;
        case ZV_MULC: return v[in.a] * in.imm;
        default: return in.imm;
    }
}

}  // namespace rv
