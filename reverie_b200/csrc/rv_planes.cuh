// rv_planes.cuh -- per-thread bodies of the prover/verifier kernels (__host__ __device__ so that tests/hostsim can replay
// the exact same code on the CPU; the product only ever runs them inside the CUDA kernels of rv_kernels.cu).
//
// Data layout in HBM, for a shard of `npi` packed instances (8 repetitions each):
//   rows      u64 [n_rows][npi]     the share tensor: row = one mask (fresh PRG mask, linear node, or the zero row);
//                                   element = the reference's ShareGF2 word: rep r, player p at bit 63-(8r+p)
//                                   (src/algebra/gf2/share.rs:23-24).  A warp reads one row as 256 coalesced bytes.
//   vals      u8  [n_vals]          value plane: plaintext bit of every wire (shared by all repetitions)
//   on/pre    u8  [8*npi][pitch]    the two hash streams of every repetition, byte t of rep r = what the reference
//                                   pushes into hash_online / hash_preprocess[r] at its t-th call
//                                   (src/algebra/gf2/share.rs:211-218, src/algebra/gf2/recon.rs:314-321)
#pragma once
#include <stdint.h>

#include "rv_aes_bs.cuh"
#include "rv_blake3.cuh"
#include "rv_compile.h"

namespace rv {

#define RV_LSB8 0x0101010101010101ull

// DomainGF2::reconstruct, src/algebra/gf2/domain.rs:47-63: per-repetition parity of the 8 player bits, smeared to 0x00/0xFF.
RV_HD uint64_t gf2_reconstruct(uint64_t t) {
    t ^= t >> 4;
    t ^= t >> 2;
    t ^= t >> 1;
    t &= RV_LSB8;
    return t * 0xFFull;
}

RV_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
#endif
}

// 4x4 byte transpose: c[k] = (a0.byte k, a1.byte k, a2.byte k, a3.byte k), byte 0 first.
RV_HD void transpose4x4(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t c[4]) {
    const uint32_t t0 = byte_perm(a0, a1, 0x5140), t1 = byte_perm(a2, a3, 0x5140);
    const uint32_t t2 = byte_perm(a0, a1, 0x7362), t3 = byte_perm(a2, a3, 0x7362);
    c[0] = byte_perm(t0, t1, 0x5410);
    c[1] = byte_perm(t0, t1, 0x7632);
    c[2] = byte_perm(t2, t3, 0x5410);
    c[3] = byte_perm(t2, t3, 0x7632);
}

// W[i] = packed word of stream position t0+i (rep r in big-endian byte r).  out[r] = the 8 bytes rep r appends to its
// stream at positions t0..t0+7, as a little-endian u64 ready to be stored at &stream[r][t0].
RV_HD void words_to_stream_bytes(const uint64_t W[8], uint64_t out[8]) {
    uint32_t c[4], d[4];
    // reps 0..3 live in the high halves (rep r = byte 3-r of the high u32)
    transpose4x4((uint32_t)(W[0] >> 32), (uint32_t)(W[1] >> 32), (uint32_t)(W[2] >> 32), (uint32_t)(W[3] >> 32), c);
    transpose4x4((uint32_t)(W[4] >> 32), (uint32_t)(W[5] >> 32), (uint32_t)(W[6] >> 32), (uint32_t)(W[7] >> 32), d);
#pragma unroll
    for (int r = 0; r < 4; r++) out[r] = ((uint64_t)d[3 - r] << 32) | c[3 - r];
    transpose4x4((uint32_t)W[0], (uint32_t)W[1], (uint32_t)W[2], (uint32_t)W[3], c);
    transpose4x4((uint32_t)W[4], (uint32_t)W[5], (uint32_t)W[6], (uint32_t)W[7], d);
#pragma unroll
    for (int r = 0; r < 4; r++) out[4 + r] = ((uint64_t)d[3 - r] << 32) | c[3 - r];
}

RV_HD uint32_t val_of(const uint8_t *vals, uint32_t vref) { return (uint32_t)(vals[vref >> 1] ^ (vref & 1)) & 1u; }
// The plaintext of a wire as a Recon-format word of packed instance pi (byte r = 0x00 / 0xFF for repetition r): shared by
// all repetitions (value plane), or per repetition (tainted plane: wires that depend on Random / B2A fresh bits).
RV_HD uint64_t val_word(const uint8_t *vals, const uint64_t *tvals, uint32_t npi, uint32_t pi, uint32_t vref) {
    if (vref & VREF_TAINT) return tvals[(size_t)((vref & ~VREF_TAINT) >> 1) * npi + pi] ^ (0ull - (uint64_t)(vref & 1));
    return 0ull - (uint64_t)val_of(vals, vref);
}
// One gate of the tainted plane for packed instance pi.
RV_HD uint64_t tainted_eval(const TGate &g, const uint64_t *rows, uint32_t npi, uint32_t pi, const uint8_t *vals, const uint64_t *tvals) {
    if (g.op == T_LEAF) return gf2_reconstruct(rows[(size_t)g.a * npi + pi]);  // corr = 0: value = reconstruct(mask)
    const uint64_t a = val_word(vals, tvals, npi, pi, g.a), b = val_word(vals, tvals, npi, pi, g.b);
    return g.op == T_AND ? (a & b) : (a ^ b);
}

// ---- item plane, prover ------------------------------------------------------------------------------------------
// The packed word one online item contributes (src/transcript/prover.rs:181-232, src/interpreter/single.rs:25-69,140-147).
// *bad is OR-ed with 1 when an AssertZero sees a non-zero plaintext (the reference's `assert!`, prover.rs:221-228).
RV_HD uint64_t prover_online_word(const Item &it, const uint64_t *rows, uint32_t npi, uint32_t pi, const uint8_t *vals, const uint64_t *tvals,
                                  int *bad) {
    if (it.kind == ITEM_MUL) {
        const uint64_t la = rows[(size_t)it.ra * npi + pi], lb = rows[(size_t)it.rb * npi + pi];
        const uint64_t mab = rows[(size_t)it.k * npi + pi], mnew = rows[(size_t)(it.k + 1) * npi + pi];
        // corr = value - reconstruct(mask)   (src/interpreter/mod.rs:17-19)
        const uint64_t ca = val_word(vals, tvals, npi, pi, it.va) ^ gf2_reconstruct(la);
        const uint64_t cb = val_word(vals, tvals, npi, pi, it.vb) ^ gf2_reconstruct(lb);
        return (lb & ca) ^ (la & cb) ^ mab ^ mnew;  // the broadcast share `s`, single.rs:41-45
    }
    if (it.kind == ITEM_INPUT) {
        const uint64_t m = rows[(size_t)it.ra * npi + pi];
        return (0ull - val_of(vals, it.va)) ^ gf2_reconstruct(m);  // masked input, prover.rs:186-195
    }
    if (it.kind == ITEM_ASSERT && val_word(vals, tvals, npi, pi, it.va) != 0) *bad |= 1;
    return rows[(size_t)it.ra * npi + pi];  // AssertZero / B2A reconstruct() broadcast the wire's mask shares, single.rs:143-144
}

// The correction word of the j-th Mul: delta = a*b - c on reconstructed masks (single.rs:35-39).
RV_HD uint64_t pre_word(const Item &it, const uint64_t *rows, uint32_t npi, uint32_t pi) {
    const uint64_t a = gf2_reconstruct(rows[(size_t)it.ra * npi + pi]);
    const uint64_t b = gf2_reconstruct(rows[(size_t)it.rb * npi + pi]);
    const uint64_t c = gf2_reconstruct(rows[(size_t)it.k * npi + pi]);
    return (a & b) ^ c;
}

// ---- online verifier (src/transcript/verifier/online.rs:25-182) -------------------------------------------------------
// One opened repetition's packed proof fields inside the uploaded proof bytes.  `eff_*` are the byte counts the reference's
// unpack actually reads for this repetition's pack of 8 (the length of the pack's FIRST repetition, src/algebra/gf2/recon.rs
// :241-259, gf2/share.rs:151-208); elements past the end read as zero (`unwrap_or_default`, online.rs:124,163,171).
struct VOpen {
    uint32_t off_recons, off_corrs, off_inputs;
    uint32_t eff_recons, eff_corrs, eff_inputs;
    uint32_t omit, pad;
};

// element e of a packed bit vector, first element = MSB of byte 0 (gf2/share.rs:66-85, gf2/recon.rs:127-148)
RV_HD uint32_t packed_bit(const uint8_t *p, uint32_t eff_bytes, uint32_t e) {
    return (e >> 3) < eff_bytes ? (uint32_t)(p[e >> 3] >> (7 - (e & 7))) & 1u : 0u;
}
// parity of the (non-omitted) player bits of repetition `rep` (0..7 within the word)
RV_HD uint32_t rho_bit(uint64_t share_word, uint32_t rep) { return (uint32_t)(gf2_reconstruct(share_word) >> (8 * (7 - rep))) & 1u; }

// u-plane leaves of opened repetition `slot`: u = corr ^ rho.
//   input k : u = inputs[k] ^ rho(mask)                                               (online.rs:123-130)
//   Mul   j : kappa = rho_a*rho_b ^ rho_ab ^ msg ^ delta, msg = the omitted player's broadcast bit, delta = proof correction
RV_HD uint8_t verify_leaf_input(const Item &it, uint32_t k, const VOpen &o, const uint8_t *proof, const uint64_t *rows, uint32_t npi, uint32_t slot) {
    const uint32_t c = packed_bit(proof + o.off_inputs, o.eff_inputs, k);
    return (uint8_t)(c ^ rho_bit(rows[(size_t)it.ra * npi + (slot >> 3)], slot & 7));
}
// Random / B2A fresh wire: corr = 0, so u = rho(mask)
RV_HD uint8_t verify_leaf_random(uint32_t row, const uint64_t *rows, uint32_t npi, uint32_t slot) {
    return (uint8_t)rho_bit(rows[(size_t)row * npi + (slot >> 3)], slot & 7);
}
RV_HD uint8_t verify_leaf_kappa(const Item &it, uint32_t recon_idx, const VOpen &o, const uint8_t *proof, const uint64_t *rows, uint32_t npi,
                                uint32_t slot) {
    const uint32_t pi = slot >> 3, r = slot & 7;
    const uint32_t ra = rho_bit(rows[(size_t)it.ra * npi + pi], r), rb = rho_bit(rows[(size_t)it.rb * npi + pi], r);
    const uint32_t rab = rho_bit(rows[(size_t)it.k * npi + pi], r);
    const uint32_t msg = packed_bit(proof + o.off_recons, o.eff_recons, recon_idx);
    const uint32_t delta = packed_bit(proof + o.off_corrs, o.eff_corrs, it.j);
    return (uint8_t)((ra & rb) ^ rab ^ msg ^ delta);
}

// The packed online word of item t for packed instance `pi` of opened repetitions (slots 8pi..8pi+7).
//   uvals: u-plane values, repetition `slot` at uvals + slot * upitch; ua/ub: the item's u-plane refs.
// *not_okay is OR-ed with 1 when an AssertZero sees a non-zero value (online.rs:176-178; unused by Proof::verify).
RV_HD uint64_t verify_online_word(const Item &it, uint32_t t, uint32_t ua, uint32_t ub, uint32_t recon_idx, const VOpen *opens,
                                  const uint8_t *proof, const uint64_t *rows, uint32_t npi, uint32_t pi, const uint8_t *uvals, size_t upitch,
                                  int *not_okay) {
    (void)t;
    uint64_t Va = 0, Vb = 0, msgw = 0, inw = 0;
#pragma unroll
    for (uint32_t r = 0; r < 8; r++) {
        const uint32_t slot = 8 * pi + r;
        const VOpen &o = opens[slot];
        const uint8_t *uv = uvals + (size_t)slot * upitch;
        const uint64_t byte = 0xFFull << (8 * (7 - r));
        if (it.kind == ITEM_INPUT) {
            if (packed_bit(proof + o.off_inputs, o.eff_inputs, it.j)) inw |= byte;
        } else {
            const uint32_t msg = packed_bit(proof + o.off_recons, o.eff_recons, recon_idx);
            msgw |= (uint64_t)msg << (63 - (8 * r + o.omit));  // the omitted player's share of the broadcast (online.rs:140-160)
            const uint32_t a = val_of(uv, ua);
            if (a) Va |= byte;
            if (it.kind == ITEM_MUL) {
                if (val_of(uv, ub)) Vb |= byte;
            } else if (it.kind == ITEM_ASSERT && (a ^ msg)) {
                *not_okay |= 1;
            }
        }
    }
    if (it.kind == ITEM_INPUT) return inw;  // the masked input from the proof is hashed as is (online.rs:126-127)
    const uint64_t la = rows[(size_t)it.ra * npi + pi];
    if (it.kind != ITEM_MUL) return la ^ msgw;  // AssertZero / B2A reconstruct()
    const uint64_t lb = rows[(size_t)it.rb * npi + pi];
    const uint64_t mab = rows[(size_t)it.k * npi + pi], mnew = rows[(size_t)(it.k + 1) * npi + pi];
    const uint64_t ca = Va ^ gf2_reconstruct(la), cb = Vb ^ gf2_reconstruct(lb);  // corr = u ^ rho
    return (lb & ca) ^ (la & cb) ^ mab ^ mnew ^ msgw;
}

// Preprocessing-stream word of opened repetitions: the proof's corrections, one 0x00/0xFF byte per repetition (online.rs:169-174).
RV_HD uint64_t verify_pre_word(uint32_t j, const VOpen *opens, const uint8_t *proof, uint32_t pi) {
    uint64_t w = 0;
#pragma unroll
    for (uint32_t r = 0; r < 8; r++) {
        const VOpen &o = opens[8 * pi + r];
        if (packed_bit(proof + o.off_corrs, o.eff_corrs, j)) w |= 0xFFull << (8 * (7 - r));
    }
    return w;
}

// ---- key setup -------------------------------------------------------------------------------------------------------
// Slice w covers packed instance w/2; odd w = the high u32 of the share word (repetitions 0..3), even w = the low u32
// (repetitions 4..7).  Bit q of a slice word belongs to stream index 31-q = 8*(rep within slice) + player.
RV_HD uint32_t slice_rep(uint32_t w, uint32_t q) { return (w >> 1) * 8 + ((w & 1) ? 0u : 4u) + ((31 - q) >> 3); }
RV_HD uint32_t slice_player(uint32_t q) { return (31 - q) & 7; }

// Player key of (rep, p) and its AES round keys.  mode[rep]==1: the key is given (verifier, online.rs:25-121); otherwise
// it is block p of AES-CTR(seed[rep]) (expand_seed, src/transcript/mod.rs:99-106).  Returns "stream is active" (the
// omitted player's tape stays zero, src/generator/batch.rs:31-34) and writes the OpenOnline.seeds entry (prover.rs:126-127).
template <typename SW = NetlistSubWord>
RV_HD bool key_setup_stream(uint32_t rep, uint32_t p, const uint8_t *seeds, const uint8_t *pkeys_in, const uint8_t *mode,
                            const uint8_t *omit, uint8_t *pkeys_out, uint32_t rk[44], SW sw = SW()) {
    uint32_t key[4];
    if (mode != nullptr && mode[rep] == 1) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(pkeys_in + ((size_t)rep * 8 + p) * 16);
        for (int i = 0; i < 4; i++) key[i] = src[i];
    } else {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(seeds + (size_t)rep * 16);
        uint32_t sk[4] = {src[0], src[1], src[2], src[3]}, in[4];
        aes128_expand_key(sk, rk, sw);
        ctr_block_words(p, in);
        aes128_encrypt_block(rk, in, key, sw);
    }
    const bool active = !(omit != nullptr && omit[rep] == p);
    uint32_t *dst = reinterpret_cast<uint32_t *>(pkeys_out + ((size_t)rep * 8 + p) * 16);
    for (int i = 0; i < 4; i++) dst[i] = active ? key[i] : 0u;
    aes128_expand_key(key, rk, sw);
    return active;
}

// Plane k = 8*B + b of counter block j holds keystream byte B, bit b.  The reference consumes keystream bits MSB-first
// within each byte (src/algebra/gf2/domain.rs:293-377), so the plane is mask number 128j + 8B + (7-b).
RV_HD uint64_t plane_to_mask_index(uint64_t j, int k) { return j * 128 + (uint32_t)(8 * (k >> 3) + 7 - (k & 7)); }

// ---- Fiat-Shamir (src/proof/mod.rs:68-83, src/crypto/ro.rs:7-20) -----------------------------------------------------
// The 56-byte message block of RandomOracle::new(CTX_CHALLENGE, comm): "random-oracle challenge" || 0x00 || comm.
RV_HD void challenge_block(const uint32_t comm[8], uint32_t m[16]) {
    const char ctx[24] = "random-oracle challenge";
    uint8_t blk[64];
    for (int i = 0; i < 64; i++) blk[i] = 0;
    for (int i = 0; i < 23; i++) blk[i] = (uint8_t)ctx[i];
    for (int i = 0; i < 8; i++)
        for (int b = 0; b < 4; b++) blk[24 + 4 * i + b] = (uint8_t)(comm[i] >> (8 * b));
    for (int i = 0; i < 16; i++)
        m[i] = (uint32_t)blk[4 * i] | ((uint32_t)blk[4 * i + 1] << 8) | ((uint32_t)blk[4 * i + 2] << 16) | ((uint32_t)blk[4 * i + 3] << 24);
}
// XOF output block t (64 bytes): the root compression with the output-block counter.
RV_HD void challenge_xof_block(const uint32_t m[16], uint64_t t, uint32_t out[16]) {
    uint32_t iv[8];
    b3_iv(iv);
    b3_compress16(iv, m, t, 56, B3_CHUNK_START | B3_CHUNK_END | B3_ROOT, out);
}
// Consume one XOF block = two (repetition, player) draws; `omit` starts as all RV_PLAYERS.  Later draws overwrite.
RV_HD void challenge_consume(const uint32_t xof[16], uint8_t *omit, int *distinct) {
    for (int h = 0; h < 2 && *distinct < RV_ONLINE_REPS; h++) {
        const uint32_t rep = xof[8 * h] & 0xff, om = xof[8 * h + 4] & 7;  // u128 LE mod 256, u128 LE mod 8
        if (omit[rep] == RV_PLAYERS) (*distinct)++;
        omit[rep] = (uint8_t)om;
    }
}

// ---- extraction (src/transcript/prover.rs:57-175) ------------------------------------------------------------------
// Byte g of a packed bit vector whose element e is bit `bit` of stream[pos[e]] (pos == NULL: e itself); elements past n
// are zero; first element -> MSB (src/algebra/gf2/share.rs:66-85, src/algebra/gf2/recon.rs:127-148).
// Bit `bit` of each of the 8 bytes of w (byte 0 = lowest address) gathered into one byte, byte 0 -> MSB.
RV_HD uint8_t gather_bit_msb_first(uint64_t w, uint32_t bit) {
    return (uint8_t)((((w >> bit) & 0x0101010101010101ull) * 0x8040201008040201ull) >> 56);
}
// 8 stream bytes starting at byte address p (any alignment) as a little-endian u64, from two aligned loads.
RV_HD uint64_t load8_unaligned(const uint8_t *p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint64_t *q = reinterpret_cast<const uint64_t *>(a & ~(uintptr_t)7);
    const uint32_t sh = 8 * (uint32_t)(a & 7);
    const uint64_t lo = q[0];
    return sh == 0 ? lo : (lo >> sh) | (q[1] << (64 - sh));
}
RV_HD uint8_t pack_bits_byte(const uint8_t *stream, const uint32_t *pos, uint32_t n, uint32_t g, uint32_t bit) {
    const uint32_t e0 = 8 * g;
    if (e0 + 8 <= n) {  // a full byte whose 8 elements sit in 8 consecutive stream bytes (always without a position table; the rule for
                        // runs of Mul / AssertZero gates with one): one or two 8-byte loads instead of 8 single ones.  The streams are
                        // padded to whole 2 KiB tiles, so the aligned loads stay inside the buffer.
        const uint32_t p0 = pos ? pos[e0] : e0;
        if (!pos || pos[e0 + 7] == p0 + 7) return gather_bit_msb_first(load8_unaligned(stream + p0), bit);
    }
    uint32_t r = 0;
#pragma unroll
    for (uint32_t i = 0; i < 8; i++) {
        const uint32_t e = e0 + i;
        uint32_t v = 0;
        if (e < n) v = (stream[pos ? pos[e] : e] >> bit) & 1u;
        r = (r << 1) | v;
    }
    return (uint8_t)r;
}

// Streaming: byte g of a packed bit vector, restricted to the elements [first, first + n) that one segment contributes (element
// first + k = bit `bit` of stream[pos ? pos[k] : k], positions relative to the segment's first stream byte); the other bits of the
// byte are zero, so the segments' contributions are OR-ed together -- a byte that straddles two segments is completed by the later one.
RV_HD uint8_t seg_pack_byte(const uint8_t *stream, const uint32_t *pos, uint64_t first, uint32_t n, uint64_t g, uint32_t bit) {
    if (8 * g >= first && 8 * g + 8 <= first + n) {  // all 8 elements belong to this segment: 8-byte loads when they sit side by side
        const uint32_t k0 = (uint32_t)(8 * g - first), p0 = pos ? pos[k0] : k0;
        if (!pos || pos[k0 + 7] == p0 + 7) return gather_bit_msb_first(load8_unaligned(stream + p0), bit);
    }
    uint32_t r = 0;
#pragma unroll
    for (uint32_t i = 0; i < 8; i++) {
        const uint64_t el = 8 * g + i;  // global element; the first one of a byte is its MSB
        uint32_t v = 0;
        if (el >= first && el - first < n) {
            const uint32_t k = (uint32_t)(el - first);
            v = (stream[pos ? pos[k] : k] >> bit) & 1u;
        }
        r = (r << 1) | v;
    }
    return (uint8_t)r;
}

// ---- bincode `Proof` layout (src/proof/mod.rs:40-66; bincode 1.3 default: LE, u64 lengths, fixed arrays inline) ----
struct ProofLayout {
    uint32_t len_recons, len_corrs, len_inputs;
    uint32_t len_zrecons = 0, len_zcorrs = 0, len_zinputs = 0;  // Z64 packs are exactly 8 n bytes (src/algebra/z64/share.rs:37-49, recon.rs:46-66)
    RV_HD size_t sz_on_g() const { return 1 + 128 + 24 + (size_t)len_recons + len_corrs + len_inputs; }
    RV_HD size_t sz_on_z() const { return 1 + 128 + 24 + (size_t)len_zrecons + len_zcorrs + len_zinputs; }
    RV_HD size_t g_base() const { return 32; }
    RV_HD size_t g_pre_base() const { return g_base() + 8 + RV_ONLINE_REPS * sz_on_g(); }
    RV_HD size_t z_base() const { return g_pre_base() + 8 + RV_PREPROCESSING_REPS * (size_t)48; }
    RV_HD size_t z_pre_base() const { return z_base() + 8 + RV_ONLINE_REPS * sz_on_z(); }
    RV_HD size_t total() const { return z_pre_base() + 8 + RV_PREPROCESSING_REPS * (size_t)48; }
};

RV_HD void put_u64le(uint8_t *p, uint64_t v) {
    for (int i = 0; i < 8; i++) p[i] = (uint8_t)(v >> (8 * i));
}

struct ExtractView {
    const uint8_t *on, *pre;       // this repetition's two streams
    const uint8_t *on_hash;        // [32] BLAKE3 of the online stream (OpenPreprocessing.comm_online, prover.rs:168)
    const uint8_t *pkeys;          // [8][16]
    const uint8_t *seed;           // [16]
    const uint8_t *comm;           // [32]
    const uint32_t *z64_empty_hash;
    const uint8_t *z_on_hash = nullptr;  // [32] BLAKE3 of this repetition's Z64 online stream; nullptr = the stream is empty
    const uint32_t *recon_pos, *input_pos;
    uint32_t n_recon, n_pre, n_inputs;
};

// Thread `tid` of `nt` writes its share of repetition `rep`'s entry (GF(2) opening + the empty Z64 opening).
RV_HD void extract_entry(const ProofLayout &L, const ExtractView &v, uint32_t rep, uint32_t omit, uint32_t rank, uint32_t tid,
                         uint32_t nt, uint8_t *P) {
    if (rep == 0 && tid == 0) {  // comm and the four Vec lengths
        for (int i = 0; i < 32; i++) P[i] = v.comm[i];
        put_u64le(P + L.g_base(), RV_ONLINE_REPS);
        put_u64le(P + L.g_pre_base(), RV_PREPROCESSING_REPS);
        put_u64le(P + L.z_base(), RV_ONLINE_REPS);
        put_u64le(P + L.z_pre_base(), RV_PREPROCESSING_REPS);
    }
    if (omit < RV_PLAYERS) {
        uint8_t *e = P + L.g_base() + 8 + rank * L.sz_on_g();
        uint8_t *z = P + L.z_base() + 8 + rank * L.sz_on_z();
        if (tid == 0) {
            e[0] = (uint8_t)omit;
            z[0] = (uint8_t)omit;
            put_u64le(e + 129, L.len_recons);
            put_u64le(e + 137 + L.len_recons, L.len_corrs);
            put_u64le(e + 145 + L.len_recons + L.len_corrs, L.len_inputs);
            put_u64le(z + 129, L.len_zrecons);  // the Z64 vectors themselves are copied by k_zextract
            put_u64le(z + 137 + L.len_zrecons, L.len_zcorrs);
            put_u64le(z + 145 + L.len_zrecons + L.len_zcorrs, L.len_zinputs);
        }
        for (uint32_t i = tid; i < 128; i += nt) {  // OpenOnline.seeds with the unopened player's key zeroed
            const uint8_t b = (i / 16 == omit) ? 0 : v.pkeys[i];
            e[1 + i] = b;
            z[1 + i] = b;
        }
        uint8_t *d = e + 137;
        for (uint32_t g = tid; g < L.len_recons; g += nt) d[g] = pack_bits_byte(v.on, v.recon_pos, v.n_recon, g, 7 - omit);
        d = e + 145 + L.len_recons;
        for (uint32_t g = tid; g < L.len_corrs; g += nt) d[g] = pack_bits_byte(v.pre, nullptr, v.n_pre, g, 0);
        d = e + 153 + L.len_recons + L.len_corrs;
        for (uint32_t g = tid; g < L.len_inputs; g += nt) d[g] = pack_bits_byte(v.on, v.input_pos, v.n_inputs, g, 0);
    } else {
        uint8_t *e = P + L.g_pre_base() + 8 + rank * (size_t)48;
        uint8_t *z = P + L.z_pre_base() + 8 + rank * (size_t)48;
        for (uint32_t i = tid; i < 48; i += nt) {
            if (i < 16) {
                e[i] = v.seed[i];
                z[i] = v.seed[i];
            } else {
                e[i] = v.on_hash[i - 16];
                z[i] = v.z_on_hash ? v.z_on_hash[i - 16] : (uint8_t)(v.z64_empty_hash[(i - 16) / 4] >> (8 * ((i - 16) & 3)));
            }
        }
    }
}

// Transcript::hash (src/transcript/mod.rs:77-96) then CombineInstance::hash (src/interpreter/combine.rs:104-118).
RV_HD void rep_join(const uint32_t h_on[8], const uint32_t h_pre[8], const uint32_t z64_rep_hash[8], uint32_t out[8]) {
    uint32_t g[8];
    b3_hash64(h_pre, h_on, g);
    b3_hash64(g, z64_rep_hash, out);
}


}  // namespace rv
