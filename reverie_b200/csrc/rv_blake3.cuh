// rv_blake3.cuh -- BLAKE3 compression, chunk hashing, tree merge and XOF for the transcript commitments.
//
// Replaces the `blake3` crate as used by the reference (paths relative to its root):
//   BufferedHasher / PackedHasher / HASH!     src/crypto/hash.rs:17-58, 61-104, 118-127
//   RandomOracle (XOF)                         src/crypto/ro.rs:7-20
// The 64 KiB buffering of BufferedHasher is transparent (plain BLAKE3 of the concatenation), so a stream can be hashed
// chunk-parallel: chunk chaining values first, then the left-heavy binary tree (BLAKE3 spec section 2.1).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RV_HD __host__ __device__ __forceinline__
#else
#define RV_HD inline
#endif

namespace rv {

enum : uint32_t { B3_CHUNK_START = 1, B3_CHUNK_END = 2, B3_PARENT = 4, B3_ROOT = 8 };
#define RV_B3_IV0 0x6A09E667u
#define RV_B3_IV1 0xBB67AE85u
#define RV_B3_IV2 0x3C6EF372u
#define RV_B3_IV3 0xA54FF53Au
#define RV_B3_IV4 0x510E527Fu
#define RV_B3_IV5 0x9B05688Cu
#define RV_B3_IV6 0x1F83D9ABu
#define RV_B3_IV7 0x5BE0CD19u

RV_HD uint32_t b3_rotr(uint32_t x, int n) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(x, x, n);
#else
    return (x >> n) | (x << (32 - n));
#endif
}

#define RV_B3_G(a, b, c, d, mx, my) \
    do {                            \
        a = a + b + (mx);           \
        d = b3_rotr(d ^ a, 16);     \
        c = c + d;                  \
        b = b3_rotr(b ^ c, 12);     \
        a = a + b + (my);           \
        d = b3_rotr(d ^ a, 8);      \
        c = c + d;                  \
        b = b3_rotr(b ^ c, 7);      \
    } while (0)

#define RV_B3_ROUND(m0, m1, m2, m3, m4, m5, m6, m7, m8, m9, m10, m11, m12, m13, m14, m15) \
    do {                                                                                   \
        RV_B3_G(v0, v4, v8, v12, m[m0], m[m1]);                                            \
        RV_B3_G(v1, v5, v9, v13, m[m2], m[m3]);                                            \
        RV_B3_G(v2, v6, v10, v14, m[m4], m[m5]);                                           \
        RV_B3_G(v3, v7, v11, v15, m[m6], m[m7]);                                           \
        RV_B3_G(v0, v5, v10, v15, m[m8], m[m9]);                                           \
        RV_B3_G(v1, v6, v11, v12, m[m10], m[m11]);                                         \
        RV_B3_G(v2, v7, v8, v13, m[m12], m[m13]);                                          \
        RV_B3_G(v3, v4, v9, v14, m[m14], m[m15]);                                          \
    } while (0)

// Full 16-word output of the compression function.  out[0..8) is the new chaining value.
RV_HD void b3_compress16(const uint32_t cv[8], const uint32_t m[16], uint64_t counter, uint32_t block_len, uint32_t flags,
                         uint32_t out[16]) {
    uint32_t v0 = cv[0], v1 = cv[1], v2 = cv[2], v3 = cv[3], v4 = cv[4], v5 = cv[5], v6 = cv[6], v7 = cv[7];
    uint32_t v8 = RV_B3_IV0, v9 = RV_B3_IV1, v10 = RV_B3_IV2, v11 = RV_B3_IV3;
    uint32_t v12 = (uint32_t)counter, v13 = (uint32_t)(counter >> 32), v14 = block_len, v15 = flags;
    // the message permutation applied 0..6 times, written out (BLAKE3 spec table 2)
    RV_B3_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    RV_B3_ROUND(2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8);
    RV_B3_ROUND(3, 4, 10, 12, 13, 2, 7, 14, 6, 5, 9, 0, 11, 15, 8, 1);
    RV_B3_ROUND(10, 7, 12, 9, 14, 3, 13, 15, 4, 0, 11, 2, 5, 8, 1, 6);
    RV_B3_ROUND(12, 13, 9, 11, 15, 10, 14, 8, 7, 2, 5, 3, 0, 1, 6, 4);
    RV_B3_ROUND(9, 14, 11, 5, 8, 12, 15, 1, 13, 3, 0, 10, 2, 6, 4, 7);
    RV_B3_ROUND(11, 15, 5, 0, 1, 9, 8, 6, 14, 10, 2, 12, 3, 4, 7, 13);
    out[0] = v0 ^ v8;
    out[1] = v1 ^ v9;
    out[2] = v2 ^ v10;
    out[3] = v3 ^ v11;
    out[4] = v4 ^ v12;
    out[5] = v5 ^ v13;
    out[6] = v6 ^ v14;
    out[7] = v7 ^ v15;
    out[8] = v8 ^ cv[0];
    out[9] = v9 ^ cv[1];
    out[10] = v10 ^ cv[2];
    out[11] = v11 ^ cv[3];
    out[12] = v12 ^ cv[4];
    out[13] = v13 ^ cv[5];
    out[14] = v14 ^ cv[6];
    out[15] = v15 ^ cv[7];
}

// In-place chaining-value update (first 8 output words only).
RV_HD void b3_compress_cv(uint32_t cv[8], const uint32_t m[16], uint64_t counter, uint32_t block_len, uint32_t flags) {
    uint32_t v0 = cv[0], v1 = cv[1], v2 = cv[2], v3 = cv[3], v4 = cv[4], v5 = cv[5], v6 = cv[6], v7 = cv[7];
    uint32_t v8 = RV_B3_IV0, v9 = RV_B3_IV1, v10 = RV_B3_IV2, v11 = RV_B3_IV3;
    uint32_t v12 = (uint32_t)counter, v13 = (uint32_t)(counter >> 32), v14 = block_len, v15 = flags;
    RV_B3_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    RV_B3_ROUND(2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8);
    RV_B3_ROUND(3, 4, 10, 12, 13, 2, 7, 14, 6, 5, 9, 0, 11, 15, 8, 1);
    RV_B3_ROUND(10, 7, 12, 9, 14, 3, 13, 15, 4, 0, 11, 2, 5, 8, 1, 6);
    RV_B3_ROUND(12, 13, 9, 11, 15, 10, 14, 8, 7, 2, 5, 3, 0, 1, 6, 4);
    RV_B3_ROUND(9, 14, 11, 5, 8, 12, 15, 1, 13, 3, 0, 10, 2, 6, 4, 7);
    RV_B3_ROUND(11, 15, 5, 0, 1, 9, 8, 6, 14, 10, 2, 12, 3, 4, 7, 13);
    cv[0] = v0 ^ v8;
    cv[1] = v1 ^ v9;
    cv[2] = v2 ^ v10;
    cv[3] = v3 ^ v11;
    cv[4] = v4 ^ v12;
    cv[5] = v5 ^ v13;
    cv[6] = v6 ^ v14;
    cv[7] = v7 ^ v15;
}

RV_HD void b3_iv(uint32_t cv[8]) {
    cv[0] = RV_B3_IV0;
    cv[1] = RV_B3_IV1;
    cv[2] = RV_B3_IV2;
    cv[3] = RV_B3_IV3;
    cv[4] = RV_B3_IV4;
    cv[5] = RV_B3_IV5;
    cv[6] = RV_B3_IV6;
    cv[7] = RV_B3_IV7;
}

// Chaining value of one chunk (<= 1024 bytes at `data`, 4-byte aligned, readable up to the next multiple of 64 bytes;
// bytes past `len` are ignored).  `root` => this chunk is the whole input (len <= 1024) and the result is the hash.
// An empty input is one chunk with a single empty block.
RV_HD void b3_chunk_cv(const uint32_t *data, uint32_t len, uint64_t chunk_index, bool root, uint32_t cv[8]) {
    b3_iv(cv);
    const uint32_t n_blocks = len == 0 ? 1 : (len + 63) / 64;
    for (uint32_t b = 0; b < n_blocks; b++) {
        uint32_t m[16];
        const uint32_t blen = (b + 1 < n_blocks) ? 64 : (len - 64 * b);
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const uint32_t off = 64 * b + 4 * i;
            uint32_t w = 0;
            if (off < len) {
                w = data[off / 4];
                if (len - off < 4) w &= (1u << (8 * (len - off))) - 1u;
            }
            m[i] = w;
        }
        uint32_t flags = (b == 0 ? B3_CHUNK_START : 0) | (b + 1 == n_blocks ? B3_CHUNK_END : 0);
        if (root && b + 1 == n_blocks) flags |= B3_ROOT;
        b3_compress_cv(cv, m, chunk_index, blen, flags);
    }
}

// Parent node: cv <- compress(IV, left || right, PARENT [| ROOT]).
RV_HD void b3_parent_cv(const uint32_t left[8], const uint32_t right[8], bool root, uint32_t out[8]) {
    uint32_t m[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        m[i] = left[i];
        m[8 + i] = right[i];
    }
    uint32_t cv[8];
    b3_iv(cv);
    b3_compress_cv(cv, m, 0, 64, B3_PARENT | (root ? B3_ROOT : 0));
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = cv[i];
}

// HASH!(a, b) of two 32-byte values = BLAKE3 of a 64-byte single-block input (src/crypto/hash.rs:118-127).
RV_HD void b3_hash64(const uint32_t a[8], const uint32_t b[8], uint32_t out[8]) {
    uint32_t m[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        m[i] = a[i];
        m[8 + i] = b[i];
    }
    uint32_t cv[8];
    b3_iv(cv);
    b3_compress_cv(cv, m, 0, 64, B3_CHUNK_START | B3_CHUNK_END | B3_ROOT);
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = cv[i];
}

}  // namespace rv
