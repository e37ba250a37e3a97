/* oracle/c/orc_blake3.c -- BLAKE3 for the CPU oracle (see orc_blake3.h).  TEST INFRASTRUCTURE ONLY.
 *
 * Portable scalar compression plus, when compiled with -mavx2, an 8-chunk-parallel path for bulk updates so that
 * the timed CPU baseline hashes its 64 KiB flushes (src/crypto/hash.rs:5,36-51) at SIMD speed like the `blake3`
 * crate does (the crate additionally has an AVX-512 path which this restatement does not).
 */
#include "orc_blake3.h"

#include <string.h>
#ifdef __AVX2__
#include <immintrin.h>
#endif

enum { CHUNK_START = 1, CHUNK_END = 2, PARENT = 4, ROOT = 8 };

static const uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                               0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
static const uint8_t PERM[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};

static uint8_t SCHED[7][16];
static int sched_ready = 0;
static void sched_init(void) {
    if (sched_ready) return;
    for (int i = 0; i < 16; i++) SCHED[0][i] = (uint8_t)i;
    for (int r = 1; r < 7; r++)
        for (int i = 0; i < 16; i++) SCHED[r][i] = SCHED[r - 1][PERM[i]];
    __atomic_store_n(&sched_ready, 1, __ATOMIC_RELEASE);
}

static inline uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static inline uint32_t ld32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static inline void st32(uint8_t *p, uint32_t v) {
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}

#define G(a, b, c, d, x, y)                  \
    do {                                     \
        a = a + b + (x); d = rotr32(d ^ a, 16); \
        c = c + d;       b = rotr32(b ^ c, 12); \
        a = a + b + (y); d = rotr32(d ^ a, 8);  \
        c = c + d;       b = rotr32(b ^ c, 7);  \
    } while (0)

/* full 16-word output of the compression function */
static void compress(const uint32_t cv[8], const uint8_t block[64], uint32_t block_len, uint64_t counter,
                     uint32_t flags, uint32_t out[16]) {
    uint32_t m[16], v[16];
    for (int i = 0; i < 16; i++) m[i] = ld32(block + 4 * i);
    for (int i = 0; i < 8; i++) v[i] = cv[i];
    v[8] = IV[0]; v[9] = IV[1]; v[10] = IV[2]; v[11] = IV[3];
    v[12] = (uint32_t)counter; v[13] = (uint32_t)(counter >> 32); v[14] = block_len; v[15] = flags;
    for (int r = 0; r < 7; r++) {
        const uint8_t *s = SCHED[r];
        G(v[0], v[4], v[8], v[12], m[s[0]], m[s[1]]);
        G(v[1], v[5], v[9], v[13], m[s[2]], m[s[3]]);
        G(v[2], v[6], v[10], v[14], m[s[4]], m[s[5]]);
        G(v[3], v[7], v[11], v[15], m[s[6]], m[s[7]]);
        G(v[0], v[5], v[10], v[15], m[s[8]], m[s[9]]);
        G(v[1], v[6], v[11], v[12], m[s[10]], m[s[11]]);
        G(v[2], v[7], v[8], v[13], m[s[12]], m[s[13]]);
        G(v[3], v[4], v[9], v[14], m[s[14]], m[s[15]]);
    }
    for (int i = 0; i < 8; i++) {
        out[i] = v[i] ^ v[i + 8];
        out[i + 8] = v[i + 8] ^ cv[i];
    }
}

static void parent_cv(const uint32_t l[8], const uint32_t r[8], uint32_t out[8]) {
    uint8_t block[64];
    uint32_t full[16];
    for (int i = 0; i < 8; i++) { st32(block + 4 * i, l[i]); st32(block + 32 + 4 * i, r[i]); }
    compress(IV, block, 64, 0, PARENT, full);
    memcpy(out, full, 32);
}

static void push_chunk_cv(orc_b3 *h, uint32_t cv[8], uint64_t total_chunks) {
    while ((total_chunks & 1) == 0) {
        uint32_t merged[8];
        parent_cv(h->cv_stack[h->cv_stack_len - 1], cv, merged);
        h->cv_stack_len--;
        memcpy(cv, merged, 32);
        total_chunks >>= 1;
    }
    memcpy(h->cv_stack[h->cv_stack_len++], cv, 32);
}

#ifdef __AVX2__
/* ---- 8 full chunks in parallel: lane c of every vector belongs to chunk c -------------------------------- */
static inline __m256i rot16(__m256i x) {
    return _mm256_shuffle_epi8(x, _mm256_set_epi8(13, 12, 15, 14, 9, 8, 11, 10, 5, 4, 7, 6, 1, 0, 3, 2,
                                                  13, 12, 15, 14, 9, 8, 11, 10, 5, 4, 7, 6, 1, 0, 3, 2));
}
static inline __m256i rot8(__m256i x) {
    return _mm256_shuffle_epi8(x, _mm256_set_epi8(12, 15, 14, 13, 8, 11, 10, 9, 4, 7, 6, 5, 0, 3, 2, 1,
                                                  12, 15, 14, 13, 8, 11, 10, 9, 4, 7, 6, 5, 0, 3, 2, 1));
}
static inline __m256i rot12(__m256i x) { return _mm256_or_si256(_mm256_srli_epi32(x, 12), _mm256_slli_epi32(x, 20)); }
static inline __m256i rot7(__m256i x) { return _mm256_or_si256(_mm256_srli_epi32(x, 7), _mm256_slli_epi32(x, 25)); }

#define GV(a, b, c, d, x, y)                                                         \
    do {                                                                             \
        a = _mm256_add_epi32(_mm256_add_epi32(a, b), x); d = rot16(_mm256_xor_si256(d, a)); \
        c = _mm256_add_epi32(c, d);                      b = rot12(_mm256_xor_si256(b, c)); \
        a = _mm256_add_epi32(_mm256_add_epi32(a, b), y); d = rot8(_mm256_xor_si256(d, a));  \
        c = _mm256_add_epi32(c, d);                      b = rot7(_mm256_xor_si256(b, c));  \
    } while (0)

static inline void transpose8(__m256i r[8]) {
    __m256i t0 = _mm256_unpacklo_epi32(r[0], r[1]), t1 = _mm256_unpackhi_epi32(r[0], r[1]);
    __m256i t2 = _mm256_unpacklo_epi32(r[2], r[3]), t3 = _mm256_unpackhi_epi32(r[2], r[3]);
    __m256i t4 = _mm256_unpacklo_epi32(r[4], r[5]), t5 = _mm256_unpackhi_epi32(r[4], r[5]);
    __m256i t6 = _mm256_unpacklo_epi32(r[6], r[7]), t7 = _mm256_unpackhi_epi32(r[6], r[7]);
    __m256i u0 = _mm256_unpacklo_epi64(t0, t2), u1 = _mm256_unpackhi_epi64(t0, t2);
    __m256i u2 = _mm256_unpacklo_epi64(t1, t3), u3 = _mm256_unpackhi_epi64(t1, t3);
    __m256i u4 = _mm256_unpacklo_epi64(t4, t6), u5 = _mm256_unpackhi_epi64(t4, t6);
    __m256i u6 = _mm256_unpacklo_epi64(t5, t7), u7 = _mm256_unpackhi_epi64(t5, t7);
    r[0] = _mm256_permute2x128_si256(u0, u4, 0x20); r[1] = _mm256_permute2x128_si256(u1, u5, 0x20);
    r[2] = _mm256_permute2x128_si256(u2, u6, 0x20); r[3] = _mm256_permute2x128_si256(u3, u7, 0x20);
    r[4] = _mm256_permute2x128_si256(u0, u4, 0x31); r[5] = _mm256_permute2x128_si256(u1, u5, 0x31);
    r[6] = _mm256_permute2x128_si256(u2, u6, 0x31); r[7] = _mm256_permute2x128_si256(u3, u7, 0x31);
}

static void hash8_chunks(const uint8_t *in, uint64_t counter, uint32_t out_cv[8][8]) {
    __m256i h[8];
    for (int i = 0; i < 8; i++) h[i] = _mm256_set1_epi32((int)IV[i]);
    uint32_t lo[8], hi[8];
    for (int c = 0; c < 8; c++) { lo[c] = (uint32_t)(counter + c); hi[c] = (uint32_t)((counter + c) >> 32); }
    const __m256i ctr_lo = _mm256_loadu_si256((const __m256i *)lo), ctr_hi = _mm256_loadu_si256((const __m256i *)hi);
    for (int b = 0; b < 16; b++) {
        __m256i m[16], v[16];
        for (int half = 0; half < 2; half++) {
            __m256i r[8];
            for (int c = 0; c < 8; c++) r[c] = _mm256_loadu_si256((const __m256i *)(in + 1024 * c + 64 * b + 32 * half));
            transpose8(r);
            for (int i = 0; i < 8; i++) m[8 * half + i] = r[i];
        }
        uint32_t flags = (b == 0 ? CHUNK_START : 0) | (b == 15 ? CHUNK_END : 0);
        for (int i = 0; i < 8; i++) v[i] = h[i];
        for (int i = 0; i < 4; i++) v[8 + i] = _mm256_set1_epi32((int)IV[i]);
        v[12] = ctr_lo; v[13] = ctr_hi; v[14] = _mm256_set1_epi32(64); v[15] = _mm256_set1_epi32((int)flags);
        for (int r = 0; r < 7; r++) {
            const uint8_t *s = SCHED[r];
            GV(v[0], v[4], v[8], v[12], m[s[0]], m[s[1]]);
            GV(v[1], v[5], v[9], v[13], m[s[2]], m[s[3]]);
            GV(v[2], v[6], v[10], v[14], m[s[4]], m[s[5]]);
            GV(v[3], v[7], v[11], v[15], m[s[6]], m[s[7]]);
            GV(v[0], v[5], v[10], v[15], m[s[8]], m[s[9]]);
            GV(v[1], v[6], v[11], v[12], m[s[10]], m[s[11]]);
            GV(v[2], v[7], v[8], v[13], m[s[12]], m[s[13]]);
            GV(v[3], v[4], v[9], v[14], m[s[14]], m[s[15]]);
        }
        for (int i = 0; i < 8; i++) h[i] = _mm256_xor_si256(v[i], v[i + 8]);
    }
    transpose8(h);
    for (int c = 0; c < 8; c++) _mm256_storeu_si256((__m256i *)out_cv[c], h[c]);
}
#endif /* __AVX2__ */

static inline size_t chunk_len(const orc_b3 *h) { return 64u * h->blocks_compressed + h->buf_len; }
static inline uint32_t start_flag(const orc_b3 *h) { return h->blocks_compressed == 0 ? CHUNK_START : 0; }

static void chunk_reset(orc_b3 *h, uint64_t counter) {
    memcpy(h->cv, IV, 32);
    h->chunk_counter = counter;
    h->buf_len = 0;
    h->blocks_compressed = 0;
}

void orc_b3_init(orc_b3 *h) {
    sched_init();
    chunk_reset(h, 0);
    h->cv_stack_len = 0;
}

static void chunk_update(orc_b3 *h, const uint8_t *in, size_t len) {
    while (len > 0) {
        if (h->buf_len == 64) {
            uint32_t full[16];
            compress(h->cv, h->buf, 64, h->chunk_counter, start_flag(h), full);
            memcpy(h->cv, full, 32);
            h->blocks_compressed++;
            h->buf_len = 0;
        }
        size_t take = 64u - h->buf_len;
        if (take > len) take = len;
        memcpy(h->buf + h->buf_len, in, take);
        h->buf_len += (uint8_t)take;
        in += take;
        len -= take;
    }
}

void orc_b3_update(orc_b3 *h, const void *data, size_t len) {
    const uint8_t *in = (const uint8_t *)data;
    while (len > 0) {
        if (chunk_len(h) == ORC_B3_CHUNK) {
            uint32_t full[16], cv[8];
            compress(h->cv, h->buf, 64, h->chunk_counter, start_flag(h) | CHUNK_END, full);
            memcpy(cv, full, 32);
            uint64_t total = h->chunk_counter + 1;
            push_chunk_cv(h, cv, total);
            chunk_reset(h, total);
        }
#ifdef __AVX2__
        while (chunk_len(h) == 0 && len > 8 * ORC_B3_CHUNK) { /* strictly more: the last chunk stays lazy */
            uint32_t cvs[8][8];
            hash8_chunks(in, h->chunk_counter, cvs);
            for (int c = 0; c < 8; c++) push_chunk_cv(h, cvs[c], h->chunk_counter + c + 1);
            chunk_reset(h, h->chunk_counter + 8);
            in += 8 * ORC_B3_CHUNK;
            len -= 8 * ORC_B3_CHUNK;
        }
#endif
        size_t want = ORC_B3_CHUNK - chunk_len(h);
        if (want > len) want = len;
        chunk_update(h, in, want);
        in += want;
        len -= want;
    }
}

/* output node = (input cv, block, block_len, counter, flags) */
typedef struct { uint32_t cv[8]; uint8_t block[64]; uint32_t block_len; uint64_t counter; uint32_t flags; } node;

static void final_node(const orc_b3 *h, node *n) {
    memcpy(n->cv, h->cv, 32);
    memset(n->block, 0, 64);
    memcpy(n->block, h->buf, h->buf_len);
    n->block_len = h->buf_len;
    n->counter = h->chunk_counter;
    n->flags = start_flag(h) | CHUNK_END;
    for (int i = h->cv_stack_len; i > 0; i--) {
        uint32_t full[16];
        compress(n->cv, n->block, n->block_len, n->counter, n->flags, full);
        for (int k = 0; k < 8; k++) { st32(n->block + 4 * k, h->cv_stack[i - 1][k]); st32(n->block + 32 + 4 * k, full[k]); }
        memcpy(n->cv, IV, 32);
        n->block_len = 64;
        n->counter = 0;
        n->flags = PARENT;
    }
}

void orc_b3_finalize_xof(const orc_b3 *h, uint64_t seek, uint8_t *out, size_t len) {
    node n;
    final_node(h, &n);
    uint64_t blk = seek / 64;
    size_t off = (size_t)(seek % 64);
    while (len > 0) {
        uint32_t full[16];
        uint8_t bytes[64];
        compress(n.cv, n.block, n.block_len, blk, n.flags | ROOT, full);
        for (int i = 0; i < 16; i++) st32(bytes + 4 * i, full[i]);
        size_t take = 64 - off;
        if (take > len) take = len;
        memcpy(out, bytes + off, take);
        out += take;
        len -= take;
        off = 0;
        blk++;
    }
}

void orc_b3_finalize(const orc_b3 *h, uint8_t out[ORC_B3_OUT]) { orc_b3_finalize_xof(h, 0, out, ORC_B3_OUT); }

void orc_b3_oneshot(const void *data, size_t len, uint8_t out[ORC_B3_OUT]) {
    orc_b3 h;
    orc_b3_init(&h);
    orc_b3_update(&h, data, len);
    orc_b3_finalize(&h, out);
}
