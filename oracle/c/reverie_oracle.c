/* oracle/c/reverie_oracle.c -- CPU oracle #2 (see reverie_oracle.h).  TEST INFRASTRUCTURE ONLY.
 *
 * A C restatement that keeps the reference's *dataflow*, so that timing it is a fair stand-in for the Rust binary
 * ("kind": "port" in bench.py): one sequential gate interpreter per packed instance (8 reps x 8 players in a u64),
 * AES-NI CTR in 16-byte calls, the AVX2 movemask 64x8 bit transpose, 128-share refill batches, 64 KiB-buffered
 * BLAKE3 per repetition lane, recorded transcripts + a second extraction pass, threads over the 32 instances.
 * Every function cites the reference file:line (relative to /root/reference) it follows.
 */
#define _GNU_SOURCE
#include "reverie_oracle.h"

#include <immintrin.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "orc_blake3.h"

/* ---- src/lib.rs:17-38 ------------------------------------------------------------------------------------- */
#define PLAYERS 8
#define PACKED 8
#define BATCH_SIZE 128
#define ONLINE_REPS 40
#define TOTAL_REPS 256
#define PREPROCESSING_REPS (TOTAL_REPS - ONLINE_REPS)
#define PACKED_REPS (TOTAL_REPS / PACKED)
#define KEY_SIZE 16
#define HASH_SIZE 32
#define LSB8 0x0101010101010101ull

enum { D_GF2 = 0, D_Z64 = 1, D_B2A = 2, D_HINT = 3 };
enum { OP_INPUT, OP_RANDOM, OP_ADD, OP_ADDC, OP_SUB, OP_SUBC, OP_MUL, OP_MULC, OP_ASSERT, OP_CONST };
enum { M_PROVER, M_ONLINE, M_PRE };

/* ============================================================================================================
 *  PRG: src/crypto/prg.rs:13-37 -- ctr::Ctr128BE<Aes128>, IV = 0, keystream only
 * ========================================================================================================== */
typedef struct { __m128i rk[11]; uint64_t ctr; } prg_t;

static inline __m128i key_step(__m128i k, __m128i assist) {
    assist = _mm_shuffle_epi32(assist, 0xff);
    k = _mm_xor_si128(k, _mm_slli_si128(k, 4));
    k = _mm_xor_si128(k, _mm_slli_si128(k, 4));
    k = _mm_xor_si128(k, _mm_slli_si128(k, 4));
    return _mm_xor_si128(k, assist);
}
static void prg_init(prg_t *p, const uint8_t key[16]) {
    __m128i k = _mm_loadu_si128((const __m128i *)key);
    p->rk[0] = k;
#define KS(i, rc) k = key_step(k, _mm_aeskeygenassist_si128(k, rc)); p->rk[i] = k;
    KS(1, 0x01) KS(2, 0x02) KS(3, 0x04) KS(4, 0x08) KS(5, 0x10) KS(6, 0x20) KS(7, 0x40) KS(8, 0x80) KS(9, 0x1b) KS(10, 0x36)
#undef KS
    p->ctr = 0;
}
static inline void prg_block(prg_t *p, uint8_t out[16]) {
    /* 128-bit big-endian counter; the high 64 bits stay zero for any stream this code can reach */
    __m128i b = _mm_set_epi64x((long long)__builtin_bswap64(p->ctr++), 0);
    b = _mm_xor_si128(b, p->rk[0]);
    for (int r = 1; r < 10; r++) b = _mm_aesenc_si128(b, p->rk[r]);
    b = _mm_aesenclast_si128(b, p->rk[10]);
    _mm_storeu_si128((__m128i *)out, b);
}
static void prg_gen(prg_t *p, uint8_t *dst, size_t n_blocks) {
    for (size_t i = 0; i < n_blocks; i++) prg_block(p, dst + 16 * i);
}

void orc_aes128_ctr(const uint8_t key[16], uint64_t first_block, uint8_t *out, size_t n_blocks) {
    prg_t p;
    prg_init(&p, key);
    p.ctr = first_block;
    prg_gen(&p, out, n_blocks);
}

/* src/transcript/mod.rs:99-106 */
static void expand_seed(const uint8_t seed[16], uint8_t keys[PLAYERS][KEY_SIZE]) {
    prg_t p;
    prg_init(&p, seed);
    for (int i = 0; i < PLAYERS; i++) prg_block(&p, keys[i]);
}

/* ============================================================================================================
 *  Hashing: src/crypto/hash.rs
 * ========================================================================================================== */
#define BUFFER_SIZE (1u << 16)
#define BUFFER_SLACK 128
typedef struct { orc_b3 h; uint8_t *buf; size_t len; } hasher_t; /* BufferedHasher, hash.rs:17-58 */

static void hasher_init(hasher_t *h) { orc_b3_init(&h->h); h->buf = (uint8_t *)malloc(BUFFER_SIZE + BUFFER_SLACK); h->len = 0; }
static void hasher_drop(hasher_t *h) { free(h->buf); h->buf = NULL; }
static inline void hasher_push(hasher_t *h, uint8_t v) { /* hash.rs:36-42 */
    h->buf[h->len++] = v;
    if (h->len >= BUFFER_SIZE) { orc_b3_update(&h->h, h->buf, h->len); h->len = 0; }
}
static inline void hasher_update(hasher_t *h, const void *b, size_t n) { /* hash.rs:44-51 */
    memcpy(h->buf + h->len, b, n);
    h->len += n;
    if (h->len >= BUFFER_SIZE) { orc_b3_update(&h->h, h->buf, h->len); h->len = 0; }
}
static void hasher_finalize(const hasher_t *h, uint8_t out[32]) { /* hash.rs:53-57 */
    orc_b3 c = h->h;
    orc_b3_update(&c, h->buf, h->len);
    orc_b3_finalize(&c, out);
}
static void hash2(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) { /* HASH!(a,b), hash.rs:118-127 */
    uint8_t cat[64];
    memcpy(cat, a, 32);
    memcpy(cat + 32, b, 32);
    orc_b3_oneshot(cat, 64, out);
}
void orc_blake3(const uint8_t *data, size_t len, uint8_t *out, size_t out_len) {
    orc_b3 h;
    orc_b3_init(&h);
    orc_b3_update(&h, data, len);
    orc_b3_finalize_xof(&h, 0, out, out_len);
}

/* growable arrays standing in for Vec<T> */
typedef struct { uint8_t *p; size_t len, cap; } vec_t;
static void vec_reserve(vec_t *v, size_t extra) {
    if (v->len + extra <= v->cap) return;
    size_t nc = v->cap ? v->cap * 2 : 256;
    while (nc < v->len + extra) nc *= 2;
    v->p = (uint8_t *)realloc(v->p, nc);
    v->cap = nc;
}
static inline void vec_push(vec_t *v, const void *e, size_t n) { vec_reserve(v, n); memcpy(v->p + v->len, e, n); v->len += n; }
static inline void vec_push_u8(vec_t *v, uint8_t b) { vec_reserve(v, 1); v->p[v->len++] = b; }
static void vec_free(vec_t *v) { free(v->p); v->p = NULL; v->len = v->cap = 0; }

/* ============================================================================================================
 *  GF(2) domain: src/algebra/gf2/{domain,share,recon,batch}.rs
 * ========================================================================================================== */
static inline uint64_t gf2_reconstruct(uint64_t t) { /* gf2/domain.rs:47-63 */
    t ^= t >> 4; t ^= t >> 2; t ^= t >> 1; t &= LSB8;
    t |= t << 1; t |= t << 2; t |= t << 4;
    return t;
}

/* byte_to_shares_avx2, gf2/domain.rs:293-378: src[8r+p] -> bit 63-(8r+p), MSB of each byte first */
static inline void byte_to_shares(uint64_t dst[8], const uint8_t src[64]) {
    uint8_t rev[64];
    for (int k = 0; k < 32; k++) { rev[31 - k] = src[k]; rev[32 + 31 - k] = src[32 + k]; } /* the _mm256_set_epi8 gathers */
    __m256i fst = _mm256_loadu_si256((const __m256i *)rev), snd = _mm256_loadu_si256((const __m256i *)(rev + 32));
    for (int j = 0; j < 8; j++) {
        uint64_t top = (uint32_t)_mm256_movemask_epi8(fst), bot = (uint32_t)_mm256_movemask_epi8(snd);
        fst = _mm256_add_epi8(fst, fst);
        snd = _mm256_add_epi8(snd, snd);
        dst[j] = (top << 32) | bot;
    }
}

typedef struct { /* ShareGen<GF2> + 8 BatchGen, generator/share.rs:8-65, generator/batch.rs:6-40 */
    prg_t prgs[PACKED][PLAYERS];
    uint8_t omit[PACKED];
    uint8_t batches[PACKED][PLAYERS][16];
    uint64_t shares[BATCH_SIZE];
    size_t next_idx;
} gen_gf2_t;

static void gen_gf2_init(gen_gf2_t *g, const uint8_t keys[PACKED][PLAYERS][KEY_SIZE], const uint8_t omit[PACKED]) {
    memset(g->batches, 0, sizeof g->batches);
    for (int r = 0; r < PACKED; r++) {
        g->omit[r] = omit[r];
        for (int p = 0; p < PLAYERS; p++) prg_init(&g->prgs[r][p], keys[r][p]);
    }
    g->next_idx = BATCH_SIZE;
}
static inline uint64_t gen_gf2_next(gen_gf2_t *g) { /* generator/share.rs:54-65 */
    if (g->next_idx >= BATCH_SIZE) {
        for (int r = 0; r < PACKED; r++)
            for (int p = 0; p < PLAYERS; p++)
                if (p != g->omit[r]) prg_gen(&g->prgs[r][p], g->batches[r][p], 1); /* gf2/batch.rs:17-21 */
        for (int i = 0; i < 16; i++) { /* batches_to_shares_x86, gf2/domain.rs:85-173 */
            uint8_t src[64];
            for (int r = 0; r < PACKED; r++)
                for (int p = 0; p < PLAYERS; p++) src[8 * r + p] = g->batches[r][p][i];
            byte_to_shares(&g->shares[8 * i], src);
        }
        g->next_idx = 0;
    }
    return g->shares[g->next_idx++];
}

/* ============================================================================================================
 *  Z64 domain: src/algebra/z64/{domain,share,recon,batch}.rs
 * ========================================================================================================== */
typedef struct { uint64_t v[PACKED][PLAYERS]; } zshare_t;
typedef struct { uint64_t v[PACKED]; } zrecon_t;

static inline zrecon_t z_reconstruct(const zshare_t *s) { /* z64/domain.rs:53-61 */
    zrecon_t r;
    for (int i = 0; i < PACKED; i++) { uint64_t a = 0; for (int j = 0; j < PLAYERS; j++) a += s->v[i][j]; r.v[i] = a; }
    return r;
}

typedef struct { /* ShareGen<Z64> */
    prg_t prgs[PACKED][PLAYERS];
    uint8_t omit[PACKED];
    uint64_t (*batches)[PLAYERS][BATCH_SIZE]; /* [PACKED][PLAYERS][128], z64/batch.rs:12-15 */
    zshare_t *shares;                          /* [128] */
    size_t next_idx;
} gen_z64_t;

static void gen_z64_init(gen_z64_t *g, const uint8_t keys[PACKED][PLAYERS][KEY_SIZE], const uint8_t omit[PACKED]) {
    g->batches = calloc(PACKED, sizeof *g->batches);
    g->shares = calloc(BATCH_SIZE, sizeof *g->shares);
    for (int r = 0; r < PACKED; r++) {
        g->omit[r] = omit[r];
        for (int p = 0; p < PLAYERS; p++) prg_init(&g->prgs[r][p], keys[r][p]);
    }
    g->next_idx = BATCH_SIZE;
}
static void gen_z64_drop(gen_z64_t *g) { free(g->batches); free(g->shares); }
static inline const zshare_t *gen_z64_next(gen_z64_t *g) {
    if (g->next_idx >= BATCH_SIZE) {
        for (int r = 0; r < PACKED; r++)
            for (int p = 0; p < PLAYERS; p++)
                if (p != g->omit[r]) prg_gen(&g->prgs[r][p], (uint8_t *)g->batches[r][p], 64); /* z64/batch.rs:25-30 */
        for (int i = 0; i < BATCH_SIZE; i++) /* z64/domain.rs:64-83 */
            for (int r = 0; r < PACKED; r++)
                for (int p = 0; p < PLAYERS; p++) g->shares[i].v[r][p] = g->batches[r][p][i];
        g->next_idx = 0;
    }
    return &g->shares[g->next_idx++];
}

/* ============================================================================================================
 *  Openings: src/proof/mod.rs:40-66
 * ========================================================================================================== */
typedef struct { uint8_t omit; uint8_t seeds[PLAYERS][KEY_SIZE]; const uint8_t *recons, *corrs, *inputs; size_t n_recons, n_corrs, n_inputs; } open_online_t;
typedef struct { uint8_t seed[KEY_SIZE]; uint8_t comm_online[HASH_SIZE]; } open_pre_t;

/* ============================================================================================================
 *  Transcripts: src/transcript/{prover.rs, verifier/online.rs, verifier/preprocess.rs} -- one struct, three modes
 * ========================================================================================================== */
typedef struct {
    int mode, err, okay;
    uint8_t seeds[PACKED][KEY_SIZE];
    gen_gf2_t gen;
    hasher_t h_on[PACKED], h_pre[PACKED];
    uint8_t comms_online[PACKED][HASH_SIZE];          /* M_PRE */
    const uint8_t *wit; size_t n_wit, i_wit;          /* M_PROVER */
    vec_t reconstructions, corrections, inputs;       /* M_PROVER: Vec<u64> each, prover.rs:29-31 */
    uint64_t *v_recons, *v_corrs, *v_inputs;          /* M_ONLINE: unpacked proof data */
    size_t nv_recons, nv_corrs, nv_inputs, i_recons, i_corrs, i_inputs;
} tr_gf2_t;

typedef struct {
    int mode, err, okay;
    uint8_t seeds[PACKED][KEY_SIZE];
    gen_z64_t gen;
    hasher_t h_on[PACKED], h_pre[PACKED];
    uint8_t comms_online[PACKED][HASH_SIZE];
    const uint64_t *wit; size_t n_wit, i_wit;
    vec_t reconstructions /* zshare_t */, corrections /* zrecon_t */, inputs /* zrecon_t */;
    zshare_t *v_recons; zrecon_t *v_corrs, *v_inputs;
    size_t nv_recons, nv_corrs, nv_inputs, i_recons, i_corrs, i_inputs;
} tr_z64_t;

static inline void gf2_hash_word(uint64_t w, hasher_t *h) { /* gf2/share.rs:211-218 (push), gf2/recon.rs:314-321 (update) */
    for (int i = 0; i < PACKED; i++) hasher_push(&h[i], (uint8_t)(w >> (56 - 8 * i)));
}
static inline void z_hash_share(const zshare_t *s, hasher_t *h) { /* z64/share.rs:100-108 */
    for (int i = 0; i < PACKED; i++)
        for (int j = 0; j < PLAYERS; j++) hasher_update(&h[i], &s->v[i][j], 8);
}
static inline void z_hash_recon(const zrecon_t *r, hasher_t *h) { /* z64/recon.rs:131-137 */
    for (int i = 0; i < PACKED; i++) hasher_update(&h[i], &r->v[i], 8);
}

static void keys_from_seeds(const uint8_t seeds[PACKED][KEY_SIZE], uint8_t keys[PACKED][PLAYERS][KEY_SIZE]) {
    for (int r = 0; r < PACKED; r++) expand_seed(seeds[r], keys[r]); /* share_gen_from_rep_seeds, transcript/mod.rs:108-122 */
}

/* ---- GF2 unpack: gf2/recon.rs:151-167,241-259 and gf2/share.rs:151-208 ---- */
static int gf2_unpack_recon(const uint8_t *src[PACKED], const size_t len[PACKED], uint64_t **out, size_t *n) {
    size_t bytes = len[0];
    for (int r = 0; r < PACKED; r++) if (len[r] < bytes) return ORC_E_FORMAT; /* index panic in the reference */
    uint64_t *o = (uint64_t *)malloc(8 * (bytes * 8 + 1));
    for (size_t i = 0; i < bytes; i++)
        for (int bit = 0; bit < 8; bit++) {
            uint64_t v = 0;
            for (int r = 0; r < PACKED; r++) if ((src[r][i] >> (7 - bit)) & 1) v |= 0xffull << (8 * (7 - r));
            o[8 * i + bit] = v;
        }
    *out = o; *n = bytes * 8;
    return ORC_OK;
}
static int gf2_unpack_selected(const uint8_t *src[PACKED], const size_t len[PACKED], const uint8_t omit[PACKED], uint64_t **out, size_t *n) {
    size_t bytes = len[0];
    for (int r = 0; r < PACKED; r++) if (len[r] != bytes) return ORC_E_FORMAT; /* assert_eq!, gf2/share.rs:158-164 */
    uint64_t *o = (uint64_t *)malloc(8 * (bytes * 8 + 1));
    uint8_t tmp[64];
    memset(tmp, 0, 64);
    for (size_t i = 0; i < bytes; i++) {
        for (int j = 0; j < PACKED; j++) tmp[omit[j] + PLAYERS * j] = src[j][i];
        byte_to_shares(&o[8 * i], tmp);
    }
    *out = o; *n = bytes * 8;
    return ORC_OK;
}

static void tr_gf2_common_init(tr_gf2_t *t, int mode) {
    memset(t, 0, sizeof *t);
    t->mode = mode;
    t->okay = 1;
    for (int i = 0; i < PACKED; i++) { hasher_init(&t->h_on[i]); hasher_init(&t->h_pre[i]); }
}
static void tr_gf2_init_prover(tr_gf2_t *t, const uint8_t seeds[PACKED][KEY_SIZE], const uint8_t *wit, size_t n_wit) { /* prover.rs:36-53 */
    tr_gf2_common_init(t, M_PROVER);
    memcpy(t->seeds, seeds, PACKED * KEY_SIZE);
    uint8_t keys[PACKED][PLAYERS][KEY_SIZE], omit[PACKED];
    keys_from_seeds(seeds, keys);
    memset(omit, PLAYERS, PACKED);
    gen_gf2_init(&t->gen, keys, omit);
    t->wit = wit; t->n_wit = n_wit;
}
static int tr_gf2_init_online(tr_gf2_t *t, const open_online_t o[PACKED]) { /* verifier/online.rs:25-121 */
    tr_gf2_common_init(t, M_ONLINE);
    const uint8_t *src[PACKED]; size_t len[PACKED]; uint8_t omit[PACKED];
    uint8_t keys[PACKED][PLAYERS][KEY_SIZE];
    for (int r = 0; r < PACKED; r++) { omit[r] = o[r].omit; if (omit[r] >= PLAYERS) return ORC_E_FORMAT; memcpy(keys[r], o[r].seeds, PLAYERS * KEY_SIZE); }
    int e;
    for (int r = 0; r < PACKED; r++) { src[r] = o[r].corrs; len[r] = o[r].n_corrs; }
    if ((e = gf2_unpack_recon(src, len, &t->v_corrs, &t->nv_corrs))) return e;
    for (int r = 0; r < PACKED; r++) { src[r] = o[r].inputs; len[r] = o[r].n_inputs; }
    if ((e = gf2_unpack_recon(src, len, &t->v_inputs, &t->nv_inputs))) return e;
    for (int r = 0; r < PACKED; r++) { src[r] = o[r].recons; len[r] = o[r].n_recons; }
    if ((e = gf2_unpack_selected(src, len, omit, &t->v_recons, &t->nv_recons))) return e;
    gen_gf2_init(&t->gen, keys, omit);
    return ORC_OK;
}
static void tr_gf2_init_pre(tr_gf2_t *t, const open_pre_t p[PACKED]) { /* verifier/preprocess.rs:17-43 */
    tr_gf2_common_init(t, M_PRE);
    uint8_t seeds[PACKED][KEY_SIZE], keys[PACKED][PLAYERS][KEY_SIZE], omit[PACKED];
    for (int r = 0; r < PACKED; r++) { memcpy(seeds[r], p[r].seed, KEY_SIZE); memcpy(t->comms_online[r], p[r].comm_online, HASH_SIZE); }
    keys_from_seeds(seeds, keys);
    memset(omit, PLAYERS, PACKED);
    gen_gf2_init(&t->gen, keys, omit);
}
static void tr_gf2_drop(tr_gf2_t *t) {
    for (int i = 0; i < PACKED; i++) { hasher_drop(&t->h_on[i]); hasher_drop(&t->h_pre[i]); }
    vec_free(&t->reconstructions); vec_free(&t->corrections); vec_free(&t->inputs);
    free(t->v_recons); free(t->v_corrs); free(t->v_inputs);
}

typedef struct { uint64_t mask, corr; } gwire_t; /* Wire<GF2>, interpreter/mod.rs:9-13 */

static inline gwire_t tr_gf2_input(tr_gf2_t *t) {
    gwire_t w;
    if (t->mode == M_PROVER) { /* prover.rs:181-199 */
        w.mask = gen_gf2_next(&t->gen);
        uint64_t lambda = gf2_reconstruct(w.mask);
        uint64_t in = 0;
        if (t->i_wit < t->n_wit) in = t->wit[t->i_wit++] ? ~0ull : 0; else t->err = ORC_E_WITNESS_SHORT;
        w.corr = in ^ lambda;
        gf2_hash_word(w.corr, t->h_on);
        vec_push(&t->inputs, &w.corr, 8);
    } else if (t->mode == M_ONLINE) { /* online.rs:123-130 */
        w.corr = t->i_inputs < t->nv_inputs ? t->v_inputs[t->i_inputs] : 0;
        t->i_inputs++;
        gf2_hash_word(w.corr, t->h_on);
        w.mask = gen_gf2_next(&t->gen);
    } else { /* preprocess.rs:46-52 */
        w.mask = gen_gf2_next(&t->gen);
        w.corr = 0;
    }
    return w;
}
static inline uint64_t tr_gf2_reconstruct(tr_gf2_t *t, uint64_t mask) {
    if (t->mode == M_PROVER) { /* prover.rs:209-213 */
        gf2_hash_word(mask, t->h_on);
        vec_push(&t->reconstructions, &mask, 8);
        return gf2_reconstruct(mask);
    } else if (t->mode == M_ONLINE) { /* online.rs:140-167 */
        uint64_t msg = t->i_recons < t->nv_recons ? t->v_recons[t->i_recons] : 0;
        t->i_recons++;
        mask ^= msg;
        gf2_hash_word(mask, t->h_on);
        return gf2_reconstruct(mask);
    }
    return 0; /* preprocess.rs:62-64 */
}
static inline uint64_t tr_gf2_correction(tr_gf2_t *t, uint64_t corr) {
    if (t->mode == M_PROVER) { /* prover.rs:215-219 */
        gf2_hash_word(corr, t->h_pre);
        vec_push(&t->corrections, &corr, 8);
        return corr;
    } else if (t->mode == M_ONLINE) { /* online.rs:169-174 */
        corr = t->i_corrs < t->nv_corrs ? t->v_corrs[t->i_corrs] : 0;
        t->i_corrs++;
        gf2_hash_word(corr, t->h_pre);
        return corr;
    }
    gf2_hash_word(corr, t->h_pre); /* preprocess.rs:66-69 */
    return corr;
}
static inline void tr_gf2_zero_check(tr_gf2_t *t, uint64_t recon) {
    if (t->mode == M_PROVER) { if (recon != 0 && !t->err) t->err = ORC_E_WITNESS_INVALID; } /* prover.rs:221-228 */
    else if (t->mode == M_ONLINE) t->okay &= (recon == 0);                                   /* online.rs:176-178 */
}
static void tr_hash_join(hasher_t *h_on, hasher_t *h_pre, const uint8_t comms_online[PACKED][HASH_SIZE], int mode, uint8_t out[PACKED][HASH_SIZE]) {
    /* Transcript::hash, transcript/mod.rs:77-96: H(preprocess || online) per rep */
    for (int i = 0; i < PACKED; i++) {
        uint8_t on[32], pre[32];
        if (mode == M_PRE) memcpy(on, comms_online[i], 32); else hasher_finalize(&h_on[i], on);
        hasher_finalize(&h_pre[i], pre);
        hash2(pre, on, out[i]);
    }
}

/* ---- Z64 transcript ---- */
static void tr_z64_common_init(tr_z64_t *t, int mode) {
    memset(t, 0, sizeof *t);
    t->mode = mode;
    t->okay = 1;
    for (int i = 0; i < PACKED; i++) { hasher_init(&t->h_on[i]); hasher_init(&t->h_pre[i]); }
}
static void tr_z64_init_prover(tr_z64_t *t, const uint8_t seeds[PACKED][KEY_SIZE], const uint64_t *wit, size_t n_wit) {
    tr_z64_common_init(t, M_PROVER);
    memcpy(t->seeds, seeds, PACKED * KEY_SIZE);
    uint8_t keys[PACKED][PLAYERS][KEY_SIZE], omit[PACKED];
    keys_from_seeds(seeds, keys); /* SAME seeds as the GF2 instance, proof/mod.rs:138,144 */
    memset(omit, PLAYERS, PACKED);
    gen_z64_init(&t->gen, keys, omit);
    t->wit = wit; t->n_wit = n_wit;
}
static int tr_z64_init_online(tr_z64_t *t, const open_online_t o[PACKED]) {
    tr_z64_common_init(t, M_ONLINE);
    uint8_t omit[PACKED], keys[PACKED][PLAYERS][KEY_SIZE];
    for (int r = 0; r < PACKED; r++) { omit[r] = o[r].omit; if (omit[r] >= PLAYERS) return ORC_E_FORMAT; memcpy(keys[r], o[r].seeds, PLAYERS * KEY_SIZE); }
    /* z64/recon.rs:68-107 and z64/share.rs:51-91: element count from rep 0; missing 8-byte chunks elsewhere read as zero */
    size_t n;
    n = o[0].n_corrs / 8; t->v_corrs = calloc(n + 1, sizeof(zrecon_t)); t->nv_corrs = n;
    for (size_t k = 0; k < n; k++) for (int r = 0; r < PACKED; r++) if (8 * k + 8 <= o[r].n_corrs) memcpy(&t->v_corrs[k].v[r], o[r].corrs + 8 * k, 8);
    n = o[0].n_inputs / 8; t->v_inputs = calloc(n + 1, sizeof(zrecon_t)); t->nv_inputs = n;
    for (size_t k = 0; k < n; k++) for (int r = 0; r < PACKED; r++) if (8 * k + 8 <= o[r].n_inputs) memcpy(&t->v_inputs[k].v[r], o[r].inputs + 8 * k, 8);
    n = o[0].n_recons / 8; t->v_recons = calloc(n + 1, sizeof(zshare_t)); t->nv_recons = n;
    for (size_t k = 0; k < n; k++) for (int r = 0; r < PACKED; r++) if (8 * k + 8 <= o[r].n_recons) memcpy(&t->v_recons[k].v[r][omit[r]], o[r].recons + 8 * k, 8);
    gen_z64_init(&t->gen, keys, omit);
    return ORC_OK;
}
static void tr_z64_init_pre(tr_z64_t *t, const open_pre_t p[PACKED]) {
    tr_z64_common_init(t, M_PRE);
    uint8_t seeds[PACKED][KEY_SIZE], keys[PACKED][PLAYERS][KEY_SIZE], omit[PACKED];
    for (int r = 0; r < PACKED; r++) { memcpy(seeds[r], p[r].seed, KEY_SIZE); memcpy(t->comms_online[r], p[r].comm_online, HASH_SIZE); }
    keys_from_seeds(seeds, keys);
    memset(omit, PLAYERS, PACKED);
    gen_z64_init(&t->gen, keys, omit);
}
static void tr_z64_drop(tr_z64_t *t) {
    for (int i = 0; i < PACKED; i++) { hasher_drop(&t->h_on[i]); hasher_drop(&t->h_pre[i]); }
    vec_free(&t->reconstructions); vec_free(&t->corrections); vec_free(&t->inputs);
    free(t->v_recons); free(t->v_corrs); free(t->v_inputs);
    gen_z64_drop(&t->gen);
}

typedef struct { zshare_t mask; zrecon_t corr; } zwire_t;

static inline void tr_z64_input(tr_z64_t *t, zwire_t *w) {
    if (t->mode == M_PROVER) {
        w->mask = *gen_z64_next(&t->gen);
        zrecon_t lambda = z_reconstruct(&w->mask);
        uint64_t in = 0;
        if (t->i_wit < t->n_wit) in = t->wit[t->i_wit++]; else t->err = ORC_E_WITNESS_SHORT;
        for (int i = 0; i < PACKED; i++) w->corr.v[i] = in - lambda.v[i];
        z_hash_recon(&w->corr, t->h_on);
        vec_push(&t->inputs, &w->corr, sizeof(zrecon_t));
    } else if (t->mode == M_ONLINE) {
        if (t->i_inputs < t->nv_inputs) w->corr = t->v_inputs[t->i_inputs]; else memset(&w->corr, 0, sizeof w->corr);
        t->i_inputs++;
        z_hash_recon(&w->corr, t->h_on);
        w->mask = *gen_z64_next(&t->gen);
    } else {
        w->mask = *gen_z64_next(&t->gen);
        memset(&w->corr, 0, sizeof w->corr);
    }
}
static inline zrecon_t tr_z64_reconstruct(tr_z64_t *t, const zshare_t *mask) {
    zrecon_t zero;
    memset(&zero, 0, sizeof zero);
    if (t->mode == M_PROVER) {
        z_hash_share(mask, t->h_on);
        vec_push(&t->reconstructions, mask, sizeof(zshare_t));
        return z_reconstruct(mask);
    } else if (t->mode == M_ONLINE) {
        zshare_t m = *mask;
        if (t->i_recons < t->nv_recons) {
            const zshare_t *msg = &t->v_recons[t->i_recons];
            for (int i = 0; i < PACKED; i++) for (int j = 0; j < PLAYERS; j++) m.v[i][j] += msg->v[i][j];
        }
        t->i_recons++;
        z_hash_share(&m, t->h_on);
        return z_reconstruct(&m);
    }
    return zero;
}
static inline zrecon_t tr_z64_correction(tr_z64_t *t, zrecon_t corr) {
    if (t->mode == M_PROVER) {
        z_hash_recon(&corr, t->h_pre);
        vec_push(&t->corrections, &corr, sizeof corr);
        return corr;
    } else if (t->mode == M_ONLINE) {
        if (t->i_corrs < t->nv_corrs) corr = t->v_corrs[t->i_corrs]; else memset(&corr, 0, sizeof corr);
        t->i_corrs++;
        z_hash_recon(&corr, t->h_pre);
        return corr;
    }
    z_hash_recon(&corr, t->h_pre);
    return corr;
}
static inline void tr_z64_zero_check(tr_z64_t *t, const zrecon_t *r) {
    int z = 1;
    for (int i = 0; i < PACKED; i++) z &= (r->v[i] == 0);
    if (t->mode == M_PROVER) { if (!z && !t->err) t->err = ORC_E_WITNESS_INVALID; }
    else if (t->mode == M_ONLINE) t->okay &= z;
}

/* ============================================================================================================
 *  Interpreter: src/interpreter/{single,combine}.rs
 * ========================================================================================================== */
typedef struct {
    tr_gf2_t tg; tr_z64_t tz;
    gwire_t *gw; size_t n_gw;
    zwire_t *zw; size_t n_zw;
    int err;
} instance_t;

static inline gwire_t g_op_mul(tr_gf2_t *t, gwire_t w1, gwire_t w2) { /* single.rs:25-69 */
    uint64_t mask_ab = gen_gf2_next(&t->gen), mask_new = gen_gf2_next(&t->gen);
    uint64_t a = gf2_reconstruct(w1.mask), b = gf2_reconstruct(w2.mask), c = gf2_reconstruct(mask_ab);
    uint64_t delta = tr_gf2_correction(t, (a & b) ^ c);
    uint64_t s = (w2.mask & w1.corr) ^ (w1.mask & w2.corr) ^ mask_ab ^ mask_new;
    uint64_t recon = tr_gf2_reconstruct(t, s) ^ delta;
    gwire_t r = {mask_new, recon ^ (w1.corr & w2.corr)};
    return r;
}
static inline gwire_t g_op_add(gwire_t a, gwire_t b) { gwire_t r = {a.mask ^ b.mask, a.corr ^ b.corr}; return r; }

static void g_step(instance_t *I, const orc_op *op) { /* single.rs:106-157 over GF2 */
    tr_gf2_t *t = &I->tg;
    gwire_t *W = I->gw;
    uint64_t v = (op->imm & 1) ? ~0ull : 0; /* bool -> Recon, gf2/recon.rs:274-287 */
    switch (op->opcode) {
        case OP_INPUT: W[op->dst] = tr_gf2_input(t); break;
        case OP_ADD: case OP_SUB: W[op->dst] = g_op_add(W[op->a], W[op->b]); break;
        case OP_MUL: W[op->dst] = g_op_mul(t, W[op->a], W[op->b]); break;
        case OP_ADDC: case OP_SUBC: { gwire_t w = W[op->a]; w.corr ^= v; W[op->dst] = w; break; }
        case OP_MULC: { gwire_t w = W[op->a]; w.mask &= v; w.corr &= v; W[op->dst] = w; break; }
        case OP_ASSERT: { gwire_t w = W[op->a]; uint64_t m = tr_gf2_reconstruct(t, w.mask); tr_gf2_zero_check(t, w.corr ^ m); break; }
        case OP_RANDOM: { gwire_t w = {gen_gf2_next(&t->gen), 0}; W[op->dst] = w; break; }
        case OP_CONST: { gwire_t w = {0, v}; W[op->dst] = w; break; }
        default: I->err = ORC_E_ARG;
    }
}

static void z_op_mul(tr_z64_t *t, const zwire_t *w1, const zwire_t *w2, zwire_t *out) { /* single.rs:25-69 over Z64 */
    zshare_t mask_ab = *gen_z64_next(&t->gen), mask_new = *gen_z64_next(&t->gen);
    zrecon_t a = z_reconstruct(&w1->mask), b = z_reconstruct(&w2->mask), c = z_reconstruct(&mask_ab), d;
    for (int i = 0; i < PACKED; i++) d.v[i] = a.v[i] * b.v[i] - c.v[i];
    zrecon_t delta = tr_z64_correction(t, d);
    zshare_t s;
    for (int i = 0; i < PACKED; i++)
        for (int j = 0; j < PLAYERS; j++)
            s.v[i][j] = w2->mask.v[i][j] * w1->corr.v[i] + w1->mask.v[i][j] * w2->corr.v[i] + mask_ab.v[i][j] - mask_new.v[i][j];
    zrecon_t recon = tr_z64_reconstruct(t, &s);
    zwire_t r;
    r.mask = mask_new;
    for (int i = 0; i < PACKED; i++) r.corr.v[i] = recon.v[i] + delta.v[i] + w1->corr.v[i] * w2->corr.v[i];
    *out = r;
}

static void z_step(instance_t *I, const orc_op *op) {
    tr_z64_t *t = &I->tz;
    zwire_t *W = I->zw;
    uint64_t v = op->imm; /* u64 -> Recon broadcast, z64/recon.rs:123-129 */
    switch (op->opcode) {
        case OP_INPUT: { zwire_t w; tr_z64_input(t, &w); W[op->dst] = w; break; }
        case OP_ADD: { zwire_t r; const zwire_t *a = &W[op->a], *b = &W[op->b];
            for (int i = 0; i < PACKED; i++) { r.corr.v[i] = a->corr.v[i] + b->corr.v[i]; for (int j = 0; j < PLAYERS; j++) r.mask.v[i][j] = a->mask.v[i][j] + b->mask.v[i][j]; }
            W[op->dst] = r; break; }
        case OP_SUB: { zwire_t r; const zwire_t *a = &W[op->a], *b = &W[op->b];
            for (int i = 0; i < PACKED; i++) { r.corr.v[i] = a->corr.v[i] - b->corr.v[i]; for (int j = 0; j < PLAYERS; j++) r.mask.v[i][j] = a->mask.v[i][j] - b->mask.v[i][j]; }
            W[op->dst] = r; break; }
        case OP_MUL: { zwire_t r; z_op_mul(t, &W[op->a], &W[op->b], &r); W[op->dst] = r; break; }
        case OP_ADDC: { zwire_t r = W[op->a]; for (int i = 0; i < PACKED; i++) r.corr.v[i] += v; W[op->dst] = r; break; }
        case OP_SUBC: { zwire_t r = W[op->a]; for (int i = 0; i < PACKED; i++) r.corr.v[i] -= v; W[op->dst] = r; break; }
        case OP_MULC: { zwire_t r = W[op->a]; for (int i = 0; i < PACKED; i++) { r.corr.v[i] *= v; for (int j = 0; j < PLAYERS; j++) r.mask.v[i][j] *= v; } W[op->dst] = r; break; }
        case OP_ASSERT: { const zwire_t *w = &W[op->a]; zrecon_t m = tr_z64_reconstruct(t, &w->mask); for (int i = 0; i < PACKED; i++) m.v[i] += w->corr.v[i]; tr_z64_zero_check(t, &m); break; }
        case OP_RANDOM: { zwire_t r; r.mask = *gen_z64_next(&t->gen); memset(&r.corr, 0, sizeof r.corr); W[op->dst] = r; break; }
        case OP_CONST: { zwire_t r; memset(&r.mask, 0, sizeof r.mask); for (int i = 0; i < PACKED; i++) r.corr.v[i] = v; W[op->dst] = r; break; }
        default: I->err = ORC_E_ARG;
    }
}

/* combine.rs:19-36: with `recon` either the plain reconstruct (transcript == NULL) or transcript.reconstruct */
static zrecon_t recon_gf2_to_z64(tr_gf2_t *t, const gwire_t bits[64]) {
    zrecon_t z;
    memset(&z, 0, sizeof z);
    for (int k = 0; k < 64; k++) {
        uint64_t r = t ? tr_gf2_reconstruct(t, bits[k].mask) : gf2_reconstruct(bits[k].mask);
        uint64_t g = (r ^ bits[k].corr) & LSB8;
        for (int j = 0; j < PACKED; j++) z.v[j] = (z.v[j] << 1) | ((g >> (56 - 8 * j)) & 0xff);
    }
    for (int j = 0; j < PACKED; j++) { /* reverse_bits */
        uint64_t x = z.v[j], y = 0;
        for (int b = 0; b < 64; b++) { y = (y << 1) | (x & 1); x >>= 1; }
        z.v[j] = y;
    }
    return z;
}

static void b2a_step(instance_t *I, const orc_op *op) { /* combine.rs:132-219, add_64 :39-93 */
    tr_gf2_t *tg = &I->tg;
    tr_z64_t *tz = &I->tz;
    gwire_t a[64], res[64];
    const gwire_t *b = &I->gw[op->a];
    for (int k = 0; k < 64; k++) { a[k].mask = gen_gf2_next(&tg->gen); a[k].corr = 0; }
    zrecon_t z64_value = recon_gf2_to_z64(NULL, a);
    zshare_t z64_mask = *gen_z64_next(&tz->gen);
    zrecon_t mr = z_reconstruct(&z64_mask), zc;
    for (int i = 0; i < PACKED; i++) zc.v[i] = z64_value.v[i] - mr.v[i];
    zc = tr_z64_correction(tz, zc);
    gwire_t carry = g_op_mul(tg, a[0], b[0]);
    res[0] = g_op_add(a[0], b[0]);
    for (int i = 1; i < 63; i++) {
        gwire_t ac = g_op_add(a[i], carry), bc = g_op_add(b[i], carry);
        gwire_t ac_bc = g_op_mul(tg, ac, bc);
        res[i] = g_op_add(ac, b[i]);
        carry = g_op_add(ac_bc, carry);
    }
    res[63] = g_op_add(carry, g_op_add(a[63], b[63]));
    zrecon_t z64_recon = recon_gf2_to_z64(tg, res);
    zwire_t out;
    for (int i = 0; i < PACKED; i++) {
        out.corr.v[i] = z64_recon.v[i] - zc.v[i];
        for (int j = 0; j < PLAYERS; j++) out.mask.v[i][j] = 0 - z64_mask.v[i][j];
    }
    I->zw[op->dst] = out;
}

static int check_op(const instance_t *I, const orc_op *op) { /* the reference would panic on an out-of-range wire */
    size_t ng = I->n_gw, nz = I->n_zw;
    switch (op->domain) {
        case D_GF2: case D_Z64: {
            size_t n = op->domain == D_GF2 ? ng : nz;
            switch (op->opcode) {
                case OP_INPUT: case OP_RANDOM: case OP_CONST: return op->dst < n;
                case OP_ADD: case OP_SUB: case OP_MUL: return op->dst < n && op->a < n && op->b < n;
                case OP_ADDC: case OP_SUBC: case OP_MULC: return op->dst < n && op->a < n;
                case OP_ASSERT: return op->a < n;
                default: return 0;
            }
        }
        case D_B2A: return op->dst < nz && (size_t)op->a + 64 <= ng;
        case D_HINT: return 1;
    }
    return 0;
}

static void run_circuit(instance_t *I, const orc_op *ops, size_t n_ops) { /* proof/mod.rs:149-152; CombineInstance::step combine.rs:120-221 */
    for (size_t k = 0; k < n_ops; k++) {
        const orc_op *op = &ops[k];
        if (op->domain == D_HINT) {
            if (I->n_zw < op->a) { I->zw = realloc(I->zw, sizeof(zwire_t) * op->a); memset(I->zw + I->n_zw, 0, sizeof(zwire_t) * (op->a - I->n_zw)); I->n_zw = op->a; }
            if (I->n_gw < op->b) { I->gw = realloc(I->gw, sizeof(gwire_t) * op->b); memset(I->gw + I->n_gw, 0, sizeof(gwire_t) * (op->b - I->n_gw)); I->n_gw = op->b; }
            continue;
        }
        if (!check_op(I, op)) { I->err = ORC_E_ARG; return; }
        if (op->domain == D_GF2) g_step(I, op);
        else if (op->domain == D_Z64) z_step(I, op);
        else b2a_step(I, op);
    }
}

static void instance_alloc(instance_t *I, size_t z64_cells, size_t gf2_cells) {
    I->n_gw = gf2_cells; I->n_zw = z64_cells;
    I->gw = calloc(gf2_cells ? gf2_cells : 1, sizeof(gwire_t));
    I->zw = calloc(z64_cells ? z64_cells : 1, sizeof(zwire_t));
    I->err = 0;
}
static void instance_hash(instance_t *I, uint8_t out[PACKED][HASH_SIZE]) { /* CombineInstance::hash, combine.rs:104-118 */
    uint8_t g[PACKED][HASH_SIZE], z[PACKED][HASH_SIZE];
    tr_hash_join(I->tg.h_on, I->tg.h_pre, I->tg.comms_online, I->tg.mode, g);
    tr_hash_join(I->tz.h_on, I->tz.h_pre, I->tz.comms_online, I->tz.mode, z);
    for (int i = 0; i < PACKED; i++) hash2(g[i], z[i], out[i]);
}
static void instance_drop(instance_t *I) { tr_gf2_drop(&I->tg); tr_z64_drop(&I->tz); free(I->gw); free(I->zw); }

/* ============================================================================================================
 *  Fiat-Shamir: src/proof/mod.rs:68-108, src/crypto/ro.rs:7-20
 * ========================================================================================================== */
void orc_challenge(const uint8_t comm[32], uint8_t omit_of_rep[TOTAL_REPS]) {
    orc_b3 h;
    orc_b3_init(&h);
    orc_b3_update(&h, "random-oracle challenge", 23); /* CTX_CHALLENGE, proof/mod.rs:18 */
    uint8_t zero = 0;
    orc_b3_update(&h, &zero, 1);
    orc_b3_update(&h, comm, 32);
    memset(omit_of_rep, PLAYERS, TOTAL_REPS);
    int distinct = 0;
    uint64_t pos = 0;
    while (distinct < ONLINE_REPS) { /* HashMap::insert overwrites, proof/mod.rs:78-81 */
        uint8_t buf[32];
        orc_b3_finalize_xof(&h, pos, buf, 32);
        pos += 32;
        unsigned rep = buf[0];        /* u128 LE mod 256 */
        unsigned omit = buf[16] & 7;  /* u128 LE mod 8   */
        if (omit_of_rep[rep] == PLAYERS) distinct++;
        omit_of_rep[rep] = (uint8_t)omit;
    }
}

/* ============================================================================================================
 *  Extraction: src/transcript/prover.rs:57-175 and the pack functions
 * ========================================================================================================== */
typedef struct { vec_t recons[PACKED], corrs[PACKED], inputs[PACKED]; } packed_out_t;

static inline uint8_t pack8_bits(const uint64_t *arr, unsigned shift) { /* gf2/share.rs:66-85 and gf2/recon.rs:127-148: first element -> MSB */
    unsigned r = 0;
    for (int k = 0; k < 8; k++) r = (r << 1) | (unsigned)((arr[k] >> shift) & 1);
    return (uint8_t)r;
}
static void gf2_pack_bits(vec_t dst[PACKED], const uint64_t *src, size_t n, const unsigned shift[PACKED], const uint8_t sel[PACKED]) {
    /* shared shape of ShareGF2::pack_selected (gf2/share.rs:87-149) and ReconGF2::pack (gf2/recon.rs:190-239):
       groups of 8 elements -> 1 byte per selected rep; the residue group is ALWAYS flushed => floor(n/8)+1 bytes */
    int any = 0;
    for (int r = 0; r < PACKED; r++) any |= sel[r];
    if (!any) return;
    size_t full = n / 8;
    for (size_t c = 0; c < full; c++)
        for (int r = 0; r < PACKED; r++) if (sel[r]) vec_push_u8(&dst[r], pack8_bits(src + 8 * c, shift[r]));
    uint64_t arr[8] = {0};
    for (size_t k = 0; k < n % 8; k++) arr[k] = src[8 * full + k];
    for (int r = 0; r < PACKED; r++) if (sel[r]) vec_push_u8(&dst[r], pack8_bits(arr, shift[r]));
}

static void extract_gf2(const tr_gf2_t *t, const uint8_t players[PACKED], packed_out_t *o) {
    unsigned sh_share[PACKED], sh_recon[PACKED];
    uint8_t sel[PACKED];
    for (int r = 0; r < PACKED; r++) {
        sel[r] = players[r] < PLAYERS;
        sh_share[r] = sel[r] ? (unsigned)((PACKED - 1 - r) * PLAYERS + (PLAYERS - 1 - players[r])) : 0;
        sh_recon[r] = (unsigned)(64 - (r + 1) * 8);
    }
    gf2_pack_bits(o->recons, (const uint64_t *)t->reconstructions.p, t->reconstructions.len / 8, sh_share, sel);
    gf2_pack_bits(o->corrs, (const uint64_t *)t->corrections.p, t->corrections.len / 8, sh_recon, sel);
    gf2_pack_bits(o->inputs, (const uint64_t *)t->inputs.p, t->inputs.len / 8, sh_recon, sel);
}
static void extract_z64(const tr_z64_t *t, const uint8_t players[PACKED], packed_out_t *o) {
    /* z64/share.rs:37-49, z64/recon.rs:46-66 */
    const zshare_t *rs = (const zshare_t *)t->reconstructions.p;
    size_t n = t->reconstructions.len / sizeof(zshare_t);
    for (size_t k = 0; k < n; k++)
        for (int r = 0; r < PACKED; r++) if (players[r] < PLAYERS) vec_push(&o->recons[r], &rs[k].v[r][players[r]], 8);
    const zrecon_t *cs = (const zrecon_t *)t->corrections.p;
    n = t->corrections.len / sizeof(zrecon_t);
    for (size_t k = 0; k < n; k++)
        for (int r = 0; r < PACKED; r++) if (players[r] < PLAYERS) vec_push(&o->corrs[r], &cs[k].v[r], 8);
    const zrecon_t *is = (const zrecon_t *)t->inputs.p;
    n = t->inputs.len / sizeof(zrecon_t);
    for (size_t k = 0; k < n; k++)
        for (int r = 0; r < PACKED; r++) if (players[r] < PLAYERS) vec_push(&o->inputs[r], &is[k].v[r], 8);
}

/* ============================================================================================================
 *  Thread pool over packed instances (stands in for rayon, proof/mod.rs:33-38,128)
 * ========================================================================================================== */
typedef struct job_s job_t;
struct job_s { void (*fn)(job_t *, int); int n_items; int next; void *ctx; };
static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    for (;;) {
        int i = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (i >= j->n_items) break;
        j->fn(j, i);
    }
    return NULL;
}
static void parallel_for(void (*fn)(job_t *, int), int n_items, void *ctx, int n_threads) {
    if (n_threads <= 0) { long c = sysconf(_SC_NPROCESSORS_ONLN); n_threads = c > 0 ? (int)c : 1; }
    if (n_threads > n_items) n_threads = n_items;
    if (n_threads > 64) n_threads = 64;
    job_t j = {fn, n_items, 0, ctx};
    pthread_t th[64];
    for (int t = 1; t < n_threads; t++) pthread_create(&th[t], NULL, worker, &j);
    worker(&j);
    for (int t = 1; t < n_threads; t++) pthread_join(th[t], NULL);
}

/* ============================================================================================================
 *  Proof::new  (src/proof/mod.rs:119-222)
 * ========================================================================================================== */
typedef struct {
    const orc_op *ops; size_t n_ops;
    const uint8_t *wit_gf2; size_t n_gf2; const uint64_t *wit_z64; size_t n_z64;
    size_t z64_cells, gf2_cells;
    const uint8_t *seeds;
    instance_t *inst;                           /* [32] */
    uint8_t (*hashes)[PACKED][HASH_SIZE];       /* [32][8][32] */
    uint8_t omit_of_rep[TOTAL_REPS];
    packed_out_t *pg, *pz;                      /* [32] */
    int err_pass1;                              /* orc_prove_lowmem: first error seen by a worker */
} prove_ctx_t;

static void prove_instance(job_t *j, int i) { /* the closure at proof/mod.rs:129-156 */
    prove_ctx_t *c = (prove_ctx_t *)j->ctx;
    instance_t *I = &c->inst[i];
    uint8_t seeds[PACKED][KEY_SIZE];
    memcpy(seeds, c->seeds + (size_t)i * PACKED * KEY_SIZE, sizeof seeds);
    instance_alloc(I, c->z64_cells, c->gf2_cells);
    tr_gf2_init_prover(&I->tg, seeds, c->wit_gf2, c->n_gf2);
    tr_z64_init_prover(&I->tz, seeds, c->wit_z64, c->n_z64);
    run_circuit(I, c->ops, c->n_ops);
    instance_hash(I, c->hashes[i]);
}
static void extract_instance(job_t *j, int i) { /* proof/mod.rs:178-196 */
    prove_ctx_t *c = (prove_ctx_t *)j->ctx;
    extract_gf2(&c->inst[i].tg, &c->omit_of_rep[i * PACKED], &c->pg[i]);
    extract_z64(&c->inst[i].tz, &c->omit_of_rep[i * PACKED], &c->pz[i]);
}

static void put_u64(vec_t *v, uint64_t x) { vec_push(v, &x, 8); } /* bincode: little-endian fixed-width u64 */

static void serialize_domain(vec_t *out, prove_ctx_t *c, int is_z64) {
    /* bincode 1.3 default of ProofSingle{online: Vec<OpenOnline>, preprocessing: Vec<OpenPreprocessing>}, proof/mod.rs:40-60 */
    put_u64(out, ONLINE_REPS);
    for (int rep = 0; rep < TOTAL_REPS; rep++) {
        uint8_t omit = c->omit_of_rep[rep];
        if (omit >= PLAYERS) continue;
        int i = rep / PACKED, r = rep % PACKED;
        packed_out_t *p = is_z64 ? &c->pz[i] : &c->pg[i];
        uint8_t keys[PLAYERS][KEY_SIZE];
        expand_seed(c->seeds + (size_t)rep * KEY_SIZE, keys); /* prover.rs:126-127 */
        memset(keys[omit], 0, KEY_SIZE);
        vec_push_u8(out, omit);
        vec_push(out, keys, sizeof keys);
        put_u64(out, p->recons[r].len); vec_push(out, p->recons[r].p, p->recons[r].len);
        put_u64(out, p->corrs[r].len);  vec_push(out, p->corrs[r].p, p->corrs[r].len);
        put_u64(out, p->inputs[r].len); vec_push(out, p->inputs[r].p, p->inputs[r].len);
    }
    put_u64(out, PREPROCESSING_REPS);
    for (int rep = 0; rep < TOTAL_REPS; rep++) {
        if (c->omit_of_rep[rep] < PLAYERS) continue;
        int i = rep / PACKED, r = rep % PACKED;
        uint8_t comm_online[32];
        hasher_finalize(is_z64 ? &c->inst[i].tz.h_on[r] : &c->inst[i].tg.h_on[r], comm_online); /* prover.rs:168 */
        vec_push(out, c->seeds + (size_t)rep * KEY_SIZE, KEY_SIZE);
        vec_push(out, comm_online, 32);
    }
}

int orc_prove(const orc_op *ops, size_t n_ops, const uint8_t *wit_gf2, size_t n_gf2, const uint64_t *wit_z64, size_t n_z64,
              size_t z64_cells, size_t gf2_cells, const uint8_t *seeds, int n_threads, uint8_t **proof, size_t *proof_len,
              uint8_t *rep_hashes) {
    prove_ctx_t c;
    memset(&c, 0, sizeof c);
    c.ops = ops; c.n_ops = n_ops; c.wit_gf2 = wit_gf2; c.n_gf2 = n_gf2; c.wit_z64 = wit_z64; c.n_z64 = n_z64;
    c.z64_cells = z64_cells; c.gf2_cells = gf2_cells; c.seeds = seeds;
    c.inst = calloc(PACKED_REPS, sizeof(instance_t));
    c.hashes = calloc(PACKED_REPS, sizeof *c.hashes);
    c.pg = calloc(PACKED_REPS, sizeof(packed_out_t));
    c.pz = calloc(PACKED_REPS, sizeof(packed_out_t));
    parallel_for(prove_instance, PACKED_REPS, &c, n_threads);
    int err = 0;
    for (int i = 0; i < PACKED_REPS && !err; i++) {
        if (c.inst[i].err) err = c.inst[i].err;
        else if (c.inst[i].tg.err) err = c.inst[i].tg.err;
        else if (c.inst[i].tz.err) err = c.inst[i].tz.err;
    }
    if (!err) {
        uint8_t comm[32];
        orc_b3_oneshot(c.hashes, (size_t)TOTAL_REPS * HASH_SIZE, comm); /* combine_hashes, proof/mod.rs:102-108,160-168 */
        if (rep_hashes) memcpy(rep_hashes, c.hashes, (size_t)TOTAL_REPS * HASH_SIZE);
        orc_challenge(comm, c.omit_of_rep);                             /* proof/mod.rs:171-172 */
        parallel_for(extract_instance, PACKED_REPS, &c, n_threads);
        vec_t out = {0};
        vec_push(&out, comm, 32);
        serialize_domain(&out, &c, 0);
        serialize_domain(&out, &c, 1);
        *proof = out.p;
        *proof_len = out.len;
    }
    for (int i = 0; i < PACKED_REPS; i++) {
        instance_drop(&c.inst[i]);
        for (int r = 0; r < PACKED; r++) {
            vec_free(&c.pg[i].recons[r]); vec_free(&c.pg[i].corrs[r]); vec_free(&c.pg[i].inputs[r]);
            vec_free(&c.pz[i].recons[r]); vec_free(&c.pz[i].corrs[r]); vec_free(&c.pz[i].inputs[r]);
        }
    }
    free(c.inst); free(c.hashes); free(c.pg); free(c.pz);
    return err;
}

/* Proof::new for circuits whose recorded transcripts (O(gates) per packed instance, transcript/prover.rs:26-34) do not fit in
 * this machine's memory at once: the same functions in two passes.  Pass 1 runs every instance for its hashes only and drops
 * it; pass 2 (after the challenge) runs each instance again, extracts its openings (proof/mod.rs:178-196) and drops the
 * recorded vectors, keeping the hashers that serialize_domain finalizes.  At most n_threads instances are alive at a time.
 * The proof bytes are identical to orc_prove's (tests/test_oracle_protocol.py pins that). */
static void prove_instance_hash_only(job_t *j, int i) {
    prove_ctx_t *c = (prove_ctx_t *)j->ctx;
    prove_instance(j, i);
    instance_t *I = &c->inst[i];
    if (I->err) c->err_pass1 = I->err; else if (I->tg.err) c->err_pass1 = I->tg.err; else if (I->tz.err) c->err_pass1 = I->tz.err;
    instance_drop(I);
    memset(I, 0, sizeof *I);
}
static void prove_instance_and_extract(job_t *j, int i) {
    prove_ctx_t *c = (prove_ctx_t *)j->ctx;
    uint8_t keep[PACKED][HASH_SIZE];
    memcpy(keep, c->hashes[i], sizeof keep);
    prove_instance(j, i);
    if (memcmp(keep, c->hashes[i], sizeof keep) != 0) c->err_pass1 = ORC_E_ARG; /* the two passes must agree */
    extract_instance(j, i);
    instance_t *I = &c->inst[i];
    vec_free(&I->tg.reconstructions); vec_free(&I->tg.corrections); vec_free(&I->tg.inputs);
    vec_free(&I->tz.reconstructions); vec_free(&I->tz.corrections); vec_free(&I->tz.inputs);
    free(I->gw); free(I->zw);
    I->gw = NULL; I->zw = NULL;
}

int orc_prove_lowmem(const orc_op *ops, size_t n_ops, const uint8_t *wit_gf2, size_t n_gf2, const uint64_t *wit_z64, size_t n_z64,
                     size_t z64_cells, size_t gf2_cells, const uint8_t *seeds, int n_threads, uint8_t **proof, size_t *proof_len) {
    prove_ctx_t c;
    memset(&c, 0, sizeof c);
    c.ops = ops; c.n_ops = n_ops; c.wit_gf2 = wit_gf2; c.n_gf2 = n_gf2; c.wit_z64 = wit_z64; c.n_z64 = n_z64;
    c.z64_cells = z64_cells; c.gf2_cells = gf2_cells; c.seeds = seeds;
    c.inst = calloc(PACKED_REPS, sizeof(instance_t));
    c.hashes = calloc(PACKED_REPS, sizeof *c.hashes);
    c.pg = calloc(PACKED_REPS, sizeof(packed_out_t));
    c.pz = calloc(PACKED_REPS, sizeof(packed_out_t));
    parallel_for(prove_instance_hash_only, PACKED_REPS, &c, n_threads);
    int err = c.err_pass1;
    if (!err) {
        uint8_t comm[32];
        orc_b3_oneshot(c.hashes, (size_t)TOTAL_REPS * HASH_SIZE, comm);
        orc_challenge(comm, c.omit_of_rep);
        parallel_for(prove_instance_and_extract, PACKED_REPS, &c, n_threads);
        err = c.err_pass1;
        if (!err) {
            vec_t out = {0};
            vec_push(&out, comm, 32);
            serialize_domain(&out, &c, 0);
            serialize_domain(&out, &c, 1);
            *proof = out.p;
            *proof_len = out.len;
        }
    }
    for (int i = 0; i < PACKED_REPS; i++) {
        if (c.inst[i].tg.h_on[0].buf) instance_drop(&c.inst[i]);
        for (int r = 0; r < PACKED; r++) {
            vec_free(&c.pg[i].recons[r]); vec_free(&c.pg[i].corrs[r]); vec_free(&c.pg[i].inputs[r]);
            vec_free(&c.pz[i].recons[r]); vec_free(&c.pz[i].corrs[r]); vec_free(&c.pz[i].inputs[r]);
        }
    }
    free(c.inst); free(c.hashes); free(c.pg); free(c.pz);
    return err;
}

/* ============================================================================================================
 *  Proof::verify  (src/proof/mod.rs:224-307)
 * ========================================================================================================== */
typedef struct { open_online_t *online; size_t n_online; open_pre_t *pre; size_t n_pre; } parsed_domain_t;

static int parse_domain(const uint8_t *buf, size_t len, size_t *pos, parsed_domain_t *d) {
#define NEED(n) do { if (*pos + (n) > len || *pos + (n) < *pos) return ORC_E_FORMAT; } while (0)
    uint64_t n;
    NEED(8); memcpy(&n, buf + *pos, 8); *pos += 8;
    if (n > (len - *pos) / 153 + 1) return ORC_E_FORMAT;
    d->n_online = (size_t)n;
    d->online = calloc(d->n_online + 1, sizeof(open_online_t));
    for (size_t k = 0; k < d->n_online; k++) {
        open_online_t *o = &d->online[k];
        NEED(1 + 128); o->omit = buf[*pos]; memcpy(o->seeds, buf + *pos + 1, 128); *pos += 129;
        const uint8_t **fp[3] = {&o->recons, &o->corrs, &o->inputs};
        size_t *fl[3] = {&o->n_recons, &o->n_corrs, &o->n_inputs};
        for (int f = 0; f < 3; f++) {
            uint64_t l;
            NEED(8); memcpy(&l, buf + *pos, 8); *pos += 8;
            if (l > len - *pos) return ORC_E_FORMAT;
            *fp[f] = buf + *pos; *fl[f] = (size_t)l; *pos += (size_t)l;
        }
    }
    NEED(8); memcpy(&n, buf + *pos, 8); *pos += 8;
    if (n > (len - *pos) / 48 + 1) return ORC_E_FORMAT;
    d->n_pre = (size_t)n;
    d->pre = calloc(d->n_pre + 1, sizeof(open_pre_t));
    for (size_t k = 0; k < d->n_pre; k++) { NEED(48); memcpy(d->pre[k].seed, buf + *pos, 16); memcpy(d->pre[k].comm_online, buf + *pos + 16, 32); *pos += 48; }
    return ORC_OK;
#undef NEED
}

typedef struct {
    const orc_op *ops; size_t n_ops; size_t z64_cells, gf2_cells;
    parsed_domain_t g, z;
    uint8_t (*hashes)[PACKED][HASH_SIZE]; /* [32]: 5 online packs then 27 preprocessing packs */
    int err[PACKED_REPS], okay[PACKED_REPS];
} verify_ctx_t;

static void verify_pack(job_t *j, int i) { /* proof/mod.rs:249-280 */
    verify_ctx_t *c = (verify_ctx_t *)j->ctx;
    instance_t I;
    memset(&I, 0, sizeof I);
    instance_alloc(&I, c->z64_cells, c->gf2_cells);
    int e = 0;
    if (i < ONLINE_REPS / PACKED) {
        e = tr_gf2_init_online(&I.tg, &c->g.online[i * PACKED]);
        int e2 = tr_z64_init_online(&I.tz, &c->z.online[i * PACKED]);
        if (!e) e = e2;
    } else {
        int k = i - ONLINE_REPS / PACKED;
        tr_gf2_init_pre(&I.tg, &c->g.pre[k * PACKED]);
        tr_z64_init_pre(&I.tz, &c->z.pre[k * PACKED]);
    }
    if (!e) {
        run_circuit(&I, c->ops, c->n_ops);
        e = I.err;
        instance_hash(&I, c->hashes[i]);
    }
    c->err[i] = e;
    c->okay[i] = I.tg.okay && I.tz.okay;
    /* a failed online init leaves generators half-built; drop only what exists */
    if (I.tz.gen.batches == NULL) { I.tz.gen.batches = NULL; I.tz.gen.shares = NULL; }
    instance_drop(&I);
}

int orc_verify(const orc_op *ops, size_t n_ops, size_t z64_cells, size_t gf2_cells, const uint8_t *proof, size_t proof_len,
               int n_threads, int *okay, uint8_t *rep_hashes) {
    verify_ctx_t c;
    memset(&c, 0, sizeof c);
    c.ops = ops; c.n_ops = n_ops; c.z64_cells = z64_cells; c.gf2_cells = gf2_cells;
    if (proof_len < 32) return ORC_E_FORMAT;
    size_t pos = 32;
    int ret = parse_domain(proof, proof_len, &pos, &c.g);
    if (!ret) ret = parse_domain(proof, proof_len, &pos, &c.z);
    if (!ret && pos != proof_len) ret = ORC_E_FORMAT;
    if (!ret) {
        /* check_format, proof/mod.rs:110-114,225-230 */
        if (c.g.n_online != ONLINE_REPS || c.g.n_pre != PREPROCESSING_REPS || c.z.n_online != ONLINE_REPS || c.z.n_pre != PREPROCESSING_REPS) ret = 0;
        else {
            c.hashes = calloc(PACKED_REPS, sizeof *c.hashes);
            parallel_for(verify_pack, PACKED_REPS, &c, n_threads);
            int e = 0, ok = 1;
            for (int i = 0; i < PACKED_REPS; i++) { if (c.err[i] && !e) e = c.err[i]; ok &= c.okay[i]; }
            if (e) ret = e;
            else {
                uint8_t omit_of_rep[TOTAL_REPS], ordered[TOTAL_REPS][HASH_SIZE], comm[32];
                orc_challenge(proof, omit_of_rep); /* proof/mod.rs:292 */
                const uint8_t *flat = (const uint8_t *)c.hashes;
                size_t on = 0, pre = ONLINE_REPS;
                for (int i = 0; i < TOTAL_REPS; i++) { /* proof/mod.rs:293-302 */
                    if (omit_of_rep[i] < PLAYERS) memcpy(ordered[i], flat + HASH_SIZE * on++, HASH_SIZE);
                    else memcpy(ordered[i], flat + HASH_SIZE * pre++, HASH_SIZE);
                }
                orc_b3_oneshot(ordered, sizeof ordered, comm);
                if (rep_hashes) memcpy(rep_hashes, ordered, sizeof ordered);
                if (okay) *okay = ok;
                ret = memcmp(comm, proof, 32) == 0; /* proof/mod.rs:305-306 */
            }
            free(c.hashes);
        }
    }
    free(c.g.online); free(c.g.pre); free(c.z.online); free(c.z.pre);
    return ret;
}

void orc_free(void *p) { free(p); }

void orc_gf2_masks(const uint8_t *seeds, const uint8_t omit[8], uint64_t *out, size_t n) {
    uint8_t s[PACKED][KEY_SIZE], keys[PACKED][PLAYERS][KEY_SIZE];
    memcpy(s, seeds, sizeof s);
    keys_from_seeds(s, keys);
    for (int r = 0; r < PACKED; r++) if (omit[r] < PLAYERS) memset(keys[r][omit[r]], 0, KEY_SIZE);
    gen_gf2_t *g = malloc(sizeof *g);
    gen_gf2_init(g, keys, omit);
    for (size_t i = 0; i < n; i++) out[i] = gen_gf2_next(g);
    free(g);
}
