/* oracle/c/reverie_oracle.h -- CPU oracle #2: C restatement of trailofbits/reverie 0.3.2's prover/verifier dataflow.
 *
 * TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load liborc.so.  The product (reverie_b200/) never links, loads or calls it.
 *
 * Parity status: "parity unpinned" by reference vectors (the reference has none and cannot be built here); pinned
 * instead by (a) primitive KATs, (b) byte equality with the independent Python restatement oracle/reverie_oracle.py,
 * (c) the reference's own test cases re-expressed.  See that file's header.
 */
#ifndef REVERIE_ORACLE_H
#define REVERIE_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Circuit op: same 24-byte layout as the product's rv_op (declared independently on purpose).
 * mcircuit::{Operation, CombineOperation} as matched at src/interpreter/single.rs:106-156, combine.rs:120-132. */
typedef struct {
    uint8_t domain; /* 0 GF2(op)  1 Z64(op)  2 B2A(dst=z64 wire, a=low gf2 wire)  3 SizeHint(a=z64 cells, b=gf2 cells) */
    uint8_t opcode; /* 0 Input(dst) 1 Random(dst) 2 Add(dst,a,b) 3 AddConst(dst,a,imm) 4 Sub(dst,a,b) 5 SubConst(dst,a,imm)
                       6 Mul(dst,a,b) 7 MulConst(dst,a,imm) 8 AssertZero(a) 9 Const(dst,imm) */
    uint16_t pad;
    uint32_t dst, a, b;
    uint64_t imm;
} orc_op;

enum { ORC_OK = 0, ORC_E_WITNESS_INVALID = -1, ORC_E_WITNESS_SHORT = -2, ORC_E_FORMAT = -3, ORC_E_ARG = -4 };

/* Proof::new (src/proof/mod.rs:119-222) with the 256 rep seeds injected in place of OsRng (:131-134).
 * *proof is malloc'd (release with orc_free) and holds the bincode bytes of `Proof`.  rep_hashes (256*32 B) optional.
 * n_threads <= 0 means "all online cores", capped at 32 like the reference's rayon fan-out (:128). */
int orc_prove(const orc_op *ops, size_t n_ops, const uint8_t *wit_gf2, size_t n_gf2, const uint64_t *wit_z64, size_t n_z64,
              size_t z64_cells, size_t gf2_cells, const uint8_t *seeds /*256*16*/, int n_threads, uint8_t **proof,
              size_t *proof_len, uint8_t *rep_hashes);

/* The same proof in two passes for circuits whose recorded transcripts exceed this machine's memory (10^8 gates = 51 GB): hashes
 * first, then each instance again for its openings; at most n_threads instances are alive at once.  Same bytes as orc_prove. */
int orc_prove_lowmem(const orc_op *ops, size_t n_ops, const uint8_t *wit_gf2, size_t n_gf2, const uint64_t *wit_z64, size_t n_z64,
                     size_t z64_cells, size_t gf2_cells, const uint8_t *seeds /*256*16*/, int n_threads, uint8_t **proof,
                     size_t *proof_len);

/* Proof::verify (src/proof/mod.rs:224-307).  Returns 1 accept / 0 reject / <0 error.  *okay (optional) receives the AND
 * of the online verifiers' zero_check flags (verifier/online.rs:176-178; unused by the reference's verify).
 * rep_hashes (256*32, optional) receives the recomputed per-repetition hashes in original rep order. */
int orc_verify(const orc_op *ops, size_t n_ops, size_t z64_cells, size_t gf2_cells, const uint8_t *proof, size_t proof_len,
               int n_threads, int *okay, uint8_t *rep_hashes);

void orc_free(void *p);

/* primitive taps for the KAT tests */
void orc_aes128_ctr(const uint8_t key[16], uint64_t first_block, uint8_t *out, size_t n_blocks);
void orc_blake3(const uint8_t *data, size_t len, uint8_t *out, size_t out_len);
/* first n GF2 / Z64 masks of one packed instance (8 rep seeds; omit[r]==8 -> none omitted; keys expanded from seeds) */
void orc_gf2_masks(const uint8_t *seeds /*8*16*/, const uint8_t omit[8], uint64_t *out, size_t n);
void orc_challenge(const uint8_t comm[32], uint8_t omit_of_rep[256] /* 8 = not opened */);

#ifdef __cplusplus
}
#endif
#endif
