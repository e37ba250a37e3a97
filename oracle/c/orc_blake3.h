/* oracle/c/orc_blake3.h -- BLAKE3 (hash + XOF) for the CPU oracle.  TEST INFRASTRUCTURE ONLY.
 *
 * The reference gets BLAKE3 from the `blake3` crate (Cargo.toml:31; call sites src/crypto/hash.rs:31,39,48,54-56,
 * 121-125 and src/crypto/ro.rs:9-14).  That crate is not vendored under /root/reference, so this is a restatement
 * of the published BLAKE3 algorithm; tests/test_oracle_primitives.py pins it against the `blake3` Python wheel
 * (which wraps the same Rust crate) on lengths straddling the 64 B / 1 KiB / 64 KiB boundaries.
 */
#ifndef ORC_BLAKE3_H
#define ORC_BLAKE3_H
#include <stddef.h>
#include <stdint.h>

#define ORC_B3_OUT 32
#define ORC_B3_CHUNK 1024
#define ORC_B3_BLOCK 64

typedef struct {
    uint32_t cv[8];          /* chaining value of the chunk in progress */
    uint64_t chunk_counter;  /* index of the chunk in progress */
    uint8_t buf[ORC_B3_BLOCK];
    uint8_t buf_len;
    uint8_t blocks_compressed;
    uint32_t cv_stack[54][8];
    uint8_t cv_stack_len;
} orc_b3;

void orc_b3_init(orc_b3 *h);
void orc_b3_update(orc_b3 *h, const void *data, size_t len);
/* does not modify h (the reference's BufferedHasher::finalize clones, src/crypto/hash.rs:53-57) */
void orc_b3_finalize(const orc_b3 *h, uint8_t out[ORC_B3_OUT]);
/* XOF: bytes [seek, seek+len) of the output stream */
void orc_b3_finalize_xof(const orc_b3 *h, uint64_t seek, uint8_t *out, size_t len);
void orc_b3_oneshot(const void *data, size_t len, uint8_t out[ORC_B3_OUT]);

#endif
