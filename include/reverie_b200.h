/* reverie_b200.h -- C ABI of libreverie_b200.so: the B200-native KKW (MPC-in-the-head) prover/verifier core.
 *
 * Drop-in boundary for trailofbits/reverie 0.3.2 (all citations relative to the reference repository root):
 *   Proof::new(circuit, wit_gf2, wit_z64, wire_counts) -> Proof      src/proof/mod.rs:119-222
 *   Proof::verify(&self, circuit, wire_counts) -> bool               src/proof/mod.rs:224-307
 * A Rust maintainer binds these entry points with a ~40-line `extern "C"` block; the proof bytes
 * crossing the boundary are exactly `bincode::serialize(&Proof)` (src/proof/mod.rs:40-66, src/main.rs:84).
 *
 * Plain pointers and sizes only.  Every call is synchronous unless its name ends in `_async`.  Nothing unwinds across
 * the boundary: the reference's panics become negative error codes.  There is no CPU fallback: without a CUDA device
 * every compute entry point returns RV_E_CUDA.
 */
#ifndef REVERIE_B200_H
#define REVERIE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- protocol constants: src/lib.rs:17-38, src/crypto/prg.rs:9, src/crypto/hash.rs:8 ---- */
#define RV_PLAYERS 8
#define RV_PACKED 8
#define RV_TOTAL_REPS 256
#define RV_ONLINE_REPS 40
#define RV_PREPROCESSING_REPS (RV_TOTAL_REPS - RV_ONLINE_REPS)
#define RV_PACKED_REPS (RV_TOTAL_REPS / RV_PACKED)
#define RV_KEY_SIZE 16
#define RV_HASH_SIZE 32

/* ---- circuit: mcircuit::{Operation, CombineOperation} as matched by the reference's interpreter
 *      (src/interpreter/single.rs:106-156, src/interpreter/combine.rs:120-132), flattened to a 24-byte POD. ---- */
enum rv_domain { RV_GF2 = 0, RV_Z64 = 1, RV_B2A = 2 /* dst = z64 wire, a = lowest of 64 gf2 wires */,
                 RV_SIZE_HINT = 3 /* a = z64 cells, b = gf2 cells */ };
enum rv_opcode {
    RV_INPUT = 0,       /* Input(dst)            */
    RV_RANDOM = 1,      /* Random(dst)           */
    RV_ADD = 2,         /* Add(dst, a, b)        */
    RV_ADDC = 3,        /* AddConst(dst, a, imm) */
    RV_SUB = 4,         /* Sub(dst, a, b)        */
    RV_SUBC = 5,        /* SubConst(dst, a, imm) */
    RV_MUL = 6,         /* Mul(dst, a, b)        */
    RV_MULC = 7,        /* MulConst(dst, a, imm) */
    RV_ASSERT_ZERO = 8, /* AssertZero(a)         */
    RV_CONST = 9        /* Const(dst, imm)       */
};
typedef struct rv_op {
    uint8_t domain;  /* enum rv_domain */
    uint8_t opcode;  /* enum rv_opcode (ignored for RV_B2A / RV_SIZE_HINT) */
    uint16_t pad;
    uint32_t dst, a, b;
    uint64_t imm;    /* GF(2): bit 0; Z64: the full word */
} rv_op;

/* ---- errors (the reference panics here) ---- */
enum rv_status {
    RV_OK = 0,
    RV_E_WITNESS_INVALID = -1, /* an AssertZero failed            src/transcript/prover.rs:221-228 */
    RV_E_WITNESS_SHORT = -2,   /* ran out of witness elements     src/transcript/prover.rs:190     */
    RV_E_FORMAT = -3,          /* malformed proof bytes           src/algebra/gf2/share.rs:158-164 */
    RV_E_ARG = -4,             /* bad argument / wire index out of range for the given wire_counts */
    RV_E_CUDA = -5,            /* CUDA runtime failure or no device */
    RV_E_NOMEM = -6,
    RV_E_UNSUPPORTED = -7,     /* a shape the device path does not serve (reported, never silently degraded): see DESIGN.md section 8 */
    RV_E_PEER = -8             /* multi-GPU: a linked session of another rank never arrived (the role SURVEY.md 8(b) gives RV_E_NCCL) */
};

typedef struct rv_circuit rv_circuit; /* a compiled circuit: device-resident gate tables, reusable across proofs  */
typedef struct rv_session rv_session; /* one in-flight prove: device buffers + stream                              */
typedef struct rv_batch rv_batch;     /* several sessions of one circuit launched as one CUDA graph               */
typedef struct rv_group rv_group;     /* a multi-GPU prover: linked sessions on every GPU it drives               */

/* Per-thread message for the last non-OK status returned on this thread. */
const char *rv_last_error(void);
/* Library/ABI version string, e.g. "reverie-b200 0.1 (sm_100a)". */
const char *rv_version(void);
/* Number of visible CUDA devices (0 without a GPU; never fails). */
int rv_device_count(void);
/* Select the CUDA device used by objects created afterwards on this thread (default 0). */
int rv_set_device(int device);

/* ---------------------------------------------------------------------------------------------------------------
 * Circuit compilation.  Stands for the `Arc<Vec<CombineOperation>>` argument of Proof::new / Proof::verify
 * (src/proof/mod.rs:120,224): compile once, share read-only across any number of proofs and threads.
 * wire counts are passed in the reference's own tuple order (z64, gf2)  -- src/proof/mod.rs:125,232.
 * ------------------------------------------------------------------------------------------------------------- */
int rv_circuit_compile(const rv_op *ops, size_t n_ops, size_t z64_cells, size_t gf2_cells, rv_circuit **out);
/* flags: RV_COMPILE_PROVE_ONLY leaves out the online verifier's tables (a third of the compile time and of the table bytes);
 * rv_verify on such a handle returns RV_E_UNSUPPORTED. */
#define RV_COMPILE_PROVE_ONLY 1u
int rv_circuit_compile_ex(const rv_op *ops, size_t n_ops, size_t z64_cells, size_t gf2_cells, unsigned flags, rv_circuit **out);
void rv_circuit_free(rv_circuit *c);

typedef struct rv_circuit_stats {
    uint64_t n_ops;
    uint64_t n_and;          /* GF(2) Mul gates                                   */
    uint64_t n_inputs;       /* GF(2) witness bits consumed                       */
    uint64_t n_assert;       /* GF(2) AssertZero                                  */
    uint64_t n_masks;        /* GF(2) PRG masks drawn per (rep, player)           */
    uint64_t n_linear;       /* materialised linear (XOR) mask nodes              */
    uint64_t value_depth;    /* levels of the plaintext plane after LUT mapping   */
    uint64_t linear_depth;   /* levels of the mask plane after XOR-cut mapping    */
    uint64_t plain_value_depth;   /* ... of the circuit's own 2-input gates       */
    uint64_t plain_linear_depth;
    uint64_t n_luts, n_lut_steps, n_vm_steps, vm_cells; /* sizes of the device programs */
    uint64_t online_bytes;   /* bytes hashed per repetition, online stream        */
    uint64_t pre_bytes;      /* bytes hashed per repetition, preprocessing stream */
    uint64_t algorithmic_bytes; /* SURVEY.md 8(d) HBM bytes for one proof (all 256 reps) */
    uint64_t device_bytes;   /* device memory held by the compiled tables         */
    /* Z64 domain (src/algebra/z64) */
    uint64_t z64_mul;           /* Z64 Mul gates                                  */
    uint64_t z64_inputs;        /* Z64 witness elements consumed                  */
    uint64_t z64_assert;
    uint64_t z64_masks;         /* Z64 PRG masks drawn per (rep, player)          */
    uint64_t z64_linear;        /* materialised Z64 linear mask nodes             */
    uint64_t z64_value_depth;   /* levels of the Z64 plaintext plane              */
    uint64_t z64_linear_depth;
    uint64_t z64_online_bytes;  /* bytes hashed per repetition, Z64 online stream */
    uint64_t z64_pre_bytes;
    uint64_t compile_ns;        /* host time rv_circuit_compile spent on this handle (compile + upload of the tables) */
    uint64_t has_verify;        /* 1 if the verifier's tables were built */
    uint64_t n_vals, n_uvals;   /* value ids of the plaintext plane / of the verifier's u-plane */
    uint64_t n_vlut_steps;      /* steps of the u-plane's LUT program */
} rv_circuit_stats;
int rv_circuit_get_stats(const rv_circuit *c, rv_circuit_stats *out);

/* Debug/test tap: copy one compiled table to the host (tests/ re-executes the tables in numpy).  `what` is one of
 * the RV_TAB_* ids; call with buf == NULL to obtain the byte length. */
enum rv_table { RV_TAB_VGATES = 0, RV_TAB_LUTS = 1, RV_TAB_XGATES = 2, RV_TAB_XLEVELS = 3, RV_TAB_ITEMS = 4,
                RV_TAB_RECON_POS = 5, RV_TAB_INPUT_POS = 6, RV_TAB_INPUT_VID = 7 };
int rv_circuit_export(const rv_circuit *c, int what, void *buf, size_t *len);

/* ---------------------------------------------------------------------------------------------------------------
 * Proof::new  (src/proof/mod.rs:119-222).
 *   wit_gf2: one byte per GF(2) witness element (0/1), consumed in Input order  (src/transcript/prover.rs:181-199)
 *   wit_z64: one u64 per Z64 witness element
 *   seeds:   256 x 16 bytes, repetition r at seeds[16r..16r+16) -- replaces the OsRng draw at src/proof/mod.rs:131-134
 *            so that proofs are reproducible; NULL = draw from the OS RNG like the reference.
 *   *proof:  library-allocated bincode bytes of `Proof`; release with rv_free.
 * ------------------------------------------------------------------------------------------------------------- */
int rv_prove(const rv_circuit *c, const uint8_t *wit_gf2, size_t n_gf2, const uint64_t *wit_z64, size_t n_z64,
             const uint8_t *seeds, uint8_t **proof, size_t *proof_len);

/* Proof::new for n independent witnesses of one circuit (a proving service's queue).  Arrays of n pointers / sizes; seeds may be
 * NULL (OS RNG for all) or hold NULL entries.  proofs[i] / proof_lens[i] / statuses[i] receive each proof's result (RV_OK,
 * RV_E_WITNESS_INVALID, RV_E_WITNESS_SHORT ...; release proofs[i] with rv_free).  The return value is RV_OK unless the call as a
 * whole failed (CUDA / memory).  Small GF(2) circuits run side by side in multi-proof sessions; other circuits one by one. */
int rv_prove_batch(const rv_circuit *c, int n, const uint8_t *const *wit_gf2, const size_t *n_gf2, const uint64_t *const *wit_z64,
                   const size_t *n_z64, const uint8_t *const *seeds, uint8_t **proofs, size_t *proof_lens, int *statuses);

/* Proof::verify  (src/proof/mod.rs:224-307).  Returns 1 accept / 0 reject / <0 error: the reference's exact verdict, i.e.
 * "the recomputed commitment equals the proof's" (src/proof/mod.rs:305-306).
 * *okay (optional) receives the AND of the online verifiers' zero_check flags, which the reference computes
 * (src/transcript/verifier/online.rs:176-178) but never reads: the reference leans on the PROVER's assert
 * (src/transcript/prover.rs:221-228) for the circuit's AssertZero constraints, so a prover that skips it passes.  A caller
 * that wants the constraints enforced accepts only when the return value is 1 AND *okay is 1; the host mirrors
 * (reverie_b200.hpp, reverie_b200/proof.py, the CLI) and rv_proof_verify do exactly that by default. */
int rv_verify(const rv_circuit *c, const uint8_t *proof, size_t proof_len, int *okay);

/* One-shot forms with the exact argument shape of the reference API (compile + run; compiled circuits are cached by content,
 * see rv_circuit_cache_*).  rv_proof_verify is strict: 1 only if the commitment matches and every AssertZero of the opened
 * repetitions holds (see rv_verify). */
int rv_proof_new(const rv_op *ops, size_t n_ops, const uint8_t *wit_gf2, size_t n_gf2, const uint64_t *wit_z64,
                 size_t n_z64, size_t z64_cells, size_t gf2_cells, const uint8_t *seeds, uint8_t **proof,
                 size_t *proof_len);
int rv_proof_verify(const rv_op *ops, size_t n_ops, size_t z64_cells, size_t gf2_cells, const uint8_t *proof,
                    size_t proof_len);
/* The reference's exact verdict (commitment equality) with the AssertZero flag reported separately, like rv_verify. */
int rv_proof_verify_ex(const rv_op *ops, size_t n_ops, size_t z64_cells, size_t gf2_cells, const uint8_t *proof,
                       size_t proof_len, int *okay);
/* The one-shot forms keep compiled circuits in a content-addressed cache (128-bit hash of the op list, wire counts, device) so
 * that a caller with the reference's call shape -- Proof::new(circuit, ...) with the circuit passed every time,
 * src/proof/mod.rs:119-124 -- compiles each circuit once.  Default: 8 entries, least recently used idle entry evicted first. */
void rv_circuit_cache_clear(void);
/* rv_proof_new on a circuit of >= n_ops ops that it has not seen before (and that streaming serves: GF(2) without Random) proves
 * in streaming mode instead of compiling the circuit for residency first -- same bytes, several times sooner for 10^8 gates; the
 * next call with that circuit compiles and caches it.  Default 2^24 ops; 0 = never. */
void rv_oneshot_streaming_min(size_t n_ops);
void rv_circuit_cache_limit(size_t max_entries);
void rv_circuit_cache_stats(uint64_t *hits, uint64_t *misses, size_t *entries);

/* Streaming Proof::new for circuits whose share tensor and transcripts do not fit in device memory (the
 * reference's README.md:14 calls it the streaming interface -- its own prover keeps O(gates) vectors, src/transcript/prover.rs:26-34).
 * The op list is proved in segments of `window_ops` ops (0 = a default of 4 M); wires that cross a segment boundary are carried
 * on the device, PRG and hash streams continue, and the circuit is walked twice (hashes, then openings).  The proof bytes are
 * identical to rv_proof_new's.  Device memory: O(window_ops) + ~45 bytes of gate tables per gate + the proof.
 * GF(2) circuits without Random / Z64 / B2A (RV_E_UNSUPPORTED otherwise); one GPU. */
int rv_prove_streaming(const rv_op *ops, size_t n_ops, size_t z64_cells, size_t gf2_cells, const uint8_t *wit_gf2, size_t n_gf2,
                       const uint64_t *wit_z64, size_t n_z64, const uint8_t *seeds, size_t window_ops, uint8_t **proof, size_t *proof_len);

/* Test hook (host only, no device needed): runs rv_prove_streaming's planner -- segmentation, liveness, cell-file slots -- and
 * checks it by a symbolic simulation of the cell file.  out: segments, slots, most imports / exports of a segment, total imports,
 * total exports. */
int rv_stream_plan_check(const rv_op *ops, size_t n_ops, size_t gf2_cells, size_t window_ops, uint64_t out[6]);

/* Releases any buffer the library handed out.  Proof buffers are slices of pooled pinned host blocks that the device wrote directly
 * (no copy on the way out); a block returns to the pool once every proof in it has been released, so release proofs when done
 * rather than keeping thousands alive. */
void rv_free(void *p);

/* ---------------------------------------------------------------------------------------------------------------
 * Phase-split proving, for sharding the 32 packed instances (src/proof/mod.rs:127-157) across GPUs and for
 * measuring with inputs resident in HBM.  A session owns a CUDA stream and all device buffers of one proof shard.
 *
 *   rv_session_create   shard = packed instances [first_instance, first_instance + n_instances)
 *   rv_session_upload   host -> device copy of witness + this shard's seeds (pinned staging, async on the stream)
 *   rv_session_commit   the closure at src/proof/mod.rs:129-156 for every instance of the shard, entirely on the
 *                       device; leaves n_instances*8 repetition hashes in device memory
 *   rv_session_hashes   device -> host copy of those hashes (n_instances * 8 * 32 bytes); synchronises
 *   rv_session_open     src/proof/mod.rs:160-196 given all 256 repetition hashes (the all-gather result, instance-major
 *                       order); computes comm + challenge on the device and extracts this shard's openings
 *   rv_session_fetch    device -> host: comm (32 B) and this shard's openings as a relocatable blob
 *   rv_proof_assemble   src/proof/mod.rs:200-221: concatenates shard blobs (any order) into the bincode `Proof`
 * ------------------------------------------------------------------------------------------------------------- */
int rv_session_create(const rv_circuit *c, int first_instance, int n_instances, rv_session **out);
/* A session that holds `n_proofs` independent proofs of the circuit SIDE BY SIDE (proof b owns its own columns of the share
 * tensor, its own streams, hashes and output buffer): every phase is the same handful of kernel launches as for one proof,
 * each launch covering all of them.  This is how many small proofs in flight keep a B200 busy -- a 22 k-gate proof alone
 * is a few hundred CTAs per kernel.  Slots are filled with rv_session_upload_slot and read with rv_session_fetch_slot; the
 * other session calls act on all slots.  rv_session_hashes_device / _all_hashes_device: [n_proofs][n_instances * 8 * 32] and,
 * gathered over ranks, [rank][n_proofs][n_instances * 8 * 32] (= one plain all-gather of the former).
 * Small GF(2) circuits only (RV_E_UNSUPPORTED otherwise: Z64 / Random / B2A, wide value planes, proofs of 4 MB and more). */
int rv_session_create_multi(const rv_circuit *c, int first_instance, int n_instances, int n_proofs, rv_session **out);
int rv_session_upload_slot(rv_session *s, int slot, const uint8_t *wit_gf2, size_t n_gf2, const uint64_t *wit_z64, size_t n_z64,
                           const uint8_t *seeds /* all 256 x 16 of that proof, or NULL = OS RNG */);
int rv_session_fetch_slot(rv_session *s, int slot, uint8_t comm[RV_HASH_SIZE], uint8_t **part, size_t *part_len); /* synchronises */
void rv_session_free(rv_session *s);
int rv_session_upload(rv_session *s, const uint8_t *wit_gf2, size_t n_gf2, const uint64_t *wit_z64, size_t n_z64,
                      const uint8_t *seeds /* all 256 x 16, or NULL = OS RNG */);
int rv_session_commit(rv_session *s);                                   /* async on the session stream; one CUDA graph launch after the first call */
int rv_session_hashes(rv_session *s, uint8_t *rep_hashes);              /* synchronises                */
/* Device pointer to this shard's n_instances * 8 * 32 bytes of repetition hashes: valid once rv_session_commit has been
 * enqueued, to be read by work ordered after it on the session stream (e.g. an NCCL all-gather launched on that stream). */
const void *rv_session_hashes_device(rv_session *s);
/* Device buffer of 256 x 32 bytes owned by the session: gather all repetition hashes straight into it (e.g. as the NCCL
 * receive buffer) and pass the same pointer to rv_session_open -- no copy, and the open phase replays as one CUDA graph. */
void *rv_session_all_hashes_device(rv_session *s);
int rv_session_open(rv_session *s, const uint8_t *all_rep_hashes);      /* async; host or device pointer; NULL = own hashes (single shard) */
int rv_session_prove(rv_session *s);                                    /* async: commit + open(own hashes) of a full shard as
                                                                           one CUDA graph launch after the first, eager, call */
int rv_session_fetch(rv_session *s, uint8_t comm[RV_HASH_SIZE], uint8_t **part, size_t *part_len); /* synchronises */
int rv_session_sync(rv_session *s);
/* Synchronises and reports the proof's status without copying it: RV_OK, or RV_E_WITNESS_INVALID if an AssertZero failed. */
int rv_session_status(rv_session *s);
/* The shard's openings in device memory (the full-length proof buffer, zero outside the shard's entries), valid after
 * rv_session_open: lets a multi-GPU caller combine the shards on the device (their non-zero bytes are disjoint, so an
 * NCCL sum-reduce of the byte buffers is the assembly of src/proof/mod.rs:200-221). */
int rv_session_proof_device(rv_session *s, void **ptr, size_t *len);
/* Multi-proof sessions: number of slots, and the byte distance between consecutive slots' proof buffers in device memory. */
int rv_session_slots(const rv_session *s);
size_t rv_session_proof_stride(const rv_session *s);
int rv_proof_assemble(const uint8_t comm[RV_HASH_SIZE], const uint8_t *const *parts, const size_t *part_lens,
                      int n_parts, uint8_t **proof, size_t *proof_len);
/* ---------------------------------------------------------------------------------------------------------------
 * Multi-GPU: linking the sessions that hold the shards of one proof on different GPUs.
 * The one exchange of the protocol -- the all-gather of the 256 x 32-byte repetition hashes, src/proof/mod.rs:160-171 -- and
 * the assembly of the openings (src/proof/mod.rs:200-221) then run over NVLink peer memory INSIDE the open phase: each rank's
 * challenge kernel stores its hashes into every rank's receive buffer and waits on per-rank flags; each rank's extraction
 * writes its entries straight into the assembling rank's (rank 0) proof buffer.  A linked shard is driven like a full one:
 * rv_session_prove / rv_batch_prove = one CUDA graph launch per rank and step, no collective library, no host hop.
 *   rv_session_peer_handle   opaque RV_PEER_HANDLE_BYTES naming this session (exchange block + proof buffer): raw device
 *                            pointers for sessions of the same process, CUDA IPC handles across processes
 *   rv_session_peer_link     rank r of `world` (2..16, world * n_instances == 32, shard r = instances [32 r / world, ...)) passes
 *                            all ranks' handles in rank order (exchanged by any host channel: MPI, torch.distributed, a pipe)
 * Every rank must then run the same sequence of prove steps; a rank that never arrives makes the others' status RV_E_PEER
 * after RV_PEER_TIMEOUT_MS (environment, default 60000).  rv_session_fetch on rank 0 returns the whole proof; on the other
 * ranks it returns the status and comm with *part = NULL.
 * ------------------------------------------------------------------------------------------------------------- */
#define RV_PEER_HANDLE_BYTES 256
int rv_session_peer_handle(rv_session *s, uint8_t handle[RV_PEER_HANDLE_BYTES]);
int rv_session_peer_link(rv_session *s, int rank, int world, const uint8_t *handles /* world x RV_PEER_HANDLE_BYTES */);
int rv_session_peer_rank(const rv_session *s, int *rank, int *world, int *assembles);

/* ---------------------------------------------------------------------------------------------------------------
 * rv_group: Proof::new on several GPUs behind one handle.  The group owns, for every GPU it drives, `n_sessions` linked
 * sessions of `slots` proofs each (n_sessions x slots proofs per step); a session's step is one CUDA graph launch per GPU.
 *   rv_group_create_local   ONE process drives `n_devices` GPUs (1, 2, 4, 8, 16): the circuit is cloned onto each device
 *                           (rv_circuit_clone: the host-side compile is shared), the shards are linked through peer access.
 *                           This is the call a Rust host makes to give Proof::new all the GPUs of a box.
 *   rv_group_create_rank    one process PER GPU (MPI / torchrun style): this process holds rank `rank` of `world`; exchange
 *                           rv_group_handles (rv_group_handles_bytes each) over any host channel, pass all of them in rank
 *                           order to rv_group_link.  Every rank then makes the same rv_group_prove* calls with the same
 *                           witnesses and seeds (seeds must be given: they have to agree across ranks); rank 0 receives
 *                           the proofs, the other ranks receive the statuses.
 *   rv_group_prove_batch    like rv_prove_batch; rv_group_prove = one proof (create the group with 1 session x 1 slot for that).
 *   rv_group_step           relaunch one step on the inputs already uploaded (asynchronous; device-resident timing)
 *   rv_group_session        session `index` of member `member` (0 for a rank group), for callers that drive uploads, steps and
 *                           fetches themselves through the session API; owned by the group.  Each session has its own stream.
 * ------------------------------------------------------------------------------------------------------------- */
int rv_circuit_clone(const rv_circuit *c, int device, rv_circuit **out);
int rv_group_create_local(const rv_circuit *c, const int *devices, int n_devices, int n_sessions, int slots, rv_group **out);
int rv_group_create_rank(const rv_circuit *c, int rank, int world, int n_sessions, int slots, rv_group **out);
size_t rv_group_handles_bytes(const rv_group *g);
int rv_group_handles(rv_group *g, uint8_t *handles);
int rv_group_link(rv_group *g, const uint8_t *all_handles /* world x rv_group_handles_bytes(g), rank order */);
int rv_group_info(const rv_group *g, int *world, int *n_members, int *n_sessions, int *slots);
rv_session *rv_group_session(rv_group *g, int member, int index);
int rv_group_step(rv_group *g);
int rv_group_prove_batch(rv_group *g, int n, const uint8_t *const *wit_gf2, const size_t *n_gf2, const uint64_t *const *wit_z64,
                         const size_t *n_z64, const uint8_t *const *seeds, uint8_t **proofs, size_t *proof_lens, int *statuses);
int rv_group_prove(rv_group *g, const uint8_t *wit_gf2, size_t n_gf2, const uint64_t *wit_z64, size_t n_z64, const uint8_t *seeds,
                   uint8_t **proof, size_t *proof_len);
/* Proof::verify for n proofs spread over the group's GPUs (whole proofs per GPU: verification has no exchange step).  results[i] =
 * 1 accept / 0 reject / < 0 error; okay[i] (optional) = the AssertZero flag, see rv_verify.  A rank group verifies the proofs
 * with i % world == rank and leaves the other entries untouched.  The group's circuit must carry the verifier's tables. */
int rv_group_verify_batch(rv_group *g, int n, const uint8_t *const *proofs, const size_t *lens, int *results, int *okay);
void rv_group_free(rv_group *g);

/* cudaStream_t of the session (as void*), so callers can bracket work with their own CUDA events. */
void *rv_session_stream(rv_session *s);

/* ---------------------------------------------------------------------------------------------------------------
 * Batches: several sessions of ONE circuit (proofs in flight, e.g. the requests a proving service has queued) driven as a
 * unit.  A phase of all of them is a single CUDA graph launch on the leader's stream (sessions[0]); the sessions' own
 * streams fork from it and join back inside the graph.  While bound to a batch a session must be driven through the batch
 * calls; rv_session_upload / _fetch / _status / _hashes_device / _all_hashes_device still address it individually and are
 * ordered on the leader's stream (rv_batch_stream), which is also where the caller enqueues the all-gather between
 * rv_batch_commit and rv_batch_open.  rv_batch_open reads every session's own rv_session_all_hashes_device buffer.
 * ------------------------------------------------------------------------------------------------------------- */
int rv_batch_create(rv_session *const *sessions, int n, rv_batch **out);
void rv_batch_free(rv_batch *b);        /* unbinds the sessions; does not free them */
int rv_batch_commit(rv_batch *b);       /* async */
int rv_batch_open(rv_batch *b);         /* async */
int rv_batch_prove(rv_batch *b);        /* async: commit + open(own hashes), full shards only */
void *rv_batch_stream(rv_batch *b);     /* cudaStream_t of the leader */

/* Per-kernel device timing (CUDA events on the session stream).  Enable, run, then read back
 * `n` (name, total ms, launches) triples accumulated since the last reset. */
typedef struct rv_kernel_time {
    char name[32];
    double ms;
    uint64_t launches;
    uint64_t algorithmic_bytes; /* SURVEY.md 8(d) bytes attributed to this kernel, summed over launches */
} rv_kernel_time;
int rv_session_timing(rv_session *s, int enable);
int rv_session_kernel_times(rv_session *s, rv_kernel_time *out, int max_out, int reset);
/* Kernels launched on this session since creation (the bench's `gpu_launches`). */
uint64_t rv_session_launch_count(const rv_session *s);

#ifdef __cplusplus
}
#endif
#endif /* REVERIE_B200_H */
