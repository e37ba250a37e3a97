// rv_kernels.cuh -- launchers of the sm_100a kernels (definitions in rv_kernels.cu).  Host-callable, all asynchronous on
// the given stream.  Nothing here falls back to the CPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rv_compile.h"

#ifndef RV_HD
#if defined(__CUDACC__)
#define RV_HD __host__ __device__ __forceinline__
#else
#define RV_HD inline
#endif
#endif

namespace rv {

// compiled tables resident in device memory
struct DevProgram {
    const XGate *xgates = nullptr;         // mapped XOR network (global-memory fallback of the mask plane)
    const uint32_t *xlevel_off = nullptr;  // n_llevels + 1
    const Item *items = nullptr;
    const uint32_t *mul_pos = nullptr;    // j -> online position of the j-th Mul
    const uint32_t *recon_pos = nullptr;  // k -> online position of the k-th reconstruct()
    const uint32_t *input_pos = nullptr;  // k -> online position of the k-th input()
    const uint32_t *input_vid = nullptr;  // k -> value id
    const LutInstr *lut_steps = nullptr;   // value-plane step stream (n_lut_steps * LUT_STEP slots)
    uint32_t n_lut_steps = 0;
    const VGate *wgates = nullptr;         // wide circuits: the level-sorted 2-input gates, one launch per level
    uint32_t n_wlevels = 0;
    // online verifier (absent for circuits of more than 4M ops)
    const LutInstr *vlut_steps = nullptr;  // u-plane step stream
    uint32_t n_vlut_steps = 0, n_uvals = 0;
    const VGate *vwgates = nullptr;        // wide circuits: level-sorted 2-input gates of the u-plane
    uint32_t n_vwlevels = 0;
    const uint32_t *vleaf_ids = nullptr;   // [n_inputs + n_and]: u-plane value id of every input, then of every Mul's kappa
    const uint32_t *item_ua = nullptr, *item_ub = nullptr, *recon_idx = nullptr;  // per online item
    const VmInstr *vm_steps = nullptr;     // mask-plane VM step stream (n_vm_steps * VM_STEP slots); empty without Add/Sub
    uint32_t n_vm_steps = 0, vm_cells = 0;
    uint32_t n_xgates = 0, n_llevels = 0;
    uint32_t n_lin = 0;
    uint32_t n_masks = 0, n_rows = 0, n_vals = 0, n_online = 0, n_pre = 0, n_inputs = 0, n_recon = 0;
    uint32_t max_llevel_width = 0;
    // tainted plane (Random / B2A): per-repetition plaintext words
    const TGate *tgates = nullptr;
    const uint32_t *tlevel_off = nullptr;
    uint32_t n_tlevels = 0, n_tvals = 0;
    const uint32_t *rand_row = nullptr;  // verifier: fresh mask row of every random leaf of the u-plane
    uint32_t n_rand = 0;
    const uint32_t *b2a_vrefs = nullptr, *b2a_urefs = nullptr;  // 64 per conversion
};

// compiled Z64 tables resident in device memory (rv_compile.h: ZProgram)
struct DevZProgram {
    const ZInstr *vprog = nullptr;
    const uint32_t *vlevel_off = nullptr;
    const ZLin *lin = nullptr;
    const ZItem *items = nullptr;
    const uint32_t *leaf_ids = nullptr, *recon_off = nullptr, *input_off = nullptr, *mul_pos = nullptr, *recon_idx = nullptr;
    const uint32_t *input_item = nullptr;  // k -> item index of the k-th input()
    uint32_t n_vlevels = 0, n_llevels = 0, n_items = 0, n_corr = 0 /* Mul + B2A */, n_inputs = 0, n_recon = 0, n_leaves = 0;
    uint32_t n_masks = 0, n_rows = 1, n_vals = 1;
    uint32_t on_bytes = 0, pre_bytes = 0;
};

// Opt-in shared-memory sizes / carveouts of every kernel for `device` (the current device); idempotent, thread-safe.
// Returns 0 or a cudaError_t.
int configure_kernels(int device);
int configure_zkernels();

// K1  seeds -> player keys -> AES round keys (src/transcript/mod.rs:99-106, src/crypto/prg.rs:16-20)
//     rk_plain: [45][32 * nslices] u32 -- the 44 round-key words of every stream (stream = 8 * rep + player), then a row of
//     all-ones / zero "stream is active" words
void launch_key_setup(const uint8_t *seeds, const uint8_t *pkeys_in, const uint8_t *mode, const uint8_t *omit, uint32_t nslices,
                      uint8_t *pkeys_out, uint32_t *rk_plain, cudaStream_t st, int *clear_flag = nullptr, uint32_t n_flags = 1, size_t flag_stride = 0);
// K2  AES-CTR mask generation straight into the share tensor (src/generator/share.rs:54-65, src/algebra/gf2/domain.rs:66-173)
//     also writes the instance-major copy `fresh_pm` [npi][pitch_pm] (u64) that the mask VM loads from (nullptr = skip)
//     busy_sms: SMs held by kernels that run alongside (the value plane's CTAs); share: sessions of the batch that run side by side
void launch_mask_gen_tt(const uint32_t *rk_plain, uint32_t nslices, uint32_t n_masks, uint64_t *rows, uint64_t *fresh_pm, size_t pitch_pm, int n_sms,
                        cudaStream_t st, uint32_t busy_sms = 0, uint32_t share = 1, bool pm_pairs = false, uint64_t mask_base = 0);
// K0  value plane (plaintext evaluation; one CTA, level-synchronous).  Returns the dynamic smem it asked for.
size_t launch_values(const LutInstr *steps, uint32_t n_steps, const uint32_t *leaf_ids, const uint8_t *leaf_vals, size_t leaf_pitch,
                     uint32_t n_leaves, uint8_t *vals, size_t vals_pitch, uint32_t n_vals, uint32_t n_instances, cudaStream_t st);
//     wide circuits (Program::values_wide): one grid-wide launch per level; returns the number of launches
int launch_values_wide(const DevProgram &P, const uint32_t *lut_level_off_host, const uint8_t *wit, uint8_t *vals, cudaStream_t st);
//     the verifier's u-plane of a wide circuit: same, for n_instances opened repetitions (grid.y)
int launch_uvalues_wide(const DevProgram &P, const uint32_t *vlut_level_off_host, const uint8_t *leaf_vals, size_t leaf_pitch, uint32_t n_leaves,
                        uint8_t *uvals, size_t upitch, uint32_t n_instances, cudaStream_t st);
// K3  mask plane (XOR network over the share tensor)
//     returns the number of kernel launches; *which (optional) names the variant: 0 VM (smem cells), 1 CTA walker, 2 per level
//     fresh_pm: instance-major copy of the fresh masks for the VM variant ([npi][pitch] u64)
int launch_linear(const DevProgram &P, const uint32_t *llevel_off_host, uint64_t *rows, uint32_t npi, const uint64_t *fresh_pm, size_t pitch_fresh,
                  cudaStream_t st, int *which = nullptr);
bool linear_uses_vm(const DevProgram &P);
//     true when the VM runs two adjacent columns per CTA for a tensor of npi columns: fresh_pm must then be the pair-interleaved copy
bool linear_vm_pairs(const DevProgram &P, uint32_t npi);
// K4  item plane: the two hash streams of every repetition
//     tvals: tainted plane [n_tvals][npi] (launch_tainted), read by items whose operands depend on Random / B2A fresh wires
void launch_tainted(const DevProgram &P, const uint64_t *rows, uint32_t npi, const uint8_t *vals, uint64_t *tvals, cudaStream_t st);
//     npi = packed instances of one proof; a session holding n_proofs proofs side by side passes their count, the pitch between
//     their value planes and the byte stride between their flags
void launch_items(const DevProgram &P, const uint64_t *rows, uint32_t npi, const uint8_t *vals, const uint64_t *tvals, uint8_t *on, size_t pitch_on,
                  uint8_t *pre, size_t pitch_pre, int *bad, cudaStream_t st, uint32_t n_proofs = 1, size_t vals_pitch = 0, size_t flag_stride = 0,
                  uint32_t base_on = 0, uint32_t base_pre = 0);
// K5  BLAKE3 chunk chaining values of `nreps` streams, then per-repetition tree + joins
void launch_chunk_cv2(const uint8_t *on, size_t pitch_on, uint32_t len_on, uint32_t *cv_on, uint32_t nreps_on, const uint8_t *pre, size_t pitch_pre,
                      uint32_t len_pre, uint32_t *cv_pre, uint32_t nreps_pre, cudaStream_t st);
void launch_chunk_cv_window(const uint8_t *on, size_t pitch_on, uint32_t len_on, uint32_t nchunks_on, uint32_t chunk0_on, uint32_t total_on, uint32_t *cv_on,
                            const uint8_t *pre, size_t pitch_pre, uint32_t len_pre, uint32_t nchunks_pre, uint32_t chunk0_pre, uint32_t total_pre,
                            uint32_t *cv_pre, uint32_t nreps, cudaStream_t st);
//     zconst: [0..8) B3(""), [8..16) H(B3("") || B3("")).  Verifier: repetitions >= first_pre use the proof's online hashes.
//     scratch (optional, prover): nreps * ceil(max(n_chunks) / 2) CVs; with it the wide lower levels of long streams' trees run grid-wide
void launch_rep_hash(uint32_t *cv_on, uint32_t n_chunks_on, uint32_t *cv_pre, uint32_t n_chunks_pre, const uint32_t *zconst, uint32_t nreps,
                     uint8_t *on_hash, uint8_t *rep_hash, cudaStream_t st, uint32_t first_pre = 0xFFFFFFFFu, const uint8_t *on_given = nullptr,
                     const uint8_t *z_on_given = nullptr, const uint32_t *zrep = nullptr, uint32_t *scratch = nullptr);
//     Z64 transcript of every repetition: roots of its two streams -> zon_hash[rep] (32 B) and zrep[rep] = H(B3(pre) || B3(on))
//     (src/transcript/mod.rs:77-96); feeds launch_rep_hash's `zrep`.  Repetitions >= first_pre take the proof's online hash.
void launch_zrep_hash(uint32_t *cv_on, uint32_t n_chunks_on, uint32_t *cv_pre, uint32_t n_chunks_pre, uint32_t nreps, uint8_t *zon_hash,
                      uint32_t *zrep, cudaStream_t st, uint32_t first_pre = 0xFFFFFFFFu, const uint8_t *z_on_given = nullptr);
// online verifier (src/transcript/verifier/online.rs)
struct VOpen;
void launch_verify_leaves(const DevProgram &P, const VOpen *opens, const uint8_t *proof, const uint64_t *rows, uint32_t npi, uint32_t n_slots,
                          uint8_t *leaf_vals, size_t leaf_pitch, cudaStream_t st);
void launch_verify_items(const DevProgram &P, const VOpen *opens, const uint8_t *proof, const uint64_t *rows, uint32_t npi, uint32_t npi_online,
                         const uint8_t *uvals, size_t upitch, uint8_t *on, size_t pitch_on, uint8_t *pre, size_t pitch_pre, int *not_okay,
                         cudaStream_t st);
void launch_items_pre_range(const DevProgram &P, const uint64_t *rows, uint32_t npi, uint32_t first_pi, uint8_t *pre, size_t pitch_pre, cudaStream_t st);
// K6  comm = H(256 rep hashes); Fiat-Shamir challenge (src/proof/mod.rs:74-108)
//     one warp per proof of the session; the hashes arrive as segments of seg_bytes per rank (see k_challenge)
//
//     Linked sessions (rv_session_peer_link): the sessions that hold the shards of one proof on different GPUs each own an
//     "exchange block" in device memory, mapped into every peer (CUDA IPC across processes, peer access inside one).
constexpr int RV_MAX_PEERS = 16;
constexpr int RV_BAD_WITNESS = 1, RV_BAD_PEER_TIMEOUT = 2;  // bits of a proof's status word
struct XchgLayout {  // byte offsets inside an exchange block of a session with n_proofs slots
    uint32_t n_proofs;
    RV_HD size_t off_flag(uint32_t parity, uint32_t src_rank) const { return ((size_t)(parity * RV_MAX_PEERS + src_rank) * n_proofs) * 4; }  // u32 [2][MAX][n_proofs]: step number of the hashes that arrived
    RV_HD size_t off_done() const { return off_flag(2, 0); }                      // u32 [MAX]: step number each rank finished extracting (read on the assembling rank)
    RV_HD size_t off_step() const { return off_done() + RV_MAX_PEERS * 4; }       // u32 [n_proofs]: this rank's own step counter per proof
    RV_HD size_t off_done_step() const { return off_step() + (size_t)n_proofs * 4; }  // u32: this rank's own counter of finished steps
    RV_HD size_t off_hash(uint32_t parity) const { return ((off_done_step() + 4 + 255) & ~(size_t)255) + (size_t)parity * n_proofs * RV_TOTAL_REPS * 32; }  // [rank][proof][seg]
    RV_HD size_t total() const { return off_hash(2); }
};
struct XchgArgs {
    uint32_t world = 1, rank = 0, dst = 0;
    uint64_t timeout_ns = 0;
    uint8_t *peer[RV_MAX_PEERS] = {};  // exchange block of every rank, as mapped on this device (peer[rank] = the own one)
};
void launch_challenge(const uint8_t *all_hashes, uint32_t seg_bytes, uint8_t *comm, size_t comm_stride, uint8_t *omit_of_rep, uint16_t *rank_of_rep,
                      uint32_t n_proofs, cudaStream_t st, const XchgArgs *x = nullptr, const uint8_t *own_hashes = nullptr);
//     last kernel of a linked open phase: "this rank's entries are in the assembling rank's proof buffer"; the assembling rank waits for all
void launch_xfinish(const XchgArgs &x, uint32_t n_proofs, int *bad, size_t flag_stride, cudaStream_t st);
// K7  openings -> bincode bytes of `Proof` (src/transcript/prover.rs:57-175, src/proof/mod.rs:40-66,200-221)
struct ExtractArgs {
    const uint8_t *on, *pre;
    size_t pitch_on, pitch_pre;
    const uint8_t *on_hash;      // [nreps][32]
    const uint8_t *pkeys;        // [nreps][8][16]
    const uint8_t *seeds;        // [nreps][16]
    const uint8_t *comm;         // [32], proof b at + b * proof_stride
    const uint8_t *omit_of_rep;  // [n_proofs][256]
    const uint16_t *rank_of_rep; // [n_proofs][256]
    uint32_t n_proofs = 1;       // proofs held side by side by the session: streams / hashes / keys are indexed by (proof, repetition)
    size_t proof_stride = 0;     // bytes between the proofs' output buffers
    const uint32_t *z64_empty_hash;  // B3("")
    const uint8_t *z_on_hash = nullptr;  // [nreps][32] BLAKE3 of the Z64 online streams (nullptr: no Z64 ops)
    uint32_t first_rep, nreps;
    uint32_t len_recons, len_corrs, len_inputs;  // packed byte lengths
    uint32_t len_zrecons = 0, len_zcorrs = 0, len_zinputs = 0;
    uint8_t *proof;
};
void launch_extract(const DevProgram &P, const ExtractArgs &a, cudaStream_t st);

// ---- streaming segments (rv_prove_streaming): carried wire state, per-segment share of the openings ----
void launch_seg_import(const uint32_t *slot, uint32_t n_imports, const uint64_t *cell_rows, const uint8_t *cell_vals, uint32_t npi, uint32_t n_prg, uint64_t *rows,
                       uint64_t *fresh_pm, size_t pitch_pm, bool pm_pairs, uint8_t *leaf_vals, cudaStream_t st);
void launch_seg_export(const uint32_t *slot, const uint32_t *row, const uint32_t *vref, uint32_t n_exports, const uint64_t *rows, const uint8_t *vals, uint32_t npi,
                       uint64_t *cell_rows, uint8_t *cell_vals, cudaStream_t st);
struct SegExtractArgs {
    const uint8_t *on, *pre;          // the segment's stream windows [rep][pitch]
    size_t pitch_on, pitch_pre;
    uint32_t base_on, base_pre;       // buffer position of the segment's first online / preprocessing byte
    const uint32_t *recon_pos, *input_pos;  // the segment's tables (positions relative to its first online byte)
    uint32_t n_recon, n_corr, n_inputs;     // elements this segment contributes
    uint64_t first_recon, first_corr, first_input;  // their global element indices
    const uint8_t *omit_of_rep;
    const uint16_t *rank_of_rep;
    uint32_t len_recons, len_corrs, len_inputs;  // packed byte lengths of the whole proof's vectors
    uint8_t *proof;
};
void launch_seg_extract(const SegExtractArgs &a, uint32_t nreps, cudaStream_t st);

// ---- Z64 domain (rv_z64.cu) ------------------------------------------------------------------------------------------
struct ZOpen;
void launch_zmask_gen_tt(const uint32_t *rk_plain, uint32_t nstreams, uint32_t n_masks, uint64_t *zrows, int n_sms, cudaStream_t st);
int launch_zlinear(const DevZProgram &Z, const uint32_t *llevel_off_host, uint64_t *zrows, uint32_t rowlen, cudaStream_t st);
//     gvals / b2a_vrefs: GF(2) value plane and B2A source refs (prover); nullptr in the verifier
void launch_zvalues(const DevZProgram &Z, const uint64_t *leaf_vals, size_t leaf_pitch, uint64_t *vals, size_t vals_pitch, uint32_t n_instances,
                    const uint8_t *gvals, const uint32_t *b2a_vrefs, cudaStream_t st);
//     grows: the GF(2) share tensor [row][npi] (B2A corrections read the 64 fresh GF(2) rows of the conversion)
void launch_zitems(const DevZProgram &Z, const uint64_t *zrows, size_t rowlen, uint32_t nreps, const uint64_t *vals, const uint64_t *grows, uint8_t *on,
                   size_t pitch_on, uint8_t *pre, size_t pitch_pre, int *bad, cudaStream_t st);
void launch_zitems_pre_range(const DevZProgram &Z, const uint64_t *zrows, size_t rowlen, uint32_t first_rep, uint32_t nreps, const uint64_t *grows,
                             uint8_t *pre, size_t pitch_pre, cudaStream_t st);
//     gopens / guvals / b2a_urefs: the GF(2) openings, u-plane and B2A result refs (leaves of B2A outputs)
void launch_zverify_leaves(const DevZProgram &Z, const ZOpen *opens, const uint8_t *proof, const uint64_t *zrows, size_t rowlen, uint32_t n_slots,
                           uint64_t *leaf_vals, size_t leaf_pitch, const VOpen *gopens, const uint8_t *guvals, size_t gupitch,
                           const uint32_t *b2a_urefs, cudaStream_t st);
void launch_zverify_items(const DevZProgram &Z, const ZOpen *opens, const uint8_t *proof, const uint64_t *zrows, size_t rowlen, uint32_t n_slots,
                          const uint64_t *uvals, size_t upitch, uint8_t *on, size_t pitch_on, uint8_t *pre, size_t pitch_pre, int *not_okay,
                          cudaStream_t st);
struct ZExtractArgs {
    const uint8_t *on, *pre;
    size_t pitch_on, pitch_pre;
    const uint8_t *omit_of_rep;   // [256]
    const uint16_t *rank_of_rep;  // [256]
    uint32_t first_rep, nreps;
    size_t z_base, sz_on_z;       // ProofLayout::z_base(), ::sz_on_z()
    uint8_t *proof;
};
void launch_zextract(const DevZProgram &Z, const ZExtractArgs &a, cudaStream_t st);

}  // namespace rv
