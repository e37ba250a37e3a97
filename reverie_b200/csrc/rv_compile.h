// rv_compile.h -- host-side circuit compiler: op list -> static schedule for the device planes.
//
// The reference interprets the op list gate by gate, once per packed instance, with every repetition carrying its own
// (mask, correction) pair per wire (src/interpreter/single.rs:106-157).  We split that state into three planes that can
// each be evaluated with far more parallelism:
//
//   value plane   the plaintext bit of every wire, shared by all 256 repetitions.  In the prover every wire satisfies
//                 corr = value - reconstruct(mask)  (src/interpreter/mod.rs:17-19, asserted at single.rs:62-66), so the
//                 per-repetition corrections never need to be carried through the circuit.
//   mask plane    the 2048 (repetition x player) mask bits of every wire.  Input/Mul/Random outputs are FRESH PRG masks
//                 (src/interpreter/single.rs:26-27, src/transcript/prover.rs:181-232), so only Add/Sub create
//                 dependencies: the plane is a XOR network whose depth is the circuit's *linear* depth.
//   item plane    one record per transcript side effect (Input / Mul / AssertZero, i.e. one byte of the online hash
//                 stream per repetition, plus one byte of the preprocessing stream per Mul); fully gate-parallel.
//
// The sequential order of the reference only fixes (i) which PRG index each gate draws and (ii) where its bytes land in
// the two hash streams; both are prefix sums over the op list and are resolved here.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/reverie_b200.h"

namespace rv {

// value-plane gate: v[dst] = (v[a>>1] ^ (a&1)) OP (v[b>>1] ^ (b&1)),  OP = xor (0) / and (1)
struct VGate {
    uint32_t dst, a, b, op;
};
// mask-plane gate after cut mapping: row[dst] = XOR of row[in[0..5]] (unused inputs name the all-zero row).  The same
// K=6 cut mapper that shortens the value plane collapses XOR chains here (a ripple carry's running XOR of AND outputs
// becomes one 6-input gate per five links), so the linear depth of SHA-256 drops from 557 to ~1/5.
struct XGate {
    uint32_t dst;
    uint32_t in[6];
    uint32_t pad;
};
// mask-plane VM instruction (shared-memory cells; see build_mask_vm).  20 bytes: the instruction stream is re-read by every
// CTA (one per packed instance), so its size is what bounds the kernel.
//   XOR : v = XOR of cell[in[0..5]]; cell[dst] = v; if (row != VM_ROW_NONE) exported[row - n_masks] = v
//   LOAD: cell[dst] <- fresh row `row`   asynchronously, at least `VM_DELTA` levels ahead of its first use (one cp.async group per level;
//         the compiler issues some earlier to fill the padding of the steps, see build_mask_vm)
struct VmInstr {
    uint32_t row;
    uint16_t dst;    // cell
    uint16_t flags;  // VM_F_*
    uint16_t in[6];  // cells
};
constexpr int VM_DELTA = 2;  // minimum prefetch distance in levels: a LOAD issued during level L-2 is awaited at the end of level L-1

// value-plane LUT instruction: v[dst] = tt >> (v[in0] | v[in1]<<1 | ... | v[in5]<<5) & 1.  Unused inputs name value 0
// (the constant 0).  Produced by the depth-oriented K=6 cut mapper (build_value_luts), which collapses cones of 2-input
// gates -- e.g. two full-adder stages of a ripple carry -- into one lookup, so the level count of the plane drops.
struct LutInstr {
    uint32_t dst;
    uint32_t in[6];
    uint32_t pad;
    uint64_t tt;
    uint64_t pad2;  // 48 bytes = three 16-byte ring units
};

// Device form of both programs: a dense "VLIW" stream of STEPS.  Every thread of the CTA executes exactly one slot per
// step (empty slots are harmless no-ops on a scratch cell), so the device loop needs no level table, no bounds checks and
// no inner loops -- the per-level dependent chain is what bounds these kernels, and it is paid in instructions per warp.
//   mask VM : step = VM_STEP slots of 20 bytes
//   LUT     : step = LUT_STEP slots of 48 bytes; `pad` = flags
// STEP_BAR on a slot means "CTA barrier after this step" (set on every slot of the last step of a level and of a chunk).
// The LUT stream carries a barrier every LUT_STEPS_PER_CHUNK steps, so it can be consumed in chunks of 4 or of 8 steps: k_values takes
// 8 when the values still fit next to the larger ring (the prover's plane), 4 otherwise (the verifier's u-plane of SHA-256).
constexpr uint32_t VM_STEP = 512, VM_STEPS_PER_CHUNK = 2, LUT_STEP = 128, LUT_STEPS_PER_CHUNK = 4, LUT_STEPS_PER_CHUNK_MAX = 8;
// Shared memory of the mask VM's CTA (rv_kernels.cu asserts that these match its own): the instruction ring + 8 bytes per cell and
// column.  vm_columns = how many tensor columns one CTA can serve at a given cell count (2 is the fast configuration, 0 = no VM).
constexpr size_t VM_SMEM_CAP = 226 * 1024, VM_STREAM_SMEM = (size_t)2 * VM_STEPS_PER_CHUNK * VM_STEP * 20 + 64;
inline int vm_columns(uint32_t cells) {
    return VM_STREAM_SMEM + ((size_t)cells + 1) * 16 <= VM_SMEM_CAP ? 2 : VM_STREAM_SMEM + ((size_t)cells + 1) * 8 <= VM_SMEM_CAP ? 1 : 0;
}
constexpr uint32_t VM_F_LOAD = 1u, VM_F_BAR = 2u, VM_F_LEVEL_END = 4u, VM_CELL_MASK = 0xFFFFu /* also "no cell yet" */,
                   VM_ROW_NONE = 0xFFFFFFFFu;
constexpr uint32_t LUT_F_BAR = 1u;

enum ItemKind : uint32_t { ITEM_INPUT = 0, ITEM_MUL = 1, ITEM_ASSERT = 2,
                           ITEM_RECON = 3 /* GF(2): reconstruct() without a zero check (B2A result bits, src/interpreter/combine.rs:204-207) */,
                           ITEM_B2A = 3 /* Z64 item list: the correction of a B2A conversion (src/interpreter/combine.rs:150-158) */ };

// Per-repetition ("tainted") values.  `Random` and the 64 fresh wires of a B2A conversion have corr = 0, so their plaintext
// is reconstruct(mask): different in every repetition (src/interpreter/single.rs:148-150, combine.rs:139-147).  Everything
// computed from them lives in the tainted plane: one Recon-format u64 per packed instance (byte r = 0x00 / 0xFF for rep r).
// A value ref with VREF_TAINT set names tainted value (ref >> 1) & 0x3FFFFFFF; bit 0 is still "negate".
constexpr uint32_t VREF_TAINT = 0x80000000u;
enum TOp : uint32_t { T_XOR = 0, T_AND = 1, T_LEAF = 2 /* a = fresh mask row: value = reconstruct(row) */ };
struct TGate {
    uint32_t op, dst, a, b;  // a, b: value refs (tainted or shared)
};
// one byte of the online stream per repetition (and, for Mul, one byte of the preprocessing stream)
struct Item {
    uint32_t kind;  // ItemKind
    uint32_t ra;    // INPUT: row of the fresh mask; MUL: row of operand a's mask; ASSERT: row of the wire's mask
    uint32_t rb;    // MUL: row of operand b's mask
    uint32_t k;     // MUL: PRG index of mask_ab (mask_new = k + 1)      src/interpreter/single.rs:26-27
    uint32_t va;    // value ref (vid << 1 | negate) of operand a / the input / the asserted wire
    uint32_t vb;    // MUL: value ref of operand b
    uint32_t j;     // MUL: position in the preprocessing stream; INPUT: witness index
    uint32_t pad;
    Item() = default;
    // emplace_back(...) builds the record in place.  push_back(Item{...}) bounces it through the stack: dword stores, then 16-byte
    // reloads that cannot be store-forwarded and so wait for every earlier table store to drain (the hottest spot of the op walk).
    Item(uint32_t kind_, uint32_t ra_, uint32_t rb_, uint32_t k_, uint32_t va_, uint32_t vb_, uint32_t j_, uint32_t pad_)
        : kind(kind_), ra(ra_), rb(rb_), k(k_), va(va_), vb(vb_), j(j_), pad(pad_) {}
};
static_assert(sizeof(TGate) == 16 && sizeof(VGate) == 16 && sizeof(XGate) == 32 && sizeof(VmInstr) == 20 && sizeof(LutInstr) == 48 && sizeof(Item) == 32, "POD layout");

// =====================================================================================================================
//  Z64 domain (src/algebra/z64/*): the same three planes over the ring Z_2^64.
//    value plane  one u64 per wire, shared by all repetitions (corr = value - reconstruct(mask) holds here too: for Mul,
//                 corr_out = (a + c1)(b + c2) - reconstruct(mask_new), src/interpreter/single.rs:25-69)
//    mask plane   512 bytes per wire per packed instance ([8 reps][8 players] u64); Add/Sub/MulConst keep a wire's mask a
//                 Z_2^64-linear combination of fresh PRG masks, carried as (row, coefficient) so MulConst costs nothing
//    item plane   Input: 8 bytes of online stream per repetition; Mul / AssertZero: 64 bytes (the 8 players' shares,
//                 src/algebra/z64/share.rs:100-108); Mul: 8 bytes of preprocessing stream (src/algebra/z64/recon.rs:131-137)
//  The same value program serves the online verifier: every Mul is v[a] * v[b] + v[kappa], with the kappa leaves zero in
//  the prover and rho_ab - rho_a * rho_b + msg + delta in the verifier.
// =====================================================================================================================
enum ZvOp : uint32_t { ZV_ADD = 0, ZV_SUB = 1, ZV_MUL = 2 /* v[a] * v[b] + v[c] */, ZV_ADDC = 3 /* v[a] + imm */, ZV_MULC = 4 /* v[a] * imm */,
                       ZV_CONST = 5 /* imm */, ZV_B2A = 6 /* v[c] + (prover: the 64 source bits of conversion `a` packed LSB first) */ };
struct ZInstr {
    uint32_t op, dst, a, b, c, pad;
    uint64_t imm;
};
// mask-plane node: zrow[dst] = ca * zrow[a] + cb * zrow[b]  (element-wise over the 64 * npi shares of a row)
struct ZLin {
    uint32_t dst, a, b, pad;
    uint64_t ca, cb;
};
struct ZItem {
    uint32_t kind;  // ItemKind
    uint32_t ra;    // row of operand a's mask (INPUT: the fresh mask; ASSERT: the wire's mask); B2A: first of the 64 fresh GF(2) rows
    uint32_t rb;    // MUL: row of operand b's mask
    uint32_t k;     // MUL: fresh-mask index of mask_ab (mask_new = k + 1); B2A: the fresh Z64 mask
    uint32_t va;    // value id of operand a / the input / the asserted wire; B2A: index of the conversion (side tables)
    uint32_t vb;    // MUL: value id of operand b; B2A: index of its first GF(2) reconstruct()
    uint32_t j;     // MUL / B2A: index among the corrections (preprocessing stream offset 8 j); INPUT: witness index
    uint32_t off;   // byte offset in the online stream (B2A: none)
    uint64_t ca;    // wire mask = ca * zrow[ra]
    uint64_t cb;
};
static_assert(sizeof(ZInstr) == 32 && sizeof(ZLin) == 32 && sizeof(ZItem) == 48, "POD layout");

struct ZProgram {
    uint64_t n_mul = 0, n_inputs = 0, n_assert = 0, n_b2a = 0;
    uint64_t n_corr = 0;   // corrections = Mul + B2A: 8 bytes of preprocessing stream each
    uint32_t n_masks = 0;  // fresh Z64 PRG masks per (rep, player): mask i = LE u64 at byte 8 i of the stream (z64/batch.rs:25-30)
    uint32_t n_lin = 0;    // linear nodes
    uint32_t n_rows = 1;   // n_masks + n_lin + 1 (the last row is all-zero)
    uint32_t n_vals = 1;   // value ids; id 0 is the constant 0
    std::vector<ZInstr> vprog;          // sorted by level
    std::vector<uint32_t> vlevel_off;   // value_depth + 1 offsets
    std::vector<ZLin> lin;              // sorted by level; dst rows are n_masks + position
    std::vector<uint32_t> llevel_off;
    std::vector<ZItem> items;           // online-stream order
    std::vector<uint32_t> leaf_ids;     // value ids of the leaves: inputs (witness order), then one per correction (Mul: kappa; B2A: u of its output)
    std::vector<uint32_t> recon_off;    // online byte offset of the k-th reconstruct() (Mul, AssertZero)
    std::vector<uint32_t> input_off;    // online byte offset of the k-th input()
    std::vector<uint32_t> mul_pos;      // correction j -> item index (Mul or B2A)
    std::vector<uint32_t> recon_idx;    // item index -> index among the reconstruct() calls
    uint64_t on_bytes = 0, pre_bytes = 0;  // per repetition
    bool any() const { return !items.empty() || n_masks != 0; }
    uint32_t zero_row() const { return n_rows - 1; }
};

struct Program {
    // GF(2) side
    uint64_t n_ops = 0, n_and = 0, n_inputs = 0, n_assert = 0;
    uint32_t n_masks = 0;   // fresh PRG masks per (rep, player)
    uint32_t n_lin = 0;     // materialised (mapped) linear nodes
    uint32_t n_rows = 0;    // n_masks + n_lin + 1; the last row is all-zero
    uint32_t n_vals = 1;    // value ids; vid 0 is the constant 0
    uint32_t n_online = 0;  // == items.size()
    uint32_t n_pre = 0;     // == n_and
    std::vector<VGate> vgates;          // plain 2-input value gates in topological order (debug/tests; dropped for huge circuits)
    std::vector<XGate> xgates;          // mapped XOR network, sorted by level; dst rows are n_masks + position
    std::vector<uint32_t> xlevel_off;   // linear_depth + 1 offsets into xgates
    std::vector<LutInstr> luts;         // mapped value-plane program, sorted by level
    std::vector<uint32_t> lut_level_off;
    std::vector<VmInstr> vm;            // mask-plane VM program over cells, sorted by VM level (= level + VM_DELTA - 1); empty if it needs > 65534 cells
    std::vector<uint32_t> vm_level_off;
    uint32_t vm_cells = 0;              // shared-memory cells the program needs (one lane word each), excluding the scratch cell
    uint32_t plain_value_depth = 0, plain_linear_depth = 0;  // depths before mapping (reported in the stats)
    std::vector<VmInstr> vm_steps;      // padded device stream of the mask VM (n_vm_steps * VM_STEP slots); cell vm_cells = scratch
    uint32_t n_vm_steps = 0;
    std::vector<LutInstr> lut_steps;    // padded device stream of the value plane (n_lut_steps * LUT_STEP slots); value n_vals = scratch
    uint32_t n_lut_steps = 0;
    bool values_wide = false;           // levels average >= WIDE_LEVEL gates: one grid-wide launch per level over `wgates` instead of the step stream
    std::vector<VGate> wgates;          // wide circuits: the 2-input gates that feed the item plane, sorted by level (no LUT mapping)
    std::vector<uint32_t> wlevel_off;
    // online-verifier value plane ("u-plane"): same circuit, every Mul is (a & b) ^ kappa_j with kappa_j a leaf
    std::vector<LutInstr> vlut_steps;   // padded device stream; value n_uvals = scratch
    uint32_t n_vlut_steps = 0, n_uvals = 0;
    std::vector<uint32_t> input_uid;    // witness index -> u-plane value id
    std::vector<uint32_t> kappa_uid;    // Mul index j -> u-plane value id of its kappa leaf
    std::vector<uint32_t> item_ua, item_ub;  // per online item: u-plane refs (id << 1 | negate) of the operands / asserted wire
    std::vector<VGate> vwgates;         // wide circuits (verify_wide): the level-sorted 2-input gates of the u-plane, one launch per level
    std::vector<uint32_t> vwlevel_off;
    bool verify_wide = false;
    bool has_verify = false;            // built for circuits of <= VERIFY_MAX_OPS ops
    std::vector<TGate> tgates;          // tainted plane, sorted by level
    std::vector<uint32_t> tlevel_off;
    uint32_t n_tvals = 0;
    std::vector<uint32_t> rand_row;     // k-th random leaf of the verifier's u-plane -> its fresh mask row (u = rho(row))
    std::vector<uint32_t> rand_uid;     //                                            -> its u-plane value id
    // B2A conversions (src/interpreter/combine.rs:132-219), 64 entries each
    std::vector<uint32_t> b2a_vrefs;    // plaintext value refs of the 64 source wires (prover: the Z64 value is their packing)
    std::vector<uint32_t> b2a_urefs;    // u-plane refs of the 64 result wires (verifier)
    std::vector<Item> items;            // online-stream order
    std::vector<uint32_t> recon_pos;    // online positions of the reconstruct() calls (Mul, AssertZero), in order
    std::vector<uint32_t> input_pos;    // online positions of the input() calls, in order
    std::vector<uint32_t> input_vid;    // witness index -> value id
    uint64_t algorithmic_bytes = 0;
    bool uses_z64 = false;
    ZProgram z;
    // Streaming segments (SegmentIO): rows [n_prg, n_masks) are IMPORTED mask rows (the carried state of wires written by earlier
    // segments), not PRG output; everywhere else they behave like fresh rows.  n_prg == n_masks for a whole circuit.
    uint32_t n_prg = 0;
    std::vector<uint32_t> export_rows;  // rows later segments need: written back by the mask VM even if no item of this segment reads them
    uint32_t zero_row() const { return n_rows - 1; }
};

// One segment of a circuit proved in streaming mode (rv_prove_streaming): the op list [a, b) of the whole circuit with its wire
// cells renumbered densely.  `import_cells` name the cells whose state (plaintext value + mask row) earlier segments left behind;
// `export_cells` the cells whose final state later segments read.  GF(2) without Random / B2A only.
struct SegmentIO {
    std::vector<uint32_t> import_cells, export_cells;  // in: local cell ids
    std::vector<uint32_t> import_vid;                  // out: value id of import j (a leaf of the value plane, after the witness inputs)
    std::vector<uint32_t> export_vref, export_row;     // out: value ref (vid << 1 | negate) and mask row of export k (zero_row = the zero mask)
};

// Returns RV_OK or a negative rv_status; `err` receives a human-readable reason.
// flags: COMPILE_PROVE_ONLY skips the online verifier's tables (u-plane): about a third of the compile time and of the table bytes.
constexpr uint32_t COMPILE_PROVE_ONLY = 1u;
int compile(const rv_op *ops, size_t n_ops, size_t z64_cells, size_t gf2_cells, Program &out, std::string &err, uint32_t flags = 0, SegmentIO *io = nullptr);

constexpr uint32_t WIDE_LEVEL = 4096;
constexpr size_t VERIFY_MAX_OPS = (size_t)1 << 28;  // the verifier's tables cost ~200 bytes of host memory per gate while compiling

// Gate-count limit above which the value plane keeps 2-input "LUTs" (the mapper's cut sets cost ~250 bytes per gate).
constexpr size_t LUT_MAP_MAX_GATES = 8u << 20;

}  // namespace rv
