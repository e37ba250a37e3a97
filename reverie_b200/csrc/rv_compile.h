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
// mask-plane gate: row[dst] = row[a] ^ row[b]   (rows of the share tensor, 8 bytes per packed instance)
struct LGate {
    uint32_t dst, a, b, pad;
};
enum ItemKind : uint32_t { ITEM_INPUT = 0, ITEM_MUL = 1, ITEM_ASSERT = 2 };
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
};
static_assert(sizeof(VGate) == 16 && sizeof(LGate) == 16 && sizeof(Item) == 32, "POD layout");

struct Program {
    // GF(2) side
    uint64_t n_ops = 0, n_and = 0, n_inputs = 0, n_assert = 0;
    uint32_t n_masks = 0;   // fresh PRG masks per (rep, player)
    uint32_t n_lin = 0;     // linear nodes
    uint32_t n_rows = 0;    // n_masks + n_lin + 1; the last row is all-zero
    uint32_t n_vals = 1;    // value ids; vid 0 is the constant 0
    uint32_t n_online = 0;  // == items.size()
    uint32_t n_pre = 0;     // == n_and
    std::vector<VGate> vgates;          // sorted by level
    std::vector<uint32_t> vlevel_off;   // value_depth + 1 offsets into vgates
    std::vector<LGate> lgates;          // sorted by level; dst rows are n_masks + position
    std::vector<uint32_t> llevel_off;   // linear_depth + 1 offsets into lgates
    std::vector<Item> items;            // online-stream order
    std::vector<uint32_t> recon_pos;    // online positions of the reconstruct() calls (Mul, AssertZero), in order
    std::vector<uint32_t> input_pos;    // online positions of the input() calls, in order
    std::vector<uint32_t> input_vid;    // witness index -> value id
    uint64_t algorithmic_bytes = 0;
    bool uses_z64 = false;
    uint32_t zero_row() const { return n_rows - 1; }
};

// Returns RV_OK or a negative rv_status; `err` receives a human-readable reason.
int compile(const rv_op *ops, size_t n_ops, size_t z64_cells, size_t gf2_cells, Program &out, std::string &err);

}  // namespace rv
