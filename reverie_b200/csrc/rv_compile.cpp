// rv_compile.cpp -- see rv_compile.h.  Pure host C++ (no CUDA), so the CPU test-suite can exercise it without a GPU.
#include "rv_compile.h"

#include <algorithm>
#include <cstring>

namespace rv {
namespace {

constexpr uint32_t ZERO_MID = 0xFFFFFFFFu;  // "the all-zero mask" until row numbers are final
constexpr uint32_t LIN_BASE = 0x80000000u;  // provisional ids of linear nodes: LIN_BASE + creation index
constexpr uint32_t VREF_ZERO = 0;           // vid 0, not negated
constexpr uint32_t VREF_ONE = 1;            // vid 0, negated

struct Cell {
    uint32_t vref;  // value id << 1 | negate
    uint32_t mid;   // fresh PRG index, LIN_BASE + linear node, or ZERO_MID
};

// algorithmic HBM bytes per gate over all 256 repetitions, plus the 16-byte descriptor
constexpr uint64_t B_AND = 2048 + 16, B_XOR = 1536 + 16, B_UNARY = 1024 + 16, B_INPUT = 768 + 16, B_ASSERT = 768 + 16,
                   B_LEAF = 512 + 16;

}  // namespace

int compile(const rv_op *ops, size_t n_ops, size_t z64_cells, size_t gf2_cells, Program &P, std::string &err) {
    P = Program();
    P.n_ops = n_ops;
    if (n_ops && !ops) {
        err = "ops is NULL";
        return RV_E_ARG;
    }
    std::vector<Cell> cells(gf2_cells, Cell{VREF_ZERO, ZERO_MID});
    std::vector<uint32_t> vlevel(1, 0);  // per value id
    std::vector<uint32_t> llevel;        // per linear node (creation order)
    std::vector<VGate> vg;               // creation order
    std::vector<LGate> lg;               // creation order, provisional ids
    uint64_t n_masks = 0;

    auto mid_level = [&](uint32_t mid) -> uint32_t { return (mid != ZERO_MID && mid >= LIN_BASE) ? llevel[mid - LIN_BASE] : 0; };
    auto new_val = [&](uint32_t level) -> uint32_t {
        vlevel.push_back(level);
        return (uint32_t)(vlevel.size() - 1);
    };
    auto bad_wire = [&](size_t i) {
        err = "op " + std::to_string(i) + ": wire index out of range for the given wire_counts";
        return RV_E_ARG;
    };

    for (size_t i = 0; i < n_ops; i++) {
        const rv_op &op = ops[i];
        if (op.domain == RV_SIZE_HINT) {  // src/interpreter/combine.rs:122-129
            if (cells.size() < op.b) cells.resize(op.b, Cell{VREF_ZERO, ZERO_MID});
            if (z64_cells < op.a) z64_cells = op.a;
            continue;
        }
        if (op.domain == RV_Z64 || op.domain == RV_B2A) {
            P.uses_z64 = true;
            err = "op " + std::to_string(i) + ": Z64 / B2A operations are not accelerated yet";
            return RV_E_UNSUPPORTED;
        }
        if (op.domain != RV_GF2) {
            err = "op " + std::to_string(i) + ": unknown domain";
            return RV_E_ARG;
        }
        const size_t nc = cells.size();
        const uint32_t c = (uint32_t)(op.imm & 1);  // bool -> Recon, src/algebra/gf2/recon.rs:274-287
        switch (op.opcode) {
            case RV_INPUT: {  // src/transcript/prover.rs:181-199
                if (op.dst >= nc) return bad_wire(i);
                uint32_t vid = new_val(0);
                Item it{ITEM_INPUT, (uint32_t)n_masks, 0, 0, vid << 1, 0, (uint32_t)P.input_vid.size(), 0};
                P.input_pos.push_back((uint32_t)P.items.size());
                P.items.push_back(it);
                P.input_vid.push_back(vid);
                cells[op.dst] = Cell{vid << 1, (uint32_t)n_masks};
                n_masks += 1;
                P.n_inputs++;
                P.algorithmic_bytes += B_INPUT;
                break;
            }
            case RV_RANDOM:
                // Wire{mask: fresh, corr: 0}: the wire's value differs per repetition, so the shared value plane does
                // not apply.  Never silently degraded: reported here.
                err = "op " + std::to_string(i) + ": Random is not accelerated yet";
                return RV_E_UNSUPPORTED;
            case RV_ADD:
            case RV_SUB: {  // src/interpreter/single.rs:71-85: mask and correction add component-wise
                if (op.dst >= nc || op.a >= nc || op.b >= nc) return bad_wire(i);
                const Cell A = cells[op.a], B = cells[op.b];
                Cell R;
                // value
                const uint32_t neg = (A.vref ^ B.vref) & 1;
                if ((A.vref >> 1) == 0) R.vref = B.vref ^ (A.vref & 1);
                else if ((B.vref >> 1) == 0) R.vref = A.vref ^ (B.vref & 1);
                else if ((A.vref >> 1) == (B.vref >> 1)) R.vref = neg;  // x ^ x (^1)
                else {
                    uint32_t vid = new_val(1 + std::max(vlevel[A.vref >> 1], vlevel[B.vref >> 1]));
                    vg.push_back(VGate{vid, A.vref & ~1u, B.vref & ~1u, 0});
                    R.vref = (vid << 1) | neg;
                }
                // mask
                if (A.mid == ZERO_MID) R.mid = B.mid;
                else if (B.mid == ZERO_MID) R.mid = A.mid;
                else if (A.mid == B.mid) R.mid = ZERO_MID;
                else {
                    if (lg.size() >= LIN_BASE - 1) {
                        err = "too many linear nodes";
                        return RV_E_UNSUPPORTED;
                    }
                    uint32_t id = LIN_BASE + (uint32_t)lg.size();
                    llevel.push_back(1 + std::max(mid_level(A.mid), mid_level(B.mid)));
                    lg.push_back(LGate{id, A.mid, B.mid, 0});
                    R.mid = id;
                }
                cells[op.dst] = R;
                P.algorithmic_bytes += B_XOR;
                break;
            }
            case RV_ADDC:
            case RV_SUBC: {  // src/interpreter/single.rs:87-95: only the correction changes
                if (op.dst >= nc || op.a >= nc) return bad_wire(i);
                Cell R = cells[op.a];
                R.vref ^= c;
                cells[op.dst] = R;
                P.algorithmic_bytes += B_UNARY;
                break;
            }
            case RV_MULC: {  // src/interpreter/single.rs:97-104
                if (op.dst >= nc || op.a >= nc) return bad_wire(i);
                cells[op.dst] = c ? cells[op.a] : Cell{VREF_ZERO, ZERO_MID};
                P.algorithmic_bytes += B_UNARY;
                break;
            }
            case RV_MUL: {  // src/interpreter/single.rs:25-69
                if (op.dst >= nc || op.a >= nc || op.b >= nc) return bad_wire(i);
                const Cell A = cells[op.a], B = cells[op.b];
                Item it{ITEM_MUL, A.mid, B.mid, (uint32_t)n_masks, A.vref, B.vref, (uint32_t)P.n_and, 0};
                P.recon_pos.push_back((uint32_t)P.items.size());
                P.items.push_back(it);
                Cell R;
                R.mid = (uint32_t)n_masks + 1;  // mask_new
                const uint32_t va = A.vref >> 1, vb = B.vref >> 1;
                if (va == 0 && vb == 0) R.vref = (A.vref & B.vref) & 1;
                else if (va == 0) R.vref = (A.vref & 1) ? B.vref : VREF_ZERO;
                else if (vb == 0) R.vref = (B.vref & 1) ? A.vref : VREF_ZERO;
                else if (A.vref == B.vref) R.vref = A.vref;
                else if (va == vb) R.vref = VREF_ZERO;  // x & ~x
                else {
                    uint32_t vid = new_val(1 + std::max(vlevel[va], vlevel[vb]));
                    vg.push_back(VGate{vid, A.vref, B.vref, 1});
                    R.vref = vid << 1;
                }
                cells[op.dst] = R;
                n_masks += 2;
                P.n_and++;
                P.algorithmic_bytes += B_AND;
                break;
            }
            case RV_ASSERT_ZERO: {  // src/interpreter/single.rs:140-147
                if (op.a >= nc) return bad_wire(i);
                const Cell A = cells[op.a];
                Item it{ITEM_ASSERT, A.mid, 0, 0, A.vref, 0, 0, 0};
                P.recon_pos.push_back((uint32_t)P.items.size());
                P.items.push_back(it);
                P.n_assert++;
                P.algorithmic_bytes += B_ASSERT;
                break;
            }
            case RV_CONST: {  // src/interpreter/single.rs:151-155
                if (op.dst >= nc) return bad_wire(i);
                cells[op.dst] = Cell{c ? VREF_ONE : VREF_ZERO, ZERO_MID};
                P.algorithmic_bytes += B_LEAF;
                break;
            }
            default:
                err = "op " + std::to_string(i) + ": unknown opcode";
                return RV_E_ARG;
        }
        if (n_masks >= LIN_BASE - 2 || P.items.size() >= 0xFFFFFFF0ull || vlevel.size() >= 0x7FFFFFF0ull) {
            err = "circuit too large for 32-bit table indices";
            return RV_E_UNSUPPORTED;
        }
    }

    P.n_masks = (uint32_t)n_masks;
    P.n_lin = (uint32_t)lg.size();
    if ((uint64_t)P.n_masks + P.n_lin + 1 >= LIN_BASE) {
        err = "circuit too large for 32-bit row indices";
        return RV_E_UNSUPPORTED;
    }
    P.n_rows = P.n_masks + P.n_lin + 1;
    P.n_vals = (uint32_t)vlevel.size();
    P.n_online = (uint32_t)P.items.size();
    P.n_pre = (uint32_t)P.n_and;

    // ---- mask plane: counting sort by level; final row of a linear node = n_masks + its rank ----
    {
        uint32_t depth = 0;
        for (uint32_t l : llevel) depth = std::max(depth, l);
        P.llevel_off.assign(depth + 1, 0);
        for (uint32_t l : llevel) P.llevel_off[l]++;  // level l >= 1 counted at index l, shifted below
        // offsets: level l (1..depth) occupies [off[l-1], off[l])
        uint32_t run = 0;
        std::vector<uint32_t> start(depth + 1, 0);
        for (uint32_t l = 1; l <= depth; l++) {
            start[l] = run;
            run += P.llevel_off[l];
        }
        for (uint32_t l = 0; l < depth; l++) P.llevel_off[l] = start[l + 1];
        P.llevel_off[depth] = run;
        std::vector<uint32_t> rank(lg.size());
        std::vector<uint32_t> cursor(start);
        for (size_t n = 0; n < lg.size(); n++) rank[n] = cursor[llevel[n]]++;
        const uint32_t zero_row = P.zero_row();
        auto row_of = [&](uint32_t mid) -> uint32_t {
            if (mid == ZERO_MID) return zero_row;
            if (mid >= LIN_BASE) return P.n_masks + rank[mid - LIN_BASE];
            return mid;
        };
        P.lgates.resize(lg.size());
        for (size_t n = 0; n < lg.size(); n++) P.lgates[rank[n]] = LGate{row_of(lg[n].dst), row_of(lg[n].a), row_of(lg[n].b), 0};
        for (Item &it : P.items) {
            it.ra = row_of(it.ra);
            if (it.kind == ITEM_MUL) it.rb = row_of(it.rb);
        }
    }
    // ---- value plane: counting sort by level (value ids keep their creation order) ----
    {
        uint32_t depth = 0;
        for (const VGate &g : vg) depth = std::max(depth, vlevel[g.dst]);
        std::vector<uint32_t> cnt(depth + 2, 0);
        for (const VGate &g : vg) cnt[vlevel[g.dst]]++;
        P.vlevel_off.assign(depth + 1, 0);
        uint32_t run = 0;
        std::vector<uint32_t> cursor(depth + 1, 0);
        for (uint32_t l = 1; l <= depth; l++) {
            cursor[l] = run;
            run += cnt[l];
            P.vlevel_off[l - 1] = cursor[l];
        }
        P.vlevel_off[depth] = run;
        P.vgates.resize(vg.size());
        for (const VGate &g : vg) P.vgates[cursor[vlevel[g.dst]]++] = g;
    }
    return RV_OK;
}

}  // namespace rv
