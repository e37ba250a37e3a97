// rv_compile.cpp -- see rv_compile.h.  Pure host C++ (no CUDA), so the CPU test-suite can exercise it without a GPU.
#include "rv_compile.h"

#include <algorithm>
#include <cstring>
#include <initializer_list>

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

// Mask-plane VM: the XOR network re-expressed over a small pool of shared-memory cells so that the dependent chain of
// the level-synchronous walk is LDS -> XOR -> STS instead of L2 round trips.  Fresh rows are brought into cells by
// asynchronous LOADs issued VM_DELTA levels early; only rows that the item plane reads are written back (`row`).
// Cells are assigned by a linear scan over levels: a cell is free again one level after its value's last use.
void build_mask_vm(Program &P) {
    P.vm.clear();
    P.vm_level_off.clear();
    P.vm_cells = 0;
    const uint32_t depth = (uint32_t)P.llevel_off.size() - 1;
    if (depth == 0) return;
    const uint32_t n_rows = P.n_rows, n_masks = P.n_masks, n_lin = P.n_lin;
    const uint32_t NEVER = 0xFFFFFFFFu;
    std::vector<uint32_t> first(n_rows, NEVER), last(n_rows, 0);
    std::vector<uint8_t> exported(n_rows, 0);
    for (uint32_t l = 0; l < depth; l++)
        for (uint32_t g = P.llevel_off[l]; g < P.llevel_off[l + 1]; g++)
            for (uint32_t r : {P.lgates[g].a, P.lgates[g].b}) {
                if (first[r] == NEVER) first[r] = l + 1;
                last[r] = l + 1;
            }
    for (const Item &it : P.items) {
        if (it.kind == ITEM_MUL) exported[it.ra] = exported[it.rb] = 1;
        else if (it.kind == ITEM_ASSERT) exported[it.ra] = 1;
    }
    // VM level of an original level L (1-based) is L + VM_DELTA - 1; the LOAD of a fresh row first used at L sits at L - 1.
    const uint32_t n_levels = depth + VM_DELTA;
    std::vector<uint32_t> cnt(n_levels + 1, 0);
    for (uint32_t r = 0; r < n_masks; r++)
        if (first[r] != NEVER) cnt[first[r] - 1]++;
    for (uint32_t l = 0; l < depth; l++) cnt[l + VM_DELTA] += P.llevel_off[l + 1] - P.llevel_off[l];
    P.vm_level_off.assign(n_levels + 1, 0);
    for (uint32_t l = 0; l < n_levels; l++) P.vm_level_off[l + 1] = P.vm_level_off[l] + cnt[l];
    P.vm.resize(P.vm_level_off[n_levels]);
    std::vector<uint32_t> cursor(P.vm_level_off.begin(), P.vm_level_off.end() - 1);
    // provisional instructions hold rows; cells are assigned in the scan below
    for (uint32_t r = 0; r < n_masks; r++)
        if (first[r] != NEVER) P.vm[cursor[first[r] - 1]++] = VmInstr{VM_LOAD, r, 0, 0};
    for (uint32_t l = 0; l < depth; l++)
        for (uint32_t g = P.llevel_off[l]; g < P.llevel_off[l + 1]; g++)
            P.vm[cursor[l + VM_DELTA]++] = VmInstr{P.lgates[g].dst, P.lgates[g].a, P.lgates[g].b, 0};
    std::vector<uint32_t> cell_of(n_rows, VM_NONE);
    std::vector<std::vector<uint32_t>> free_at(n_levels + 2);
    std::vector<uint32_t> free_list;
    uint32_t n_cells = 0;
    auto alloc = [&]() -> uint32_t {
        if (!free_list.empty()) {
            uint32_t c = free_list.back();
            free_list.pop_back();
            return c;
        }
        return n_cells++;
    };
    for (uint32_t l = 0; l < n_levels; l++) {
        for (uint32_t c : free_at[l]) free_list.push_back(c);
        free_at[l].clear();
        free_at[l].shrink_to_fit();
        for (uint32_t k = P.vm_level_off[l]; k < P.vm_level_off[l + 1]; k++) {
            VmInstr &in = P.vm[k];
            if (in.dst == VM_LOAD) {
                const uint32_t r = in.a, c = alloc();
                cell_of[r] = c;
                free_at[last[r] + VM_DELTA].push_back(c);  // last use at VM level last + DELTA - 1
                in.dst = VM_LOAD | c;
            } else {
                const uint32_t r = in.dst;
                in.a = cell_of[in.a];
                in.b = cell_of[in.b];
                in.row = exported[r] ? r : VM_NONE;
                if (first[r] != NEVER) {
                    const uint32_t c = alloc();
                    cell_of[r] = c;
                    free_at[last[r] + VM_DELTA].push_back(c);
                    in.dst = c;
                } else {
                    in.dst = VM_NONE;
                }
            }
        }
    }
    (void)n_lin;
    P.vm_cells = n_cells;
}

// Split levels wider than `maxw` into consecutive sub-levels (always legal: instructions of a level are independent).
static void split_levels(std::vector<uint32_t> &off, uint32_t maxw) {
    if (off.size() < 2) return;
    std::vector<uint32_t> out;
    out.push_back(off[0]);
    for (size_t l = 0; l + 1 < off.size(); l++) {
        uint32_t s = off[l];
        const uint32_t e = off[l + 1];
        while (e - s > maxw) {
            s += maxw;
            out.push_back(s);
        }
        out.push_back(e);
    }
    off.swap(out);
}

// ---- value-plane technology mapping: K-feasible cuts, depth first (the FPGA "priority cuts" scheme) ------------------
constexpr int LUT_K = 6, CUTS_PER_NODE = 6;
struct Cut {
    uint32_t leaf[LUT_K];
    uint8_t n;
    uint32_t depth;
};

static bool merge_cuts(const Cut &a, const Cut &b, Cut &o) {
    int i = 0, j = 0, n = 0;
    while (i < a.n || j < b.n) {
        uint32_t v;
        if (j >= b.n || (i < a.n && a.leaf[i] < b.leaf[j])) v = a.leaf[i++];
        else if (i >= a.n || b.leaf[j] < a.leaf[i]) v = b.leaf[j++];
        else {
            v = a.leaf[i];
            i++;
            j++;
        }
        if (n == LUT_K) return false;
        o.leaf[n++] = v;
    }
    o.n = (uint8_t)n;
    return true;
}

void build_value_luts(Program &P, const std::vector<VGate> &vg /* creation (topological) order */, bool map) {
    const uint32_t n_vals = P.n_vals;
    std::vector<uint32_t> gate_of(n_vals, 0xFFFFFFFFu);  // vid -> index into vg, or none for leaves (inputs, constant)
    for (uint32_t g = 0; g < vg.size(); g++) gate_of[vg[g].dst] = g;
    std::vector<uint8_t> required(n_vals, 0);
    for (const Item &it : P.items) {
        required[it.va >> 1] = 1;
        if (it.kind == ITEM_MUL) required[it.vb >> 1] = 1;
    }
    std::vector<uint32_t> depth(n_vals, 0);
    std::vector<Cut> best(vg.size());  // chosen cut per gate
    if (map) {
        std::vector<Cut> cuts((size_t)vg.size() * CUTS_PER_NODE);
        std::vector<uint8_t> ncuts(vg.size(), 0);
        auto cut_set = [&](uint32_t vid, Cut *tmp, int &n) {  // the node's stored cuts plus its trivial cut
            n = 0;
            if (vid == 0) {  // the constant contributes no leaf
                tmp[n].n = 0;
                tmp[n].depth = 0;
                n++;
                return;
            }
            const uint32_t g = gate_of[vid];
            if (g != 0xFFFFFFFFu)
                for (int i = 0; i < ncuts[g]; i++) tmp[n++] = cuts[(size_t)g * CUTS_PER_NODE + i];
            tmp[n].n = 1;
            tmp[n].leaf[0] = vid;
            tmp[n].depth = 0;
            n++;
        };
        Cut ca[CUTS_PER_NODE + 1], cb[CUTS_PER_NODE + 1], cand[(CUTS_PER_NODE + 1) * (CUTS_PER_NODE + 1)];
        for (uint32_t g = 0; g < vg.size(); g++) {
            int na, nb, nc = 0;
            cut_set(vg[g].a >> 1, ca, na);
            cut_set(vg[g].b >> 1, cb, nb);
            for (int i = 0; i < na; i++)
                for (int j = 0; j < nb; j++) {
                    Cut &o = cand[nc];
                    if (!merge_cuts(ca[i], cb[j], o)) continue;
                    uint32_t d = 0;
                    for (int k = 0; k < o.n; k++) d = std::max(d, depth[o.leaf[k]]);
                    o.depth = d + 1;
                    bool dup = false;
                    for (int k = 0; k < nc && !dup; k++) dup = cand[k].n == o.n && std::memcmp(cand[k].leaf, o.leaf, o.n * 4) == 0;
                    if (!dup) nc++;
                }
            std::sort(cand, cand + nc, [](const Cut &x, const Cut &y) { return x.depth != y.depth ? x.depth < y.depth : x.n < y.n; });
            const int keep = std::min(nc, CUTS_PER_NODE);
            for (int i = 0; i < keep; i++) cuts[(size_t)g * CUTS_PER_NODE + i] = cand[i];
            ncuts[g] = (uint8_t)keep;
            best[g] = cand[0];
            depth[vg[g].dst] = cand[0].depth;
        }
    } else {
        for (uint32_t g = 0; g < vg.size(); g++) {
            Cut c;
            c.n = 0;
            uint32_t a = vg[g].a >> 1, b = vg[g].b >> 1;
            if (a > b) std::swap(a, b);
            if (a) c.leaf[c.n++] = a;
            if (b && b != a) c.leaf[c.n++] = b;
            uint32_t d = 0;
            for (int k = 0; k < c.n; k++) d = std::max(d, depth[c.leaf[k]]);
            c.depth = d + 1;
            best[g] = c;
            depth[vg[g].dst] = c.depth;
        }
    }
    // cover: walk backwards from the values the item plane reads
    for (size_t g = vg.size(); g-- > 0;) {
        if (!required[vg[g].dst]) continue;
        for (int k = 0; k < best[g].n; k++) required[best[g].leaf[k]] = 1;
    }
    // final levels (over the chosen cover) and truth tables
    static const uint64_t PAT[6] = {0xAAAAAAAAAAAAAAAAull, 0xCCCCCCCCCCCCCCCCull, 0xF0F0F0F0F0F0F0F0ull,
                                    0xFF00FF00FF00FF00ull, 0xFFFF0000FFFF0000ull, 0xFFFFFFFF00000000ull};
    std::vector<uint32_t> level(n_vals, 0), stamp(n_vals, 0);
    std::vector<uint64_t> tmp(n_vals, 0);
    std::vector<LutInstr> luts;
    std::vector<uint32_t> lut_level;
    std::vector<uint32_t> stack;
    uint32_t epoch = 0, max_level = 0;
    for (uint32_t g = 0; g < vg.size(); g++) {
        const uint32_t out = vg[g].dst;
        if (!required[out]) continue;
        const Cut &c = best[g];
        epoch++;
        uint32_t lv = 0;
        for (int k = 0; k < c.n; k++) {
            stamp[c.leaf[k]] = epoch;
            tmp[c.leaf[k]] = PAT[k];
            lv = std::max(lv, level[c.leaf[k]]);
        }
        stamp[0] = epoch;
        tmp[0] = 0;
        // evaluate the cone bottom-up with an explicit stack (post-order)
        stack.clear();
        stack.push_back(out);
        while (!stack.empty()) {
            const uint32_t v = stack.back();
            if (stamp[v] == epoch) {
                stack.pop_back();
                continue;
            }
            const VGate &gt = vg[gate_of[v]];
            const uint32_t a = gt.a >> 1, b = gt.b >> 1;
            const bool ra = stamp[a] == epoch, rb = stamp[b] == epoch;
            if (ra && rb) {
                const uint64_t x = tmp[a] ^ (0ull - (gt.a & 1)), y = tmp[b] ^ (0ull - (gt.b & 1));
                tmp[v] = gt.op ? (x & y) : (x ^ y);
                stamp[v] = epoch;
                stack.pop_back();
            } else {
                if (!ra) stack.push_back(a);
                if (!rb) stack.push_back(b);
            }
        }
        LutInstr li;
        li.dst = out;
        for (int k = 0; k < 6; k++) li.in[k] = k < c.n ? c.leaf[k] : 0;
        li.pad = 0;
        li.pad2 = 0;
        li.tt = tmp[out];
        level[out] = lv + 1;
        max_level = std::max(max_level, lv + 1);
        luts.push_back(li);
        lut_level.push_back(lv + 1);
    }
    // counting sort by level
    P.lut_depth = max_level;
    P.lut_level_off.assign(max_level + 1, 0);
    std::vector<uint32_t> cursor(max_level + 2, 0);
    for (uint32_t l : lut_level) cursor[l]++;
    uint32_t run = 0;
    for (uint32_t l = 1; l <= max_level; l++) {
        const uint32_t c = cursor[l];
        cursor[l] = run;
        P.lut_level_off[l - 1] = run;
        run += c;
    }
    P.lut_level_off[max_level] = run;
    P.luts.resize(luts.size());
    for (size_t i = 0; i < luts.size(); i++) P.luts[cursor[lut_level[i]]++] = luts[i];
    split_levels(P.lut_level_off, LUT_LEVEL_MAX);
}

// ---- step streams (see rv_compile.h) ---------------------------------------------------------------------------------
static void emit_vm_steps(Program &P) {
    P.vm_steps.clear();
    P.n_vm_steps = 0;
    if (P.vm.empty()) return;
    const uint32_t scratch = P.vm_cells;  // one extra cell absorbs the writes of empty slots and of export-only XORs
    const VmInstr nop{scratch, scratch, scratch, VM_ROW_NONE};
    const size_t n_levels = P.vm_level_off.size() - 1;
    for (size_t l = 0; l < n_levels; l++) {
        const uint32_t s = P.vm_level_off[l], e = P.vm_level_off[l + 1];
        if (e == s) continue;
        const uint32_t steps = (e - s + VM_STEP - 1) / VM_STEP;
        for (uint32_t k = 0; k < steps; k++) {
            const bool last = k + 1 == steps;
            const bool chunk_end = (P.n_vm_steps + 1) % VM_STEPS_PER_CHUNK == 0;
            for (uint32_t t = 0; t < VM_STEP; t++) {
                const uint32_t g = s + k * VM_STEP + t;
                VmInstr o = nop;
                if (g < e) {
                    const VmInstr &in = P.vm[g];
                    if (in.dst & VM_LOAD) o = VmInstr{VM_F_LOAD | (in.dst & ~VM_LOAD), in.a, 0, VM_ROW_NONE};
                    else o = VmInstr{in.dst == VM_NONE ? scratch : in.dst, in.a, in.b, in.row == VM_NONE ? VM_ROW_NONE : in.row};
                }
                if (last || chunk_end) o.dst |= VM_F_BAR;
                P.vm_steps.push_back(o);
            }
            P.n_vm_steps++;
        }
    }
}

static void emit_lut_steps(Program &P) {
    P.lut_steps.clear();
    P.n_lut_steps = 0;
    if (P.luts.empty()) return;
    LutInstr nop;
    std::memset(&nop, 0, sizeof nop);
    nop.dst = P.n_vals;  // scratch value slot
    const size_t n_levels = P.lut_level_off.size() - 1;
    for (size_t l = 0; l < n_levels; l++) {
        const uint32_t s = P.lut_level_off[l], e = P.lut_level_off[l + 1];
        if (e == s) continue;
        const uint32_t steps = (e - s + LUT_STEP - 1) / LUT_STEP;
        for (uint32_t k = 0; k < steps; k++) {
            const bool last = k + 1 == steps;
            const bool chunk_end = (P.n_lut_steps + 1) % LUT_STEPS_PER_CHUNK == 0;
            for (uint32_t t = 0; t < LUT_STEP; t++) {
                const uint32_t g = s + k * LUT_STEP + t;
                LutInstr o = g < e ? P.luts[g] : nop;
                o.pad = (last || chunk_end) ? LUT_F_BAR : 0;
                P.lut_steps.push_back(o);
            }
            P.n_lut_steps++;
        }
    }
}

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
    build_mask_vm(P);
    split_levels(P.vm_level_off, VM_LEVEL_MAX);
    build_value_luts(P, vg, vg.size() <= LUT_MAP_MAX_GATES);
    emit_vm_steps(P);
    emit_lut_steps(P);
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
