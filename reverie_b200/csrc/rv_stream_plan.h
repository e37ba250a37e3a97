// rv_stream_plan.h -- host-side planning of a streaming proof (rv_prove_streaming): segmentation of the op list, liveness of the
// wires that cross segment boundaries, allocation (and recycling) of their slots in the device-resident cell file.  Pure host C++:
// shared by the product (rv_api.cu) and by the CPU replay of the test-suite (tests/hostsim).
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <new>
#include <string>
#include <vector>

#include "rv_compile.h"

namespace rv {

constexpr uint32_t NO_SLOT = 0xFFFFFFFFu;

struct Segment {
    size_t a = 0, b = 0;               // op range in the whole circuit
    std::vector<rv_op> ops;            // its ops with wire cells renumbered densely (dropped after compilation)
    uint32_t n_local = 0;
    SegmentIO io;
    std::vector<uint32_t> import_slot, export_slot;
    std::vector<uint32_t> import_global, export_global;  // the same wires by their cell index in the whole circuit (consistency check)
    rv_circuit *c = nullptr;
    uint64_t mask0 = 0, on0 = 0, pre0 = 0, wit0 = 0, recon0 = 0;  // what the ops before this segment drew / emitted
    uint32_t n_on = 0, n_pre = 0, n_in = 0, n_recon = 0, n_imp = 0, n_exp = 0;
    const uint32_t *d_leaf_ids = nullptr, *d_imp_slot = nullptr, *d_exp_slot = nullptr, *d_exp_row = nullptr, *d_exp_vref = nullptr;
    int rc = RV_OK;
    std::string err;
};
// Liveness + segmentation of a streaming proof (host only; rv_stream_plan_check exposes it to the CPU test-suite).
struct StreamPlan {
    std::vector<Segment> segs;
    uint32_t n_slots = 0;
    uint64_t masks = 0, tot_on = 0, tot_pre = 0, tot_inputs = 0, tot_recon = 0;
    size_t gf2_cells = 0;
};
// `planned` (optional) is advanced to k + 1 as soon as segment k is final, so that its compilation can start while the rest is
// still being planned; plan.segs is sized before the first segment is published and never reallocated.
inline size_t stream_segments(size_t n_ops, size_t window_ops) { return std::max<size_t>(1, (n_ops + window_ops - 1) / window_ops); }
inline int plan_stream(const rv_op *ops, size_t n_ops, size_t gf2_cells, size_t window_ops, StreamPlan &plan, std::string &err,
                       std::atomic<size_t> *planned = nullptr) {
    const size_t n_seg = stream_segments(n_ops, window_ops);
    auto reads = [](const rv_op &op, uint32_t r[2]) -> int {
        switch (op.opcode) {
            case RV_ADD: case RV_SUB: case RV_MUL: r[0] = op.a; r[1] = op.b; return 2;
            case RV_ADDC: case RV_SUBC: case RV_MULC: case RV_ASSERT_ZERO: r[0] = op.a; return 1;
            default: return 0;
        }
    };
    auto writes = [](const rv_op &op) { return op.opcode != RV_ASSERT_ZERO; };
    auto bad_wire = [&](size_t i) {
        err = "op " + std::to_string(i) + ": wire index out of range for the given wire_counts";
        return RV_E_ARG;
    };
    // per wire cell: segment (+1) of its last read, segment (+1) in which `local` is valid, its dense index there, its slot in the
    // cell file, and 0 = never written / 1 = written by an earlier segment / 2 = written in the segment being scanned
    std::vector<uint32_t> last_read_seg, stamp, local, slot_of;
    std::vector<uint8_t> written;
    try {
        // ---- pass 1: validation, SizeHint, liveness ----
        last_read_seg.assign(gf2_cells, 0);
        for (size_t i = 0; i < n_ops; i++) {
            const rv_op &op = ops[i];
            if (op.domain == RV_SIZE_HINT) {  // src/interpreter/combine.rs:122-129
                if (op.b > gf2_cells) {
                    gf2_cells = op.b;
                    last_read_seg.resize(gf2_cells, 0);
                }
                continue;
            }
            if (op.domain != RV_GF2 || op.opcode == RV_RANDOM || op.opcode > RV_CONST) {
                err = "op " + std::to_string(i) + ": streaming mode serves GF(2) circuits without Random / Z64 / B2A";
                return RV_E_UNSUPPORTED;
            }
            uint32_t r[2];
            const int nr = reads(op, r);
            for (int k = 0; k < nr; k++) {
                if (r[k] >= gf2_cells) return bad_wire(i);
                last_read_seg[r[k]] = (uint32_t)(i / window_ops) + 1;
            }
            if (writes(op) && op.dst >= gf2_cells) return bad_wire(i);
        }
        // a wire index may exceed a later SizeHint's predecessor: the reference grows its wire file at the hint, we size it once
        stamp.assign(gf2_cells, 0);
        local.assign(gf2_cells, 0);
        slot_of.assign(gf2_cells, NO_SLOT);
        written.assign(gf2_cells, 0);
    } catch (const std::bad_alloc &) {
        err = "out of host memory";
        return RV_E_NOMEM;
    }
    // ---- pass 2: the segments, in order ----
    std::vector<Segment> &segs = plan.segs;
    if (segs.size() != n_seg) segs.resize(n_seg);  // (a caller that compiles concurrently has sized it already)
    std::vector<uint32_t> free_slots, to_free, wrote;
    uint32_t n_slots = 0;
    uint64_t masks = 0, on = 0, pre = 0, wit = 0, recon = 0;
    for (size_t sidx = 0; sidx < n_seg; sidx++) {
        Segment &S = segs[sidx];
        S.a = sidx * window_ops;
        S.b = std::min(n_ops, S.a + window_ops);
        S.mask0 = masks, S.on0 = on, S.pre0 = pre, S.wit0 = wit, S.recon0 = recon;
        const uint32_t tag = (uint32_t)sidx + 1;
        to_free.clear();
        wrote.clear();
        auto local_of = [&](uint32_t c, bool is_read) -> uint32_t {
            if (stamp[c] != tag) {
                stamp[c] = tag;
                local[c] = S.n_local++;
                if (is_read && written[c]) {  // first access is a read of a wire an earlier segment wrote: carried in
                    S.io.import_cells.push_back(local[c]);
                    S.import_slot.push_back(slot_of[c]);
                    S.import_global.push_back(c);
                    if (last_read_seg[c] == tag) to_free.push_back(c);  // ... for the last time: its slot is free after this segment
                }
            }
            return local[c];
        };
        S.ops.reserve(S.b - S.a);
        for (size_t i = S.a; i < S.b; i++) {
            rv_op op = ops[i];
            if (op.domain != RV_GF2) continue;
            uint32_t r[2];
            const int nr = reads(op, r);
            if (nr >= 1) op.a = local_of(r[0], true);
            if (nr >= 2) op.b = local_of(r[1], true);
            if (writes(op)) {
                const uint32_t c = op.dst;
                op.dst = local_of(c, false);
                if (written[c] != 2) {  // first write of this segment
                    written[c] = 2;
                    wrote.push_back(c);
                }
            }
            S.ops.push_back(op);
            switch (op.opcode) {
                case RV_INPUT: masks += 1, on += 1, wit += 1; break;
                case RV_MUL: masks += 2, on += 1, pre += 1, recon += 1; break;
                case RV_ASSERT_ZERO: on += 1, recon += 1; break;
                default: break;
            }
        }
        for (uint32_t c : to_free) {
            free_slots.push_back(slot_of[c]);
            slot_of[c] = NO_SLOT;
        }
        // wires written here and read by a later segment leave through the cell file (a slot freed above may be reused at once:
        // imports are read at the start of the segment, exports written at its end)
        for (uint32_t c : wrote) {
            written[c] = 1;
            if (last_read_seg[c] <= tag) continue;
            if (slot_of[c] == NO_SLOT) {
                if (!free_slots.empty()) {
                    slot_of[c] = free_slots.back();
                    free_slots.pop_back();
                } else slot_of[c] = n_slots++;
            }
            S.io.export_cells.push_back(local[c]);
            S.export_slot.push_back(slot_of[c]);
            S.export_global.push_back(c);
        }
        S.n_local = std::max<uint32_t>(S.n_local, 1);
        if (planned) planned->store(sidx + 1, std::memory_order_release);
    }
    plan.n_slots = n_slots;
    plan.masks = masks, plan.tot_on = on, plan.tot_pre = pre, plan.tot_inputs = wit, plan.tot_recon = recon;
    plan.gf2_cells = gf2_cells;
    return RV_OK;
}

}  // namespace rv
