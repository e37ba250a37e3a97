// rv_stream_plan.h -- host-side planning of a streaming proof (rv_prove_streaming): segmentation of the op list, liveness of the
// wires that cross segment boundaries, allocation (and recycling) of their slots in the device-resident cell file.  Pure host C++:
// shared by the product (rv_api.cu) and by the CPU replay of the test-suite (tests/hostsim).
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <new>
#include <string>
#include <thread>
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
inline int plan_stream_serial(const rv_op *ops, size_t n_ops, size_t gf2_cells, size_t window_ops, StreamPlan &plan, std::string &err,
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

// The planner above is one thread walking 10^8 ... 10^9 ops twice, several dependent cache misses per operand on a circuit with
// many wires: for a streaming proof it, not the GPU, sets the time (3 x 10^8 flat gates: 9 of 15 s).  This one produces the same
// plan with the per-op work spread over threads:
//   pass 1   liveness: the op list is cut into one range per thread; a wire's last-read segment is a running maximum and its
//            first-write segment a running minimum, so the ranges combine through atomic max / min;
//   pass 2a  per segment, any thread: dense renumbering of the segment's wires in order of first access, the renumbered copy of
//            its ops, and -- both decided by pass 1's two numbers alone -- its imports (first access is a read of a wire an
//            earlier segment wrote) and exports (written here, read by a later segment), in the serial planner's order;
//   pass 2b  the calling thread, segments in order, per CARRIED wire only: the slots of the cell file (taken at a wire's
//            export, returned after the segment of its last read), the one thing that has to be decided in sequence.
// A SizeHint that grows the wire file past the caller's gf2_cells (never the case behind largest_wires()) sends the call to
// the serial planner, which resizes its tables at the hint like the reference does; so does any op it would refuse.
inline int plan_stream(const rv_op *ops, size_t n_ops, size_t gf2_cells, size_t window_ops, StreamPlan &plan, std::string &err,
                       std::atomic<size_t> *planned = nullptr, unsigned n_threads = 0) {
    const size_t n_seg = stream_segments(n_ops, window_ops);
    if (n_threads == 0) n_threads = std::max(1u, std::min(8u, std::thread::hardware_concurrency() / 2));
    // the per-thread tables of pass 2a take 8 bytes per wire: keep them within 16 GB
    while (n_threads > 1 && (double)n_threads * (double)gf2_cells * 8.0 > 16e9) n_threads--;
    n_threads = (unsigned)std::min<size_t>(n_threads, n_seg);
    constexpr uint32_t HI = 0x80000000u;
    if (n_threads <= 1 || n_ops < ((size_t)1 << 16) || gf2_cells >= HI || n_seg >= HI - 2)
        return plan_stream_serial(ops, n_ops, gf2_cells, window_ops, plan, err, planned);

    struct Wire {              // what pass 1 learns about a wire cell
        uint32_t last_read;    // segment (+1) of its last read, 0 = never read
        uint32_t first_write;  // segment (+1) of its first write, 0 = never written
    };
    struct Loc {               // one thread's view of a wire inside the segment it is renumbering
        uint32_t stamp;        // segment (+1) in which `local` is valid
        uint32_t local;        // dense index there; bit 31: written in this segment
    };
    struct Scratch {           // what pass 2a leaves for pass 2b besides the Segment itself
        uint64_t masks = 0, on = 0, pre = 0, wit = 0, recon = 0;  // drawn / emitted by this segment
        std::atomic<int> done{0};  // 1 ready, -1 out of memory
    };
    auto reads = [](const rv_op &op, uint32_t r[2]) -> int {
        switch (op.opcode) {
            case RV_ADD: case RV_SUB: case RV_MUL: r[0] = op.a; r[1] = op.b; return 2;
            case RV_ADDC: case RV_SUBC: case RV_MULC: case RV_ASSERT_ZERO: r[0] = op.a; return 1;
            default: return 0;
        }
    };
    auto writes = [](const rv_op &op) { return op.opcode != RV_ASSERT_ZERO; };
    struct Free {
        void operator()(void *p) const { std::free(p); }
    };
    // calloc: the tables of a circuit with 10^8 wires stay zero pages until touched
    std::unique_ptr<Wire[], Free> wire((Wire *)std::calloc(std::max<size_t>(gf2_cells, 1), sizeof(Wire)));
    std::unique_ptr<uint32_t[], Free> slot1((uint32_t *)std::calloc(std::max<size_t>(gf2_cells, 1), 4));  // 1 + slot while carried, 0 = none
    std::unique_ptr<Scratch[]> scratch(new (std::nothrow) Scratch[n_seg]);
    if (!wire || !slot1 || !scratch) {
        err = "out of host memory";
        return RV_E_NOMEM;
    }
    const bool trace = std::getenv("RV_TRACE") != nullptr;  // phase times on stderr, like the compiler's
    auto t_last = std::chrono::steady_clock::now();
    double t_wait = 0;
    auto mark = [&](const char *what) {
        const auto n = std::chrono::steady_clock::now();
        if (trace) std::fprintf(stderr, "[rv_plan] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(n - t_last).count());
        t_last = n;
    };

    // ---- pass 1 ----
    std::atomic<bool> serial{false};
    {
        std::vector<std::thread> th;
        const size_t per = (n_ops + n_threads - 1) / n_threads;
        for (unsigned t = 0; t < n_threads; t++)
            th.emplace_back([&, t]() {
                const size_t a = std::min(n_ops, t * per), b = std::min(n_ops, a + per);
                for (size_t i = a; i < b; i++) {
                    const rv_op &op = ops[i];
                    if (op.domain == RV_SIZE_HINT) {
                        if (op.b <= gf2_cells) continue;
                        serial.store(true);
                        return;
                    }
                    uint32_t r[2];
                    const int nr = (op.domain != RV_GF2 || op.opcode == RV_RANDOM || op.opcode > RV_CONST) ? -1 : reads(op, r);
                    const bool wr = writes(op);
                    bool bad = nr < 0 || (wr && op.dst >= gf2_cells);
                    for (int k = 0; k < nr && !bad; k++) bad = r[k] >= gf2_cells;
                    if (bad) {  // the serial walk finds the first offending op and words the message
                        serial.store(true);
                        return;
                    }
                    const uint32_t seg1 = (uint32_t)(i / window_ops) + 1;
                    for (int k = 0; k < nr; k++) {
                        uint32_t *p = &wire[r[k]].last_read, cur = __atomic_load_n(p, __ATOMIC_RELAXED);
                        while (cur < seg1 && !__atomic_compare_exchange_n(p, &cur, seg1, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
                    }
                    if (wr) {
                        uint32_t *p = &wire[op.dst].first_write, cur = __atomic_load_n(p, __ATOMIC_RELAXED);
                        while ((cur == 0 || cur > seg1) && !__atomic_compare_exchange_n(p, &cur, seg1, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
                    }
                }
            });
        for (std::thread &t : th) t.join();
    }
    mark("pass 1 (liveness)");
    if (serial.load()) return plan_stream_serial(ops, n_ops, gf2_cells, window_ops, plan, err, planned);

    // ---- pass 2a (threads) + 2b (this thread) ----
    std::vector<Segment> &segs = plan.segs;
    if (segs.size() != n_seg) segs.resize(n_seg);
    std::atomic<size_t> next_seg{0};
    std::atomic<bool> stop{false};
    auto renumber = [&]() {
        std::unique_ptr<Loc[], Free> loc((Loc *)std::calloc(std::max<size_t>(gf2_cells, 1), sizeof(Loc)));
        for (;;) {
            const size_t sidx = next_seg.fetch_add(1);
            if (sidx >= n_seg || stop.load()) return;
            Segment &S = segs[sidx];
            Scratch &X = scratch[sidx];
            if (!loc) {
                X.done.store(-1, std::memory_order_release);
                continue;
            }
            try {
                S.a = sidx * window_ops;
                S.b = std::min(n_ops, S.a + window_ops);
                const uint32_t tag = (uint32_t)sidx + 1;
                uint32_t n_local = 0;
                auto local_of = [&](uint32_t c, bool is_read) -> uint32_t {
                    Loc &l = loc[c];
                    if (l.stamp != tag) {
                        l.stamp = tag;
                        l.local = n_local++;
                        const Wire w = wire[c];
                        if (is_read && w.first_write != 0 && w.first_write < tag) {  // first access is a read of a wire an earlier segment wrote: carried in
                            S.io.import_cells.push_back(l.local);
                            S.import_global.push_back(c | (w.last_read == tag ? HI : 0u));  // bit 31 (cleared in 2b): read here for the last time
                        }
                    }
                    return l.local & ~HI;
                };
                S.ops.reserve(S.b - S.a);  // (not resize: the zero fill would write the 96 MB of a default window twice)
                uint64_t masks = 0, on = 0, pre = 0, wit = 0, recon = 0;
                for (size_t i = S.a; i < S.b; i++) {
                    if (i + 16 < S.b) {  // a wide circuit's records are cache misses: ask for them a few ops ahead
                        const rv_op &f = ops[i + 16];
                        if (f.a < gf2_cells) __builtin_prefetch(&loc[f.a]);
                        if (f.b < gf2_cells) __builtin_prefetch(&loc[f.b]);
                        if (f.dst < gf2_cells) __builtin_prefetch(&loc[f.dst]);
                    }
                    rv_op op = ops[i];
                    if (op.domain != RV_GF2) continue;
                    uint32_t r[2];
                    const int nr = reads(op, r);
                    if (nr >= 1) op.a = local_of(r[0], true);
                    if (nr >= 2) op.b = local_of(r[1], true);
                    if (writes(op)) {
                        const uint32_t c = op.dst;
                        op.dst = local_of(c, false);
                        if (!(loc[c].local & HI)) {  // first write of this segment
                            loc[c].local |= HI;
                            if (wire[c].last_read > tag) {  // read by a later segment: leaves through the cell file
                                S.io.export_cells.push_back(op.dst);
                                S.export_global.push_back(c);
                            }
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
                S.n_local = std::max<uint32_t>(n_local, 1);
                S.import_slot.resize(S.import_global.size());
                S.export_slot.resize(S.export_global.size());
                X.masks = masks, X.on = on, X.pre = pre, X.wit = wit, X.recon = recon;
                X.done.store(1, std::memory_order_release);
            } catch (const std::bad_alloc &) {
                X.done.store(-1, std::memory_order_release);
            }
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 0; t < n_threads; t++) th.emplace_back(renumber);
    struct Join {
        std::vector<std::thread> &th;
        std::atomic<bool> &stop;
        ~Join() {
            stop.store(true);
            for (std::thread &t : th) t.join();
        }
    } join{th, stop};

    std::vector<uint32_t> free_slots, to_free;
    uint32_t n_slots = 0;
    uint64_t masks = 0, on = 0, pre = 0, wit = 0, recon = 0;
    try {
        for (size_t sidx = 0; sidx < n_seg; sidx++) {
            Segment &S = segs[sidx];
            Scratch &X = scratch[sidx];
            int st;
            const auto w0 = std::chrono::steady_clock::now();
            while ((st = X.done.load(std::memory_order_acquire)) == 0) std::this_thread::sleep_for(std::chrono::microseconds(100));
            t_wait += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
            if (st < 0) throw std::bad_alloc();
            S.mask0 = masks, S.on0 = on, S.pre0 = pre, S.wit0 = wit, S.recon0 = recon;
            masks += X.masks, on += X.on, pre += X.pre, wit += X.wit, recon += X.recon;
            to_free.clear();
            const size_t ni = S.import_global.size(), ne = S.export_global.size();
            for (size_t j = 0; j < ni; j++) {
                if (j + 16 < ni) __builtin_prefetch(&slot1[S.import_global[j + 16] & ~HI]);
                const bool last = (S.import_global[j] & HI) != 0;
                const uint32_t c = (S.import_global[j] &= ~HI);
                S.import_slot[j] = slot1[c] - 1;
                if (last) to_free.push_back(c);
            }
            for (uint32_t c : to_free) {  // (a slot freed here may be reused at once: imports are read at the start of the segment, exports written at its end)
                free_slots.push_back(slot1[c] - 1);
                slot1[c] = 0;
            }
            for (size_t j = 0; j < ne; j++) {
                if (j + 16 < ne) __builtin_prefetch(&slot1[S.export_global[j + 16]]);
                const uint32_t c = S.export_global[j];
                if (slot1[c] == 0) {
                    if (!free_slots.empty()) {
                        slot1[c] = free_slots.back() + 1;
                        free_slots.pop_back();
                    } else slot1[c] = ++n_slots;
                }
                S.export_slot[j] = slot1[c] - 1;
            }
            if (planned) planned->store(sidx + 1, std::memory_order_release);
        }
    } catch (const std::bad_alloc &) {
        err = "out of host memory";
        return RV_E_NOMEM;
    }
    mark("pass 2 (segments)");
    if (trace) std::fprintf(stderr, "[rv_plan]   of which waiting for 2a   %8.1f ms\n", t_wait);
    plan.n_slots = n_slots;
    plan.masks = masks, plan.tot_on = on, plan.tot_pre = pre, plan.tot_inputs = wit, plan.tot_recon = recon;
    plan.gf2_cells = gf2_cells;
    return RV_OK;
}

}  // namespace rv
