// rv_compile.cpp -- see rv_compile.h.  Pure host C++ (no CUDA), so the CPU test-suite can exercise it without a GPU.
#include "rv_compile.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <thread>
#include <initializer_list>
#ifdef __linux__
#include <sys/mman.h>
#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23  // Linux 5.14
#endif
#endif

namespace rv {
namespace {

constexpr uint32_t ZERO_MID = 0xFFFFFFFFu;  // "the all-zero mask" until row numbers are final
constexpr uint32_t LIN_BASE = 0x80000000u;  // provisional ids of linear nodes: LIN_BASE + creation index
constexpr uint32_t IMP_BASE = 0x70000000u;  // provisional ids of imported mask rows (streaming segments): IMP_BASE + import index
constexpr uint32_t VREF_ZERO = 0;           // vid 0, not negated
constexpr uint32_t VREF_ONE = 1;            // vid 0, negated
constexpr uint32_t NONE32 = 0xFFFFFFFFu;

struct Cell {
    uint32_t vref;  // value id << 1 | negate
    uint32_t mid;   // fresh PRG index, LIN_BASE + linear node, or ZERO_MID
    uint32_t uref;  // u-plane (online verifier) value ref
};

// algorithmic HBM bytes per gate over all 256 repetitions, plus the 16-byte descriptor
constexpr uint64_t B_AND = 2048 + 16, B_XOR = 1536 + 16, B_UNARY = 1024 + 16, B_INPUT = 768 + 16, B_ASSERT = 768 + 16,
                   B_LEAF = 512 + 16;

// =====================================================================================================================
//  Technology mapping: K-feasible cuts, depth first (the FPGA "priority cuts" scheme), shared by both planes.
//  A network is a topologically ordered list of 2-input gates over node ids; id 0 is the constant 0 and is free.
// =====================================================================================================================
struct Trace {  // RV_TRACE=1: phase times of the compiler on stderr
    bool on = std::getenv("RV_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void mark(const char *what) {
        if (!on) return;
        const auto n = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[rv_compile] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

// The tables of a 10^8-gate circuit are gigabytes that are written exactly once, front to back, and every first touch of a
// page is a fault: measured, 40-50 % of the op walk (0.3-0.5 s per GiB).  Helper threads populate the reserved ranges ahead of
// the writer, in transparent huge pages where the kernel grants them.  MADV_POPULATE_WRITE maps pages without changing their
// contents, so running beside the writer (or after the memory is gone: it fails with ENOMEM) is harmless; on kernels without it
// the helpers stop at the first EINVAL and the walk takes its faults itself, as before.
// It pays when the walk is bound by its own stores (flat 2 x 10^7: 2.3 -> 1.3 s; SSA circuits with narrow layers: 2.0 -> 1.3 s).
// It loses when the walk is bound by cache misses on its operands (layers of 2^20 wires: 3.2 -> 4.6 s): there a page zeroed by
// the fault handler arrives cache-hot, a pre-populated one costs a read-for-ownership miss per line on top of the operand
// misses.  So compile() measures the operands' mean distance from the destination in its counting pass and asks for the
// helpers only when the cells an op reads are likely to sit in the L2 (RV_PREFAULT=0 / 1 overrides, for measurements).
struct Prefault {
    struct Range {
        uintptr_t a, e;
    };
    std::vector<Range> ranges;
    std::vector<std::thread> th;
    std::atomic<size_t> next{0};
    std::atomic<bool> stop{false};
    size_t rounds = 1;
    static constexpr size_t MIN_BYTES = 16u << 20, SLICE = 8u << 20, HUGE = 2u << 20;
    void add(const void *p, size_t bytes) {
#ifdef __linux__
        if (bytes < MIN_BYTES) return;
        const uintptr_t a = ((uintptr_t)p + 4095) & ~(uintptr_t)4095, e = ((uintptr_t)p + bytes) & ~(uintptr_t)4095;
        const uintptr_t ha = (a + HUGE - 1) & ~(uintptr_t)(HUGE - 1), he = e & ~(uintptr_t)(HUGE - 1);
        if (he > ha) madvise((void *)ha, he - ha, MADV_HUGEPAGE);
        ranges.push_back(Range{a, e});
        rounds = std::max(rounds, std::min<size_t>(1024, (e - a) / SLICE));
#else
        (void)p, (void)bytes;
#endif
    }
    template <class T>
    void add(const std::vector<T> &v) {  // the reserved capacity of a table the caller is about to fill
        add(v.data(), v.capacity() * sizeof(T));
    }
    // every table grows at its own constant rate, so slice r of each range is populated in round r
    void start(unsigned n_threads) {
#ifdef __linux__
        if (ranges.empty() || n_threads == 0) return;
        const size_t n_jobs = rounds * ranges.size();
        for (unsigned t = 0; t < n_threads; t++)
            th.emplace_back([this, n_jobs]() {
                for (size_t j; !stop.load(std::memory_order_relaxed) && (j = next.fetch_add(1)) < n_jobs;) {
                    const Range &R = ranges[j % ranges.size()];
                    const size_t r = j / ranges.size(), len = R.e - R.a;
                    const uintptr_t b = R.a + ((len / rounds * r) & ~(uintptr_t)(HUGE - 1));
                    const uintptr_t f = r + 1 == rounds ? R.e : R.a + ((len / rounds * (r + 1)) & ~(uintptr_t)(HUGE - 1));
                    if (f > b && madvise((void *)b, f - b, MADV_POPULATE_WRITE) != 0) stop.store(true);
                }
            });
#else
        (void)n_threads;
#endif
    }
    void finish() {
        stop.store(true);
        for (std::thread &t : th) t.join();
        th.clear();
    }
    ~Prefault() { finish(); }
};
constexpr uint32_t VM_LOAD_WINDOW = 4;  // how many levels early (beyond VM_DELTA) a LOAD may be issued to balance the steps
constexpr size_t PREFAULT_MIN_OPS = 1u << 20, PREFAULT_MAX_REACH_BYTES = 1u << 20;

constexpr int MAP_K = 6, CUTS_PER_NODE = 4;  // 4 kept cuts map SHA-256 / AES-128 exactly as deep as 6 do, in 60 % of the time
struct MGate {
    uint32_t out, a, b;  // a, b = id << 1 | negate
    uint32_t op;         // 0 xor, 1 and
    MGate() = default;
    MGate(uint32_t out_, uint32_t a_, uint32_t b_, uint32_t op_) : out(out_), a(a_), b(b_), op(op_) {}  // for emplace_back, like Item
};
struct MNode {  // one mapped node: out = f(leaf[0..n)), f given by its truth table over the leaves
    uint32_t out;
    uint32_t leaf[MAP_K];
    uint32_t n;
    uint32_t level;
    uint64_t tt;
};
struct Cut {
    uint32_t leaf[MAP_K];  // sorted
    uint32_t depth;
    uint32_t n;
    uint64_t sig;  // OR of 1 << (leaf & 63): popcount(sig_a | sig_b) > K proves a merge infeasible without doing it
};

inline bool merge_cuts(const Cut &a, const Cut &b, Cut &o) {
    uint32_t i = 0, j = 0, n = 0;
    while (i < a.n && j < b.n) {
        const uint32_t x = a.leaf[i], y = b.leaf[j];
        if (n == MAP_K) return false;
        o.leaf[n++] = x < y ? x : y;
        i += x <= y;
        j += y <= x;
    }
    while (i < a.n) {
        if (n == MAP_K) return false;
        o.leaf[n++] = a.leaf[i++];
    }
    while (j < b.n) {
        if (n == MAP_K) return false;
        o.leaf[n++] = b.leaf[j++];
    }
    o.n = n;
    return true;
}

// `required` marks the nodes some consumer outside the network reads; on return it also marks every leaf the chosen
// cover uses.  With map == false every gate keeps its own two inputs as its cut (no collapsing).
void map_network(uint32_t n_ids, const std::vector<MGate> &g, std::vector<uint8_t> &required, bool map, std::vector<MNode> &out) {
    Trace mt;
    std::vector<uint32_t> gate_of(n_ids, NONE32);
    for (uint32_t i = 0; i < g.size(); i++) gate_of[g[i].out] = i;
    std::vector<uint32_t> depth(n_ids, 0);
    std::vector<Cut> best(g.size());
    std::vector<Cut> cuts;
    std::vector<uint8_t> ncuts;
    if (map) {
        // priority cuts: every node keeps its CUTS_PER_NODE best (depth, then size) cuts; candidates = pairwise merges of the
        // fan-ins' cut sets (stored cuts + the trivial cut).  The kept set is maintained by insertion, so no candidate
        // array and no sort; duplicates can only tie with a kept cut and are dropped there.
        cuts.resize((size_t)g.size() * CUTS_PER_NODE);
        ncuts.assign(g.size(), 0);
        // a fan-in's cut set = its stored cuts (by pointer: copying five 40-byte records per fan-in and gate was a tenth of the
        // mapper) plus its trivial cut, built in `triv`
        auto cut_set = [&](uint32_t id, const Cut **set, Cut &triv, int &n) {
            n = 0;
            if (id == 0) {  // the constant contributes no leaf
                triv.n = 0;
                triv.depth = 1;
                triv.sig = 0;
                set[n++] = &triv;
                return;
            }
            const uint32_t gi = gate_of[id];
            if (gi != NONE32)
                for (int i = 0; i < ncuts[gi]; i++) set[n++] = &cuts[(size_t)gi * CUTS_PER_NODE + i];
            triv.n = 1;
            triv.leaf[0] = id;
            triv.depth = depth[id] + 1;  // as a fan-in cut: (max depth of its leaves) + 1, like the stored ones
            triv.sig = 1ull << (id & 63);
            set[n++] = &triv;
        };
        const Cut *ca[CUTS_PER_NODE + 1], *cb[CUTS_PER_NODE + 1];
        Cut ta, tb, keep[CUTS_PER_NODE];
        for (uint32_t gi = 0; gi < g.size(); gi++) {
            int na, nb, nk = 0;
            cut_set(g[gi].a >> 1, ca, ta, na);
            cut_set(g[gi].b >> 1, cb, tb, nb);
            for (int i = 0; i < na; i++)
                for (int j = 0; j < nb; j++) {
                    const Cut &A = *ca[i], &B = *cb[j];
                    // the merged cut's depth and a lower bound on its size are known before merging: most candidates lose
                    // against the kept set right here
                    const uint32_t d = std::max(A.depth, B.depth);
                    if (nk == CUTS_PER_NODE) {
                        const Cut &w = keep[CUTS_PER_NODE - 1];
                        if (d > w.depth || (d == w.depth && std::max(A.n, B.n) >= w.n)) continue;
                    }
                    const uint64_t sig = A.sig | B.sig;
                    if (__builtin_popcountll(sig) > MAP_K) continue;
                    Cut o;
                    if (!merge_cuts(A, B, o)) continue;
                    o.depth = d;
                    o.sig = sig;
                    // position among the kept cuts (ties keep the earlier candidate first)
                    int pos = nk;
                    while (pos > 0 && (o.depth < keep[pos - 1].depth || (o.depth == keep[pos - 1].depth && o.n < keep[pos - 1].n))) pos--;
                    if (pos == CUTS_PER_NODE) continue;
                    bool dup = false;
                    for (int k = 0; k < nk && !dup; k++)
                        dup = keep[k].sig == sig && keep[k].n == o.n && std::memcmp(keep[k].leaf, o.leaf, o.n * 4) == 0;
                    if (dup) continue;
                    if (nk < CUTS_PER_NODE) nk++;
                    for (int k = nk - 1; k > pos; k--) keep[k] = keep[k - 1];
                    keep[pos] = o;
                }
            for (int i = 0; i < nk; i++) cuts[(size_t)gi * CUTS_PER_NODE + i] = keep[i];
            ncuts[gi] = (uint8_t)nk;
            best[gi] = keep[0];
            depth[g[gi].out] = keep[0].depth;
        }
    } else {
        for (uint32_t gi = 0; gi < g.size(); gi++) {
            Cut c;
            c.n = 0;
            c.sig = 0;
            uint32_t a = g[gi].a >> 1, b = g[gi].b >> 1;
            if (a > b) std::swap(a, b);
            if (a) c.leaf[c.n++] = a;
            if (b && b != a) c.leaf[c.n++] = b;
            uint32_t d = 0;
            for (uint32_t k = 0; k < c.n; k++) d = std::max(d, depth[c.leaf[k]]);
            c.depth = d + 1;
            best[gi] = c;
            depth[g[gi].out] = c.depth;
        }
    }
    mt.mark("    map: cuts");
    // cover: walk backwards from the required nodes.  Every required node is read only after its plane has finished, so the one
    // deadline is the depth D of the deepest of them: a node is free to use any kept cut that meets the budget its consumers leave
    // it, and takes the one that pulls the fewest nodes into the cover that are not in it yet (SHA-256: 74 939 -> 70 062 XOR nodes
    // and 8751 -> 7903 live cells in the mask VM, 82 018 -> 80 007 LUTs; same depths).
    if (map) {
        uint32_t D = 0;
        for (size_t gi = 0; gi < g.size(); gi++)
            if (required[g[gi].out]) D = std::max(D, best[gi].depth);
        std::vector<uint32_t> req(n_ids, D);
        for (size_t gi = g.size(); gi-- > 0;) {
            const uint32_t o = g[gi].out;
            if (!required[o]) continue;
            const uint32_t budget = req[o];
            int pick = -1;
            uint32_t pick_new = ~0u;
            for (int k = 0; k < ncuts[gi]; k++) {
                const Cut &c = cuts[gi * CUTS_PER_NODE + k];
                if (c.depth > budget) continue;
                uint32_t nw = 0;
                for (uint32_t q = 0; q < c.n; q++) nw += !required[c.leaf[q]] && gate_of[c.leaf[q]] != NONE32;
                if (nw < pick_new) pick_new = nw, pick = k;
            }
            if (pick >= 0) best[gi] = cuts[gi * CUTS_PER_NODE + pick];
            for (uint32_t k = 0; k < best[gi].n; k++) {
                const uint32_t lf = best[gi].leaf[k];
                required[lf] = 1;
                req[lf] = std::min(req[lf], budget - 1);
            }
        }
    } else {
        for (size_t gi = g.size(); gi-- > 0;) {
            if (!required[g[gi].out]) continue;
            for (uint32_t k = 0; k < best[gi].n; k++) required[best[gi].leaf[k]] = 1;
        }
    }
    // levels over the chosen cover and truth tables (cone simulated on the 64 input patterns at once)
    static const uint64_t PAT[6] = {0xAAAAAAAAAAAAAAAAull, 0xCCCCCCCCCCCCCCCCull, 0xF0F0F0F0F0F0F0F0ull,
                                    0xFF00FF00FF00FF00ull, 0xFFFF0000FFFF0000ull, 0xFFFFFFFF00000000ull};
    std::vector<uint32_t> level(n_ids, 0), stamp(n_ids, 0);
    std::vector<uint64_t> tmp(n_ids, 0);
    std::vector<uint32_t> stack;
    uint32_t epoch = 0;
    out.clear();
    for (uint32_t gi = 0; gi < g.size(); gi++) {
        const uint32_t o = g[gi].out;
        if (!required[o]) continue;
        const Cut &c = best[gi];
        epoch++;
        uint32_t lv = 0;
        for (uint32_t k = 0; k < c.n; k++) {
            stamp[c.leaf[k]] = epoch;
            tmp[c.leaf[k]] = PAT[k];
            lv = std::max(lv, level[c.leaf[k]]);
        }
        stamp[0] = epoch;
        tmp[0] = 0;
        stack.clear();
        stack.push_back(o);
        while (!stack.empty()) {  // post-order evaluation of the cone
            const uint32_t v = stack.back();
            if (stamp[v] == epoch) {
                stack.pop_back();
                continue;
            }
            const MGate &gt = g[gate_of[v]];
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
        MNode m;
        m.out = o;
        m.n = c.n;
        for (int k = 0; k < MAP_K; k++) m.leaf[k] = k < (int)c.n ? c.leaf[k] : 0;
        m.level = lv + 1;
        m.tt = tmp[o];
        level[o] = lv + 1;
        out.push_back(m);
    }
    mt.mark("    map: cover + truth tables");
}

// Drops every gate no required node depends on (walks backwards marking the transitive fan-in in `needed`, then compacts).
// The reference's bench circuit, for one, multiplies the same two inputs 10^8 times and reads none of the products.
void prune_network(std::vector<MGate> &g, const std::vector<uint8_t> &required, std::vector<uint8_t> &needed) {
    needed = required;
    size_t n_req = 0;
    for (size_t gi = g.size(); gi-- > 0;) {
        if (!needed[g[gi].out]) continue;
        n_req++;
        needed[g[gi].a >> 1] = 1;
        needed[g[gi].b >> 1] = 1;
    }
    if (n_req == g.size()) return;
    size_t w = 0;
    for (size_t gi = 0; gi < g.size(); gi++)
        if (needed[g[gi].out]) g[w++] = g[gi];
    g.resize(w);
    if (w < g.capacity() / 2) g.shrink_to_fit();
}

// Wide networks (average level width >= WIDE_LEVEL) are evaluated with one grid-wide launch per level, so collapsing cones into
// LUTs buys nothing there and the cut enumeration would dominate the compile time (microseconds per gate on AND-heavy
// layers).  They keep their 2-input gates (the network is already pruned to the ones that feed a required node), sorted by
// level.  Returns false (and leaves `out` empty) when the network is not wide.
bool build_wide(uint32_t n_ids, const std::vector<MGate> &g, std::vector<VGate> &out, std::vector<uint32_t> &level_off) {
    out.clear();
    level_off.assign(1, 0);
    const size_t n_req = g.size();  // pruned: every gate feeds a required node
    if (n_req == 0) return false;
    std::vector<uint32_t> level(n_ids, 0);
    uint32_t depth = 0;
    for (const MGate &gt : g) {
        const uint32_t l = 1 + std::max(level[gt.a >> 1], level[gt.b >> 1]);
        level[gt.out] = l;
        depth = std::max(depth, l);
    }
    if (n_req / depth < WIDE_LEVEL) return false;
    std::vector<uint32_t> cnt(depth + 1, 0);
    for (const MGate &gt : g) cnt[level[gt.out]]++;
    level_off.assign(depth + 1, 0);  // levels 1..depth -> [level_off[l - 1], level_off[l])
    uint32_t run = 0;
    for (uint32_t l = 1; l <= depth; l++) {
        level_off[l - 1] = run;
        run += cnt[l];
    }
    level_off[depth] = run;
    out.resize(n_req);
    std::vector<uint32_t> pos(level_off.begin(), level_off.end());
    for (const MGate &gt : g) out[pos[level[gt.out] - 1]++] = VGate{gt.out, gt.a, gt.b, gt.op};
    return true;
}

// counting sort of mapped nodes by level; returns offsets (levels 1..depth -> [off[l-1], off[l]))
std::vector<uint32_t> sort_by_level(const std::vector<MNode> &nodes, std::vector<uint32_t> &order) {
    uint32_t depth = 0;
    for (const MNode &m : nodes) depth = std::max(depth, m.level);
    std::vector<uint32_t> cnt(depth + 2, 0), off(depth + 1, 0);
    for (const MNode &m : nodes) cnt[m.level]++;
    uint32_t run = 0;
    std::vector<uint32_t> cursor(depth + 1, 0);
    for (uint32_t l = 1; l <= depth; l++) {
        cursor[l] = run;
        off[l - 1] = run;
        run += cnt[l];
    }
    off[depth] = run;
    order.resize(nodes.size());
    for (uint32_t i = 0; i < nodes.size(); i++) order[i] = cursor[nodes[i].level]++;  // position of node i in level order
    return off;
}

// =====================================================================================================================
//  Mask-plane VM: the mapped XOR network re-expressed over a small pool of shared-memory cells so that the dependent
//  chain of the level-synchronous walk is LDS -> XOR -> STS instead of L2 round trips.  Fresh rows are brought into cells
//  by asynchronous LOADs issued VM_DELTA levels early; only rows that the item plane reads are written back (`row`).
//  Cells are assigned by a linear scan over levels: a cell is free again one level after its value's last use.
// =====================================================================================================================
void build_mask_vm(Program &P, uint32_t load_window) {
    P.vm.clear();
    P.vm_level_off.clear();
    P.vm_cells = 0;
    const uint32_t depth = (uint32_t)P.xlevel_off.size() - 1;
    if (depth == 0) return;
    const uint32_t n_rows = P.n_rows, n_masks = P.n_masks, zero = P.zero_row();
    std::vector<uint32_t> first(n_rows, NONE32), last(n_rows, 0);
    std::vector<uint8_t> exported(n_rows, 0);
    for (uint32_t l = 0; l < depth; l++)
        for (uint32_t gi = P.xlevel_off[l]; gi < P.xlevel_off[l + 1]; gi++)
            for (uint32_t r : P.xgates[gi].in) {
                if (r == zero) continue;
                if (first[r] == NONE32) first[r] = l + 1;
                last[r] = l + 1;
            }
    for (const Item &it : P.items) {
        if (it.kind == ITEM_MUL) exported[it.ra] = exported[it.rb] = 1;
        else if (it.kind != ITEM_INPUT) exported[it.ra] = 1;
    }
    for (uint32_t r : P.export_rows) exported[r] = 1;
    // VM level of an original level L (1-based) is L + VM_DELTA - 1; the LOAD of a fresh row first used at L sits at L - 1.
    struct Tmp {  // provisional instruction: rows until the scan below assigns cells
        uint32_t dst, in[6];
        bool load;
    };
    const uint32_t n_levels = depth + VM_DELTA;
    // A LOAD must sit at least VM_DELTA levels before the row's first use; it may sit earlier.  Every level is padded to whole
    // steps of VM_STEP slots, and the kernel's time is its number of (dependent) steps: a level that spills a few dozen
    // instructions into one more step hands that many of its LOADs to the spare slots of the levels before it
    // (SHA-256: 262 -> ~205 steps for +1.5 % cells; the XORs themselves sit ALAP and have next to no room to move).
    std::vector<uint32_t> load_level(n_masks, NONE32);
    std::vector<uint32_t> cnt(n_levels + 1, 0);
    {
        std::vector<std::vector<uint32_t>> loads_at(n_levels);
        for (uint32_t r = 0; r < n_masks; r++)
            if (first[r] != NONE32) loads_at[first[r] - 1].push_back(r);
        for (uint32_t l = 0; l < depth; l++) cnt[l + VM_DELTA] += P.xlevel_off[l + 1] - P.xlevel_off[l];
        auto room_of = [&](uint32_t v) {
            const uint32_t c = cnt[v] + (uint32_t)loads_at[v].size();
            return std::max<uint32_t>(1, (c + VM_STEP - 1) / VM_STEP) * VM_STEP - c;
        };
        for (uint32_t v = n_levels; v-- > 1;) {
            const uint32_t c = cnt[v] + (uint32_t)loads_at[v].size();
            const uint32_t e = c % VM_STEP;
            if (c <= VM_STEP || e == 0 || e > loads_at[v].size()) continue;
            uint32_t have = 0;
            for (uint32_t d = 1; d <= load_window && d <= v && have < e; d++) have += room_of(v - d);
            if (have < e) continue;  // would only move the spill somewhere else
            uint32_t need = e;
            for (uint32_t d = 1; d <= load_window && d <= v && need; d++) {
                const uint32_t take = std::min(room_of(v - d), need);
                for (uint32_t k = 0; k < take; k++) {
                    loads_at[v - d].push_back(loads_at[v].back());
                    loads_at[v].pop_back();
                }
                need -= take;
            }
        }
        for (uint32_t v = 0; v < n_levels; v++) {
            cnt[v] += (uint32_t)loads_at[v].size();
            for (uint32_t r : loads_at[v]) load_level[r] = v;
        }
    }
    P.vm_level_off.assign(n_levels + 1, 0);
    for (uint32_t l = 0; l < n_levels; l++) P.vm_level_off[l + 1] = P.vm_level_off[l] + cnt[l];
    std::vector<Tmp> tmp(P.vm_level_off[n_levels]);
    std::vector<uint32_t> cursor(P.vm_level_off.begin(), P.vm_level_off.end() - 1);
    for (uint32_t r = 0; r < n_masks; r++)
        if (first[r] != NONE32) {
            Tmp t{};
            t.load = true;
            t.in[0] = r;
            tmp[cursor[load_level[r]]++] = t;
        }
    for (uint32_t l = 0; l < depth; l++)
        for (uint32_t gi = P.xlevel_off[l]; gi < P.xlevel_off[l + 1]; gi++) {
            Tmp t{};
            t.load = false;
            t.dst = P.xgates[gi].dst;
            for (int k = 0; k < 6; k++) t.in[k] = P.xgates[gi].in[k];
            tmp[cursor[l + VM_DELTA]++] = t;
        }
    std::vector<uint32_t> cell_of(n_rows, NONE32);
    cell_of[zero] = 0;
    std::vector<std::vector<uint32_t>> free_at(n_levels + 2);
    std::vector<uint32_t> free_list;
    uint32_t n_cells = 1;  // cell 0 = zero
    auto alloc = [&]() -> uint32_t {
        if (!free_list.empty()) {
            uint32_t c = free_list.back();
            free_list.pop_back();
            return c;
        }
        return n_cells++;
    };
    P.vm.resize(tmp.size());
    for (uint32_t l = 0; l < n_levels; l++) {
        for (uint32_t c : free_at[l]) free_list.push_back(c);
        std::vector<uint32_t>().swap(free_at[l]);
        for (uint32_t k = P.vm_level_off[l]; k < P.vm_level_off[l + 1]; k++) {
            const Tmp &t = tmp[k];
            VmInstr out;
            std::memset(&out, 0, sizeof out);
            if (t.load) {
                const uint32_t r = t.in[0], c = alloc();
                cell_of[r] = c;
                free_at[last[r] + VM_DELTA].push_back(c);  // last use at VM level last + DELTA - 1
                out.flags = (uint16_t)VM_F_LOAD;
                out.dst = (uint16_t)c;
                out.row = r;
            } else {
                const uint32_t r = t.dst;
                for (int q = 0; q < 6; q++) out.in[q] = (uint16_t)cell_of[t.in[q]];
                out.row = exported[r] ? r : VM_ROW_NONE;
                if (first[r] != NONE32) {
                    const uint32_t c = alloc();
                    cell_of[r] = c;
                    free_at[last[r] + VM_DELTA].push_back(c);
                    out.dst = (uint16_t)c;
                } else {
                    out.dst = (uint16_t)VM_CELL_MASK;  // no consumer in the network: resolved to the scratch cell when the steps are emitted
                }
            }
            P.vm[k] = out;
        }
        if (n_cells >= VM_CELL_MASK) {  // 16-bit cell ids: such a network does not fit in shared memory anyway; use the fallbacks
            P.vm.clear();
            P.vm_level_off.clear();
            P.vm_cells = 0;
            return;
        }
    }
    P.vm_cells = n_cells;
}

// Balanced LOADs live a little longer; if the extra cells cost the VM a column (or the VM altogether), the plain placement wins.
void build_mask_vm(Program &P) {
    build_mask_vm(P, VM_LOAD_WINDOW);
    const int cols = P.vm.empty() ? 0 : vm_columns(P.vm_cells);
    if (cols == 2) return;
    std::vector<VmInstr> vm = std::move(P.vm);
    std::vector<uint32_t> off = std::move(P.vm_level_off);
    const uint32_t cells = P.vm_cells;
    build_mask_vm(P, 0);
    if ((P.vm.empty() ? 0 : vm_columns(P.vm_cells)) > cols) return;
    P.vm = std::move(vm), P.vm_level_off = std::move(off), P.vm_cells = cells;
}

// ---- step streams (see rv_compile.h) ---------------------------------------------------------------------------------
void emit_vm_steps(Program &P) {
    P.vm_steps.clear();
    P.n_vm_steps = 0;
    if (P.vm.empty()) return;
    const uint32_t scratch = P.vm_cells;  // one extra cell absorbs the writes of empty slots and of export-only XORs
    VmInstr nop;
    std::memset(&nop, 0, sizeof nop);  // XOR of six zero cells
    nop.dst = (uint16_t)scratch;
    nop.row = VM_ROW_NONE;
    // Every VM level becomes at least one step, even an empty one: the device waits for a LOAD by counting cp.async groups
    // (one per level, committed at the level's last step), so a LOAD must stay VM_DELTA levels ahead of its first use.
    const size_t n_levels = P.vm_level_off.size() - 1;
    for (size_t l = 0; l < n_levels; l++) {
        const uint32_t s = P.vm_level_off[l], e = P.vm_level_off[l + 1];
        const uint32_t steps = std::max<uint32_t>(1, (e - s + VM_STEP - 1) / VM_STEP);
        for (uint32_t k = 0; k < steps; k++) {
            const bool last = k + 1 == steps;
            const bool chunk_end = (P.n_vm_steps + 1) % VM_STEPS_PER_CHUNK == 0;
            for (uint32_t t = 0; t < VM_STEP; t++) {
                const uint32_t gi = s + k * VM_STEP + t;
                VmInstr o = gi < e ? P.vm[gi] : nop;
                if (!(o.flags & VM_F_LOAD) && o.dst == VM_CELL_MASK) o.dst = (uint16_t)scratch;
                if (last || chunk_end) o.flags |= (uint16_t)VM_F_BAR;
                if (last) o.flags |= (uint16_t)VM_F_LEVEL_END;
                P.vm_steps.push_back(o);
            }
            P.n_vm_steps++;
        }
    }
}

void emit_lut_steps(const std::vector<LutInstr> &luts, const std::vector<uint32_t> &level_off, uint32_t scratch, std::vector<LutInstr> &steps,
                    uint32_t &n_steps) {
    steps.clear();
    n_steps = 0;
    if (luts.empty()) return;
    LutInstr nop;
    std::memset(&nop, 0, sizeof nop);
    nop.dst = scratch;  // scratch value slot
    const size_t n_levels = level_off.size() - 1;
    for (size_t l = 0; l < n_levels; l++) {
        const uint32_t s = level_off[l], e = level_off[l + 1];
        if (e == s) continue;
        const uint32_t cnt = (e - s + LUT_STEP - 1) / LUT_STEP;
        for (uint32_t k = 0; k < cnt; k++) {
            const bool last = k + 1 == cnt;
            const bool chunk_end = (n_steps + 1) % LUT_STEPS_PER_CHUNK == 0;
            for (uint32_t t = 0; t < LUT_STEP; t++) {
                const uint32_t gi = s + k * LUT_STEP + t;
                LutInstr o = gi < e ? luts[gi] : nop;
                o.pad = (last || chunk_end) ? LUT_F_BAR : 0;
                steps.push_back(o);
            }
            n_steps++;
        }
    }
}

// The value planes run as dependent steps of LUT_STEP slots, every level padded to whole steps.  Mapped nodes sit at their earliest
// level; one with slack (all its consumers sit more than a level later; outputs are read after the plane) may sit later.  Walking
// the levels from the last one down, a level that spills a few nodes into one more step hands them to spare slots of later levels
// within its nodes' slack (SHA-256: 1185 -> 1036 steps for 1028 levels; the verifier's u-plane 1486 -> 1356 with the area cover).  Levels keep their numbers; only MNode::level changes.
void balance_lut_levels(uint32_t n_ids, std::vector<MNode> &nodes, uint32_t step) {
    uint32_t depth = 0;
    for (const MNode &m : nodes) depth = std::max(depth, m.level);
    if (depth < 2) return;
    constexpr uint32_t WINDOW = 256;  // how far ahead a node may be pushed (bounds the search)
    std::vector<uint32_t> cnt(depth + 2, 0), off(depth + 2, 0), prod(n_ids, NONE32), latest(nodes.size(), depth);
    for (const MNode &m : nodes) cnt[m.level]++;
    for (uint32_t l = 1; l <= depth; l++) off[l + 1] = off[l] + cnt[l];
    std::vector<uint32_t> by_level(nodes.size()), cur(off.begin(), off.end());
    for (uint32_t i = 0; i < nodes.size(); i++) {
        by_level[cur[nodes[i].level]++] = i;
        prod[nodes[i].out] = i;
    }
    auto room = [&](uint32_t l) { return cnt[l] == 0 ? 0u : (cnt[l] + step - 1) / step * step - cnt[l]; };
    std::vector<std::pair<uint32_t, uint32_t>> plan;
    std::vector<uint32_t> taken(depth + 2, 0);
    for (uint32_t l = depth; l >= 1; l--) {
        const uint32_t need = cnt[l] % step;
        if (need && cnt[l] > need) {  // (a level that fits one step stays: emptying it would need every node to have slack)
            plan.clear();
            for (uint32_t k = off[l]; k < off[l + 1] && plan.size() < need; k++) {
                const uint32_t i = by_level[k], hi = std::min(latest[i], l + WINDOW);
                for (uint32_t L = l + 1; L <= hi; L++)
                    if (room(L) > taken[L]) {
                        taken[L]++;
                        plan.emplace_back(i, L);
                        break;
                    }
            }
            for (const auto &mv : plan) taken[mv.second] = 0;
            if (plan.size() == need)
                for (const auto &mv : plan) {
                    nodes[mv.first].level = mv.second;
                    cnt[mv.second]++;
                    cnt[l]--;
                }
        }
        // the nodes that started at this level are final now: their producers must stay below them
        for (uint32_t k = off[l]; k < off[l + 1]; k++) {
            const MNode &m = nodes[by_level[k]];
            for (uint32_t q = 0; q < m.n; q++) {
                const uint32_t pi = m.leaf[q] < n_ids ? prod[m.leaf[q]] : NONE32;
                if (pi != NONE32) latest[pi] = std::min(latest[pi], m.level - 1);
            }
        }
    }
}

// network -> level-sorted LUT list + level offsets
void map_to_luts(uint32_t n_ids, std::vector<MGate> &net, std::vector<uint8_t> &required, std::vector<LutInstr> &luts, std::vector<uint32_t> &level_off) {
    std::vector<MNode> nodes;
    map_network(n_ids, net, required, net.size() <= LUT_MAP_MAX_GATES, nodes);
    balance_lut_levels(n_ids, nodes, LUT_STEP);
    std::vector<uint32_t> order;
    level_off = sort_by_level(nodes, order);
    luts.resize(nodes.size());
    for (uint32_t i = 0; i < nodes.size(); i++) {
        LutInstr li;
        std::memset(&li, 0, sizeof li);
        li.dst = nodes[i].out;
        for (int k = 0; k < 6; k++) li.in[k] = nodes[i].leaf[k];
        li.tt = nodes[i].tt;
        luts[order[i]] = li;
    }
}

// Value algebra shared by the prover's plaintext plane and the verifier's u-plane: constant folding on refs.
struct ValueNet {
    std::vector<MGate> g;
    uint32_t n_ids = 1;  // id 0 = constant 0
    uint32_t fresh() { return n_ids++; }
    uint32_t vxor(uint32_t a, uint32_t b) {
        const uint32_t neg = (a ^ b) & 1;
        if ((a >> 1) == 0) return b ^ (a & 1);
        if ((b >> 1) == 0) return a ^ (b & 1);
        if ((a >> 1) == (b >> 1)) return neg;
        const uint32_t id = fresh();
        g.emplace_back(id, a & ~1u, b & ~1u, 0u);
        return (id << 1) | neg;
    }
    uint32_t vand(uint32_t a, uint32_t b) {
        const uint32_t va = a >> 1, vb = b >> 1;
        if (va == 0 && vb == 0) return (a & b) & 1;
        if (va == 0) return (a & 1) ? b : 0u;
        if (vb == 0) return (b & 1) ? a : 0u;
        if (a == b) return a;
        if (va == vb) return 0u;  // x & ~x
        const uint32_t id = fresh();
        g.emplace_back(id, a, b, 1u);
        return id << 1;
    }
};


// =====================================================================================================================
//  Z64 domain front end (see rv_compile.h).  Cells carry (value id, mask row, coefficient); rows are provisional ids until
//  finish() sorts the linear nodes by level.
// =====================================================================================================================
constexpr uint64_t ZB_WIRE = (512 + 64) * 32;  // one Z64 wire over all 256 repetitions = share 512 B + corr 64 B per instance
constexpr uint64_t ZB_MUL = 73728 + 16, ZB_BIN = 3 * ZB_WIRE + 16, ZB_UNARY = 2 * ZB_WIRE + 16, ZB_INPUT = ZB_WIRE + 2048 + 16,
                   ZB_ASSERT = ZB_WIRE + 16384 + 16, ZB_LEAF = ZB_WIRE + 16;

struct ZCell {
    uint32_t vid;
    uint32_t mid;   // fresh index, LIN_BASE + node, or ZERO_MID
    uint64_t coef;  // wire mask = coef * row
};

struct ZBuilder {
    ZProgram &Z;
    std::vector<ZCell> cells;
    std::vector<uint32_t> vlevel;  // per value id
    std::vector<uint32_t> llevel;  // per linear node
    std::vector<ZLin> lin;         // provisional ids
    std::vector<ZInstr> prog;      // creation order
    std::vector<uint32_t> kappa_ids;
    std::vector<uint32_t> input_ids;
    uint64_t n_masks = 0;
    uint64_t alg_bytes = 0;

    explicit ZBuilder(ZProgram &z, size_t n_cells) : Z(z), cells(n_cells, ZCell{0, ZERO_MID, 0}), vlevel(1, 0) {}

    uint32_t mid_level(uint32_t mid) const { return (mid != ZERO_MID && mid >= LIN_BASE) ? llevel[mid - LIN_BASE] : 0; }
    uint32_t new_val(uint32_t level) {
        vlevel.push_back(level);
        return (uint32_t)(vlevel.size() - 1);
    }
    uint32_t emit(uint32_t op, uint32_t a, uint32_t b, uint32_t c, uint64_t imm) {
        const uint32_t lv = 1 + std::max(vlevel[a], std::max(vlevel[b], vlevel[c]));
        const uint32_t id = new_val(lv);
        prog.push_back(ZInstr{op, id, a, b, c, 0, imm});
        return id;
    }
    static bool is_zero(const ZCell &c) { return c.mid == ZERO_MID || c.coef == 0; }

    // returns RV_OK or an error; `i` is the op index for messages
    int step(const rv_op &op, size_t i, std::string &err) {
        const size_t nc = cells.size();
        auto bad_wire = [&]() {
            err = "op " + std::to_string(i) + ": Z64 wire index out of range for the given wire_counts";
            return (int)RV_E_ARG;
        };
        const uint64_t c = op.imm;  // u64 -> Recon broadcast, src/algebra/z64/recon.rs:123-129
        switch (op.opcode) {
            case RV_INPUT: {  // src/transcript/prover.rs:181-199
                if (op.dst >= nc) return bad_wire();
                const uint32_t vid = new_val(0);
                ZItem it{ITEM_INPUT, (uint32_t)n_masks, 0, 0, vid, 0, (uint32_t)input_ids.size(), (uint32_t)Z.on_bytes, 1, 0};
                Z.input_off.push_back((uint32_t)Z.on_bytes);
                Z.recon_idx.push_back(0);
                Z.items.push_back(it);
                input_ids.push_back(vid);
                cells[op.dst] = ZCell{vid, (uint32_t)n_masks, 1};
                n_masks += 1;
                Z.on_bytes += 8;
                Z.n_inputs++;
                alg_bytes += ZB_INPUT;
                break;
            }
            case RV_RANDOM:
                err = "op " + std::to_string(i) + ": Random is not accelerated yet";
                return RV_E_UNSUPPORTED;
            case RV_ADD:
            case RV_SUB: {  // src/interpreter/single.rs:71-85
                if (op.dst >= nc || op.a >= nc || op.b >= nc) return bad_wire();
                const ZCell A = cells[op.a];
                ZCell B = cells[op.b];
                const bool sub = op.opcode == RV_SUB;
                if (sub) B.coef = 0 - B.coef;
                ZCell R;
                R.vid = emit(sub ? ZV_SUB : ZV_ADD, A.vid, B.vid, 0, 0);
                if (is_zero(A) && is_zero(B)) R.mid = ZERO_MID, R.coef = 0;
                else if (is_zero(A)) R.mid = B.mid, R.coef = B.coef;
                else if (is_zero(B)) R.mid = A.mid, R.coef = A.coef;
                else if (A.mid == B.mid) {
                    R.mid = A.mid;
                    R.coef = A.coef + B.coef;
                    if (R.coef == 0) R.mid = ZERO_MID;
                } else {
                    if (lin.size() >= LIN_BASE - 2) {
                        err = "too many Z64 linear nodes";
                        return RV_E_UNSUPPORTED;
                    }
                    const uint32_t id = LIN_BASE + (uint32_t)lin.size();
                    llevel.push_back(1 + std::max(mid_level(A.mid), mid_level(B.mid)));
                    lin.push_back(ZLin{id, A.mid, B.mid, 0, A.coef, B.coef});
                    R.mid = id;
                    R.coef = 1;
                }
                cells[op.dst] = R;
                alg_bytes += ZB_BIN;
                break;
            }
            case RV_ADDC:
            case RV_SUBC: {  // src/interpreter/single.rs:87-95: only the correction changes
                if (op.dst >= nc || op.a >= nc) return bad_wire();
                ZCell R = cells[op.a];
                R.vid = emit(ZV_ADDC, R.vid, 0, 0, op.opcode == RV_ADDC ? c : 0 - c);
                cells[op.dst] = R;
                alg_bytes += ZB_UNARY;
                break;
            }
            case RV_MULC: {  // src/interpreter/single.rs:97-104: mask and correction are both scaled
                if (op.dst >= nc || op.a >= nc) return bad_wire();
                ZCell R = cells[op.a];
                R.vid = emit(ZV_MULC, R.vid, 0, 0, c);
                R.coef *= c;
                if (R.coef == 0) R.mid = ZERO_MID;
                cells[op.dst] = R;
                alg_bytes += ZB_UNARY;
                break;
            }
            case RV_MUL: {  // src/interpreter/single.rs:25-69
                if (op.dst >= nc || op.a >= nc || op.b >= nc) return bad_wire();
                const ZCell A = cells[op.a], B = cells[op.b];
                ZItem it{ITEM_MUL, A.mid, B.mid, (uint32_t)n_masks, A.vid, B.vid, (uint32_t)Z.n_corr, (uint32_t)Z.on_bytes,
                         is_zero(A) ? 0 : A.coef, is_zero(B) ? 0 : B.coef};
                Z.recon_off.push_back((uint32_t)Z.on_bytes);
                Z.recon_idx.push_back((uint32_t)Z.recon_off.size() - 1);
                Z.mul_pos.push_back((uint32_t)Z.items.size());
                Z.items.push_back(it);
                const uint32_t kid = new_val(0);  // kappa leaf: 0 in the prover
                kappa_ids.push_back(kid);
                ZCell R;
                R.vid = emit(ZV_MUL, A.vid, B.vid, kid, 0);
                R.mid = (uint32_t)n_masks + 1;  // mask_new
                R.coef = 1;
                cells[op.dst] = R;
                n_masks += 2;
                Z.on_bytes += 64;
                Z.pre_bytes += 8;
                Z.n_mul++;
                Z.n_corr++;
                alg_bytes += ZB_MUL;
                break;
            }
            case RV_ASSERT_ZERO: {  // src/interpreter/single.rs:140-147
                if (op.a >= nc) return bad_wire();
                const ZCell A = cells[op.a];
                ZItem it{ITEM_ASSERT, A.mid, 0, 0, A.vid, 0, 0, (uint32_t)Z.on_bytes, is_zero(A) ? 0 : A.coef, 0};
                Z.recon_off.push_back((uint32_t)Z.on_bytes);
                Z.recon_idx.push_back((uint32_t)Z.recon_off.size() - 1);
                Z.items.push_back(it);
                Z.on_bytes += 64;
                Z.n_assert++;
                alg_bytes += ZB_ASSERT;
                break;
            }
            case RV_CONST: {  // src/interpreter/single.rs:151-155
                if (op.dst >= nc) return bad_wire();
                cells[op.dst] = ZCell{emit(ZV_CONST, 0, 0, 0, c), ZERO_MID, 0};
                alg_bytes += ZB_LEAF;
                break;
            }
            default:
                err = "op " + std::to_string(i) + ": unknown opcode";
                return RV_E_ARG;
        }
        if (n_masks >= LIN_BASE - 2 || Z.on_bytes >= 0xFFFFFF00ull || vlevel.size() >= 0x7FFFFFF0ull) {
            err = "Z64 circuit too large for 32-bit table indices / stream offsets";
            return RV_E_UNSUPPORTED;
        }
        return RV_OK;
    }

    // The Z64 half of a B2A conversion (src/interpreter/combine.rs:148-158,208-218): a fresh Z64 mask, the correction
    // r - reconstruct(mask) into the preprocessing stream (r = the 64 fresh GF(2) wires' per-repetition plaintext), and the
    // output wire Wire{mask: -z64_mask, corr: recon - corr}.  g0 = first of the 64 fresh GF(2) rows, grecon0 = index of the
    // first of the 64 GF(2) reconstruct() calls.
    void b2a(uint32_t dst, uint32_t g0, uint32_t grecon0) {
        const uint32_t idx = (uint32_t)Z.n_b2a;
        ZItem it{ITEM_B2A, g0, 0, (uint32_t)n_masks, idx, grecon0, (uint32_t)Z.n_corr, (uint32_t)Z.on_bytes, 1, 0};
        Z.recon_idx.push_back(0);
        Z.mul_pos.push_back((uint32_t)Z.items.size());
        Z.items.push_back(it);
        const uint32_t leaf = new_val(0);  // verifier: u of the output wire; prover: 0
        kappa_ids.push_back(leaf);
        cells[dst] = ZCell{emit(ZV_B2A, 0, 0, leaf, 0), (uint32_t)n_masks, 0 - (uint64_t)1};
        prog.back().a = idx;  // `a` names the conversion, not a value id
        n_masks += 1;
        Z.pre_bytes += 8;
        Z.n_b2a++;
        Z.n_corr++;
    }

    void finish() {
        Z.n_masks = (uint32_t)n_masks;
        Z.n_lin = (uint32_t)lin.size();
        Z.n_rows = Z.n_masks + Z.n_lin + 1;
        Z.n_vals = (uint32_t)vlevel.size();
        Z.leaf_ids = input_ids;
        Z.leaf_ids.insert(Z.leaf_ids.end(), kappa_ids.begin(), kappa_ids.end());
        // value program by level (counting sort; stable, so creation order is kept inside a level)
        {
            uint32_t depth = 0;
            for (const ZInstr &in : prog) depth = std::max(depth, vlevel[in.dst]);
            Z.vlevel_off.assign(depth + 1, 0);
            std::vector<uint32_t> cnt(depth + 2, 0);
            for (const ZInstr &in : prog) cnt[vlevel[in.dst]]++;
            uint32_t run = 0;
            std::vector<uint32_t> cur(depth + 1, 0);
            for (uint32_t l = 1; l <= depth; l++) {
                cur[l] = run;
                Z.vlevel_off[l - 1] = run;
                run += cnt[l];
            }
            Z.vlevel_off[depth] = run;
            Z.vprog.resize(prog.size());
            for (const ZInstr &in : prog) Z.vprog[cur[vlevel[in.dst]]++] = in;
            std::vector<ZInstr>().swap(prog);
        }
        // linear nodes by level; final row numbers
        {
            uint32_t depth = 0;
            for (uint32_t l : llevel) depth = std::max(depth, l);
            Z.llevel_off.assign(depth + 1, 0);
            std::vector<uint32_t> cnt(depth + 2, 0), cur(depth + 1, 0), row_of(lin.size());
            for (uint32_t l : llevel) cnt[l]++;
            uint32_t run = 0;
            for (uint32_t l = 1; l <= depth; l++) {
                cur[l] = run;
                Z.llevel_off[l - 1] = run;
                run += cnt[l];
            }
            Z.llevel_off[depth] = run;
            for (size_t k = 0; k < lin.size(); k++) row_of[k] = Z.n_masks + cur[llevel[k]]++;
            const uint32_t zero = Z.zero_row();
            auto row = [&](uint32_t mid) -> uint32_t {
                if (mid == ZERO_MID) return zero;
                if (mid >= LIN_BASE) return row_of[mid - LIN_BASE];
                return mid;
            };
            Z.lin.resize(lin.size());
            for (size_t k = 0; k < lin.size(); k++) {
                ZLin n = lin[k];
                n.dst = row_of[k];
                n.a = row(n.a);
                n.b = row(n.b);
                Z.lin[n.dst - Z.n_masks] = n;
            }
            for (ZItem &it : Z.items) {
                if (it.kind == ITEM_B2A) continue;  // ra names GF(2) rows there
                it.ra = it.ca == 0 && it.kind != ITEM_INPUT ? zero : row(it.ra);
                if (it.kind == ITEM_MUL) it.rb = it.cb == 0 ? zero : row(it.rb);
            }
        }
    }
};

}  // namespace

int compile(const rv_op *ops, size_t n_ops, size_t z64_cells, size_t gf2_cells, Program &P, std::string &err, uint32_t flags, SegmentIO *io) {
    Trace tr;
    P = Program();
    P.n_ops = n_ops;
    if (n_ops && !ops) {
        err = "ops is NULL";
        return RV_E_ARG;
    }
    std::vector<Cell> cells(gf2_cells, Cell{VREF_ZERO, ZERO_MID, VREF_ZERO});
    ValueNet un;  // u-plane network (built only while the circuit stays small)
    const bool want_verify = n_ops <= VERIFY_MAX_OPS && !(flags & COMPILE_PROVE_ONLY);
    uint64_t n_vids = 1;  // value ids handed out; id 0 is the constant 0.  (Their plain 2-input depth, a statistic, is computed by the
                          // value-plane job from the gate list: two dependent cache misses per gate that the walk can do without.)
    std::vector<uint32_t> llevel;        // per linear node (plain depth)
    std::vector<uint32_t> tlevel;        // per tainted value
    std::vector<MGate> vg;               // value network, topological; ids = value ids
    std::vector<MGate> lg;               // mask network, topological; provisional ids (see mask_id below)
    std::vector<TGate> tg;               // tainted plane, creation order
    uint64_t n_masks = 0;
    ZBuilder zb(P.z, z64_cells);
    Prefault pf;  // after the tables it serves: joined before they go away
    {  // one counting pass sizes the tables the walk appends to (a std::vector that doubles its way to 10^8 entries copies them all twice)
        size_t c_mul = 0, c_lin = 0, c_in = 0, c_as = 0, c_b2a = 0, c_z = 0;
        uint64_t reach = 0, n_bin = 0;  // sum over the two-operand GF(2) ops of |dst - a| + |dst - b|; their number
        for (size_t i = 0; i < n_ops; i++) {
            const rv_op &op = ops[i];
            if (op.domain == RV_GF2) {
                const bool mul = op.opcode == RV_MUL, lin = op.opcode == RV_ADD || op.opcode == RV_SUB;
                c_mul += mul;
                c_lin += lin;
                if (mul || lin) n_bin++, reach += (uint64_t)(op.dst > op.a ? op.dst - op.a : op.a - op.dst) + (op.dst > op.b ? op.dst - op.b : op.b - op.dst);
                c_in += op.opcode == RV_INPUT;
                c_as += op.opcode == RV_ASSERT_ZERO;
            } else if (op.domain == RV_B2A) c_b2a++;
            else c_z += op.domain == RV_Z64;
        }
        c_mul += 63 * c_b2a;
        c_lin += 189 * c_b2a;
        c_as += 64 * c_b2a;
        const size_t n_items = c_mul + c_in + c_as;
        P.items.reserve(n_items);
        P.recon_pos.reserve(c_mul + c_as);
        P.input_pos.reserve(c_in);
        P.input_vid.reserve(c_in);
        vg.reserve(c_mul + c_lin);
        lg.reserve(c_lin);
        llevel.reserve(c_lin);
        if (want_verify) {
            un.g.reserve(2 * c_mul + c_lin);
            P.kappa_uid.reserve(c_mul);
            P.item_ua.reserve(n_items);
            P.item_ub.reserve(n_items);
            P.input_uid.reserve(c_in);
        }
        zb.prog.reserve(c_z);
        zb.vlevel.reserve(1 + 2 * c_z);
        zb.Z.items.reserve(c_z / 2);
        const char *pf_env = std::getenv("RV_PREFAULT");
        const double mean_reach = (double)reach / (double)std::max<uint64_t>(1, 2 * n_bin);
        const bool local_reads = mean_reach * sizeof(Cell) <= PREFAULT_MAX_REACH_BYTES;
        if (n_ops >= PREFAULT_MIN_OPS && (pf_env ? pf_env[0] == '1' : local_reads)) {
            pf.add(P.items), pf.add(P.recon_pos), pf.add(vg), pf.add(lg), pf.add(llevel), pf.add(cells);
            if (want_verify) pf.add(un.g), pf.add(P.kappa_uid), pf.add(P.item_ua), pf.add(P.item_ub);
            pf.add(zb.prog), pf.add(zb.vlevel), pf.add(zb.Z.items);
            pf.start(io ? 2 : 4);  // streaming segments are compiled several at a time already
        }
    }

    auto mid_level = [&](uint32_t mid) -> uint32_t { return (mid != ZERO_MID && mid >= LIN_BASE) ? llevel[mid - LIN_BASE] : 0; };
    auto new_val = [&]() -> uint32_t { return (uint32_t)n_vids++; };
    auto bad_wire = [&](size_t i) {
        err = "op " + std::to_string(i) + ": wire index out of range for the given wire_counts";
        return RV_E_ARG;
    };
    auto tainted = [](uint32_t vref) { return (vref & VREF_TAINT) != 0; };
    auto t_level = [&](uint32_t vref) -> uint32_t { return tainted(vref) ? tlevel[(vref & ~VREF_TAINT) >> 1] : 0; };
    auto new_tval = [&](uint32_t op, uint32_t a, uint32_t b) -> uint32_t {  // returns the (non-negated) ref
        const uint32_t tid = (uint32_t)tlevel.size();
        tlevel.push_back(op == T_LEAF ? 1 : 1 + std::max(t_level(a), t_level(b)));
        tg.push_back(TGate{op, tid, a, b});
        return VREF_TAINT | (tid << 1);
    };
    // ---- value algebra on refs: shared plaintext (constant folding as before) or tainted (per repetition) ----
    auto v_xor = [&](uint32_t a, uint32_t b) -> uint32_t {
        const uint32_t neg = (a ^ b) & 1;
        if (!tainted(a) && (a >> 1) == 0) return b ^ (a & 1);
        if (!tainted(b) && (b >> 1) == 0) return a ^ (b & 1);
        if ((a & ~1u) == (b & ~1u)) return neg;  // x ^ x (^1)
        if (tainted(a) || tainted(b)) return new_tval(T_XOR, a & ~1u, b & ~1u) | neg;
        const uint32_t vid = new_val();
        vg.emplace_back(vid, a & ~1u, b & ~1u, 0u);
        return (vid << 1) | neg;
    };
    auto v_and = [&](uint32_t a, uint32_t b) -> uint32_t {
        if (!tainted(a) && (a >> 1) == 0) return (a & 1) ? b : VREF_ZERO;
        if (!tainted(b) && (b >> 1) == 0) return (b & 1) ? a : VREF_ZERO;
        if (a == b) return a;
        if ((a & ~1u) == (b & ~1u)) return VREF_ZERO;  // x & ~x
        if (tainted(a) || tainted(b)) return new_tval(T_AND, a, b);
        const uint32_t vid = new_val();
        vg.emplace_back(vid, a, b, 1u);
        return vid << 1;
    };
    // ---- cell-level gates (shared by the op loop and by the adder inside B2A) ----
    auto cell_xor = [&](const Cell &A, const Cell &B, Cell &R) -> int {  // src/interpreter/single.rs:71-85
        R.vref = v_xor(A.vref, B.vref);
        if (A.mid == ZERO_MID) R.mid = B.mid;
        else if (B.mid == ZERO_MID) R.mid = A.mid;
        else if (A.mid == B.mid) R.mid = ZERO_MID;
        else {
            if (lg.size() >= LIN_BASE - 1) {
                err = "too many linear nodes";
                return RV_E_UNSUPPORTED;
            }
            const uint32_t id = LIN_BASE + (uint32_t)lg.size();
            llevel.push_back(1 + std::max(mid_level(A.mid), mid_level(B.mid)));
            lg.emplace_back(id, A.mid, B.mid, 0u);  // provisional ids; renumbered below
            R.mid = id;
        }
        R.uref = want_verify ? un.vxor(A.uref, B.uref) : 0;
        return RV_OK;
    };
    auto cell_and = [&](const Cell &A, const Cell &B, Cell &R) {  // src/interpreter/single.rs:25-69
        P.recon_pos.push_back((uint32_t)P.items.size());
        P.items.emplace_back((uint32_t)ITEM_MUL, A.mid, B.mid, (uint32_t)n_masks, A.vref, B.vref, (uint32_t)P.n_and, 0u);
        R.mid = (uint32_t)n_masks + 1;  // mask_new
        R.vref = v_and(A.vref, B.vref);
        R.uref = 0;
        if (want_verify) {  // u_out = u_a & u_b ^ kappa_j  
            const uint32_t kid = un.fresh();
            P.kappa_uid.push_back(kid);
            P.item_ua.push_back(A.uref);
            P.item_ub.push_back(B.uref);
            R.uref = un.vxor(un.vand(A.uref, B.uref), kid << 1);
        }
        n_masks += 2;
        P.n_and++;
        P.algorithmic_bytes += B_AND;
    };
    auto cell_random = [&](Cell &R) {  // Wire{mask: fresh, corr: 0}: value = reconstruct(mask), src/interpreter/single.rs:148-150
        R.mid = (uint32_t)n_masks;
        R.vref = new_tval(T_LEAF, (uint32_t)n_masks, 0);
        R.uref = 0;
        if (want_verify) {  // corr = 0  =>  u = rho(mask): a leaf computed from the opened players' shares
            const uint32_t uid = un.fresh();
            P.rand_row.push_back((uint32_t)n_masks);
            P.rand_uid.push_back(uid);
            R.uref = uid << 1;
        }
        n_masks += 1;
    };
    auto push_recon = [&](const Cell &A, uint32_t kind) {  // reconstruct(mask): one online byte per repetition, recorded for the opening
        P.recon_pos.push_back((uint32_t)P.items.size());
        P.items.emplace_back(kind, A.mid, 0u, 0u, A.vref, 0u, 0u, 0u);
        if (want_verify) {
            P.item_ua.push_back(A.uref);
            P.item_ub.push_back(0);
        }
    };

    const uint32_t n_imports = io ? (uint32_t)io->import_cells.size() : 0;
    if (io) {  // streaming segment: the carried wires are leaves of both planes
        io->import_vid.clear();
        io->export_vref.clear();
        io->export_row.clear();
        for (uint32_t j = 0; j < n_imports; j++) {
            const uint32_t c = io->import_cells[j];
            if (c >= cells.size()) return bad_wire(0);
            const uint32_t vid = new_val();
            io->import_vid.push_back(vid);
            cells[c] = Cell{vid << 1, IMP_BASE + j, 0};
        }
    }

    for (size_t i = 0; i < n_ops; i++) {
        const rv_op &op = ops[i];
        if (io && (op.domain != RV_GF2 || op.opcode == RV_RANDOM)) {
            err = "op " + std::to_string(i) + ": streaming mode serves GF(2) circuits without Random / Z64 / B2A";
            return RV_E_UNSUPPORTED;
        }
        if (op.domain == RV_SIZE_HINT) {  // src/interpreter/combine.rs:122-129
            if (cells.size() < op.b) cells.resize(op.b, Cell{VREF_ZERO, ZERO_MID, VREF_ZERO});
            if (z64_cells < op.a) z64_cells = op.a;
            if (zb.cells.size() < z64_cells) zb.cells.resize(z64_cells, ZCell{0, ZERO_MID, 0});
            continue;
        }
        if (op.domain == RV_Z64) {
            P.uses_z64 = true;
            const int zrc = zb.step(op, i, err);
            if (zrc != RV_OK) return zrc;
            continue;
        }
        if (op.domain == RV_B2A) {  // src/interpreter/combine.rs:132-219: z64[dst] <- the 64 GF(2) wires a .. a+63, LSB first
            P.uses_z64 = true;
            if (op.dst >= zb.cells.size() || (uint64_t)op.a + 64 > cells.size()) return bad_wire(i);
            Cell rnd[64], src[64], res[64];
            const uint32_t g0 = (uint32_t)n_masks;
            for (int k = 0; k < 64; k++) cell_random(rnd[k]);  // 64 fresh GF(2) masks, then the Z64 mask + correction (combine.rs:139-158)
            for (int k = 0; k < 64; k++) {
                src[k] = cells[op.a + k];
                if (tainted(src[k].vref)) {
                    err = "op " + std::to_string(i) + ": B2A of wires that depend on Random / another B2A's fresh bits is not accelerated yet";
                    return RV_E_UNSUPPORTED;
                }
                P.b2a_vrefs.push_back(src[k].vref);
            }
            // add_64 (combine.rs:39-93): ripple adder, 63 ANDs, no carry out
            Cell carry, ac, bc, acbc, t;
            int rc2;
            cell_and(rnd[0], src[0], carry);
            if ((rc2 = cell_xor(rnd[0], src[0], res[0]))) return rc2;
            for (int k = 1; k < 63; k++) {
                if ((rc2 = cell_xor(rnd[k], carry, ac)) || (rc2 = cell_xor(src[k], carry, bc))) return rc2;
                cell_and(ac, bc, acbc);
                if ((rc2 = cell_xor(ac, src[k], res[k])) || (rc2 = cell_xor(acbc, carry, t))) return rc2;
                carry = t;
            }
            if ((rc2 = cell_xor(rnd[63], src[63], t)) || (rc2 = cell_xor(carry, t, res[63]))) return rc2;
            const uint32_t grecon0 = (uint32_t)P.recon_pos.size();
            for (int k = 0; k < 64; k++) {  // recon_gf2_to_z64 over the sum: 64 reconstruct() calls (combine.rs:204-207)
                push_recon(res[k], ITEM_RECON);
                P.b2a_urefs.push_back(res[k].uref);
            }
            zb.b2a(op.dst, g0, grecon0);
            P.algorithmic_bytes += 64 * B_LEAF + 63 * B_AND + 189 * B_XOR + 64 * B_ASSERT + ZB_MUL;
            continue;
        }
        if (op.domain != RV_GF2) {
            err = "op " + std::to_string(i) + ": unknown domain";
            return RV_E_ARG;
        }
        const size_t nc = cells.size();
        if (i + 48 < n_ops) {  // the operands' cell records of a wide circuit are cache misses: ask for them a few ops ahead
            const rv_op &f = ops[i + 48];
            if (f.a < nc) __builtin_prefetch(&cells[f.a]);
            if (f.b < nc) __builtin_prefetch(&cells[f.b]);
        }
        const uint32_t c = (uint32_t)(op.imm & 1);  // bool -> Recon, src/algebra/gf2/recon.rs:274-287
        switch (op.opcode) {
            case RV_INPUT: {  // src/transcript/prover.rs:181-199
                if (op.dst >= nc) return bad_wire(i);
                uint32_t vid = new_val();
                P.input_pos.push_back((uint32_t)P.items.size());
                P.items.emplace_back((uint32_t)ITEM_INPUT, (uint32_t)n_masks, 0u, 0u, vid << 1, 0u, (uint32_t)P.input_vid.size(), 0u);
                P.input_vid.push_back(vid);
                uint32_t uid = 0;
                if (want_verify) {
                    uid = un.fresh();
                    P.input_uid.push_back(uid);
                    P.item_ua.push_back(uid << 1);
                    P.item_ub.push_back(0);
                }
                cells[op.dst] = Cell{vid << 1, (uint32_t)n_masks, uid << 1};
                n_masks += 1;
                P.n_inputs++;
                P.algorithmic_bytes += B_INPUT;
                break;
            }
            case RV_RANDOM: {
                if (op.dst >= nc) return bad_wire(i);
                Cell R;
                cell_random(R);
                cells[op.dst] = R;
                P.algorithmic_bytes += B_LEAF;
                break;
            }
            case RV_ADD:
            case RV_SUB: {  // src/interpreter/single.rs:71-85: mask and correction add component-wise
                if (op.dst >= nc || op.a >= nc || op.b >= nc) return bad_wire(i);
                const Cell A = cells[op.a], B = cells[op.b];  // (copies: dst may be one of them)
                const int rc2 = cell_xor(A, B, cells[op.dst]);  // written in place: a Cell returned through the stack is reloaded with a
                if (rc2) return rc2;                            // wider load than the stores that built it, a failed store-to-load forward
                P.algorithmic_bytes += B_XOR;
                break;
            }
            case RV_ADDC:
            case RV_SUBC: {  // src/interpreter/single.rs:87-95: only the correction changes
                if (op.dst >= nc || op.a >= nc) return bad_wire(i);
                Cell R = cells[op.a];
                R.vref ^= c;
                R.uref ^= c;
                cells[op.dst] = R;
                P.algorithmic_bytes += B_UNARY;
                break;
            }
            case RV_MULC: {  // src/interpreter/single.rs:97-104
                if (op.dst >= nc || op.a >= nc) return bad_wire(i);
                cells[op.dst] = c ? cells[op.a] : Cell{VREF_ZERO, ZERO_MID, VREF_ZERO};
                P.algorithmic_bytes += B_UNARY;
                break;
            }
            case RV_MUL: {  // src/interpreter/single.rs:25-69
                if (op.dst >= nc || op.a >= nc || op.b >= nc) return bad_wire(i);
                const Cell A = cells[op.a], B = cells[op.b];
                cell_and(A, B, cells[op.dst]);
                break;
            }
            case RV_ASSERT_ZERO: {  // src/interpreter/single.rs:140-147
                if (op.a >= nc) return bad_wire(i);
                push_recon(cells[op.a], ITEM_ASSERT);
                P.n_assert++;
                P.algorithmic_bytes += B_ASSERT;
                break;
            }
            case RV_CONST: {  // src/interpreter/single.rs:151-155
                if (op.dst >= nc) return bad_wire(i);
                cells[op.dst] = Cell{c ? VREF_ONE : VREF_ZERO, ZERO_MID, c ? VREF_ONE : VREF_ZERO};
                P.algorithmic_bytes += B_LEAF;
                break;
            }
            default:
                err = "op " + std::to_string(i) + ": unknown opcode";
                return RV_E_ARG;
        }
        if (n_masks >= IMP_BASE - 130 || P.items.size() >= 0xFFFFFF00ull || n_vids >= 0x3FFFFFF0ull || tlevel.size() >= 0x3FFFFFF0ull) {
            err = "circuit too large for 32-bit table indices";
            return RV_E_UNSUPPORTED;
        }
    }
    pf.finish();
    tr.mark("op walk");
    // tainted plane by level
    {
        uint32_t depth = 0;
        for (uint32_t l : tlevel) depth = std::max(depth, l);
        P.n_tvals = (uint32_t)tlevel.size();
        P.tlevel_off.assign(depth + 1, 0);
        std::vector<uint32_t> cnt(depth + 2, 0), cur(depth + 1, 0);
        for (uint32_t l : tlevel) cnt[l]++;
        uint32_t run = 0;
        for (uint32_t l = 1; l <= depth; l++) {
            cur[l] = run;
            P.tlevel_off[l - 1] = run;
            run += cnt[l];
        }
        P.tlevel_off[depth] = run;
        P.tgates.resize(tg.size());
        for (const TGate &g : tg) P.tgates[cur[tlevel[g.dst]]++] = g;
        std::vector<TGate>().swap(tg);
    }
    std::vector<uint32_t> export_mid;
    if (io)
        for (uint32_t c : io->export_cells) {
            if (c >= cells.size()) return bad_wire(n_ops);
            io->export_vref.push_back(cells[c].vref);
            export_mid.push_back(cells[c].mid);
        }
    std::vector<Cell>().swap(cells);
    zb.finish();
    P.algorithmic_bytes += zb.alg_bytes;

    P.n_prg = (uint32_t)n_masks;
    P.n_masks = (uint32_t)n_masks + n_imports;  // imported rows sit right behind the PRG rows and behave like fresh rows from here on
    P.n_vals = (uint32_t)n_vids;
    P.n_online = (uint32_t)P.items.size();
    P.n_pre = (uint32_t)P.n_and;
    for (uint32_t l : llevel) P.plain_linear_depth = std::max(P.plain_linear_depth, l);
    const bool small = n_ops <= (io ? (1u << 20) : (4u << 20));  // debug tables only where tests can use them (not for a streaming segment of the default window: 64 MB each)

    // The three planes are independent from here on (they only share read-only parts of P.items): for circuits big enough
    // to care they are built side by side.
    int mask_rc = RV_OK;
    std::string mask_err;
    // ---- mask plane: map the XOR network, number the rows -----------------------------------------------------------
    auto job_mask = [&]() {
        Trace jt;
        // mapper ids: 0 = zero mask, 1 + i = fresh PRG mask i, 1 + n_masks + k = linear node k
        const uint32_t n_lin_all = (uint32_t)lg.size();
        if ((uint64_t)P.n_masks + n_lin_all + 2 >= LIN_BASE) {
            mask_err = "circuit too large for 32-bit row indices";
            mask_rc = RV_E_UNSUPPORTED;
            return;
        }
        auto mask_id = [&](uint32_t mid) -> uint32_t {
            if (mid == ZERO_MID) return 0;
            if (mid >= LIN_BASE) return 1 + P.n_masks + (mid - LIN_BASE);
            if (mid >= IMP_BASE) return 1 + P.n_prg + (mid - IMP_BASE);
            return 1 + mid;
        };
        if (n_lin_all == 0) {  // no Add/Sub of masked wires (e.g. the reference's bench circuit): rows are the fresh masks
            P.xlevel_off.assign(1, 0);
            P.n_lin = 0;
            P.n_rows = P.n_masks + 1;
            const uint32_t zero_row = P.zero_row();
            auto row_of_mid = [&](uint32_t mid) -> uint32_t { return mid == ZERO_MID ? zero_row : mask_id(mid) - 1; };
            for (Item &it : P.items) {
                it.ra = row_of_mid(it.ra);
                if (it.kind == ITEM_MUL) it.rb = row_of_mid(it.rb);
            }
            if (io)
                for (uint32_t mid : export_mid) io->export_row.push_back(row_of_mid(mid));
            return;
        }
        for (MGate &g : lg) {
            g.out = mask_id(g.out);
            g.a = mask_id(g.a) << 1;
            g.b = mask_id(g.b) << 1;
        }
        const uint32_t n_ids = 1 + P.n_masks + n_lin_all;
        std::vector<uint8_t> required(n_ids, 0);
        for (const Item &it : P.items) {
            required[mask_id(it.ra)] = 1;
            if (it.kind == ITEM_MUL) required[mask_id(it.rb)] = 1;
        }
        for (uint32_t mid : export_mid) required[mask_id(mid)] = 1;  // later segments read these wires' masks
        {
            std::vector<uint8_t> needed;
            prune_network(lg, required, needed);
        }
        std::vector<MNode> nodes;
        map_network(n_ids, lg, required, lg.size() <= LUT_MAP_MAX_GATES, nodes);
        std::vector<MGate>().swap(lg);
        // ALAP: a node is computed just before its earliest consumer instead of as early as possible.  The depth is unchanged
        // but values wait far less in the VM's shared-memory cells (SHA-256: 21.8k -> ~10k live cells).
        {
            std::vector<uint32_t> cons_min(n_ids, NONE32);
            for (size_t i = nodes.size(); i-- > 0;) {
                MNode &m = nodes[i];
                if (cons_min[m.out] != NONE32) m.level = std::max(m.level, cons_min[m.out] - 1);
                for (uint32_t k = 0; k < m.n; k++) cons_min[m.leaf[k]] = std::min(cons_min[m.leaf[k]], m.level);
            }
        }
        std::vector<uint32_t> order;
        P.xlevel_off = sort_by_level(nodes, order);
        P.n_lin = (uint32_t)nodes.size();
        P.n_rows = P.n_masks + P.n_lin + 1;
        const uint32_t zero_row = P.zero_row();
        std::vector<uint32_t> row_of_lin(n_lin_all, NONE32);  // linear node k -> row (only for materialised nodes)
        for (uint32_t i = 0; i < nodes.size(); i++) row_of_lin[nodes[i].out - 1 - P.n_masks] = P.n_masks + order[i];
        auto row_of_id = [&](uint32_t id) -> uint32_t {
            if (id == 0) return zero_row;
            if (id <= P.n_masks) return id - 1;
            return row_of_lin[id - 1 - P.n_masks];
        };
        P.xgates.resize(nodes.size());
        for (uint32_t i = 0; i < nodes.size(); i++) {
            XGate x;
            x.dst = P.n_masks + order[i];
            x.pad = 0;
            int q = 0;
            for (uint32_t k = 0; k < nodes[i].n; k++)
                if ((nodes[i].tt >> (1u << k)) & 1) x.in[q++] = row_of_id(nodes[i].leaf[k]);  // leaves that cancel (x ^ x) drop out
            for (; q < 6; q++) x.in[q] = zero_row;
            P.xgates[order[i]] = x;
        }
        for (Item &it : P.items) {
            it.ra = row_of_id(mask_id(it.ra));
            if (it.kind == ITEM_MUL) it.rb = row_of_id(mask_id(it.rb));
        }
        if (io)
            for (uint32_t mid : export_mid) {
                const uint32_t row = row_of_id(mask_id(mid));
                io->export_row.push_back(row);
                if (row >= P.n_masks && row != zero_row) P.export_rows.push_back(row);
            }
        jt.mark("  mask: map + rows");
        build_mask_vm(P);
        emit_vm_steps(P);
        jt.mark("  mask: VM");
        if (!small) {
            std::vector<VmInstr>().swap(P.vm);
            std::vector<uint32_t>().swap(P.vm_level_off);
        }
    };

    // ---- value plane: 6-input LUT step stream, or (wide circuits) the level-sorted 2-input gates --------------------------
    auto job_value = [&]() {
        Trace jt;
        std::vector<uint8_t> required(P.n_vals, 0);
        auto need = [&](uint32_t vref) {
            if (!(vref & VREF_TAINT)) required[vref >> 1] = 1;
        };
        for (const Item &it : P.items) {
            need(it.va);
            if (it.kind == ITEM_MUL) need(it.vb);
        }
        for (const TGate &g : P.tgates)
            if (g.op != T_LEAF) {
                need(g.a);
                need(g.b);
            }
        for (uint32_t r : P.b2a_vrefs) need(r);
        if (io)
            for (uint32_t r : io->export_vref) need(r);
        {  // the statistic the walk no longer keeps: depth of the unpruned 2-input network
            std::vector<uint32_t> lv(P.n_vals, 0);
            uint32_t depth = 0;
            for (const MGate &g : vg) depth = std::max(depth, lv[g.out] = 1 + std::max(lv[g.a >> 1], lv[g.b >> 1]));
            P.plain_value_depth = depth;
        }
        if (small) {
            P.vgates.resize(vg.size());
            for (size_t i = 0; i < vg.size(); i++) P.vgates[i] = VGate{vg[i].out, vg[i].a, vg[i].b, vg[i].op};
        }
        std::vector<uint8_t> needed;
        prune_network(vg, required, needed);
        std::vector<uint8_t>().swap(needed);
        P.values_wide = build_wide(P.n_vals, vg, P.wgates, P.wlevel_off);
        if (!P.values_wide && !vg.empty()) {
            map_to_luts(P.n_vals, vg, required, P.luts, P.lut_level_off);
            emit_lut_steps(P.luts, P.lut_level_off, P.n_vals, P.lut_steps, P.n_lut_steps);
            if (!small) std::vector<LutInstr>().swap(P.luts);
        }
        std::vector<MGate>().swap(vg);
        jt.mark("  value plane");
    };

    // ---- online verifier's u-plane ---------------------------------------------------------------------------------
    auto job_u = [&]() {
        if (!want_verify) return;
        Trace jt;
        P.n_uvals = un.n_ids;
        std::vector<uint8_t> required(P.n_uvals, 0);
        auto mark = [&]() {
            for (size_t t = 0; t < P.items.size(); t++) {
                required[P.item_ua[t] >> 1] = 1;
                required[P.item_ub[t] >> 1] = 1;
            }
        };
        mark();
        std::vector<uint8_t> needed;
        prune_network(un.g, required, needed);
        std::vector<uint8_t>().swap(needed);
        P.verify_wide = build_wide(P.n_uvals, un.g, P.vwgates, P.vwlevel_off);
        if (!P.verify_wide && !un.g.empty()) {
            std::vector<LutInstr> vluts;
            std::vector<uint32_t> vlut_level_off;
            map_to_luts(P.n_uvals, un.g, required, vluts, vlut_level_off);
            emit_lut_steps(vluts, vlut_level_off, P.n_uvals, P.vlut_steps, P.n_vlut_steps);
        }
        std::vector<MGate>().swap(un.g);
        P.has_verify = true;
        jt.mark("  u-plane");
    };

    if (n_ops >= 20000) {  // below that the thread start-up is not worth it
        std::exception_ptr ex[2] = {nullptr, nullptr};
        auto guarded = [](auto &job, std::exception_ptr &e) {
            return [&job, &e]() {
                try {
                    job();
                } catch (...) {
                    e = std::current_exception();
                }
            };
        };
        std::thread t1(guarded(job_value, ex[0])), t2(guarded(job_u, ex[1]));
        std::exception_ptr ex0 = nullptr;
        try {
            job_mask();
        } catch (...) {
            ex0 = std::current_exception();
        }
        t1.join();
        t2.join();
        for (std::exception_ptr e : {ex0, ex[0], ex[1]})
            if (e) std::rethrow_exception(e);
    } else {
        job_mask();
        job_value();
        job_u();
    }
    if (mask_rc != RV_OK) {
        err = mask_err;
        return mask_rc;
    }
    tr.mark("planes (side by side)");
    return RV_OK;
}

}  // namespace rv
