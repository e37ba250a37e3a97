// The streaming planner on an op file: plan_stream_serial vs plan_stream (times, plans compared field by field) and, with a
// fifth argument, the plan + compile pipeline of rv_prove_streaming (segments compiled by N threads as the planner publishes them).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "rv_stream_plan.h"
#include "sigprof.h"
using namespace rv;
struct rv_circuit {};
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv) {
    if (argc < 5) return std::fprintf(stderr, "usage: %s ops.bin gf2_cells window_ops planner_threads [compile_threads]\n", argv[0]), 2;
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) return std::perror(argv[1]), 2;
    std::fseek(f, 0, SEEK_END);
    const size_t bytes = (size_t)std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<rv_op> ops(bytes / sizeof(rv_op));
    if (std::fread(ops.data(), 1, bytes, f) != bytes) return 2;
    const size_t gc = std::strtoull(argv[2], 0, 10), window = std::strtoull(argv[3], 0, 10);
    const unsigned nt = (unsigned)std::atoi(argv[4]);
    {
        StreamPlan A, B;
        std::string ea, eb;
        const double t0 = now();
        const int ra = plan_stream_serial(ops.data(), ops.size(), gc, window, A, ea);
        const double t1 = now();
        const int rb = plan_stream(ops.data(), ops.size(), gc, window, B, eb, nullptr, nt);
        const double t2 = now();
        std::printf("serial rc=%d %.3f s (%.1f ns/op); %u threads rc=%d %.3f s (%.1f ns/op)\n", ra, t1 - t0, (t1 - t0) * 1e9 / ops.size(), nt, rb, t2 - t1, (t2 - t1) * 1e9 / ops.size());
        bool ok = ra == rb && ea == eb;
        if (ok && ra == RV_OK) {
            ok = A.n_slots == B.n_slots && A.masks == B.masks && A.tot_on == B.tot_on && A.tot_pre == B.tot_pre && A.tot_inputs == B.tot_inputs && A.tot_recon == B.tot_recon &&
                 A.segs.size() == B.segs.size();
            for (size_t k = 0; ok && k < A.segs.size(); k++) {
                const Segment &x = A.segs[k], &y = B.segs[k];
                ok = x.a == y.a && x.b == y.b && x.n_local == y.n_local && x.ops.size() == y.ops.size() && (x.ops.empty() || !std::memcmp(x.ops.data(), y.ops.data(), x.ops.size() * sizeof(rv_op))) &&
                     x.io.import_cells == y.io.import_cells && x.io.export_cells == y.io.export_cells && x.import_slot == y.import_slot && x.export_slot == y.export_slot &&
                     x.import_global == y.import_global && x.export_global == y.export_global && x.mask0 == y.mask0 && x.on0 == y.on0 && x.pre0 == y.pre0 && x.wit0 == y.wit0 && x.recon0 == y.recon0;
            }
        }
        std::printf(ok ? "plans identical (%zu segments, %u slots)\n" : "PLANS DIFFER\n", A.segs.size(), A.n_slots);
        if (!ok) return 1;
    }
    if (argc > 5) {
        const unsigned nc = (unsigned)std::atoi(argv[5]);
        if (std::getenv("PROF")) prof_start();
        const double t0 = now();
        StreamPlan plan;
        std::string err;
        const size_t n_seg = stream_segments(ops.size(), window);
        plan.segs.resize(n_seg);
        std::atomic<size_t> planned{0}, next{0};
        auto worker = [&]() {
            for (;;) {
                const size_t k = next.fetch_add(1);
                if (k >= n_seg) return;
                while (planned.load(std::memory_order_acquire) <= k) std::this_thread::sleep_for(std::chrono::microseconds(200));
                Segment &S = plan.segs[k];
                Program P;
                S.rc = compile(S.ops.data(), S.ops.size(), 0, S.n_local, P, S.err, COMPILE_PROVE_ONLY, &S.io);
                std::vector<rv_op>().swap(S.ops);
            }
        };
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nc; t++) pool.emplace_back(worker);
        const int rc = plan_stream(ops.data(), ops.size(), gc, window, plan, err, &planned, nt);
        const double t1 = now();
        for (auto &t : pool) t.join();
        const double t2 = now();
        if (std::getenv("PROF")) prof_stop("sigprof.samples");
        std::printf("pipeline rc=%d: %zu segments, planner %.3f s, planner + %u compile threads %.3f s (%.1f ns/op)\n", rc, n_seg, t1 - t0, nc, t2 - t0, (t2 - t0) * 1e9 / ops.size());
    }
    return 0;
}
