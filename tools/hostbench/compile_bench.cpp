// rv::compile on an op file (tools/hostbench/dump_ops.py): wall time per run, RV_TRACE=1 for the phases, PROF=1 for a SIGPROF profile.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "rv_compile.h"
#include "sigprof.h"

int main(int argc, char **argv) {
    if (argc < 4) return std::fprintf(stderr, "usage: %s ops.bin z64_cells gf2_cells [flags: 1 = prove only] [reps]\n", argv[0]), 2;
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) return std::perror(argv[1]), 2;
    std::fseek(f, 0, SEEK_END);
    const size_t bytes = (size_t)std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<rv_op> ops(bytes / sizeof(rv_op));
    if (std::fread(ops.data(), 1, bytes, f) != bytes) return 2;
    const size_t zc = std::strtoull(argv[2], 0, 10), gc = std::strtoull(argv[3], 0, 10);
    const uint32_t flags = argc > 4 ? (uint32_t)std::atoi(argv[4]) : 0;
    const int reps = argc > 5 ? std::atoi(argv[5]) : 1;
    if (std::getenv("PROF")) prof_start();
    for (int r = 0; r < reps; r++) {
        rv::Program P;
        std::string err;
        const auto t0 = std::chrono::steady_clock::now();
        const int rc = rv::compile(ops.data(), ops.size(), zc, gc, P, err, flags, nullptr);
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::printf("rc=%d %s  %.3f s  %.1f ns/op\n", rc, err.c_str(), s, s * 1e9 / (double)ops.size());
    }
    if (std::getenv("PROF")) prof_stop("sigprof.samples");
    return 0;
}
