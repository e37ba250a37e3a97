// tests/cpp/test_proof_api.cpp -- the reference's own proof tests (src/proof/mod.rs:311-428) re-expressed against the C++ host
// mirror include/reverie_b200.hpp, plus byte-for-byte comparison with the CPU oracle (oracle/c, test infrastructure only).
// Built and run by tests/test_gpu_parity.py::test_cpp_host_mirror on a GPU box; compiled (not run) by the CPU suite.
#include <cstdio>
#include <cstring>
#include <memory>

#include "../../include/reverie_b200.hpp"
#include "../../oracle/c/reverie_oracle.h"

using namespace reverie;
using Circ = std::vector<CombineOperation>;

#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) {                                                     \
            std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);    \
            return 1;                                                      \
        }                                                                  \
    } while (0)

static std::vector<uint8_t> seeds_of(uint32_t salt) {  // any fixed 256 x 16 bytes: both sides get the same ones
    std::vector<uint8_t> s(256 * 16);
    uint32_t x = 0x9E3779B9u ^ salt;
    for (auto &b : s) {
        x = x * 1664525u + 1013904223u;
        b = (uint8_t)(x >> 24);
    }
    return s;
}

static std::vector<uint8_t> oracle_proof(const Circ &c, const std::vector<bool> &wg, const std::vector<uint64_t> &wz, std::pair<size_t, size_t> wc,
                                         const std::vector<uint8_t> &seeds, int *rc_out = nullptr) {
    std::vector<uint8_t> w(wg.begin(), wg.end());
    uint8_t *p = nullptr;
    size_t n = 0;
    static_assert(sizeof(orc_op) == sizeof(rv_op), "same record layout");
    const int rc = orc_prove(reinterpret_cast<const orc_op *>(c.data()), c.size(), w.data(), w.size(), wz.data(), wz.size(), wc.first, wc.second, seeds.data(), 0,
                             &p, &n, nullptr);
    if (rc_out) *rc_out = rc;
    std::vector<uint8_t> out;
    if (rc == 0) {
        out.assign(p, p + n);
        orc_free(p);
    }
    return out;
}

// src/proof/mod.rs:397-427
static int test_prover_gf2_mul() {
    Circ circuit;
    for (int i = 2; i < 66; i++) circuit.push_back(CombineOperation::GF2(Operation<bool>::Input(1)));
    circuit.push_back(CombineOperation::B2A(0, 2));
    circuit.push_back(CombineOperation::GF2(Operation<bool>::Input(0)));
    circuit.push_back(CombineOperation::GF2(Operation<bool>::Input(1)));
    circuit.push_back(CombineOperation::GF2(Operation<bool>::Mul(2, 0, 1)));
    circuit.push_back(CombineOperation::GF2(Operation<bool>::Add(3, 0, 1)));
    circuit.push_back(CombineOperation::GF2(Operation<bool>::Mul(2, 2, 3)));
    auto arc = std::make_shared<const Circ>(circuit);
    auto wit_gf2 = std::make_shared<const std::vector<bool>>(128, true);
    auto wit_z64 = std::make_shared<const std::vector<uint64_t>>(1, 0);
    Proof proof = Proof::new_(arc, wit_gf2, wit_z64, {128, 128});  // OS RNG seeds, like the reference
    std::printf("size = %zu\n", proof.serialize().size());
    CHECK(proof.verify(arc, {128, 128}));
    CHECK(orc_verify(reinterpret_cast<const orc_op *>(circuit.data()), circuit.size(), 128, 128, proof.serialize().data(), proof.serialize().size(), 0,
                     nullptr, nullptr) == 1);
    // reproducible seeds: the bytes must be the oracle's
    const auto seeds = seeds_of(1);
    Circuit compiled(circuit, {128, 128});
    Proof fixed = Proof::new_(compiled, *wit_gf2, *wit_z64, seeds.data());
    CHECK(fixed.serialize() == oracle_proof(circuit, *wit_gf2, *wit_z64, {128, 128}, seeds));
    CHECK(fixed.verify(compiled));
    return 0;
}

// src/proof/mod.rs:318-395 (bench_prover / bench_verifier circuits)
static int test_bench_circuits() {
    for (size_t n_mul : {size_t(1), size_t(100000)}) {
        Circ circuit = {CombineOperation::GF2(Operation<bool>::Input(0)), CombineOperation::GF2(Operation<bool>::Input(1))};
        circuit.insert(circuit.end(), n_mul, CombineOperation::GF2(Operation<bool>::Mul(2, 0, 1)));
        const std::vector<bool> wit_gf2 = {true, true};
        const std::vector<uint64_t> wit_z64 = {0};
        const auto seeds = seeds_of((uint32_t)n_mul);
        Circuit compiled(circuit, {128, 128});
        Proof proof = Proof::new_(compiled, wit_gf2, wit_z64, seeds.data());
        CHECK(proof.serialize() == oracle_proof(circuit, wit_gf2, wit_z64, {128, 128}, seeds));
        CHECK(proof.verify(compiled));
        CHECK(Proof::deserialize(proof.serialize()).verify(compiled));
        // the same proof in streaming mode (segments of 30 000 ops) and through a two-member group (both on device 0 here)
        CHECK(Proof::new_streaming(circuit, wit_gf2, {128, 128}, 30000, seeds.data()).serialize() == proof.serialize());
        CHECK(Proof::new_on(compiled, {0, 0}, wit_gf2, wit_z64, seeds.data()).serialize() == proof.serialize());
    }
    return 0;
}

// Z64 arithmetic with the reference's error behaviour: prover.rs:190 (short witness), :223 (invalid witness), malformed / tampered proofs
static int test_z64_and_errors() {
    const uint64_t x = 0x0123456789ABCDEFull, y = 0xFEDCBA9876543210ull;
    Circ circuit = {CombineOperation::Z64(Operation<uint64_t>::Input(0)), CombineOperation::Z64(Operation<uint64_t>::Input(1)),
                    CombineOperation::Z64(Operation<uint64_t>::Mul(2, 0, 1)), CombineOperation::Z64(Operation<uint64_t>::AddConst(3, 2, 7)),
                    CombineOperation::Z64(Operation<uint64_t>::MulConst(3, 3, 3)), CombineOperation::Z64(Operation<uint64_t>::SubConst(4, 3, (x * y + 7) * 3)),
                    CombineOperation::Z64(Operation<uint64_t>::AssertZero(4))};
    const auto wc = largest_wires(circuit);
    CHECK(wc.first == 5 && wc.second == 0);
    const auto seeds = seeds_of(7);
    Circuit compiled(circuit, wc);
    Proof proof = Proof::new_(compiled, {}, {x, y}, seeds.data());
    CHECK(proof.serialize() == oracle_proof(circuit, {}, {x, y}, wc, seeds));
    CHECK(proof.verify(compiled));
    bool threw = false;
    try {
        Proof::new_(compiled, {}, {x, y + 1}, seeds.data());
    } catch (const WitnessError &e) {
        threw = e.code == RV_E_WITNESS_INVALID;
    }
    CHECK(threw);
    threw = false;
    try {
        Proof::new_(compiled, {}, {x}, seeds.data());
    } catch (const WitnessError &e) {
        threw = e.code == RV_E_WITNESS_SHORT;
    }
    CHECK(threw);
    std::vector<uint8_t> bad = proof.serialize();
    bad[bad.size() / 2] ^= 0x20;
    CHECK(!Proof::deserialize(bad).verify(compiled));
    bad = proof.serialize();
    bad.resize(bad.size() - 9);
    threw = false;
    try {
        Proof::deserialize(bad).verify(compiled);
    } catch (const FormatError &) {
        threw = true;
    }
    CHECK(threw);
    return 0;
}

int main(int argc, char **argv) {
    if (argc > 1 && !std::strcmp(argv[1], "--compile-check")) return 0;
    if (rv_device_count() < 1) {
        std::printf("no CUDA device: reverie-b200 has no CPU fallback\n");
        return 2;
    }
    if (test_prover_gf2_mul() || test_bench_circuits() || test_z64_and_errors()) return 1;
    std::printf("cpp host mirror ok\n");
    return 0;
}
