// reverie_b200.hpp -- header-only C++17 host mirror of trailofbits/reverie's public API for the KKW hot path, over the C ABI of
// libreverie_b200.so (include/reverie_b200.h).  The reference is Rust; cargo/rustc are absent from this environment, so this is
// the compiled-language face of the drop-in: same type names, constructor shapes, argument order and error behaviour as
//
//   mcircuit::Operation<T> / CombineOperation     matched at src/interpreter/single.rs:106-156, src/interpreter/combine.rs:120-221
//   mcircuit::largest_wires                       consumed at src/proof/mod.rs:125  -> (z64_cells, gf2_cells)
//   reverie::Proof::new(circuit, wit_gf2, wit_z64, wire_counts)    src/proof/mod.rs:119-222
//   Proof::verify(&self, circuit, wire_counts) -> bool             src/proof/mod.rs:224-307
//   bincode::serialize(&proof) / bincode::deserialize              src/main.rs:84,103
//
// Where the reference panics (invalid / short witness, malformed proof lengths) this mirror throws; nothing is computed on
// the host: every call goes to the GPU through the C ABI, and fails loudly (Error, RV_E_CUDA) when there is no device.
#pragma once
#include <algorithm>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "reverie_b200.h"

namespace reverie {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
struct WitnessError : Error {  // "witness is invalid!" / witness too short: src/transcript/prover.rs:190,223
    using Error::Error;
};
struct FormatError : Error {  // malformed proof bytes: src/algebra/gf2/share.rs:158-164
    using Error::Error;
};
inline int check(int rc) {
    if (rc >= 0) return rc;
    const std::string msg = rv_last_error();
    if (rc == RV_E_WITNESS_INVALID || rc == RV_E_WITNESS_SHORT) throw WitnessError(rc, msg);
    if (rc == RV_E_FORMAT) throw FormatError(rc, msg);
    throw Error(rc, msg);
}

// mcircuit::Operation<T>, T = bool (GF2) or uint64_t (Z64)
template <typename T>
struct Operation {
    rv_opcode kind;
    size_t dst = 0, a = 0, b = 0;
    T c = T();
    static Operation Input(size_t dst) { return {RV_INPUT, dst, 0, 0, T()}; }
    static Operation Random(size_t dst) { return {RV_RANDOM, dst, 0, 0, T()}; }
    static Operation Add(size_t dst, size_t a, size_t b) { return {RV_ADD, dst, a, b, T()}; }
    static Operation AddConst(size_t dst, size_t a, T c) { return {RV_ADDC, dst, a, 0, c}; }
    static Operation Sub(size_t dst, size_t a, size_t b) { return {RV_SUB, dst, a, b, T()}; }
    static Operation SubConst(size_t dst, size_t a, T c) { return {RV_SUBC, dst, a, 0, c}; }
    static Operation Mul(size_t dst, size_t a, size_t b) { return {RV_MUL, dst, a, b, T()}; }
    static Operation MulConst(size_t dst, size_t a, T c) { return {RV_MULC, dst, a, 0, c}; }
    static Operation AssertZero(size_t src) { return {RV_ASSERT_ZERO, 0, src, 0, T()}; }
    static Operation Const(size_t dst, T c) { return {RV_CONST, dst, 0, 0, c}; }
};

// mcircuit::CombineOperation
struct CombineOperation {
    rv_op op;
    static CombineOperation GF2(const Operation<bool> &o) { return lower(RV_GF2, o.kind, o.dst, o.a, o.b, o.c ? 1u : 0u); }
    static CombineOperation Z64(const Operation<uint64_t> &o) { return lower(RV_Z64, o.kind, o.dst, o.a, o.b, o.c); }
    static CombineOperation B2A(size_t z64_dst, size_t gf2_low) { return lower(RV_B2A, RV_INPUT, z64_dst, gf2_low, 0, 0); }
    static CombineOperation SizeHint(size_t z64_cells, size_t gf2_cells) { return lower(RV_SIZE_HINT, RV_INPUT, 0, z64_cells, gf2_cells, 0); }

  private:
    static CombineOperation lower(rv_domain d, rv_opcode k, size_t dst, size_t a, size_t b, uint64_t imm) {
        if (dst > 0xFFFFFFFFull || a > 0xFFFFFFFFull || b > 0xFFFFFFFFull) throw Error(RV_E_ARG, "wire index does not fit 32 bits");
        CombineOperation c;
        c.op = rv_op{(uint8_t)d, (uint8_t)k, 0, (uint32_t)dst, (uint32_t)a, (uint32_t)b, imm};
        return c;
    }
};
static_assert(sizeof(CombineOperation) == sizeof(rv_op), "a Vec<CombineOperation> is passed to the C ABI as is");

// mcircuit::largest_wires -> (z64_cells, gf2_cells): 1 + the largest wire index touched per domain (B2A touches z64 dst and 64 gf2 wires)
inline std::pair<size_t, size_t> largest_wires(const std::vector<CombineOperation> &circuit) {
    size_t z = 0, g = 0;
    for (const CombineOperation &c : circuit) {
        const rv_op &o = c.op;
        if (o.domain == RV_SIZE_HINT) {
            z = std::max<size_t>(z, o.a);
            g = std::max<size_t>(g, o.b);
        } else if (o.domain == RV_B2A) {
            z = std::max<size_t>(z, (size_t)o.dst + 1);
            g = std::max<size_t>(g, (size_t)o.a + 64);
        } else {
            size_t &m = o.domain == RV_GF2 ? g : z;
            const bool binary = o.opcode == RV_ADD || o.opcode == RV_SUB || o.opcode == RV_MUL;
            if (o.opcode != RV_ASSERT_ZERO) m = std::max<size_t>(m, (size_t)o.dst + 1);
            if (o.opcode != RV_INPUT && o.opcode != RV_RANDOM && o.opcode != RV_CONST) m = std::max<size_t>(m, (size_t)o.a + 1);
            if (binary) m = std::max<size_t>(m, (size_t)o.b + 1);
        }
    }
    return {z, g};
}

// A compiled circuit: plays the role of the reference's Arc<Vec<CombineOperation>> (compile once, prove / verify many times).
class Circuit {
  public:
    Circuit(const std::vector<CombineOperation> &ops, std::pair<size_t, size_t> wire_counts) : wire_counts_(wire_counts) {
        check(rv_circuit_compile(ops.empty() ? nullptr : &ops[0].op, ops.size(), wire_counts.first, wire_counts.second, &h_));
    }
    ~Circuit() { rv_circuit_free(h_); }
    Circuit(const Circuit &) = delete;
    Circuit &operator=(const Circuit &) = delete;
    const rv_circuit *handle() const { return h_; }
    std::pair<size_t, size_t> wire_counts() const { return wire_counts_; }

  private:
    rv_circuit *h_ = nullptr;
    std::pair<size_t, size_t> wire_counts_;
};

class Proof {
  public:
    // Proof::new (src/proof/mod.rs:119-124).  `seeds` (256 x 16 bytes) replaces the OsRng draw at :131-134; nullptr = OS RNG.
    static Proof new_(const Circuit &circuit, const std::vector<bool> &wit_gf2, const std::vector<uint64_t> &wit_z64, const uint8_t *seeds = nullptr) {
        std::vector<uint8_t> w(wit_gf2.begin(), wit_gf2.end());
        uint8_t *p = nullptr;
        size_t n = 0;
        check(rv_prove(circuit.handle(), w.data(), w.size(), wit_z64.data(), wit_z64.size(), seeds, &p, &n));
        Proof out;
        out.bytes_.assign(p, p + n);
        rv_free(p);
        return out;
    }
    // The reference's own call shape: the op list travels with every call; the library compiles it once and keeps it in its
    // content-addressed cache (rv_proof_new).
    static Proof new_(std::shared_ptr<const std::vector<CombineOperation>> circuit, std::shared_ptr<const std::vector<bool>> wit_gf2,
                      std::shared_ptr<const std::vector<uint64_t>> wit_z64, std::pair<size_t, size_t> wire_counts, const uint8_t *seeds = nullptr) {
        std::vector<uint8_t> w(wit_gf2->begin(), wit_gf2->end());
        uint8_t *p = nullptr;
        size_t n = 0;
        check(rv_proof_new(circuit->empty() ? nullptr : &(*circuit)[0].op, circuit->size(), w.data(), w.size(), wit_z64->data(), wit_z64->size(),
                           wire_counts.first, wire_counts.second, seeds, &p, &n));
        Proof out;
        out.bytes_.assign(p, p + n);
        rv_free(p);
        return out;
    }
    // Proof::new on several GPUs of this box: one process drives them all (rv_group_create_local: the circuit is cloned onto each
    // device, the shards exchange their repetition hashes and assemble the proof over NVLink peer memory).  The group is built per
    // call here for brevity; keep an rv_group next to the circuit when proving repeatedly.
    static Proof new_on(const Circuit &circuit, const std::vector<int> &devices, const std::vector<bool> &wit_gf2, const std::vector<uint64_t> &wit_z64,
                        const uint8_t *seeds = nullptr) {
        std::vector<uint8_t> w(wit_gf2.begin(), wit_gf2.end());
        rv_group *g = nullptr;
        check(rv_group_create_local(circuit.handle(), devices.data(), (int)devices.size(), 1, 1, &g));
        uint8_t *p = nullptr;
        size_t n = 0;
        const int rc = rv_group_prove(g, w.data(), w.size(), wit_z64.data(), wit_z64.size(), seeds, &p, &n);
        rv_group_free(g);
        check(rc);
        Proof out;
        out.bytes_.assign(p, p + n);
        rv_free(p);
        return out;
    }
    // Proof::new for circuits larger than device memory (rv_prove_streaming): proved in segments of window_ops ops, same bytes.
    static Proof new_streaming(const std::vector<CombineOperation> &circuit, const std::vector<bool> &wit_gf2, std::pair<size_t, size_t> wire_counts,
                               size_t window_ops = 0, const uint8_t *seeds = nullptr) {
        std::vector<uint8_t> w(wit_gf2.begin(), wit_gf2.end());
        uint8_t *p = nullptr;
        size_t n = 0;
        check(rv_prove_streaming(circuit.empty() ? nullptr : &circuit[0].op, circuit.size(), wire_counts.first, wire_counts.second, w.data(), w.size(),
                                 nullptr, 0, seeds, window_ops, &p, &n));
        Proof out;
        out.bytes_.assign(p, p + n);
        rv_free(p);
        return out;
    }
    // Proof::verify (src/proof/mod.rs:224).  strict (default): the commitment must match AND every AssertZero of the opened
    // repetitions must hold -- the reference computes that flag (src/transcript/verifier/online.rs:176-178) and never reads it,
    // so a prover that skips its own assert (src/transcript/prover.rs:221-228) would be accepted.  strict = false is the
    // reference's exact verdict (commitment equality only).
    bool verify(const Circuit &circuit, bool strict = true) const {
        int okay = 1;
        const bool accept = check(rv_verify(circuit.handle(), bytes_.data(), bytes_.size(), &okay)) == 1;
        return accept && (!strict || okay != 0);
    }
    bool verify(std::shared_ptr<const std::vector<CombineOperation>> circuit, std::pair<size_t, size_t> wire_counts, bool strict = true) const {
        int okay = 1;
        const bool accept = check(rv_proof_verify_ex(circuit->empty() ? nullptr : &(*circuit)[0].op, circuit->size(), wire_counts.first,
                                                     wire_counts.second, bytes_.data(), bytes_.size(), &okay)) == 1;
        return accept && (!strict || okay != 0);
    }
    // bincode::serialize(&proof) / bincode::deserialize::<Proof>(bytes)
    const std::vector<uint8_t> &serialize() const { return bytes_; }
    static Proof deserialize(std::vector<uint8_t> bytes) {
        Proof p;
        p.bytes_ = std::move(bytes);
        return p;
    }

  private:
    std::vector<uint8_t> bytes_;
};

}  // namespace reverie
