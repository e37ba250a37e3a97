// rv_api.cu -- the extern "C" boundary (include/reverie_b200.h): circuit handles, sessions, prove / verify.
// Host orchestration only; every byte of proof data is produced by the kernels in rv_kernels.cu.
#include <cuda_runtime.h>
#include <sys/random.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "rv_bincode.h"
#include "rv_kernels.cuh"
#include "rv_planes.cuh"
#include "rv_stream_plan.h"
#include "rv_zplanes.cuh"

using namespace rv;

// ---------------------------------------------------------------------------------------------------------------------
//  errors
// ---------------------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define CU(call)                                                                                                     \
    do {                                                                                                             \
        cudaError_t e_ = (call);                                                                                     \
        if (e_ != cudaSuccess) return fail(RV_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
    } while (0)

extern "C" const char *rv_last_error(void) { return g_err.c_str(); }
extern "C" const char *rv_version(void) { return "reverie-b200 0.1 (sm_100a)"; }
extern "C" int rv_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
static thread_local int g_device = 0;
extern "C" int rv_set_device(int device) {
    if (device < 0 || device >= rv_device_count()) return fail(RV_E_CUDA, "no such CUDA device");
    g_device = device;
    CU(cudaSetDevice(device));
    return RV_OK;
}
// Large proofs (hundreds of MB for Z64 / 10^8-gate circuits) are returned in pinned host buffers the device copies into
// directly; rv_free recognises them and parks up to a few in a pool, because pinning a gigabyte costs more than proving.
static std::mutex g_pin_mu;
static std::vector<std::pair<void *, size_t>> g_pin_live, g_pin_free;
static constexpr size_t PIN_THRESHOLD = 4u << 20, PIN_POOL_MAX = 4;
static void *pinned_get(size_t bytes) {
    {
        std::lock_guard<std::mutex> g(g_pin_mu);
        for (size_t i = 0; i < g_pin_free.size(); i++)
            if (g_pin_free[i].second >= bytes && g_pin_free[i].second <= 2 * bytes) {
                auto e = g_pin_free[i];
                g_pin_free.erase(g_pin_free.begin() + i);
                g_pin_live.push_back(e);
                return e.first;
            }
    }
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    std::lock_guard<std::mutex> g(g_pin_mu);
    g_pin_live.push_back({p, bytes});
    return p;
}
// Small proofs (the common case: 264 KB for SHA-256) are handed to the caller WITHOUT a copy: a session's device-to-host copy lands
// in a pinned "output block" (all its slots back to back, like the device buffer), rv_session_fetch returns pointers into it, and
// the session moves on to a fresh block for its next step (the CUDA graph's memcpy node is retargeted).  A block returns to the
// pool when the session has let go of it and every pointer handed out has been rv_free'd.
namespace {
struct OutBlock {
    uint8_t *base;
    size_t bytes;
    int refs;  // pointers handed out + 1 while a session targets it
};
std::mutex g_ob_mu;
std::vector<OutBlock> g_ob_live;
std::vector<std::pair<uint8_t *, size_t>> g_ob_free;
constexpr size_t OB_POOL_MAX = 48;

uint8_t *outblock_get(size_t bytes) {  // refs = 1 (the session)
    {
        std::lock_guard<std::mutex> g(g_ob_mu);
        for (size_t i = 0; i < g_ob_free.size(); i++)
            if (g_ob_free[i].second == bytes) {
                uint8_t *b = g_ob_free[i].first;
                g_ob_free.erase(g_ob_free.begin() + i);
                g_ob_live.push_back({b, bytes, 1});
                return b;
            }
    }
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    std::lock_guard<std::mutex> g(g_ob_mu);
    g_ob_live.push_back({(uint8_t *)p, bytes, 1});
    return (uint8_t *)p;
}
// +1 / -1 on the block that contains p; returns false if p is not inside an output block
bool outblock_ref(const void *p, int delta) {
    uint8_t *to_free = nullptr;
    {
        std::lock_guard<std::mutex> g(g_ob_mu);
        const uint8_t *q = (const uint8_t *)p;
        size_t i = 0;
        for (; i < g_ob_live.size(); i++)
            if (q >= g_ob_live[i].base && q < g_ob_live[i].base + g_ob_live[i].bytes) break;
        if (i == g_ob_live.size()) return false;
        g_ob_live[i].refs += delta;
        if (g_ob_live[i].refs <= 0) {
            const OutBlock b = g_ob_live[i];
            g_ob_live.erase(g_ob_live.begin() + i);
            if (g_ob_free.size() < OB_POOL_MAX) g_ob_free.push_back({b.base, b.bytes});
            else to_free = b.base;
        }
    }
    if (to_free) cudaFreeHost(to_free);
    return true;
}
}  // namespace

extern "C" void rv_free(void *p) {
    if (!p) return;
    if (outblock_ref(p, -1)) return;
    {
        std::lock_guard<std::mutex> g(g_pin_mu);
        for (size_t i = 0; i < g_pin_live.size(); i++)
            if (g_pin_live[i].first == p) {
                auto e = g_pin_live[i];
                g_pin_live.erase(g_pin_live.begin() + i);
                if (g_pin_free.size() < PIN_POOL_MAX) {
                    g_pin_free.push_back(e);
                    return;
                }
                // pool full: keep the larger buffers (a caller that moves on to bigger proofs must not re-pin a gigabyte per call)
                size_t small = 0;
                for (size_t k = 1; k < g_pin_free.size(); k++)
                    if (g_pin_free[k].second < g_pin_free[small].second) small = k;
                if (g_pin_free[small].second < e.second) std::swap(g_pin_free[small], e);
                cudaFreeHost(e.first);
                return;
            }
    }
    free(p);
}

// ---------------------------------------------------------------------------------------------------------------------
//  circuit
// ---------------------------------------------------------------------------------------------------------------------
static size_t round_up(size_t x, size_t m) { return (x + m - 1) / m * m; }

struct rv_circuit {
    std::shared_ptr<Program> progp;  // the compiled host tables, shared by the per-device clones of a circuit (rv_circuit_clone)
    Program &prog;
    rv_circuit() : progp(std::make_shared<Program>()), prog(*progp) {}
    explicit rv_circuit(std::shared_ptr<Program> p) : progp(std::move(p)), prog(*progp) {}
    DevProgram dev;
    std::vector<void *> allocs;
    uint8_t *arena = nullptr;  // optional: one device allocation the table uploads are carved from (streaming segments: 1 cudaMalloc / cudaFree instead of 25)
    size_t arena_cap = 0, arena_off = 0;
    std::vector<uint32_t> mul_pos;
    std::vector<uint32_t> recon_idx;  // online item -> index among the reconstruct() calls (verifier)
    std::vector<uint32_t> vleaf_ids;  // u-plane value ids of the verifier's leaves: inputs then kappas
    DevZProgram zdev;
    std::vector<uint32_t> z_input_item;  // k -> item index of the k-th Z64 input()
    int device = 0, n_sms = 148;
    uint64_t device_bytes = 0;
    uint64_t compile_ns = 0;  // host compile + table upload
    uint32_t z64_empty_hash[8];  // B3("")
    uint32_t z64_rep_hash[8];    // Transcript::hash of an empty Z64 transcript: H(B3("") || B3(""))
    // idle full-shard sessions kept for rv_prove, so that back-to-back proofs reuse their device buffers
    mutable std::mutex pool_mu;
    mutable std::vector<rv_session *> pool;
    mutable std::vector<rv_session *> multi_pool;  // idle multi-proof sessions (rv_prove_batch), any slot counts
};

// A thread's pinned double buffer for big table uploads.  cudaMemcpy from pageable memory moves ~3 GB/s and does not get faster
// with more threads (one staging path inside the driver); a memcpy into pinned memory by the uploading thread followed by an
// async copy does, thread by thread, until the link is full.  Used by the streaming prover, whose compile threads upload 12 GB of
// segment tables for 3 x 10^8 gates.
struct UploadStage {
    static constexpr size_t CHUNK = 16u << 20;
    uint8_t *buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaStream_t st = nullptr;
    bool ok = false;
    void init() {
        ok = cudaMallocHost(&buf[0], CHUNK) == cudaSuccess && cudaMallocHost(&buf[1], CHUNK) == cudaSuccess &&
             cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming) == cudaSuccess &&
             cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
        if (!ok) cudaGetLastError();  // (the plain path serves)
    }
    cudaError_t copy(void *d, const void *h, size_t bytes) {
        const uint8_t *src = (const uint8_t *)h;
        uint8_t *dst = (uint8_t *)d;
        cudaError_t e = cudaSuccess;
        for (size_t off = 0, k = 0; off < bytes && e == cudaSuccess; off += CHUNK, k ^= 1) {
            const size_t n = std::min(CHUNK, bytes - off);
            if (off >= 2 * CHUNK) e = cudaEventSynchronize(ev[k]);  // the copy that last read this half has left it
            if (e != cudaSuccess) break;
            memcpy(buf[k], src + off, n);
            e = cudaMemcpyAsync(dst + off, buf[k], n, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaEventRecord(ev[k], st);
        }
        const cudaError_t e2 = cudaStreamSynchronize(st);
        return e != cudaSuccess ? e : e2;
    }
    ~UploadStage() {
        for (int k = 0; k < 2; k++) {
            if (buf[k]) cudaFreeHost(buf[k]);
            if (ev[k]) cudaEventDestroy(ev[k]);
        }
        if (st) cudaStreamDestroy(st);
    }
};
static thread_local UploadStage *g_stage = nullptr;  // set by a thread that wants its uploads staged

template <typename T>
static int upload(rv_circuit *c, const std::vector<T> &v, const T **out) {
    *out = nullptr;
    const size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T), padded = round_up(bytes, 256);
    void *d = nullptr;
    if (c->arena && c->arena_off + padded <= c->arena_cap) {
        d = c->arena + c->arena_off;
        c->arena_off += padded;
    } else {
        CU(cudaMalloc(&d, bytes));
        c->allocs.push_back(d);
    }
    c->device_bytes += bytes;
    if (!v.empty()) {
        if (g_stage && g_stage->ok && v.size() * sizeof(T) >= (1u << 20)) CU(g_stage->copy(d, v.data(), v.size() * sizeof(T)));
        else CU(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    }
    *out = reinterpret_cast<const T *>(d);
    return RV_OK;
}

extern "C" void rv_circuit_free(rv_circuit *c) {
    if (!c) return;
    for (rv_session *s : c->pool) rv_session_free(s);
    c->pool.clear();
    for (rv_session *s : c->multi_pool) rv_session_free(s);
    c->multi_pool.clear();
    for (void *p : c->allocs) cudaFree(p);
    delete c;
}

// Derived host tables + the device-resident copy of every table on `device`.  Without a device the handle still carries the host
// tables (stats / export for the CPU test-suite), but it cannot prove or verify: there is no CPU fallback.
static int circuit_to_device(rv_circuit *c, int device) {
    int rc;
    Program &P = c->prog;
    c->mul_pos.reserve(P.n_and);
    for (uint32_t t = 0; t < P.n_online; t++)
        if (P.items[t].kind == ITEM_MUL) c->mul_pos.push_back(t);
    c->recon_idx.assign(P.n_online, 0);
    for (uint32_t k = 0; k < P.recon_pos.size(); k++) c->recon_idx[P.recon_pos[k]] = k;
    if (P.has_verify) {
        c->vleaf_ids = P.input_uid;
        c->vleaf_ids.insert(c->vleaf_ids.end(), P.kappa_uid.begin(), P.kappa_uid.end());
        c->vleaf_ids.insert(c->vleaf_ids.end(), P.rand_uid.begin(), P.rand_uid.end());
    }
    {
        uint32_t e[8];
        b3_chunk_cv(nullptr, 0, 0, true, e);
        memcpy(c->z64_empty_hash, e, 32);
        b3_hash64(e, e, c->z64_rep_hash);
    }
    // Device tables.  Without a device the handle still carries the host tables (stats / export for the CPU test-suite),
    // but it cannot prove or verify: there is no CPU fallback and rv_session_create reports RV_E_CUDA.
    if (rv_device_count() == 0) {
        c->device = -1;
        return RV_OK;
    }
    c->device = device;
    cudaSetDevice(c->device);
    cudaDeviceGetAttribute(&c->n_sms, cudaDevAttrMultiProcessorCount, c->device);
    DevProgram &D = c->dev;
    if ((rc = upload(c, P.xgates, &D.xgates)) || (rc = upload(c, P.xlevel_off, &D.xlevel_off)) || (rc = upload(c, P.items, &D.items)) ||
        (rc = upload(c, c->mul_pos, &D.mul_pos)) || (rc = upload(c, P.recon_pos, &D.recon_pos)) || (rc = upload(c, P.input_pos, &D.input_pos)) ||
        (rc = upload(c, P.input_vid, &D.input_vid)) || (rc = upload(c, P.vm_steps, &D.vm_steps)) || (rc = upload(c, P.lut_steps, &D.lut_steps)) || (rc = upload(c, P.vlut_steps, &D.vlut_steps)) ||
        (rc = upload(c, c->vleaf_ids, &D.vleaf_ids)) || (rc = upload(c, P.item_ua, &D.item_ua)) || (rc = upload(c, P.item_ub, &D.item_ub)) ||
        (rc = upload(c, c->recon_idx, &D.recon_idx)) || (rc = upload(c, P.tgates, &D.tgates)) || (rc = upload(c, P.tlevel_off, &D.tlevel_off)) ||
        (rc = upload(c, P.rand_row, &D.rand_row)) || (rc = upload(c, P.b2a_vrefs, &D.b2a_vrefs)) || (rc = upload(c, P.b2a_urefs, &D.b2a_urefs))) {
        return rc;
    }
    D.n_tlevels = P.tlevel_off.empty() ? 0 : (uint32_t)P.tlevel_off.size() - 1;
    D.n_tvals = P.n_tvals;
    D.n_rand = (uint32_t)P.rand_row.size();
    if (P.z.any()) {
        const ZProgram &Z = P.z;
        DevZProgram &DZ = c->zdev;
        for (uint32_t t = 0; t < Z.items.size(); t++)
            if (Z.items[t].kind == ITEM_INPUT) c->z_input_item.push_back(t);
        if ((rc = upload(c, Z.vprog, &DZ.vprog)) || (rc = upload(c, Z.vlevel_off, &DZ.vlevel_off)) || (rc = upload(c, Z.lin, &DZ.lin)) ||
            (rc = upload(c, Z.items, &DZ.items)) || (rc = upload(c, Z.leaf_ids, &DZ.leaf_ids)) || (rc = upload(c, Z.recon_off, &DZ.recon_off)) ||
            (rc = upload(c, Z.input_off, &DZ.input_off)) || (rc = upload(c, Z.mul_pos, &DZ.mul_pos)) || (rc = upload(c, Z.recon_idx, &DZ.recon_idx)) ||
            (rc = upload(c, c->z_input_item, &DZ.input_item))) {
            rv_circuit_free(c);
            return rc;
        }
        DZ.n_vlevels = Z.vlevel_off.empty() ? 0 : (uint32_t)Z.vlevel_off.size() - 1;
        DZ.n_llevels = Z.llevel_off.empty() ? 0 : (uint32_t)Z.llevel_off.size() - 1;
        DZ.n_items = (uint32_t)Z.items.size();
        DZ.n_corr = (uint32_t)Z.n_corr;
        DZ.n_inputs = (uint32_t)Z.n_inputs;
        DZ.n_recon = (uint32_t)Z.recon_off.size();
        DZ.n_leaves = (uint32_t)Z.leaf_ids.size();
        DZ.n_masks = Z.n_masks;
        DZ.n_rows = Z.n_rows;
        DZ.n_vals = Z.n_vals;
        DZ.on_bytes = (uint32_t)Z.on_bytes;
        DZ.pre_bytes = (uint32_t)Z.pre_bytes;
    }
    D.n_xgates = (uint32_t)P.xgates.size();
    D.n_llevels = (uint32_t)P.xlevel_off.size() - 1;
    D.n_lut_steps = P.n_lut_steps;
    if (P.values_wide) {
        if ((rc = upload(c, P.wgates, &D.wgates))) {
            rv_circuit_free(c);
            return rc;
        }
        D.n_wlevels = (uint32_t)P.wlevel_off.size() - 1;
    }
    D.n_vlut_steps = P.n_vlut_steps;
    if (P.verify_wide) {
        if ((rc = upload(c, P.vwgates, &D.vwgates))) {
            rv_circuit_free(c);
            return rc;
        }
        D.n_vwlevels = (uint32_t)P.vwlevel_off.size() - 1;
    }
    D.n_uvals = P.n_uvals;
    D.n_vm_steps = P.n_vm_steps;
    D.vm_cells = P.vm_cells;
    D.n_lin = P.n_lin;
    D.n_masks = P.n_masks;
    D.n_rows = P.n_rows;
    D.n_vals = P.n_vals;
    D.n_online = P.n_online;
    D.n_pre = P.n_pre;
    D.n_inputs = (uint32_t)P.n_inputs;
    D.n_recon = (uint32_t)P.recon_pos.size();
    cudaSetDevice(g_device);
    return RV_OK;
}


extern "C" int rv_circuit_compile(const rv_op *ops, size_t n_ops, size_t z64_cells, size_t gf2_cells, rv_circuit **out) {
    return rv_circuit_compile_ex(ops, n_ops, z64_cells, gf2_cells, 0, out);
}

extern "C" int rv_circuit_compile_ex(const rv_op *ops, size_t n_ops, size_t z64_cells, size_t gf2_cells, unsigned flags, rv_circuit **out) {
    if (!out) return fail(RV_E_ARG, "out is NULL");
    *out = nullptr;
    if (flags & ~(unsigned)RV_COMPILE_PROVE_ONLY) return fail(RV_E_ARG, "unknown compile flag");
    rv_circuit *c = new (std::nothrow) rv_circuit();
    if (!c) return fail(RV_E_NOMEM, "out of memory");
    std::string err;
    int rc;
    const auto t0 = std::chrono::steady_clock::now();
    try {
        rc = compile(ops, n_ops, z64_cells, gf2_cells, c->prog, err, (flags & RV_COMPILE_PROVE_ONLY) ? COMPILE_PROVE_ONLY : 0);
    } catch (const std::bad_alloc &) {
        delete c;
        return fail(RV_E_NOMEM, "out of host memory while compiling the circuit");
    }
    if (rc != RV_OK) {
        delete c;
        return fail(rc, err);
    }
    const int rc2 = circuit_to_device(c, g_device);
    if (rc2 != RV_OK) {
        rv_circuit_free(c);
        return rc2;
    }
    c->compile_ns = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    *out = c;
    return RV_OK;
}

// The same compiled circuit resident on another device: shares the host tables, uploads its own device tables.
extern "C" int rv_circuit_clone(const rv_circuit *src, int device, rv_circuit **out) {
    if (!src || !out) return fail(RV_E_ARG, "NULL argument");
    *out = nullptr;
    if (device < 0 || device >= rv_device_count()) return fail(RV_E_CUDA, "no such CUDA device");
    rv_circuit *c = new (std::nothrow) rv_circuit(src->progp);
    if (!c) return fail(RV_E_NOMEM, "out of memory");
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = circuit_to_device(c, device);
    if (rc != RV_OK) {
        rv_circuit_free(c);
        return rc;
    }
    c->compile_ns = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    *out = c;
    return RV_OK;
}

extern "C" int rv_circuit_get_stats(const rv_circuit *c, rv_circuit_stats *o) {
    if (!c || !o) return fail(RV_E_ARG, "NULL argument");
    const Program &P = c->prog;
    o->n_ops = P.n_ops;
    o->n_and = P.n_and;
    o->n_inputs = P.n_inputs;
    o->n_assert = P.n_assert;
    o->n_masks = P.n_masks;
    o->n_linear = P.n_lin;
    o->value_depth = P.values_wide ? P.wlevel_off.size() - 1 : (P.lut_level_off.empty() ? 0 : P.lut_level_off.size() - 1);
    o->linear_depth = P.xlevel_off.size() - 1;
    o->plain_value_depth = P.plain_value_depth;
    o->plain_linear_depth = P.plain_linear_depth;
    o->n_luts = P.values_wide ? P.wgates.size() : (P.n_lut_steps ? P.lut_steps.size() : 0);
    o->n_lut_steps = P.n_lut_steps;
    o->n_vm_steps = P.n_vm_steps;
    o->vm_cells = P.vm_cells;
    o->online_bytes = P.n_online;
    o->pre_bytes = P.n_pre;
    o->algorithmic_bytes = P.algorithmic_bytes;
    o->device_bytes = c->device_bytes;
    const ZProgram &Z = P.z;
    o->z64_mul = Z.n_mul;
    o->z64_inputs = Z.n_inputs;
    o->z64_assert = Z.n_assert;
    o->z64_masks = Z.n_masks;
    o->z64_linear = Z.n_lin;
    o->z64_value_depth = Z.vlevel_off.empty() ? 0 : Z.vlevel_off.size() - 1;
    o->z64_linear_depth = Z.llevel_off.empty() ? 0 : Z.llevel_off.size() - 1;
    o->z64_online_bytes = Z.on_bytes;
    o->z64_pre_bytes = Z.pre_bytes;
    o->compile_ns = c->compile_ns;
    o->has_verify = P.has_verify ? 1 : 0;
    o->n_vals = P.n_vals;
    o->n_uvals = P.n_uvals;
    o->n_vlut_steps = P.n_vlut_steps;
    return RV_OK;
}

extern "C" int rv_circuit_export(const rv_circuit *c, int what, void *buf, size_t *len) {
    if (!c || !len) return fail(RV_E_ARG, "NULL argument");
    const Program &P = c->prog;
    const void *src = nullptr;
    size_t n = 0;
    switch (what) {
        case RV_TAB_VGATES: src = P.vgates.data(); n = P.vgates.size() * sizeof(VGate); break;
        case RV_TAB_LUTS: src = P.luts.data(); n = P.luts.size() * sizeof(LutInstr); break;
        case RV_TAB_XGATES: src = P.xgates.data(); n = P.xgates.size() * sizeof(XGate); break;
        case RV_TAB_XLEVELS: src = P.xlevel_off.data(); n = P.xlevel_off.size() * 4; break;
        case RV_TAB_ITEMS: src = P.items.data(); n = P.items.size() * sizeof(Item); break;
        case RV_TAB_RECON_POS: src = P.recon_pos.data(); n = P.recon_pos.size() * 4; break;
        case RV_TAB_INPUT_POS: src = P.input_pos.data(); n = P.input_pos.size() * 4; break;
        case RV_TAB_INPUT_VID: src = P.input_vid.data(); n = P.input_vid.size() * 4; break;
        default: return fail(RV_E_ARG, "unknown table id");
    }
    if (buf) {
        if (*len < n) return fail(RV_E_ARG, "buffer too small");
        if (n) memcpy(buf, src, n);
    }
    *len = n;
    return RV_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
//  session
// ---------------------------------------------------------------------------------------------------------------------

struct KTimer {
    std::string name;
    double ms = 0;
    uint64_t launches = 0, bytes = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
};

struct rv_session {
    const rv_circuit *c = nullptr;
    uint32_t first_instance = 0, npi = 0, nreps = 0, first_rep = 0;  // npi / nreps: columns / streams of the WHOLE session
    // several proofs side by side (rv_session_create_multi): proof b owns columns [b * npi1, (b + 1) * npi1) of the share tensor
    uint32_t n_proofs = 1, npi1 = 0, nreps1 = 0;
    uint32_t share = 1;  // sessions that run side by side with this one (its batch / group): sizes grids that want one wave
    size_t wit_pitch = 0, vals_pitch = 0, proof_pitch = 0, in_pitch = 0;
    cudaStream_t st = nullptr, st_val = nullptr;
    bool own_stream = true;
    cudaEvent_t ev_fork = nullptr, ev_vals = nullptr;
    // CUDA graphs: after one eager run a phase is captured once and replayed as a single launch
    struct GraphSlot {
        cudaGraphExec_t exec = nullptr;
        int calls = 0;
        uint64_t kernels = 0;
        cudaGraph_t graph = nullptr;          // kept for g_prove: its device-to-host copy node is retargeted to a fresh output block
        cudaGraphNode_t d2h_node = nullptr;
    } g_prove /* commit + open of a full shard */, g_commit, g_open /* open from the session's own all-gather buffer */,
      g_open_x /* open of a linked shard: exchange over peer memory */, g_verify /* rv_verify: upload .. repetition hashes */;
    size_t verify_graph_need = 0;  // staging bytes the captured verify graph copies
    // device buffers
    uint8_t *d_wit = nullptr, *d_seeds = nullptr, *d_pkeys = nullptr, *d_vals = nullptr;
    uint64_t *d_tvals = nullptr;  // tainted plane [n_tvals][npi]
    uint64_t *d_rows = nullptr;
    uint64_t *d_fresh_sm = nullptr;  // instance-major copy of the fresh masks for the mask VM
    size_t pitch_fresh = 0;
    uint8_t *d_on = nullptr, *d_pre = nullptr;
    size_t pitch_on = 0, pitch_pre = 0;
    uint32_t *d_cv_on = nullptr, *d_cv_pre = nullptr, n_chunks_on = 1, n_chunks_pre = 1;
    uint32_t *d_cv_scratch = nullptr;  // long streams: ping-pong buffer of the grid-wide tree levels
    uint8_t *d_on_hash = nullptr, *d_rep_hash = nullptr, *d_all_hashes = nullptr, *d_comm = nullptr, *d_omit = nullptr;
    uint16_t *d_rank = nullptr;
    uint32_t *d_zconst = nullptr;  // [0..8) B3(""), [8..16) H(B3("")||B3(""))
    int *d_bad = nullptr;
    uint8_t *d_proof = nullptr;
    size_t proof_len = 0;
    // Z64 domain (allocated only when the circuit has Z64 ops)
    bool has_z = false;
    size_t zrowlen = 0;  // 64 * npi
    uint64_t *d_zrows = nullptr, *d_zleaf = nullptr, *d_zvals = nullptr;
    uint32_t *d_rk_plain = nullptr;  // [45][64 * npi]: plain round keys of every PRG stream for the T-table Z64 generator
    uint8_t *d_zon = nullptr, *d_zpre = nullptr;
    size_t pitch_zon = 0, pitch_zpre = 0;
    uint32_t *d_zcv_on = nullptr, *d_zcv_pre = nullptr, n_chunks_zon = 1, n_chunks_zpre = 1, *d_zrep = nullptr;
    uint8_t *d_zon_hash = nullptr;
    uint32_t len_zrecons = 0, len_zcorrs = 0, len_zinputs = 0;
    uint64_t *d_zleaf_v = nullptr, *d_zuvals = nullptr;  // verifier: [40][pitch]
    size_t zleaf_pitch = 0, zupitch = 0;
    // verifier-only buffers (allocated on first rv_verify)
    uint8_t *d_vin = nullptr, *h_vin = nullptr;  // one staging blob: see VerifyLayout
    size_t vin_bytes = 0;
    uint8_t *d_leaf_vals = nullptr, *d_uvals = nullptr;
    size_t leaf_pitch = 0, upitch = 0;
    uint8_t *h_vout = nullptr;  // rep hashes (8 KB) + not_okay
    uint32_t len_recons = 0, len_corrs = 0, len_inputs = 0;
    // pinned host staging
    uint8_t *h_in = nullptr;   // witness || seeds(256*16)
    uint8_t *h_out = nullptr;  // proof || bad(4) || comm(32); big proofs (>= PIN_THRESHOLD) skip the proof part: rv_session_fetch copies
                               // them from d_proof straight into the pinned buffer it returns
    size_t out_off = 0;        // offset of bad / comm inside h_out
    bool out_handed = false;   // a pointer into the current h_out block has been given to the caller: the next step needs a fresh block
    bool out_direct = false;   // the last step went through rv_session_prove (its graph's copy node follows h_out): fetch may hand out pointers
    bool in_batch = false;     // bound to an rv_batch (whose graph bakes h_out): never retargeted
    size_t tail_off = 0;       // offset of bad / comm behind the proof bytes in d_proof
    size_t h_in_bytes = 0;
    std::vector<void *> allocs;
    bool timing = false;
    std::vector<KTimer> timers;
    uint64_t launches = 0;
    bool committed = false, opened = false, ever_committed = false;
    // batching (rv_batch): host-initiated work of a bound session goes to the leader's stream; the phases fork from / join into it
    rv_session *lead = nullptr;
    cudaEvent_t ev_bjoin = nullptr;
    // exchange block (flags + double-buffered receive buffers of the repetition hashes; d_all_hashes points into it) and, once
    // rv_session_peer_link has run, the blocks of the sessions that hold the other shards of the same proofs on other GPUs
    uint8_t *d_xchg = nullptr;
    size_t xchg_bytes = 0, proof_alloc_bytes = 0;
    XchgArgs x;                      // world == 1: not linked
    uint8_t *dst_proof = nullptr;    // the assembling rank's proof buffer as mapped here (== d_proof on that rank)
    std::vector<void *> ipc_opened;  // cudaIpcOpenMemHandle mappings to close
    bool linked() const { return x.world > 1; }
    bool assembles() const { return x.world == 1 || x.rank == x.dst; }
};
static cudaStream_t host_stream(const rv_session *s) { return s->lead ? s->lead->st : s->st; }

template <typename T>
static int dalloc(rv_session *s, T **p, size_t count) {
    void *d = nullptr;
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    cudaError_t e = cudaMalloc(&d, bytes);
    if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? RV_E_NOMEM : RV_E_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    s->allocs.push_back(d);
    *p = reinterpret_cast<T *>(d);
    return RV_OK;
}

extern "C" void rv_session_free(rv_session *s) {
    if (!s) return;
    cudaSetDevice(s->c->device);
    if (s->st) cudaStreamSynchronize(s->st);
    if (s->st_val) cudaStreamSynchronize(s->st_val);
    for (auto &t : s->timers)
        for (auto &p : t.pending) {
            cudaEventDestroy(p.first);
            cudaEventDestroy(p.second);
        }
    for (void *p : s->ipc_opened) cudaIpcCloseMemHandle(p);
    for (void *p : s->allocs) cudaFree(p);
    if (s->h_in) cudaFreeHost(s->h_in);
    if (s->h_out) outblock_ref(s->h_out, -1);
    if (s->h_vin) cudaFreeHost(s->h_vin);
    if (s->h_vout) cudaFreeHost(s->h_vout);
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->ev_bjoin) cudaEventDestroy(s->ev_bjoin);
    for (rv_session::GraphSlot *g : {&s->g_prove, &s->g_commit, &s->g_open, &s->g_open_x, &s->g_verify}) {
        if (g->exec) cudaGraphExecDestroy(g->exec);
        if (g->graph) cudaGraphDestroy(g->graph);
    }
    if (s->ev_vals) cudaEventDestroy(s->ev_vals);
    if (s->st && s->own_stream) cudaStreamDestroy(s->st);
    if (s->st_val) cudaStreamDestroy(s->st_val);
    delete s;
}

extern "C" int rv_session_create(const rv_circuit *c, int first_instance, int n_instances, rv_session **out) {
    return rv_session_create_multi(c, first_instance, n_instances, 1, out);
}

extern "C" int rv_session_create_multi(const rv_circuit *c, int first_instance, int n_instances, int n_proofs, rv_session **out) {
    if (!c || !out) return fail(RV_E_ARG, "NULL argument");
    *out = nullptr;
    if (first_instance < 0 || n_instances <= 0 || first_instance + n_instances > RV_PACKED_REPS)
        return fail(RV_E_ARG, "shard must be a non-empty range of the 32 packed instances");
    if (n_proofs < 1 || n_proofs > 128) return fail(RV_E_ARG, "a session holds between 1 and 128 proofs");
    // the item-plane tiles and the rank-major layout of gathered hashes are built for power-of-two shards at aligned positions
    // (32 / G instances per rank, G in {1, 2, 4, 8, 16, 32}); anything else is refused rather than silently mis-tiled
    if ((n_instances & (n_instances - 1)) != 0 || first_instance % n_instances != 0)
        return fail(RV_E_ARG, "a shard is 1, 2, 4, 8, 16 or 32 packed instances starting at a multiple of its size");
    if (n_proofs > 1 && n_instances < 4)
        return fail(RV_E_UNSUPPORTED, "multi-proof sessions need shards of at least 4 packed instances (a 1 KiB hash chunk must not straddle ranks)");
    if (n_proofs > 1 && (c->prog.z.any() || c->prog.n_tvals || c->prog.values_wide ||
                         ProofLayout{(uint32_t)(c->prog.recon_pos.size() / 8 + 1), c->prog.n_pre / 8 + 1, (uint32_t)(c->prog.n_inputs / 8 + 1)}.total() >= PIN_THRESHOLD))
        return fail(RV_E_UNSUPPORTED, "multi-proof sessions serve small GF(2) circuits (no Z64 / Random / B2A, proofs below 4 MB): use one session per proof");
    if (c->device < 0) return fail(RV_E_CUDA, "no CUDA device: reverie-b200 has no CPU fallback");
    CU(cudaSetDevice(c->device));
    if (const int ce = configure_kernels(c->device)) return fail(RV_E_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString((cudaError_t)ce));
    rv_session *s = new (std::nothrow) rv_session();
    if (!s) return fail(RV_E_NOMEM, "out of memory");
    s->c = c;
    s->first_instance = (uint32_t)first_instance;
    s->n_proofs = (uint32_t)n_proofs;
    s->npi1 = (uint32_t)n_instances;
    s->nreps1 = 8 * s->npi1;
    s->npi = s->npi1 * s->n_proofs;
    s->nreps = 8 * s->npi;
    s->first_rep = 8 * s->first_instance;
    const Program &P = c->prog;
    int rc = RV_OK;
    auto bail = [&](int code) {
        rv_session_free(s);
        return code;
    };
    if (cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&s->st_val, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&s->ev_vals, cudaEventDisableTiming) != cudaSuccess)
        return bail(fail(RV_E_CUDA, "stream/event creation failed"));
    s->pitch_on = round_up(std::max<size_t>(P.n_online, 1), 2048);  // whole tiles of the item plane (k_items: T <= 2048)
    s->pitch_pre = round_up(std::max<size_t>(P.n_pre, 1), 2048);
    s->n_chunks_on = P.n_online == 0 ? 1 : (P.n_online + 1023) / 1024;
    s->n_chunks_pre = P.n_pre == 0 ? 1 : (P.n_pre + 1023) / 1024;
    s->len_recons = (uint32_t)(P.recon_pos.size() / 8 + 1);  // floor(n/8)+1: the residue group is always flushed (gf2/share.rs:131-138)
    s->len_corrs = P.n_pre / 8 + 1;
    s->len_inputs = (uint32_t)(P.n_inputs / 8 + 1);
    const ZProgram &Z = P.z;
    s->has_z = Z.any();
    if (s->has_z) {
        s->len_zrecons = (uint32_t)(8 * Z.recon_off.size());  // exactly 8 n bytes: src/algebra/z64/share.rs:37-49, recon.rs:46-66
        s->len_zcorrs = (uint32_t)(8 * Z.n_corr);
        s->len_zinputs = (uint32_t)(8 * Z.n_inputs);
    }
    s->proof_len = ProofLayout{s->len_recons, s->len_corrs, s->len_inputs, s->len_zrecons, s->len_zcorrs, s->len_zinputs}.total();
    if (s->has_z) {
        s->zrowlen = (size_t)64 * s->npi;
        s->pitch_zon = round_up(std::max<size_t>(Z.on_bytes, 1), 64);
        s->pitch_zpre = round_up(std::max<size_t>(Z.pre_bytes, 1), 64);
        s->n_chunks_zon = Z.on_bytes == 0 ? 1 : (uint32_t)((Z.on_bytes + 1023) / 1024);
        s->n_chunks_zpre = Z.pre_bytes == 0 ? 1 : (uint32_t)((Z.pre_bytes + 1023) / 1024);
        if ((rc = dalloc(s, &s->d_zrows, (size_t)Z.n_rows * s->zrowlen)) || (rc = dalloc(s, &s->d_zleaf, Z.leaf_ids.size())) ||
            (rc = dalloc(s, &s->d_zvals, (size_t)Z.n_vals + 1)) || (rc = dalloc(s, &s->d_zon, s->pitch_zon * s->nreps)) ||
            (rc = dalloc(s, &s->d_zpre, s->pitch_zpre * s->nreps)) || (rc = dalloc(s, &s->d_zcv_on, (size_t)s->n_chunks_zon * s->nreps * 8)) ||
            (rc = dalloc(s, &s->d_zcv_pre, (size_t)s->n_chunks_zpre * s->nreps * 8)) || (rc = dalloc(s, &s->d_zrep, (size_t)s->nreps * 8)) ||
            (rc = dalloc(s, &s->d_zon_hash, (size_t)s->nreps * 32)))
            return bail(rc);
        if (cudaMemset(s->d_zleaf, 0, std::max<size_t>(Z.leaf_ids.size(), 1) * 8) != cudaSuccess ||
            cudaMemset(s->d_zrows + (size_t)Z.zero_row() * s->zrowlen, 0, s->zrowlen * 8) != cudaSuccess)
            return bail(fail(RV_E_CUDA, "cudaMemset failed"));
    }
    if (linear_uses_vm(c->dev)) {
        s->pitch_fresh = round_up((size_t)P.n_masks + 128, 128);
        if ((rc = dalloc(s, &s->d_fresh_sm, s->pitch_fresh * s->npi))) return bail(rc);
    }
    s->wit_pitch = round_up(std::max<size_t>(P.n_inputs, 1), 16);
    s->vals_pitch = round_up((size_t)P.n_vals + 1, 16);
    s->tail_off = round_up(s->proof_len, 16);
    s->proof_pitch = s->tail_off + 64;
    if ((rc = dalloc(s, &s->d_wit, s->wit_pitch * s->n_proofs)) || (rc = dalloc(s, &s->d_seeds, (size_t)s->nreps * 16)) ||
        (rc = dalloc(s, &s->d_pkeys, (size_t)s->nreps * 128)) || (rc = dalloc(s, &s->d_vals, s->vals_pitch * s->n_proofs)) ||
        (rc = dalloc(s, &s->d_rk_plain, (size_t)45 * 64 * s->npi)) || (rc = dalloc(s, &s->d_tvals, (size_t)P.n_tvals * s->npi)) || 
        (rc = dalloc(s, &s->d_rows, (size_t)P.n_rows * s->npi)) || (rc = dalloc(s, &s->d_on, s->pitch_on * s->nreps)) ||
        (rc = dalloc(s, &s->d_pre, s->pitch_pre * s->nreps)) || (rc = dalloc(s, &s->d_cv_on, (size_t)s->n_chunks_on * s->nreps * 8)) ||
        (rc = dalloc(s, &s->d_cv_pre, (size_t)s->n_chunks_pre * s->nreps * 8)) || (rc = dalloc(s, &s->d_on_hash, (size_t)s->nreps * 32)) ||
        (rc = dalloc(s, &s->d_rep_hash, (size_t)s->nreps * 32)) ||
        (rc = dalloc(s, &s->d_omit, (size_t)RV_TOTAL_REPS * s->n_proofs)) || (rc = dalloc(s, &s->d_rank, (size_t)RV_TOTAL_REPS * s->n_proofs)) ||
        (rc = dalloc(s, &s->d_zconst, 16)))
        return bail(rc);
    if (std::max(s->n_chunks_on, s->n_chunks_pre) > 2048 && (rc = dalloc(s, &s->d_cv_scratch, (size_t)(std::max(s->n_chunks_on, s->n_chunks_pre) + 1) / 2 * s->nreps * 8)))
        return bail(rc);
    // The exchange block and the proof buffer can be mapped into other processes (rv_session_peer_handle): each is its own
    // allocation of whole 2 MiB pages, so that an IPC mapping exposes nothing else.
    s->xchg_bytes = round_up(XchgLayout{s->n_proofs}.total(), (size_t)2 << 20);
    s->proof_alloc_bytes = round_up(s->proof_pitch * s->n_proofs, (size_t)2 << 20);
    if ((rc = dalloc(s, &s->d_xchg, s->xchg_bytes)) || (rc = dalloc(s, &s->d_proof, s->proof_alloc_bytes))) return bail(rc);
    if (cudaMemset(s->d_xchg, 0, s->xchg_bytes) != cudaSuccess) return bail(fail(RV_E_CUDA, "cudaMemset failed"));
    s->d_all_hashes = s->d_xchg + XchgLayout{s->n_proofs}.off_hash(0);
    // the status flag and comm live right behind the proof bytes, so one device-to-host copy returns all three
    s->d_bad = reinterpret_cast<int *>(s->d_proof + s->tail_off);
    s->d_comm = s->d_proof + s->tail_off + 4;
    if (cudaMemset(s->d_rows + (size_t)P.zero_row() * s->npi, 0, (size_t)s->npi * 8) != cudaSuccess) return bail(fail(RV_E_CUDA, "cudaMemset failed"));
    s->in_pitch = s->wit_pitch + RV_TOTAL_REPS * 16 + 8 * (size_t)Z.n_inputs;
    s->h_in_bytes = s->in_pitch * s->n_proofs;
    s->out_off = s->proof_len >= PIN_THRESHOLD ? 0 : s->tail_off;
    if (cudaMallocHost(&s->h_in, s->h_in_bytes) != cudaSuccess || (s->h_out = outblock_get((s->out_off + 64) * s->n_proofs)) == nullptr)
        return bail(fail(RV_E_NOMEM, "pinned host allocation failed"));
    uint32_t zc[16];
    memcpy(zc, c->z64_empty_hash, 32);
    memcpy(zc + 8, c->z64_rep_hash, 32);
    if (cudaMemcpy(s->d_zconst, zc, 64, cudaMemcpyHostToDevice) != cudaSuccess) return bail(fail(RV_E_CUDA, "cudaMemcpy failed"));
    *out = s;
    return RV_OK;
}

extern "C" void *rv_session_stream(rv_session *s) { return s ? (void *)host_stream(s) : nullptr; }
extern "C" uint64_t rv_session_launch_count(const rv_session *s) { return s ? s->launches : 0; }
extern "C" int rv_session_sync(rv_session *s) {
    if (!s) return fail(RV_E_ARG, "NULL session");
    CU(cudaStreamSynchronize(host_stream(s)));
    return RV_OK;
}

// ---- per-kernel timing ---------------------------------------------------------------------------------------------
struct Scope {
    rv_session *s;
    KTimer *t = nullptr;
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t stream;
    const char *label;
    Scope(rv_session *s_, const char *name, uint64_t bytes, uint64_t n_launches = 1, cudaStream_t on = nullptr) : s(s_), stream(on ? on : s_->st), label(name) {
        s->launches += n_launches;
        if (!s->timing) return;
        for (auto &k : s->timers)
            if (k.name == name) t = &k;
        if (!t) {
            s->timers.push_back(KTimer());
            t = &s->timers.back();
            t->name = name;
        }
        t->launches += n_launches;
        t->bytes += bytes;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a, stream);
    }
    ~Scope() {
        static const bool debug = getenv("RV_DEBUG") != nullptr;
        if (debug) {  // name the launch that failed (development aid; the error stays pending for the caller's check)
            const cudaError_t e = cudaPeekAtLastError();
            if (e != cudaSuccess) fprintf(stderr, "[reverie_b200] launch '%s' on device %d: %s\n", label, s->c->device, cudaGetErrorString(e));
        }
        if (!t) return;
        cudaEventRecord(b, stream);
        t->pending.push_back({a, b});
    }
};

extern "C" int rv_session_timing(rv_session *s, int enable) {
    if (!s) return fail(RV_E_ARG, "NULL session");
    s->timing = enable != 0;
    return RV_OK;
}

extern "C" int rv_session_kernel_times(rv_session *s, rv_kernel_time *out, int max_out, int reset) {
    if (!s) return fail(RV_E_ARG, "NULL session");
    CU(cudaStreamSynchronize(s->st));
    CU(cudaStreamSynchronize(s->st_val));
    int n = 0;
    for (auto &t : s->timers) {
        for (auto &p : t.pending) {
            float ms = 0;
            cudaEventElapsedTime(&ms, p.first, p.second);
            t.ms += ms;
            cudaEventDestroy(p.first);
            cudaEventDestroy(p.second);
        }
        t.pending.clear();
        if (out && n < max_out) {
            memset(&out[n], 0, sizeof(rv_kernel_time));
            strncpy(out[n].name, t.name.c_str(), sizeof(out[n].name) - 1);
            out[n].ms = t.ms;
            out[n].launches = t.launches;
            out[n].algorithmic_bytes = t.bytes;
        }
        n++;
    }
    if (reset) s->timers.clear();
    return n;
}

// ---- upload / commit / open / fetch ----------------------------------------------------------------------------------
extern "C" int rv_session_upload(rv_session *s, const uint8_t *wit_gf2, size_t n_gf2, const uint64_t *wit_z64, size_t n_z64,
                                 const uint8_t *seeds) {
    return rv_session_upload_slot(s, 0, wit_gf2, n_gf2, wit_z64, n_z64, seeds);
}

// Host half of an upload: the slot's witness and seeds into the session's pinned staging buffer.  `first` = the caller is about to
// refill the buffer: wait until the copies of the previous proof have left it.
static int stage_slot(rv_session *s, int slot, const uint8_t *wit_gf2, size_t n_gf2, const uint64_t *wit_z64, size_t n_z64, const uint8_t *seeds, bool first) {
    if (!s) return fail(RV_E_ARG, "NULL session");
    if (slot < 0 || (uint32_t)slot >= s->n_proofs) return fail(RV_E_ARG, "no such proof slot");
    const Program &P = s->c->prog;
    if (n_gf2 < P.n_inputs || n_z64 < P.z.n_inputs) return fail(RV_E_WITNESS_SHORT, "witness is too short");  // prover.rs:190
    if ((P.n_inputs && !wit_gf2) || (P.z.n_inputs && !wit_z64)) return fail(RV_E_ARG, "witness pointer is NULL");
    if (first) {
        CU(cudaSetDevice(s->c->device));
        CU(cudaStreamSynchronize(host_stream(s)));  // the staging buffer may still be in flight from a previous proof
    }
    uint8_t *hin = s->h_in + (size_t)slot * s->in_pitch;
    if (P.n_inputs) memcpy(hin, wit_gf2, P.n_inputs);
    uint8_t *hs = hin + s->wit_pitch;
    if (seeds) memcpy(hs, seeds, RV_TOTAL_REPS * 16);
    else {  // OsRng, src/proof/mod.rs:131-134
        size_t got = 0;
        while (got < RV_TOTAL_REPS * 16) {
            ssize_t r = getrandom(hs + got, RV_TOTAL_REPS * 16 - got, 0);
            if (r <= 0) return fail(RV_E_ARG, "getrandom failed");
            got += (size_t)r;
        }
    }
    if (P.z.n_inputs) memcpy(hs + RV_TOTAL_REPS * 16, wit_z64, 8 * (size_t)P.z.n_inputs);
    return RV_OK;
}
// Device half: slots [slot0, slot0 + n) of the staging buffer to the device -- one strided copy for the witnesses, one for this
// shard's seeds (instead of two small copies per slot: a 32-proof batch issues 8 copies, not 64).
static int flush_slots(rv_session *s, int slot0, int n) {
    const Program &P = s->c->prog;
    cudaStream_t hst = host_stream(s);
    const uint8_t *hin = s->h_in + (size_t)slot0 * s->in_pitch;
    if (P.n_inputs) CU(cudaMemcpy2DAsync(s->d_wit + (size_t)slot0 * s->wit_pitch, s->wit_pitch, hin, s->in_pitch, P.n_inputs, n, cudaMemcpyHostToDevice, hst));
    if (P.z.n_inputs)  // (Z64 circuits: one proof per session) the Z64 witness fills the first leaves of the value plane; the kappa leaves stay zero
        CU(cudaMemcpyAsync(s->d_zleaf, hin + s->wit_pitch + RV_TOTAL_REPS * 16, 8 * (size_t)P.z.n_inputs, cudaMemcpyHostToDevice, hst));
    CU(cudaMemcpy2DAsync(s->d_seeds + (size_t)slot0 * s->nreps1 * 16, (size_t)s->nreps1 * 16, hin + s->wit_pitch + (size_t)s->first_rep * 16, s->in_pitch,
                         (size_t)s->nreps1 * 16, n, cudaMemcpyHostToDevice, hst));
    s->committed = s->opened = false;
    return RV_OK;
}

extern "C" int rv_session_upload_slot(rv_session *s, int slot, const uint8_t *wit_gf2, size_t n_gf2, const uint64_t *wit_z64, size_t n_z64,
                                      const uint8_t *seeds) {
    const int rc = stage_slot(s, slot, wit_gf2, n_gf2, wit_z64, n_z64, seeds, true);
    return rc != RV_OK ? rc : flush_slots(s, slot, 1);
}

// Runs `body` (a sequence of asynchronous launches on the session's streams) eagerly the first time and whenever per-kernel
// timing is on; the second plain call captures it into a CUDA graph, later calls replay that graph with one launch.
template <typename F>
static int run_graphed(rv_session *s, rv_session::GraphSlot &g, F &&body, bool track_d2h = false) {
    int rc;
    if (s->timing || g.calls == 0) {
        g.calls++;
        const uint64_t before = s->launches;
        rc = body();
        g.kernels = s->launches - before;
        return rc;
    }
    if (!g.exec) {
        cudaGraph_t graph = nullptr;
        const uint64_t before = s->launches;
        CU(cudaStreamBeginCapture(s->st, cudaStreamCaptureModeThreadLocal));
        rc = body();
        cudaError_t e = cudaStreamEndCapture(s->st, &graph);
        s->launches = before;
        if (rc != RV_OK || e != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            return rc != RV_OK ? rc : fail(RV_E_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
        }
        e = cudaGraphInstantiate(&g.exec, graph, 0);
        if (e == cudaSuccess && track_d2h) {  // the node that copies the proofs into h_out: retargeted when the caller keeps a block
            size_t n = 0;
            cudaGraphGetNodes(graph, nullptr, &n);
            std::vector<cudaGraphNode_t> nodes(n);
            if (n) cudaGraphGetNodes(graph, nodes.data(), &n);
            for (cudaGraphNode_t nd : nodes) {
                cudaGraphNodeType t;
                cudaMemcpy3DParms mp;
                if (cudaGraphNodeGetType(nd, &t) == cudaSuccess && t == cudaGraphNodeTypeMemcpy && cudaGraphMemcpyNodeGetParams(nd, &mp) == cudaSuccess &&
                    mp.dstPtr.ptr == (void *)s->h_out)
                    g.d2h_node = nd;
            }
            cudaGetLastError();
            g.graph = graph;  // node handles belong to the graph: keep it
        } else {
            cudaGraphDestroy(graph);
        }
        if (e != cudaSuccess) return fail(RV_E_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    }
    CU(cudaGraphLaunch(g.exec, s->st));
    s->launches += g.kernels;
    return RV_OK;
}

static int commit_body(rv_session *s) {
    const rv_circuit *c = s->c;
    const Program &P = c->prog;
    const DevProgram &D = c->dev;
    CU(cudaSetDevice(c->device));
    const uint32_t nslices = 2 * s->npi;
    // value plane on its own stream: it depends only on the witness and overlaps the whole mask pipeline.  Forking from the
    // main stream orders it after the upload and after the previous proof's item plane (which still reads d_vals), and
    // makes the whole proof capturable as one CUDA graph.
    CU(cudaEventRecord(s->ev_fork, s->st));
    CU(cudaStreamWaitEvent(s->st_val, s->ev_fork, 0));
    if (P.values_wide) {
        Scope k(s, "values", (uint64_t)P.wgates.size() * sizeof(VGate), 1 + D.n_wlevels, s->st_val);
        launch_values_wide(D, P.wlevel_off.data(), s->d_wit, s->d_vals, s->st_val);
    } else {
        Scope k(s, "values", (uint64_t)P.lut_steps.size() * sizeof(LutInstr), 1, s->st_val);
        launch_values(D.lut_steps, D.n_lut_steps, D.input_vid, s->d_wit, s->wit_pitch, D.n_inputs, s->d_vals, s->vals_pitch, D.n_vals, s->n_proofs, s->st_val);
    }
    const DevZProgram &DZ = c->zdev;
    if (s->has_z) {
        Scope k(s, "z.values", (uint64_t)P.z.vprog.size() * sizeof(ZInstr), 1, s->st_val);
        launch_zvalues(DZ, s->d_zleaf, 0, s->d_zvals, 0, 1, s->d_vals, D.b2a_vrefs, s->st_val);
    }
    CU(cudaEventRecord(s->ev_vals, s->st_val));
    {
        Scope k(s, "key_setup", 0);  // also clears the proof's "an AssertZero failed" flag
        launch_key_setup(s->d_seeds, nullptr, nullptr, nullptr, nslices, s->d_pkeys, s->d_rk_plain, s->st, s->d_bad, s->n_proofs, s->proof_pitch);
    }
    {
        Scope k(s, "mask_gen", (uint64_t)P.n_masks * s->npi * 8);
        launch_mask_gen_tt(s->d_rk_plain, nslices, P.n_masks, s->d_rows, s->d_fresh_sm, s->pitch_fresh, c->n_sms, s->st,
                           P.values_wide ? 0 : s->n_proofs /* k_values: one CTA (one SM) per proof, on the side stream */, s->share, linear_vm_pairs(D, s->npi));
    }
    if (D.n_llevels) {
        const double avg_width = (double)D.n_xgates / D.n_llevels;
        Scope k(s, "linear", (uint64_t)P.n_lin * s->npi * 8 * 3, avg_width < 4096.0 ? 1 : D.n_llevels);
        launch_linear(D, P.xlevel_off.data(), s->d_rows, s->npi, s->d_fresh_sm, s->pitch_fresh, s->st, nullptr);
    }
    if (s->has_z) {
        {
            Scope k(s, "z.mask_gen", (uint64_t)P.z.n_masks * s->zrowlen * 8);
            launch_zmask_gen_tt(s->d_rk_plain, (uint32_t)s->zrowlen, P.z.n_masks, s->d_zrows, c->n_sms, s->st);
        }
        if (DZ.n_llevels) {
            Scope k(s, "z.linear", (uint64_t)P.z.n_lin * s->zrowlen * 8 * 3, DZ.n_llevels);
            launch_zlinear(DZ, P.z.llevel_off.data(), s->d_zrows, (uint32_t)s->zrowlen, s->st);
        }
    }
    CU(cudaStreamWaitEvent(s->st, s->ev_vals, 0));
    if (s->has_z) {
        {
            // per Mul: 4 row segments of 64 B read, 64 + 8 stream bytes written, per repetition; pre: 3 segments + 8 bytes
            Scope k(s, "z.items", ((uint64_t)P.z.n_mul * (7 * 64 + 72) + P.z.n_inputs * 72 + P.z.n_assert * 128) * s->nreps, 2);
            launch_zitems(DZ, s->d_zrows, s->zrowlen, s->nreps, s->d_zvals, s->d_rows, s->d_zon, s->pitch_zon, s->d_zpre, s->pitch_zpre, s->d_bad, s->st);
        }
        {
            Scope k(s, "z.chunk_cv", (P.z.on_bytes + P.z.pre_bytes) * s->nreps, 1);
            launch_chunk_cv2(s->d_zon, s->pitch_zon, (uint32_t)P.z.on_bytes, s->d_zcv_on, s->nreps, s->d_zpre, s->pitch_zpre, (uint32_t)P.z.pre_bytes,
                             s->d_zcv_pre, s->nreps, s->st);
        }
        {
            Scope k(s, "z.rep_hash", ((uint64_t)s->n_chunks_zon + s->n_chunks_zpre) * s->nreps * 32);
            launch_zrep_hash(s->d_zcv_on, s->n_chunks_zon, s->d_zcv_pre, s->n_chunks_zpre, s->nreps, s->d_zon_hash, s->d_zrep, s->st);
        }
    }
    {
        // per Mul: 4 row reads + 2 stream bytes per rep (online) and 3 row reads + 1 byte per rep (pre)
        Scope k(s, "items", ((uint64_t)P.n_and * 7 + P.n_inputs + P.n_assert) * s->npi * 8 + ((uint64_t)P.n_online + P.n_pre) * s->nreps, 1);
        if (D.n_tlevels) launch_tainted(D, s->d_rows, s->npi, s->d_vals, s->d_tvals, s->st);
        launch_items(D, s->d_rows, s->npi1, s->d_vals, s->d_tvals, s->d_on, s->pitch_on, s->d_pre, s->pitch_pre, s->d_bad, s->st, s->n_proofs, s->vals_pitch,
                     s->proof_pitch);
    }
    {
        Scope k(s, "chunk_cv", ((uint64_t)P.n_online + P.n_pre) * s->nreps, 1);
        launch_chunk_cv2(s->d_on, s->pitch_on, P.n_online, s->d_cv_on, s->nreps, s->d_pre, s->pitch_pre, P.n_pre, s->d_cv_pre, s->nreps, s->st);
    }
    {
        Scope k(s, "rep_hash", ((uint64_t)s->n_chunks_on + s->n_chunks_pre) * s->nreps * 32);
        launch_rep_hash(s->d_cv_on, s->n_chunks_on, s->d_cv_pre, s->n_chunks_pre, s->d_zconst, s->nreps, s->d_on_hash, s->d_rep_hash, s->st, 0xFFFFFFFFu,
                        nullptr, nullptr, s->has_z ? s->d_zrep : nullptr, s->d_cv_scratch);
    }
    CU(cudaGetLastError());
    return RV_OK;
}

extern "C" int rv_session_commit(rv_session *s) {
    if (!s) return fail(RV_E_ARG, "NULL session");
    CU(cudaSetDevice(s->c->device));
    const int rc = run_graphed(s, s->g_commit, [&] { return commit_body(s); });
    if (rc == RV_OK) s->committed = s->ever_committed = true;
    return rc;
}

extern "C" int rv_session_hashes(rv_session *s, uint8_t *rep_hashes) {
    if (!s || !rep_hashes) return fail(RV_E_ARG, "NULL argument");
    if (!s->committed) return fail(RV_E_ARG, "rv_session_commit has not run");
    CU(cudaSetDevice(s->c->device));
    CU(cudaMemcpyAsync(rep_hashes, s->d_rep_hash, (size_t)s->nreps * 32, cudaMemcpyDeviceToHost, host_stream(s)));
    CU(cudaStreamSynchronize(host_stream(s)));
    return RV_OK;
}

extern "C" const void *rv_session_hashes_device(rv_session *s) { return (s && s->committed) ? s->d_rep_hash : nullptr; }

extern "C" void *rv_session_all_hashes_device(rv_session *s) { return s ? s->d_all_hashes : nullptr; }

static int open_body(rv_session *s, const uint8_t *all_rep_hashes) {
    const rv_circuit *c = s->c;
    const DevProgram &D = c->dev;
    const uint8_t *hashes = s->d_all_hashes;
    const bool linked = s->linked() && all_rep_hashes == nullptr;  // the exchange happens inside k_challenge, over peer memory
    if (linked) {
        hashes = nullptr;
    } else if (all_rep_hashes == s->d_all_hashes) {
        // gathered in place (rv_session_all_hashes_device): nothing to copy
    } else if (all_rep_hashes) {
        cudaPointerAttributes at;
        const bool on_device = cudaPointerGetAttributes(&at, all_rep_hashes) == cudaSuccess && at.type == cudaMemoryTypeDevice;
        cudaGetLastError();
        CU(cudaMemcpyAsync(s->d_all_hashes, all_rep_hashes, (size_t)RV_TOTAL_REPS * 32 * s->n_proofs, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                           s->st));
    } else {
        if (s->npi1 != RV_PACKED_REPS) return fail(RV_E_ARG, "a partial shard needs the all-gathered repetition hashes (or rv_session_peer_link)");
        hashes = s->d_rep_hash;
    }
    {
        // the gathered hashes are rank-major over the session's proofs: [rank][proof][this rank's repetitions x 32 bytes]
        Scope k(s, "challenge", (uint64_t)RV_TOTAL_REPS * 32 * s->n_proofs);
        launch_challenge(hashes, s->nreps1 * 32, s->d_comm, s->proof_pitch, s->d_omit, s->d_rank, s->n_proofs, s->st, linked ? &s->x : nullptr, s->d_rep_hash);
    }
    // a full shard writes every byte of the proof, and so do the linked shards together (each straight into the assembling rank's buffer)
    if (s->npi1 != RV_PACKED_REPS && !linked) CU(cudaMemset2DAsync(s->d_proof, s->proof_pitch, 0, s->proof_len, s->n_proofs, s->st));
    uint8_t *proof_out = linked ? s->dst_proof : s->d_proof;
    {
        Scope k(s, "extract", s->proof_len);
        ExtractArgs a;
        a.on = s->d_on;
        a.pre = s->d_pre;
        a.pitch_on = s->pitch_on;
        a.pitch_pre = s->pitch_pre;
        a.on_hash = s->d_on_hash;
        a.pkeys = s->d_pkeys;
        a.seeds = s->d_seeds;
        a.comm = s->d_comm;
        a.omit_of_rep = s->d_omit;
        a.rank_of_rep = s->d_rank;
        a.z64_empty_hash = s->d_zconst;
        a.first_rep = s->first_rep;
        a.nreps = s->nreps1;
        a.n_proofs = s->n_proofs;
        a.proof_stride = s->proof_pitch;
        a.len_recons = s->len_recons;
        a.len_corrs = s->len_corrs;
        a.len_inputs = s->len_inputs;
        a.len_zrecons = s->len_zrecons;
        a.len_zcorrs = s->len_zcorrs;
        a.len_zinputs = s->len_zinputs;
        a.z_on_hash = s->has_z ? s->d_zon_hash : nullptr;
        a.proof = proof_out;
        launch_extract(D, a, s->st);
    }
    if (s->has_z) {
        Scope k(s, "z.extract", (uint64_t)RV_ONLINE_REPS * (s->len_zrecons + s->len_zcorrs + s->len_zinputs));
        const ProofLayout L{s->len_recons, s->len_corrs, s->len_inputs, s->len_zrecons, s->len_zcorrs, s->len_zinputs};
        ZExtractArgs a;
        a.on = s->d_zon;
        a.pre = s->d_zpre;
        a.pitch_on = s->pitch_zon;
        a.pitch_pre = s->pitch_zpre;
        a.omit_of_rep = s->d_omit;
        a.rank_of_rep = s->d_rank;
        a.first_rep = s->first_rep;
        a.nreps = s->nreps;
        a.z_base = L.z_base();
        a.sz_on_z = L.sz_on_z();
        a.proof = proof_out;
        launch_zextract(c->zdev, a, s->st);
    }
    if (linked) {
        Scope k(s, "xfinish", 0);
        launch_xfinish(s->x, s->n_proofs, s->d_bad, s->proof_pitch, s->st);
    }
    if (s->out_off && (!linked || s->assembles()))
        CU(cudaMemcpyAsync(s->h_out, s->d_proof, s->proof_pitch * s->n_proofs, cudaMemcpyDeviceToHost, s->st));  // h_out mirrors d_proof
    else  // status word + comm of every proof (big proofs are copied by rv_session_fetch; a non-assembling rank has no proof bytes)
        CU(cudaMemcpy2DAsync(s->h_out + s->out_off, s->out_off ? s->proof_pitch : 64, s->d_proof + s->tail_off, s->proof_pitch, 36, s->n_proofs,
                             cudaMemcpyDeviceToHost, s->st));
    CU(cudaGetLastError());
    return RV_OK;
}

extern "C" int rv_session_open(rv_session *s, const uint8_t *all_rep_hashes) {
    if (!s) return fail(RV_E_ARG, "NULL session");
    if (!s->committed) return fail(RV_E_ARG, "rv_session_commit has not run");
    CU(cudaSetDevice(s->c->device));
    int rc;
    if (all_rep_hashes == s->d_all_hashes) rc = run_graphed(s, s->g_open, [&] { return open_body(s, all_rep_hashes); });
    else if (!all_rep_hashes && s->linked()) rc = run_graphed(s, s->g_open_x, [&] { return open_body(s, nullptr); });
    else rc = open_body(s, all_rep_hashes);  // caller-owned buffer: its address may change from call to call
    s->out_direct = false;
    if (rc == RV_OK) s->opened = true;
    return rc;
}

// commit + open(own hashes) of a full-shard session.  After one eager run the sequence (12 kernels on two streams, memsets,
// the device-to-host copy of the proof) is captured once and replayed as a single CUDA graph launch.
extern "C" int rv_session_prove(rv_session *s) {
    if (!s) return fail(RV_E_ARG, "NULL session");
    if (s->npi1 != RV_PACKED_REPS && !s->linked()) return fail(RV_E_ARG, "rv_session_prove needs a full shard, or a shard linked to its peers (rv_session_peer_link)");
    CU(cudaSetDevice(s->c->device));
    // a session bound to a batch but launched on its own: its stream forks from the leader's (where its uploads were ordered) and
    // joins back (where its fetch synchronises)
    const bool bound = s->lead && s->lead != s;
    if (bound) {
        CU(cudaEventRecord(s->ev_bjoin, s->lead->st));
        CU(cudaStreamWaitEvent(s->st, s->ev_bjoin, 0));
    }
    const bool direct = s->out_off != 0 && !s->in_batch && s->assembles();  // small proofs: the caller gets pointers into h_out
    if (direct && s->out_handed) {  // the caller still holds the previous step's proofs: this step writes a fresh block
        const size_t bytes = (s->out_off + 64) * s->n_proofs;
        uint8_t *nb = outblock_get(bytes);
        if (!nb) return fail(RV_E_NOMEM, "pinned host allocation failed");
        if (s->g_prove.exec && s->g_prove.d2h_node &&
            cudaGraphExecMemcpyNodeSetParams1D(s->g_prove.exec, s->g_prove.d2h_node, nb, s->d_proof, s->proof_pitch * s->n_proofs, cudaMemcpyDeviceToHost) != cudaSuccess) {
            cudaGetLastError();  // could not retarget: drop the graph, it is captured again with the new block
            cudaGraphExecDestroy(s->g_prove.exec);
            if (s->g_prove.graph) cudaGraphDestroy(s->g_prove.graph);
            s->g_prove = rv_session::GraphSlot();
            s->g_prove.calls = 1;
        }
        for (rv_session::GraphSlot *g : {&s->g_open, &s->g_open_x})  // these bake the old block's address: captured again on demand
            if (g->exec) {
                cudaGraphExecDestroy(g->exec);
                *g = rv_session::GraphSlot();
                g->calls = 1;
            }
        outblock_ref(s->h_out, -1);
        s->h_out = nb;
        s->out_handed = false;
    }
    const int rc = run_graphed(s, s->g_prove, [&] {
        const int r = commit_body(s);
        return r != RV_OK ? r : open_body(s, nullptr);
    }, direct);
    s->out_direct = direct && rc == RV_OK && (s->g_prove.exec == nullptr || s->g_prove.d2h_node != nullptr);
    if (bound) {
        CU(cudaEventRecord(s->ev_bjoin, s->st));
        CU(cudaStreamWaitEvent(s->lead->st, s->ev_bjoin, 0));
    }
    if (rc == RV_OK) s->committed = s->opened = s->ever_committed = true;
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------------
//  rv_batch: several sessions of one circuit driven as one unit -- a phase of ALL of them is one CUDA graph launch on the
//  leader's stream (the sessions' own streams fork from it and join back inside the graph), so a proving service with many
//  small proofs in flight pays three launches per step instead of three per proof.
// ---------------------------------------------------------------------------------------------------------------------
struct rv_batch {
    std::vector<rv_session *> ss;
    rv_session::GraphSlot g[3];  // commit, open, prove
    cudaEvent_t ev_fork = nullptr;
};

extern "C" int rv_batch_create(rv_session *const *ss, int n, rv_batch **out) {
    if (!ss || n <= 0 || !out) return fail(RV_E_ARG, "bad argument");
    for (int i = 0; i < n; i++) {
        if (!ss[i] || ss[i]->c != ss[0]->c) return fail(RV_E_ARG, "the sessions of a batch must share one circuit");
        if (i && ss[i]->lead && ss[i]->lead != ss[0]) return fail(RV_E_ARG, "session already belongs to another batch");
    }
    rv_batch *b = new (std::nothrow) rv_batch();
    if (!b) return fail(RV_E_NOMEM, "out of memory");
    CU(cudaSetDevice(ss[0]->c->device));
    if (cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming) != cudaSuccess) {
        delete b;
        return fail(RV_E_CUDA, "event creation failed");
    }
    for (int i = 0; i < n; i++) {
        rv_session *s = ss[i];
        if (i) {
            CU(cudaStreamSynchronize(s->st));
            s->lead = ss[0];
            if (!s->ev_bjoin && cudaEventCreateWithFlags(&s->ev_bjoin, cudaEventDisableTiming) != cudaSuccess) {
                for (rv_session *f : b->ss) f->lead = nullptr;
                s->lead = nullptr;
                cudaEventDestroy(b->ev_fork);
                delete b;
                return fail(RV_E_CUDA, "event creation failed");
            }
        }
        b->ss.push_back(s);
    }
    for (rv_session *s : b->ss) {
        s->share = (uint32_t)n;
        s->in_batch = true;
    }
    *out = b;
    return RV_OK;
}

extern "C" void rv_batch_free(rv_batch *b) {
    if (!b) return;
    if (!b->ss.empty()) {
        cudaSetDevice(b->ss[0]->c->device);
        cudaStreamSynchronize(b->ss[0]->st);
    }
    for (size_t i = 1; i < b->ss.size(); i++) b->ss[i]->lead = nullptr;
    for (rv_session *s : b->ss) {
        s->share = 1;
        s->in_batch = false;
    }
    for (auto &g : b->g)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    if (b->ev_fork) cudaEventDestroy(b->ev_fork);
    delete b;
}

static int batch_run(rv_batch *b, int kind) {
    rv_session *lead = b->ss[0];
    CU(cudaSetDevice(lead->c->device));
    if (kind == 1)
        for (rv_session *s : b->ss)
            if (!s->committed) return fail(RV_E_ARG, "rv_batch_commit has not run");
    if (kind == 2)
        for (rv_session *s : b->ss)
            if (s->npi1 != RV_PACKED_REPS && !s->linked()) return fail(RV_E_ARG, "rv_batch_prove needs full shards, or shards linked to their peers");
    bool timing = false;
    std::vector<uint64_t> before;
    for (rv_session *s : b->ss) {
        timing |= s->timing;
        before.push_back(s->launches);
    }
    const bool keep = lead->timing;
    lead->timing = timing;  // any session being timed keeps the whole phase eager
    const int rc = run_graphed(lead, b->g[kind], [&] {
        if (cudaEventRecord(b->ev_fork, lead->st) != cudaSuccess) return fail(RV_E_CUDA, "cudaEventRecord failed");
        for (size_t i = 1; i < b->ss.size(); i++)
            if (cudaStreamWaitEvent(b->ss[i]->st, b->ev_fork, 0) != cudaSuccess) return fail(RV_E_CUDA, "cudaStreamWaitEvent failed");
        for (rv_session *s : b->ss) {
            int r = RV_OK;
            if (kind != 1) r = commit_body(s);
            if (r == RV_OK && kind != 0) r = open_body(s, kind == 1 ? s->d_all_hashes : nullptr);
            if (r != RV_OK) return r;
        }
        for (size_t i = 1; i < b->ss.size(); i++) {
            rv_session *f = b->ss[i];
            if (cudaEventRecord(f->ev_bjoin, f->st) != cudaSuccess || cudaStreamWaitEvent(lead->st, f->ev_bjoin, 0) != cudaSuccess)
                return fail(RV_E_CUDA, "stream join failed");
        }
        return (int)RV_OK;
    });
    lead->timing = keep;
    if (rc != RV_OK) return rc;
    const uint64_t per = b->g[kind].kernels;  // the leader's own launches of this phase = every session's (same circuit)
    for (size_t i = 0; i < b->ss.size(); i++) {
        rv_session *s = b->ss[i];
        s->launches = before[i] + per;
        if (kind != 1) s->committed = s->ever_committed = true;
        if (kind != 0) s->opened = true;
        s->out_direct = false;
    }
    return RV_OK;
}
extern "C" int rv_batch_commit(rv_batch *b) { return b ? batch_run(b, 0) : fail(RV_E_ARG, "NULL batch"); }
extern "C" int rv_batch_open(rv_batch *b) { return b ? batch_run(b, 1) : fail(RV_E_ARG, "NULL batch"); }
extern "C" int rv_batch_prove(rv_batch *b) { return b ? batch_run(b, 2) : fail(RV_E_ARG, "NULL batch"); }
extern "C" void *rv_batch_stream(rv_batch *b) { return b ? (void *)b->ss[0]->st : nullptr; }

// the status word of a proof: RV_BAD_WITNESS (an AssertZero saw a non-zero value), RV_BAD_PEER_TIMEOUT (a linked rank never arrived)
static int status_of(int bad) {
    if (bad & RV_BAD_PEER_TIMEOUT) return fail(RV_E_PEER, "a linked session of another rank did not arrive in time (rv_session_peer_link: every rank must run the same steps)");
    if (bad) return fail(RV_E_WITNESS_INVALID, "witness is invalid!");  // prover.rs:223
    return RV_OK;
}

extern "C" int rv_session_fetch(rv_session *s, uint8_t comm[RV_HASH_SIZE], uint8_t **part, size_t *part_len) {
    return rv_session_fetch_slot(s, 0, comm, part, part_len);
}

extern "C" int rv_session_fetch_slot(rv_session *s, int slot, uint8_t comm[RV_HASH_SIZE], uint8_t **part, size_t *part_len) {
    if (!s || !part || !part_len) return fail(RV_E_ARG, "NULL argument");
    if (slot < 0 || (uint32_t)slot >= s->n_proofs) return fail(RV_E_ARG, "no such proof slot");
    if (!s->opened) return fail(RV_E_ARG, "rv_session_open has not run");
    CU(cudaSetDevice(s->c->device));
    CU(cudaStreamSynchronize(host_stream(s)));
    const uint8_t *hout = s->h_out + (size_t)slot * (s->out_off ? s->proof_pitch : 64);
    int bad;
    memcpy(&bad, hout + s->out_off, 4);
    if (const int st = status_of(bad)) return st;
    if (s->linked() && !s->assembles()) {  // the proof bytes live on the assembling rank; this rank reports its status and comm
        if (comm) memcpy(comm, hout + s->out_off + 4, 32);
        *part = nullptr;
        *part_len = 0;
        return RV_OK;
    }
    uint8_t *p;
    if (s->out_off && s->out_direct) {  // no copy: the caller owns this slot of the output block until rv_free
        p = const_cast<uint8_t *>(hout);
        outblock_ref(p, +1);
        s->out_handed = true;
    } else if (s->out_off) {
        p = (uint8_t *)malloc(s->proof_len);
        if (!p) return fail(RV_E_NOMEM, "out of memory");
        memcpy(p, hout, s->proof_len);
    } else {  // big proof: device -> the returned pinned buffer, no staging copy
        p = (uint8_t *)pinned_get(s->proof_len);
        if (!p) return fail(RV_E_NOMEM, "pinned host allocation failed");
        cudaError_t e = cudaMemcpyAsync(p, s->d_proof, s->proof_len, cudaMemcpyDeviceToHost, host_stream(s));
        if (e == cudaSuccess) e = cudaStreamSynchronize(host_stream(s));
        if (e != cudaSuccess) {
            rv_free(p);
            return fail(RV_E_CUDA, std::string("proof copy: ") + cudaGetErrorString(e));
        }
    }
    if (comm) memcpy(comm, hout + s->out_off + 4, 32);
    *part = p;
    *part_len = s->proof_len;
    return RV_OK;
}

extern "C" int rv_session_status(rv_session *s) {
    if (!s) return fail(RV_E_ARG, "NULL session");
    if (!s->opened) return fail(RV_E_ARG, "rv_session_open has not run");
    CU(cudaSetDevice(s->c->device));
    CU(cudaStreamSynchronize(host_stream(s)));
    for (uint32_t b = 0; b < s->n_proofs; b++) {
        int bad;
        memcpy(&bad, s->h_out + (size_t)b * (s->out_off ? s->proof_pitch : 64) + s->out_off, 4);
        if (const int st = status_of(bad)) return st;
    }
    return RV_OK;
}

extern "C" size_t rv_session_proof_stride(const rv_session *s) { return s ? s->proof_pitch : 0; }
extern "C" int rv_session_slots(const rv_session *s) { return s ? (int)s->n_proofs : 0; }

extern "C" int rv_session_proof_device(rv_session *s, void **ptr, size_t *len) {
    if (!s || !ptr || !len) return fail(RV_E_ARG, "NULL argument");
    *ptr = s->d_proof;
    *len = s->proof_len;
    return RV_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
//  Linking the shards of one proof across GPUs (the exchange of src/proof/mod.rs:160-171 and the assembly of
//  :200-221 over NVLink peer memory instead of a collective library + host hop).  A handle names a session's exchange block and
//  proof buffer: raw pointers for sessions of the same process (peer access), CUDA IPC handles across processes.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
struct PeerHandle {
    uint64_t magic;
    int32_t pid, device;
    uint64_t xchg_ptr, proof_ptr, xchg_bytes, proof_bytes, proof_len;
    uint32_t n_proofs, npi1, first_instance, pad;
    cudaIpcMemHandle_t ipc_xchg, ipc_proof;
};
static_assert(sizeof(PeerHandle) <= RV_PEER_HANDLE_BYTES, "handle layout");
constexpr uint64_t PEER_MAGIC = 0x3130424b50565252ull;  // "RRVPKB01"
}  // namespace

extern "C" int rv_session_peer_handle(rv_session *s, uint8_t *out) {
    if (!s || !out) return fail(RV_E_ARG, "NULL argument");
    CU(cudaSetDevice(s->c->device));
    PeerHandle h;
    memset(&h, 0, sizeof h);
    h.magic = PEER_MAGIC;
    h.pid = (int32_t)getpid();
    h.device = s->c->device;
    h.xchg_ptr = (uint64_t)(uintptr_t)s->d_xchg;
    h.proof_ptr = (uint64_t)(uintptr_t)s->d_proof;
    h.xchg_bytes = s->xchg_bytes;
    h.proof_bytes = s->proof_alloc_bytes;
    h.proof_len = s->proof_len;
    h.n_proofs = s->n_proofs;
    h.npi1 = s->npi1;
    h.first_instance = s->first_instance;
    CU(cudaIpcGetMemHandle(&h.ipc_xchg, s->d_xchg));
    CU(cudaIpcGetMemHandle(&h.ipc_proof, s->d_proof));
    memset(out, 0, RV_PEER_HANDLE_BYTES);
    memcpy(out, &h, sizeof h);
    return RV_OK;
}

extern "C" int rv_session_peer_link(rv_session *s, int rank, int world, const uint8_t *handles) {
    if (!s || !handles) return fail(RV_E_ARG, "NULL argument");
    if (world < 2 || world > RV_MAX_PEERS || rank < 0 || rank >= world) return fail(RV_E_ARG, "a link joins 2..16 ranks");
    if (s->linked()) return fail(RV_E_ARG, "session is already linked");
    if ((uint32_t)world * s->npi1 != RV_PACKED_REPS || s->first_instance != (uint32_t)rank * s->npi1)
        return fail(RV_E_ARG, "rank r of a link holds packed instances [32 r / world, 32 (r + 1) / world)");
    CU(cudaSetDevice(s->c->device));
    CU(cudaStreamSynchronize(host_stream(s)));
    XchgArgs x;
    x.world = (uint32_t)world;
    x.rank = (uint32_t)rank;
    x.dst = 0;
    uint64_t tmo_ms = 60000;
    if (const char *e = getenv("RV_PEER_TIMEOUT_MS")) tmo_ms = strtoull(e, nullptr, 10);
    x.timeout_ns = tmo_ms * 1000000ull;
    uint8_t *dst_proof = nullptr;
    std::vector<void *> opened;
    auto undo = [&](int code) {
        for (void *p : opened) cudaIpcCloseMemHandle(p);
        return code;
    };
    for (int r = 0; r < world; r++) {
        PeerHandle h;
        memcpy(&h, handles + (size_t)r * RV_PEER_HANDLE_BYTES, sizeof h);
        if (h.magic != PEER_MAGIC) return undo(fail(RV_E_ARG, "not a peer handle"));
        if (h.n_proofs != s->n_proofs || h.npi1 != s->npi1 || h.first_instance != (uint32_t)r * s->npi1 || h.proof_len != s->proof_len ||
            h.xchg_bytes != s->xchg_bytes || h.proof_bytes != s->proof_alloc_bytes)
            return undo(fail(RV_E_ARG, "the linked sessions must hold the same circuit, slot count and consecutive shards in rank order"));
        uint8_t *xp = nullptr, *pp = nullptr;
        if (r == rank) {
            if (h.xchg_ptr != (uint64_t)(uintptr_t)s->d_xchg) return undo(fail(RV_E_ARG, "handles[rank] is not this session's handle"));
            xp = s->d_xchg;
            pp = s->d_proof;
        } else if (h.pid == (int32_t)getpid()) {  // same process: plain pointers, peer access between the two devices
            if (h.device != s->c->device) {
                int can = 0;
                CU(cudaDeviceCanAccessPeer(&can, s->c->device, h.device));
                if (!can) return undo(fail(RV_E_UNSUPPORTED, "no peer access between the linked devices"));
                const cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return undo(fail(RV_E_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)));
                cudaGetLastError();
            }
            xp = reinterpret_cast<uint8_t *>((uintptr_t)h.xchg_ptr);
            pp = reinterpret_cast<uint8_t *>((uintptr_t)h.proof_ptr);
        } else {  // another process: CUDA IPC
            void *m = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&m, h.ipc_xchg, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) return undo(fail(RV_E_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e)));
            opened.push_back(m);
            xp = reinterpret_cast<uint8_t *>(m);
            if (r == (int)x.dst) {
                e = cudaIpcOpenMemHandle(&m, h.ipc_proof, cudaIpcMemLazyEnablePeerAccess);
                if (e != cudaSuccess) return undo(fail(RV_E_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e)));
                opened.push_back(m);
                pp = reinterpret_cast<uint8_t *>(m);
            }
        }
        x.peer[r] = xp;
        if (r == (int)x.dst) dst_proof = pp;
    }
    s->x = x;
    s->dst_proof = dst_proof;
    s->ipc_opened = opened;
    // graphs captured before the link describe the unlinked phases
    for (rv_session::GraphSlot *g : {&s->g_prove, &s->g_open_x}) {
        if (g->exec) cudaGraphExecDestroy(g->exec);
        if (g->graph) cudaGraphDestroy(g->graph);
        *g = rv_session::GraphSlot();
    }
    return RV_OK;
}

extern "C" int rv_session_peer_rank(const rv_session *s, int *rank, int *world, int *assembles) {
    if (!s) return fail(RV_E_ARG, "NULL session");
    if (rank) *rank = (int)s->x.rank;
    if (world) *world = (int)s->x.world;
    if (assembles) *assembles = s->assembles() ? 1 : 0;
    return RV_OK;
}

// Shard blobs are full-length proofs with only the shard's entries filled in (zero elsewhere); entries never overlap,
// so the assembly of src/proof/mod.rs:200-221 is a byte-wise OR.
extern "C" int rv_proof_assemble(const uint8_t comm[RV_HASH_SIZE], const uint8_t *const *parts, const size_t *part_lens, int n_parts,
                                 uint8_t **proof, size_t *proof_len) {
    if (!parts || !part_lens || n_parts <= 0 || !proof || !proof_len) return fail(RV_E_ARG, "bad argument");
    const size_t n = part_lens[0];
    for (int i = 1; i < n_parts; i++)
        if (part_lens[i] != n) return fail(RV_E_FORMAT, "shard blobs differ in length");
    if (n < 32) return fail(RV_E_FORMAT, "shard blob too short");
    uint8_t *p = (uint8_t *)calloc(n, 1);
    if (!p) return fail(RV_E_NOMEM, "out of memory");
    for (int i = 0; i < n_parts; i++)
        for (size_t k = 0; k < n; k++) p[k] |= parts[i][k];
    if (comm) memcpy(p, comm, 32);
    *proof = p;
    *proof_len = n;
    return RV_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
//  Proof::new / Proof::verify
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int rv_prove(const rv_circuit *c, const uint8_t *wit_gf2, size_t n_gf2, const uint64_t *wit_z64, size_t n_z64,
                        const uint8_t *seeds, uint8_t **proof, size_t *proof_len) {
    if (!c || !proof || !proof_len) return fail(RV_E_ARG, "NULL argument");
    rv_session *s = nullptr;
    {
        std::lock_guard<std::mutex> g(c->pool_mu);
        if (!c->pool.empty()) {
            s = c->pool.back();
            c->pool.pop_back();
        }
    }
    int rc = RV_OK;
    if (!s && (rc = rv_session_create(c, 0, RV_PACKED_REPS, &s))) return rc;
    if ((rc = rv_session_upload(s, wit_gf2, n_gf2, wit_z64, n_z64, seeds)) == RV_OK && (rc = rv_session_prove(s)) == RV_OK)
        rc = rv_session_fetch(s, nullptr, proof, proof_len);
    if (rc == RV_E_CUDA) {
        rv_session_free(s);
        return rc;
    }
    std::lock_guard<std::mutex> g(c->pool_mu);
    if (c->pool.size() < 32) c->pool.push_back(s);
    else rv_session_free(s);
    return rc;
}

// Proof::new for n independent witnesses of one circuit, the way a proving service sees its queue: the proofs are packed side
// by side into multi-proof sessions (RV_BATCH_SLOTS per session, each on its own stream), everything is enqueued before the
// first synchronisation, and every session's phase is one launch per kernel for all of its proofs.
static constexpr int RV_BATCH_SLOTS = 8;
extern "C" int rv_prove_batch(const rv_circuit *c, int n, const uint8_t *const *wit_gf2, const size_t *n_gf2, const uint64_t *const *wit_z64,
                              const size_t *n_z64, const uint8_t *const *seeds, uint8_t **proofs, size_t *proof_lens, int *statuses) {
    if (!c || n <= 0 || !proofs || !proof_lens || !statuses) return fail(RV_E_ARG, "bad argument");
    for (int i = 0; i < n; i++) {
        proofs[i] = nullptr;
        proof_lens[i] = 0;
        statuses[i] = RV_OK;
    }
    // groups of up to RV_BATCH_SLOTS witnesses, each in a multi-proof session with exactly as many slots as it has witnesses
    // (idle sessions are pooled per slot count); circuits a multi-proof session does not serve (Z64 / Random / B2A / big
    // proofs) take one rv_prove per witness
    const int n_sess = (n + RV_BATCH_SLOTS - 1) / RV_BATCH_SLOTS;
    auto slots_of = [&](int k) { return std::min(RV_BATCH_SLOTS, n - k * RV_BATCH_SLOTS); };
    auto take = [&](int slots, rv_session **out) -> int {
        {
            std::lock_guard<std::mutex> g(c->pool_mu);
            for (size_t i = 0; i < c->multi_pool.size(); i++)
                if ((int)c->multi_pool[i]->n_proofs == slots) {
                    *out = c->multi_pool[i];
                    c->multi_pool.erase(c->multi_pool.begin() + i);
                    return RV_OK;
                }
        }
        return rv_session_create_multi(c, 0, RV_PACKED_REPS, slots, out);
    };
    std::vector<rv_session *> used;
    int rc = RV_OK;
    for (int k = 0; k < n_sess && rc == RV_OK; k++) {
        rv_session *s = nullptr;
        rc = take(slots_of(k), &s);
        if (rc == RV_E_UNSUPPORTED && used.empty()) {
            int worst = RV_OK;
            for (int i = 0; i < n; i++) {
                statuses[i] = rv_prove(c, wit_gf2 ? wit_gf2[i] : nullptr, n_gf2 ? n_gf2[i] : 0, wit_z64 ? wit_z64[i] : nullptr, n_z64 ? n_z64[i] : 0,
                                       seeds ? seeds[i] : nullptr, &proofs[i], &proof_lens[i]);
                if (statuses[i] != RV_OK && worst == RV_OK) worst = statuses[i];
            }
            return worst == RV_E_CUDA ? worst : RV_OK;
        }
        if (s) used.push_back(s);
    }
    // fill the slots, launch, then collect
    for (int k = 0; k < (int)used.size() && rc == RV_OK; k++) {
        for (int b = 0; b < slots_of(k) && rc == RV_OK; b++) {
            const int i = k * RV_BATCH_SLOTS + b;
            const int r = stage_slot(used[k], b, wit_gf2 ? wit_gf2[i] : nullptr, n_gf2 ? n_gf2[i] : 0, wit_z64 ? wit_z64[i] : nullptr, n_z64 ? n_z64[i] : 0,
                                     seeds ? seeds[i] : nullptr, b == 0);
            if (r == RV_E_WITNESS_SHORT || r == RV_E_ARG) statuses[i] = r;  // the slot keeps whatever it held (its output is dropped)
            else if (r != RV_OK) rc = r;
        }
        if (rc == RV_OK) rc = flush_slots(used[k], 0, slots_of(k));
        if (rc == RV_OK) rc = rv_session_prove(used[k]);
    }
    for (int k = 0; k < (int)used.size() && rc == RV_OK; k++)  // collect in launch order: session k's copies overlap the later sessions' tails
        for (int b = 0; b < slots_of(k); b++) {
            const int i = k * RV_BATCH_SLOTS + b;
            if (statuses[i] != RV_OK) continue;
            statuses[i] = rv_session_fetch_slot(used[k], b, nullptr, &proofs[i], &proof_lens[i]);
            if (statuses[i] == RV_E_CUDA) rc = RV_E_CUDA;
        }
    {
        std::lock_guard<std::mutex> g(c->pool_mu);
        for (rv_session *s : used) {
            if (rc != RV_E_CUDA && c->multi_pool.size() < 24) c->multi_pool.push_back(s);
            else rv_session_free(s);
        }
    }
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------------
//  rv_group: Proof::new on several GPUs behind one handle (SURVEY.md 8(b) "one host thread drives G GPUs ... or one thread
//  per GPU", 8(e)).  A group owns, per GPU it drives ("member"), n_sessions linked multi-proof sessions of `slots` proofs and
//  the rv_batch that launches them as one CUDA graph.  Local groups hold all `world` members in this process (the circuit is
//  cloned onto each device, the shards are linked by peer access); rank groups hold one member and are linked to the other
//  processes' groups through their handles (CUDA IPC).
// ---------------------------------------------------------------------------------------------------------------------
struct rv_group {
    struct Member {
        int rank = 0;
        const rv_circuit *c = nullptr;
        rv_circuit *owned = nullptr;  // clone made for this member's device
        std::vector<rv_session *> ss;  // each on its own stream, launched one by one: uploads of the next overlap the previous one's work
    };
    std::vector<Member> members;
    int world = 1, n_sessions = 1, slots = 1;
    bool linked = false, local = false;
};

extern "C" void rv_group_free(rv_group *g) {
    if (!g) return;
    for (auto &m : g->members) {
        for (rv_session *s : m.ss) rv_session_free(s);
        if (m.owned) rv_circuit_free(m.owned);
    }
    delete g;
}

static int group_add_member(rv_group *g, const rv_circuit *c, rv_circuit *owned, int rank) {
    g->members.emplace_back();
    rv_group::Member &m = g->members.back();
    m.rank = rank;
    m.c = c;
    m.owned = owned;
    const int per = RV_PACKED_REPS / g->world;
    for (int i = 0; i < g->n_sessions; i++) {
        rv_session *s = nullptr;
        const int rc = rv_session_create_multi(c, rank * per, per, g->slots, &s);
        if (rc) return rc;
        s->share = (uint32_t)g->n_sessions;
        m.ss.push_back(s);
    }
    return RV_OK;
}

static int group_check_shape(int world, int n_sessions, int slots) {
    if (world < 1 || world > RV_MAX_PEERS || RV_PACKED_REPS % world) return fail(RV_E_ARG, "the number of GPUs must divide the 32 packed instances (1, 2, 4, 8, 16)");
    if (n_sessions < 1 || n_sessions > 64 || slots < 1 || slots > 128) return fail(RV_E_ARG, "a group holds 1..64 sessions of 1..128 proofs");
    return RV_OK;
}

static int group_link_all(rv_group *g, const uint8_t *all, size_t stride) {  // all: [rank][session][RV_PEER_HANDLE_BYTES]
    std::vector<uint8_t> hs((size_t)g->world * RV_PEER_HANDLE_BYTES);
    for (auto &m : g->members)
        for (int i = 0; i < g->n_sessions; i++) {
            for (int r = 0; r < g->world; r++) memcpy(hs.data() + (size_t)r * RV_PEER_HANDLE_BYTES, all + (size_t)r * stride + (size_t)i * RV_PEER_HANDLE_BYTES, RV_PEER_HANDLE_BYTES);
            const int rc = rv_session_peer_link(m.ss[i], m.rank, g->world, hs.data());
            if (rc) return rc;
        }
    g->linked = true;
    return RV_OK;
}

extern "C" int rv_group_create_local(const rv_circuit *c, const int *devices, int n_devices, int n_sessions, int slots, rv_group **out) {
    if (!c || !out || (n_devices > 0 && !devices)) return fail(RV_E_ARG, "NULL argument");
    *out = nullptr;
    if (const int rc = group_check_shape(n_devices, n_sessions, slots)) return rc;
    if (c->device < 0) return fail(RV_E_CUDA, "no CUDA device: reverie-b200 has no CPU fallback");
    rv_group *g = new (std::nothrow) rv_group();
    if (!g) return fail(RV_E_NOMEM, "out of memory");
    g->world = n_devices;
    g->n_sessions = n_sessions;
    g->slots = slots;
    g->local = true;
    g->members.reserve(n_devices);
    int rc = RV_OK;
    for (int r = 0; r < n_devices && rc == RV_OK; r++) {
        const rv_circuit *mc = c;
        rv_circuit *owned = nullptr;
        if (devices[r] != c->device) {
            rc = rv_circuit_clone(c, devices[r], &owned);
            mc = owned;
        }
        if (rc == RV_OK) rc = group_add_member(g, mc, owned, r);
        else if (owned) rv_circuit_free(owned);
    }
    if (rc == RV_OK && n_devices > 1) {
        const size_t stride = (size_t)n_sessions * RV_PEER_HANDLE_BYTES;
        std::vector<uint8_t> all((size_t)n_devices * stride);
        for (int r = 0; r < n_devices && rc == RV_OK; r++)
            for (int i = 0; i < n_sessions && rc == RV_OK; i++) rc = rv_session_peer_handle(g->members[r].ss[i], all.data() + (size_t)r * stride + (size_t)i * RV_PEER_HANDLE_BYTES);
        if (rc == RV_OK) rc = group_link_all(g, all.data(), stride);
    }
    if (rc != RV_OK) {
        rv_group_free(g);
        return rc;
    }
    *out = g;
    return RV_OK;
}

extern "C" int rv_group_create_rank(const rv_circuit *c, int rank, int world, int n_sessions, int slots, rv_group **out) {
    if (!c || !out) return fail(RV_E_ARG, "NULL argument");
    *out = nullptr;
    if (const int rc = group_check_shape(world, n_sessions, slots)) return rc;
    if (rank < 0 || rank >= world) return fail(RV_E_ARG, "rank out of range");
    if (c->device < 0) return fail(RV_E_CUDA, "no CUDA device: reverie-b200 has no CPU fallback");
    rv_group *g = new (std::nothrow) rv_group();
    if (!g) return fail(RV_E_NOMEM, "out of memory");
    g->world = world;
    g->n_sessions = n_sessions;
    g->slots = slots;
    g->members.reserve(1);
    const int rc = group_add_member(g, c, nullptr, rank);
    if (rc != RV_OK) {
        rv_group_free(g);
        return rc;
    }
    *out = g;
    return RV_OK;
}

extern "C" size_t rv_group_handles_bytes(const rv_group *g) { return g ? (size_t)g->n_sessions * RV_PEER_HANDLE_BYTES : 0; }

extern "C" int rv_group_handles(rv_group *g, uint8_t *out) {
    if (!g || !out) return fail(RV_E_ARG, "NULL argument");
    if (g->local) return fail(RV_E_ARG, "a local group is linked at creation");
    for (int i = 0; i < g->n_sessions; i++)
        if (const int rc = rv_session_peer_handle(g->members[0].ss[i], out + (size_t)i * RV_PEER_HANDLE_BYTES)) return rc;
    return RV_OK;
}

extern "C" int rv_group_link(rv_group *g, const uint8_t *all_handles) {
    if (!g || !all_handles) return fail(RV_E_ARG, "NULL argument");
    if (g->local || g->linked) return fail(RV_E_ARG, "group is already linked");
    if (g->world == 1) return RV_OK;
    return group_link_all(g, all_handles, rv_group_handles_bytes(g));
}

extern "C" int rv_group_info(const rv_group *g, int *world, int *n_members, int *n_sessions, int *slots) {
    if (!g) return fail(RV_E_ARG, "NULL group");
    if (world) *world = g->world;
    if (n_members) *n_members = (int)g->members.size();
    if (n_sessions) *n_sessions = g->n_sessions;
    if (slots) *slots = g->slots;
    return RV_OK;
}
extern "C" rv_session *rv_group_session(rv_group *g, int member, int index) {
    if (!g || member < 0 || member >= (int)g->members.size() || index < 0 || index >= g->n_sessions) return nullptr;
    return g->members[member].ss[index];
}

// Launches one step (commit + exchange + open) of session `i` of every member, asynchronously: one CUDA graph launch per GPU.
static int group_launch(rv_group *g, int i) {
    for (auto &m : g->members)
        if (const int rc = rv_session_prove(m.ss[i])) return rc;
    return RV_OK;
}

extern "C" int rv_group_step(rv_group *g) {
    if (!g) return fail(RV_E_ARG, "NULL group");
    if (g->world > 1 && !g->linked) return fail(RV_E_ARG, "rv_group_link has not run");
    for (int i = 0; i < g->n_sessions; i++)
        if (const int rc = group_launch(g, i)) return rc;
    return RV_OK;
}

extern "C" int rv_group_prove_batch(rv_group *g, int n, const uint8_t *const *wit_gf2, const size_t *n_gf2, const uint64_t *const *wit_z64,
                                    const size_t *n_z64, const uint8_t *const *seeds, uint8_t **proofs, size_t *proof_lens, int *statuses) {
    if (!g || n <= 0 || !statuses) return fail(RV_E_ARG, "bad argument");
    if (g->world > 1 && !g->linked) return fail(RV_E_ARG, "rv_group_link has not run");
    const bool assembles = g->members[0].rank == 0;
    if (assembles && (!proofs || !proof_lens)) return fail(RV_E_ARG, "the assembling rank needs the output arrays");
    // every rank of a proof must use the same 256 seeds: a local group draws them here, a rank group needs them from the caller
    std::vector<uint8_t> drawn;
    if (g->world > 1) {
        bool missing = !seeds;
        for (int i = 0; i < n && !missing; i++) missing = !seeds[i];
        if (missing) {
            if (!g->local) return fail(RV_E_ARG, "a rank group needs the proofs' seeds from the caller (the same on every rank: draw them on rank 0 and broadcast)");
            drawn.resize((size_t)n * RV_TOTAL_REPS * 16);
            size_t got = 0;
            while (got < drawn.size()) {
                ssize_t r = getrandom(drawn.data() + got, drawn.size() - got, 0);
                if (r <= 0) return fail(RV_E_ARG, "getrandom failed");
                got += (size_t)r;
            }
        }
    }
    auto seed_of = [&](int i) -> const uint8_t * {
        if (seeds && seeds[i]) return seeds[i];
        return drawn.empty() ? nullptr : drawn.data() + (size_t)i * RV_TOTAL_REPS * 16;
    };
    for (int i = 0; i < n; i++) {
        statuses[i] = RV_OK;
        if (proofs) proofs[i] = nullptr;
        if (proof_lens) proof_lens[i] = 0;
    }
    const int cap = g->n_sessions * g->slots;
    int rc = RV_OK;
    for (int base = 0; base < n && rc == RV_OK; base += cap) {  // waves of up to `cap` proofs
        const int cnt = std::min(cap, n - base), n_used = (cnt + g->slots - 1) / g->slots;
        // session by session: fill its slots on every member, launch it; the next session's uploads overlap its work
        for (int si = 0; si < n_used && rc == RV_OK; si++) {
            const int k0 = si * g->slots, k1 = std::min(cnt, (si + 1) * g->slots);
            for (auto &m : g->members) {
                for (int k = k0; k < k1 && rc == RV_OK; k++) {
                    const int i = base + k;
                    const int r = stage_slot(m.ss[si], k - k0, wit_gf2 ? wit_gf2[i] : nullptr, n_gf2 ? n_gf2[i] : 0, wit_z64 ? wit_z64[i] : nullptr,
                                             n_z64 ? n_z64[i] : 0, seed_of(i), k == k0);
                    if (r == RV_E_WITNESS_SHORT || r == RV_E_ARG) statuses[i] = r;  // the slot keeps whatever it held (its output is dropped)
                    else if (r != RV_OK) rc = r;
                }
                if (rc == RV_OK) rc = flush_slots(m.ss[si], 0, k1 - k0);
            }
            if (rc == RV_OK) rc = group_launch(g, si);
        }
        for (auto &m : g->members)
            for (int k = 0; k < cnt && rc == RV_OK; k++) {
                const int i = base + k;
                if (statuses[i] != RV_OK && statuses[i] != RV_E_WITNESS_INVALID) continue;
                uint8_t *p = nullptr;
                size_t len = 0;
                const int r = rv_session_fetch_slot(m.ss[k / g->slots], k % g->slots, nullptr, &p, &len);
                if (r == RV_E_CUDA || r == RV_E_PEER) rc = r;
                if (r != RV_OK) {
                    statuses[i] = r;
                    if (p) rv_free(p);
                } else if (m.rank == 0 && proofs) {
                    proofs[i] = p;
                    proof_lens[i] = len;
                } else if (p) rv_free(p);
            }
    }
    return rc;
}

// Proof::verify for a queue of proofs, spread over the group's GPUs.  Verification has no exchange step (the 32 packs of a proof
// are independent, src/proof/mod.rs:234-280), so whole proofs go to whole GPUs: member m takes the proofs i = m (mod members) --
// a rank group those with i = rank (mod world), the other entries are left untouched -- with a few verifications in flight per GPU.
extern "C" int rv_group_verify_batch(rv_group *g, int n, const uint8_t *const *proofs, const size_t *lens, int *results, int *okay) {
    if (!g || n <= 0 || !proofs || !lens || !results) return fail(RV_E_ARG, "bad argument");
    const int stride = g->local ? (int)g->members.size() : g->world;
    constexpr int IN_FLIGHT = 8;
    std::vector<std::thread> pool;
    std::atomic<int> worst{RV_OK};
    for (size_t mi = 0; mi < g->members.size(); mi++) {
        const rv_circuit *c = g->members[mi].c;
        const int first = g->local ? (int)mi : g->members[mi].rank;
        for (int t = 0; t < IN_FLIGHT; t++)
            pool.emplace_back([=, &worst]() {
                for (int i = first + t * stride; i < n; i += IN_FLIGHT * stride) {
                    int ok = 1;
                    results[i] = rv_verify(c, proofs[i], lens[i], &ok);
                    if (okay) okay[i] = ok;
                    if (results[i] == RV_E_CUDA) worst = RV_E_CUDA;
                }
            });
    }
    for (auto &t : pool) t.join();
    return worst.load();
}

extern "C" int rv_group_prove(rv_group *g, const uint8_t *wit_gf2, size_t n_gf2, const uint64_t *wit_z64, size_t n_z64, const uint8_t *seeds,
                              uint8_t **proof, size_t *proof_len) {
    int status = RV_OK;
    uint8_t *p = nullptr;
    size_t len = 0;
    const int rc = rv_group_prove_batch(g, 1, &wit_gf2, &n_gf2, &wit_z64, &n_z64, seeds ? &seeds : nullptr, &p, &len, &status);
    if (proof) *proof = p;
    else if (p) rv_free(p);
    if (proof_len) *proof_len = len;
    return rc != RV_OK ? rc : status;
}

// ---------------------------------------------------------------------------------------------------------------------
//  Streaming Proof::new (the "streaming interface" of the reference's README.md:14): circuits whose share tensor
//  and transcripts (~1.1 KB of device memory per gate) do not fit in HBM.  The op list is cut into segments of `window_ops` ops.
//  Wires that cross a segment boundary are carried in a "cell file" on the device (one slot per LIVE wire: plaintext bit + mask
//  row), PRG streams and both hash streams simply continue (random-access CTR; BLAKE3 chunk CVs are independent, a partial chunk
//  is carried), and because the openings can only be extracted once the challenge is known -- which needs every repetition's
//  hash -- the circuit is walked twice: pass 1 hashes, pass 2 recomputes the streams and packs the opened repetitions' bits.
//  Device memory: O(window) buffers + the segments' gate tables (~45 bytes per gate) + 32 bytes of chunk CVs per 1024 gates and
//  repetition + the proof.  GF(2) circuits without Random / Z64 / B2A; one GPU.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
struct DevBuf {  // frees what a failed or finished streaming call allocated
    std::vector<void *> dev, host;
    std::vector<rv_circuit *> circuits;
    cudaStream_t st = nullptr;
    ~DevBuf() {
        if (st) {
            cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
        }
        for (void *p : dev) cudaFree(p);
        for (void *p : host) cudaFreeHost(p);
        for (rv_circuit *c : circuits) rv_circuit_free(c);
    }
    template <typename T>
    int alloc(T **p, size_t count) {
        void *d = nullptr;
        const cudaError_t e = cudaMalloc(&d, std::max<size_t>(count, 1) * sizeof(T));
        if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? RV_E_NOMEM : RV_E_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
        dev.push_back(d);
        *p = reinterpret_cast<T *>(d);
        return RV_OK;
    }
};
}  // namespace

// Test hook (CPU): the streaming planner alone, checked by a symbolic simulation of the cell file -- every wire a segment reads
// before writing it (and that an earlier segment wrote) must be imported from a slot that holds exactly the state its last
// writer exported, however slots were recycled in between.  out: segments, slots, most imports / exports of a segment, totals.
extern "C" int rv_stream_plan_check(const rv_op *ops, size_t n_ops, size_t gf2_cells, size_t window_ops, uint64_t out[6]) {
    if ((n_ops && !ops) || !out) return fail(RV_E_ARG, "NULL argument");
    window_ops = std::max<size_t>(window_ops ? window_ops : ((size_t)1 << 22), 64);
    StreamPlan plan;
    std::string perr;
    if (const int rc = plan_stream(ops, n_ops, gf2_cells, window_ops, plan, perr)) return fail(rc, perr);
    constexpr int64_t NEVER = -1;
    std::vector<int64_t> last_write(plan.gf2_cells, NEVER);
    std::vector<uint32_t> seen(plan.gf2_cells, 0);  // segment tag of the wire's first access
    struct Held {
        uint32_t cell;
        int64_t version;
    };
    std::vector<Held> slot(plan.n_slots, Held{0xFFFFFFFFu, NEVER});
    uint64_t max_imp = 0, max_exp = 0, tot_imp = 0, tot_exp = 0;
    for (size_t k = 0; k < plan.segs.size(); k++) {
        const Segment &S = plan.segs[k];
        const uint32_t tag = (uint32_t)k + 1;
        std::vector<uint8_t> imported;  // by cell, sparse: mark through `seen` with a second pass instead of a map
        for (size_t j = 0; j < S.import_global.size(); j++) {
            const uint32_t c = S.import_global[j], sl = S.import_slot[j];
            if (sl >= plan.n_slots || slot[sl].cell != c || slot[sl].version != last_write[c] || last_write[c] == NEVER)
                return fail(RV_E_ARG, "plan check: segment " + std::to_string(k) + " imports wire " + std::to_string(c) + " from a slot that does not hold its latest state");
        }
        // completeness: a wire whose first access in the segment is a read, and that has been written before, is among the imports
        size_t n_need = 0;
        for (size_t i = S.a; i < S.b; i++) {
            const rv_op &op = ops[i];
            if (op.domain != RV_GF2) continue;
            uint32_t r[2] = {0, 0};
            int nr = 0;
            switch (op.opcode) {
                case RV_ADD: case RV_SUB: case RV_MUL: r[0] = op.a, r[1] = op.b, nr = 2; break;
                case RV_ADDC: case RV_SUBC: case RV_MULC: case RV_ASSERT_ZERO: r[0] = op.a, nr = 1; break;
                default: break;
            }
            for (int q = 0; q < nr; q++)
                if (seen[r[q]] != tag) {
                    seen[r[q]] = tag;
                    if (last_write[r[q]] != NEVER && last_write[r[q]] < (int64_t)S.a) n_need++;
                }
            if (op.opcode != RV_ASSERT_ZERO) {
                seen[op.dst] = tag;
                last_write[op.dst] = (int64_t)i;
            }
        }
        if (n_need != S.import_global.size())
            return fail(RV_E_ARG, "plan check: segment " + std::to_string(k) + " needs " + std::to_string(n_need) + " carried wires but imports " +
                                      std::to_string(S.import_global.size()));
        for (size_t j = 0; j < S.export_global.size(); j++) {
            if (S.export_slot[j] >= plan.n_slots) return fail(RV_E_ARG, "plan check: export slot out of range");
            slot[S.export_slot[j]] = Held{S.export_global[j], last_write[S.export_global[j]]};
        }
        max_imp = std::max<uint64_t>(max_imp, S.import_global.size());
        max_exp = std::max<uint64_t>(max_exp, S.export_global.size());
        tot_imp += S.import_global.size();
        tot_exp += S.export_global.size();
    }
    out[0] = plan.segs.size(), out[1] = plan.n_slots, out[2] = max_imp, out[3] = max_exp, out[4] = tot_imp, out[5] = tot_exp;
    return RV_OK;
}

// A compiled segment's tables go to the device from the thread that compiled it (the derived host tables and the pageable copies
// of 72 segments of 4 M gates cost 3.9 s when done one after the other, more than the planner and both GPU passes together).
static int segment_to_device(Segment &S, int device) {
    rv_circuit *c = S.c;
    Program &P = c->prog;
    if (P.n_tvals || P.z.any()) return fail(RV_E_UNSUPPORTED, "streaming mode serves GF(2) circuits without Random / Z64 / B2A");
    S.n_on = P.n_online, S.n_pre = P.n_pre, S.n_in = (uint32_t)P.n_inputs, S.n_recon = (uint32_t)P.recon_pos.size();
    S.n_imp = (uint32_t)S.io.import_cells.size(), S.n_exp = (uint32_t)S.io.export_cells.size();
    {  // one arena for everything this segment uploads (a table that should not fit falls back to its own allocation)
        size_t cap = 0;
        auto add = [&](size_t n, size_t elem) { cap += round_up(std::max<size_t>(n, 1) * elem, 256); };
        add(P.xgates.size(), sizeof(XGate)), add(P.xlevel_off.size(), 4), add(P.items.size(), sizeof(Item)), add(P.n_and, 4), add(P.recon_pos.size(), 4);
        add(P.input_pos.size(), 4), add(P.input_vid.size(), 4), add(P.vm_steps.size(), sizeof(VmInstr)), add(P.lut_steps.size(), sizeof(LutInstr));
        add(P.vlut_steps.size(), sizeof(LutInstr)), add(P.input_uid.size() + P.kappa_uid.size() + P.rand_uid.size(), 4), add(P.item_ua.size(), 4), add(P.item_ub.size(), 4);
        add(P.n_online, 4), add(P.tgates.size(), sizeof(TGate)), add(P.tlevel_off.size(), 4), add(P.rand_row.size(), 4), add(P.b2a_vrefs.size(), 4), add(P.b2a_urefs.size(), 4);
        add(P.wgates.size(), sizeof(VGate)), add(P.vwgates.size(), sizeof(VGate));
        add(P.input_vid.size() + S.io.import_vid.size(), 4), add(S.import_slot.size(), 4), add(S.export_slot.size(), 4), add(S.io.export_row.size(), 4), add(S.io.export_vref.size(), 4);
        void *a = nullptr;
        if (cudaSetDevice(device) == cudaSuccess && cudaMalloc(&a, cap) == cudaSuccess) {
            c->allocs.push_back(a);
            c->arena = (uint8_t *)a, c->arena_cap = cap, c->arena_off = 0;
        } else cudaGetLastError();
    }
    int rc = circuit_to_device(c, device);
    std::vector<uint32_t> leaf_ids(P.input_vid);
    leaf_ids.insert(leaf_ids.end(), S.io.import_vid.begin(), S.io.import_vid.end());
    if (rc == RV_OK) rc = upload(c, leaf_ids, &S.d_leaf_ids);
    if (rc == RV_OK) rc = upload(c, S.import_slot, &S.d_imp_slot);
    if (rc == RV_OK) rc = upload(c, S.export_slot, &S.d_exp_slot);
    if (rc == RV_OK) rc = upload(c, S.io.export_row, &S.d_exp_row);
    if (rc == RV_OK) rc = upload(c, S.io.export_vref, &S.d_exp_vref);
    if (rc != RV_OK) return rc;
    // the host copies of the big tables are not needed any more (launches read the device copies and the small level offsets)
    std::vector<Item>().swap(P.items);
    std::vector<XGate>().swap(P.xgates);
    std::vector<LutInstr>().swap(P.lut_steps);
    std::vector<VmInstr>().swap(P.vm_steps);
    std::vector<VGate>().swap(P.wgates);
    std::vector<uint32_t>().swap(P.recon_pos);
    std::vector<uint32_t>().swap(P.input_pos);
    std::vector<uint32_t>().swap(P.input_vid);
    std::vector<uint32_t>().swap(c->mul_pos);
    std::vector<uint32_t>().swap(c->recon_idx);
    return RV_OK;
}

extern "C" int rv_prove_streaming(const rv_op *ops, size_t n_ops, size_t z64_cells, size_t gf2_cells, const uint8_t *wit_gf2, size_t n_gf2,
                                  const uint64_t *wit_z64, size_t n_z64, const uint8_t *seeds, size_t window_ops, uint8_t **proof, size_t *proof_len) {
    (void)wit_z64;
    (void)n_z64;
    (void)z64_cells;
    if (!proof || !proof_len || (n_ops && !ops)) return fail(RV_E_ARG, "NULL argument");
    *proof = nullptr;
    *proof_len = 0;
    if (window_ops == 0) window_ops = (size_t)1 << 22;  // ~5 GB of window buffers
    window_ops = std::max<size_t>(window_ops, 64);
    // ---- 1 + 2. planning (this thread) and compilation of the segments (host threads), pipelined: a segment is compiled as soon
    //      as the planner has published it ----
    if (rv_device_count() == 0) {  // no device: the arguments are still checked
        StreamPlan probe;
        std::string perr0;
        if (const int prc = plan_stream(ops, n_ops, gf2_cells, window_ops, probe, perr0)) return fail(prc, perr0);
        return fail(RV_E_CUDA, "no CUDA device: reverie-b200 has no CPU fallback");
    }
    const bool trace = std::getenv("RV_TRACE") != nullptr;  // stage times on stderr (the GPU stages are synchronised for it)
    auto t_last = std::chrono::steady_clock::now();
    CU(cudaSetDevice(g_device));
    if (const int ce = configure_kernels(g_device)) return fail(RV_E_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString((cudaError_t)ce));
    auto mark = [&](const char *what, cudaStream_t sync = nullptr) {
        if (!trace) return;
        if (sync) cudaStreamSynchronize(sync);
        const auto n = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[rv_stream] %-34s %9.1f ms\n", what, std::chrono::duration<double, std::milli>(n - t_last).count());
        t_last = n;
    };
    struct Teardown {  // (declared before what it times: destroyed last)
        bool on;
        std::chrono::steady_clock::time_point t0;
        ~Teardown() {
            if (on && t0.time_since_epoch().count())
                std::fprintf(stderr, "[rv_stream] %-34s %9.1f ms\n", "teardown", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        }
    } teardown{trace, {}};
    StreamPlan plan;
    DevBuf B;
    const size_t n_seg = stream_segments(n_ops, window_ops);
    plan.segs.resize(n_seg);
    mark("device, kernels");
    std::vector<Segment> &segs = plan.segs;
    std::string perr;
    int plan_rc = RV_OK;
    {
        std::atomic<size_t> planned{0}, next{0};
        std::atomic<bool> abort{false};
        std::atomic<uint64_t> us_compile{0}, us_upload{0}, us_wait{0};  // summed over the workers (RV_TRACE)
        auto us_since = [](std::chrono::steady_clock::time_point t) {
            return (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t).count();
        };
        const int device = g_device;  // (thread_local: the workers get the caller's)
        auto worker = [&]() {
            g_device = device;
            cudaSetDevice(device);
            UploadStage stage;  // (allocated while the planner works on the first segments)
            stage.init();
            g_stage = &stage;
            struct Unset {
                ~Unset() { g_stage = nullptr; }
            } unset;
            for (;;) {
                const size_t k = next.fetch_add(1);
                if (k >= n_seg) return;
                auto t0 = std::chrono::steady_clock::now();
                while (planned.load(std::memory_order_acquire) <= k) {
                    if (abort.load()) return;
                    std::this_thread::sleep_for(std::chrono::microseconds(200));  // (not a spin: the planner needs its core)
                }
                us_wait += us_since(t0);
                Segment &S = segs[k];
                S.c = new (std::nothrow) rv_circuit();
                if (!S.c) {
                    S.rc = RV_E_NOMEM;
                    continue;
                }
                try {
                    t0 = std::chrono::steady_clock::now();
                    S.rc = compile(S.ops.data(), S.ops.size(), 0, S.n_local, S.c->prog, S.err, COMPILE_PROVE_ONLY, &S.io);
                    std::vector<rv_op>().swap(S.ops);
                    us_compile += us_since(t0);
                    t0 = std::chrono::steady_clock::now();
                    if (S.rc == RV_OK && (S.rc = segment_to_device(S, device)) != RV_OK) S.err = g_err;  // (this thread's g_err)
                    us_upload += us_since(t0);
                } catch (const std::bad_alloc &) {
                    S.rc = RV_E_NOMEM;
                    S.err = "out of host memory while compiling a segment";
                }
                std::vector<rv_op>().swap(S.ops);
            }
        };
        const unsigned nt = (unsigned)std::min<size_t>(n_seg, std::max(1u, std::min(16u, std::thread::hardware_concurrency())));
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nt; t++) pool.emplace_back(worker);
        plan_rc = plan_stream(ops, n_ops, gf2_cells, window_ops, plan, perr, &planned);
        if (plan_rc != RV_OK) abort.store(true);
        mark("planner");
        for (auto &t : pool) t.join();
        mark("compile + upload threads' tail");
        if (trace)
            std::fprintf(stderr, "[rv_stream]   %u workers, summed: compile %.1f s, tables to the device %.1f s, waiting for the planner %.1f s\n", nt, us_compile.load() * 1e-6,
                         us_upload.load() * 1e-6, us_wait.load() * 1e-6);
    }
    for (Segment &S : segs)
        if (S.c) B.circuits.push_back(S.c);
    if (plan_rc != RV_OK) return fail(plan_rc, perr);
    for (Segment &S : segs)
        if (S.rc != RV_OK) return fail(S.rc, "segment [" + std::to_string(S.a) + ", " + std::to_string(S.b) + "): " + S.err);
    const uint32_t n_slots = plan.n_slots;
    const uint64_t masks = plan.masks, tot_on = plan.tot_on, tot_pre = plan.tot_pre, tot_inputs = plan.tot_inputs, tot_recon = plan.tot_recon;
    if (n_gf2 < tot_inputs) return fail(RV_E_WITNESS_SHORT, "witness is too short");  // prover.rs:190
    if (tot_inputs && !wit_gf2) return fail(RV_E_ARG, "witness pointer is NULL");
    if (tot_on / 1024 >= 0xFFFFFFF0ull || masks / 128 >= 0xFFFFFFF0ull) return fail(RV_E_UNSUPPORTED, "circuit too large for the 32-bit block counters of the streaming path");
    size_t max_rows = 1, max_vals = 1, max_leaves = 1, max_on = 1, max_pre = 1, max_masks = 1;
    bool any_vm = false;
    for (Segment &S : segs) {
        const Program &P = S.c->prog;
        max_rows = std::max<size_t>(max_rows, P.n_rows);
        max_masks = std::max<size_t>(max_masks, P.n_masks);
        max_vals = std::max<size_t>(max_vals, (size_t)P.n_vals + 1);
        max_leaves = std::max<size_t>(max_leaves, (size_t)S.n_in + S.n_imp);
        max_on = std::max<size_t>(max_on, (size_t)(S.on0 % 1024) + S.n_on);
        max_pre = std::max<size_t>(max_pre, (size_t)(S.pre0 % 1024) + S.n_pre);
        any_vm |= linear_uses_vm(S.c->dev);
    }

    // ---- 3. window buffers, cell file, hash state ----
    constexpr uint32_t NPI = RV_PACKED_REPS, NREPS = RV_TOTAL_REPS;
    const uint32_t nslices = 2 * NPI;
    const size_t pitch_on = round_up(max_on, 2048), pitch_pre = round_up(max_pre, 2048), pitch_fresh = round_up(max_masks + 128, 128);
    const uint32_t tot_chunks_on = tot_on == 0 ? 1 : (uint32_t)((tot_on + 1023) / 1024), tot_chunks_pre = tot_pre == 0 ? 1 : (uint32_t)((tot_pre + 1023) / 1024);
    const ProofLayout L{(uint32_t)(tot_recon / 8 + 1), (uint32_t)(tot_pre / 8 + 1), (uint32_t)(tot_inputs / 8 + 1)};
    const size_t plen = L.total(), tail_off = round_up(plen, 16);
    uint64_t *d_rows = nullptr, *d_fresh = nullptr, *d_cell_rows = nullptr;
    uint8_t *d_vals = nullptr, *d_leaf = nullptr, *d_on = nullptr, *d_pre = nullptr, *d_cell_vals = nullptr, *d_carry_on = nullptr, *d_carry_pre = nullptr, *d_seeds = nullptr,
            *d_pkeys = nullptr, *d_on_hash = nullptr, *d_rep_hash = nullptr, *d_omit = nullptr, *d_proof = nullptr;
    uint32_t *d_cv_on = nullptr, *d_cv_pre = nullptr, *d_rk = nullptr, *d_zconst = nullptr, *d_cv_scratch = nullptr;
    uint16_t *d_rank = nullptr;
    int rc;
    if ((rc = B.alloc(&d_rows, max_rows * NPI)) || (any_vm && (rc = B.alloc(&d_fresh, pitch_fresh * NPI))) || (rc = B.alloc(&d_cell_rows, (size_t)std::max(n_slots, 1u) * NPI)) ||
        (rc = B.alloc(&d_cell_vals, std::max(n_slots, 1u))) || (rc = B.alloc(&d_vals, round_up(max_vals, 16))) || (rc = B.alloc(&d_leaf, max_leaves + 16)) ||
        (rc = B.alloc(&d_on, pitch_on * NREPS)) || (rc = B.alloc(&d_pre, pitch_pre * NREPS)) || (rc = B.alloc(&d_carry_on, (size_t)1024 * NREPS)) ||
        (rc = B.alloc(&d_carry_pre, (size_t)1024 * NREPS)) || (rc = B.alloc(&d_cv_on, (size_t)tot_chunks_on * NREPS * 8)) ||
        (rc = B.alloc(&d_cv_pre, (size_t)tot_chunks_pre * NREPS * 8)) || (rc = B.alloc(&d_seeds, (size_t)NREPS * 16)) || (rc = B.alloc(&d_pkeys, (size_t)NREPS * 128)) ||
        (rc = B.alloc(&d_rk, (size_t)45 * 64 * NPI)) || (rc = B.alloc(&d_on_hash, (size_t)NREPS * 32)) || (rc = B.alloc(&d_rep_hash, (size_t)NREPS * 32)) ||
        (rc = B.alloc(&d_omit, NREPS)) || (rc = B.alloc(&d_rank, NREPS)) || (rc = B.alloc(&d_zconst, 16)) || (rc = B.alloc(&d_proof, tail_off + 64)) ||
        (std::max(tot_chunks_on, tot_chunks_pre) > 2048 && (rc = B.alloc(&d_cv_scratch, (size_t)(std::max(tot_chunks_on, tot_chunks_pre) + 1) / 2 * NREPS * 8))))
        return rc;
    int *d_bad = reinterpret_cast<int *>(d_proof + tail_off);
    uint8_t *d_comm = d_proof + tail_off + 4;
    CU(cudaStreamCreateWithFlags(&B.st, cudaStreamNonBlocking));
    cudaStream_t st = B.st;
    uint8_t h_seeds[RV_TOTAL_REPS * 16];
    if (seeds) memcpy(h_seeds, seeds, sizeof h_seeds);
    else {  // OsRng, src/proof/mod.rs:131-134
        size_t got = 0;
        while (got < sizeof h_seeds) {
            const ssize_t r = getrandom(h_seeds + got, sizeof h_seeds - got, 0);
            if (r <= 0) return fail(RV_E_ARG, "getrandom failed");
            got += (size_t)r;
        }
    }
    {
        uint32_t zc[16];
        memcpy(zc, segs[0].c->z64_empty_hash, 32);
        memcpy(zc + 8, segs[0].c->z64_rep_hash, 32);
        CU(cudaMemcpyAsync(d_zconst, zc, 64, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_seeds, h_seeds, sizeof h_seeds, cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));
    }
    launch_key_setup(d_seeds, nullptr, nullptr, nullptr, nslices, d_pkeys, d_rk, st, d_bad, 1, 0);

    mark("buffers, seeds, keys", st);
    // ---- 4. the two passes ----
    const int n_sms = segs[0].c->n_sms;
    for (int pass = 1; pass <= 2; pass++) {
        if (pass == 2) {
            mark("pass 1 (hashes)", st);
            // every repetition's hash is known: comm, challenge, then the proof's headers, keys, hashes and zeroed vectors
            launch_rep_hash(d_cv_on, tot_chunks_on, d_cv_pre, tot_chunks_pre, d_zconst, NREPS, d_on_hash, d_rep_hash, st, 0xFFFFFFFFu, nullptr, nullptr, nullptr, d_cv_scratch);
            launch_challenge(d_rep_hash, NREPS * 32, d_comm, 0, d_omit, d_rank, 1, st);
            ExtractArgs a;
            a.on = d_on, a.pre = d_pre, a.pitch_on = pitch_on, a.pitch_pre = pitch_pre, a.on_hash = d_on_hash, a.pkeys = d_pkeys, a.seeds = d_seeds, a.comm = d_comm;
            a.omit_of_rep = d_omit, a.rank_of_rep = d_rank, a.z64_empty_hash = d_zconst, a.first_rep = 0, a.nreps = NREPS, a.n_proofs = 1, a.proof_stride = 0;
            a.len_recons = L.len_recons, a.len_corrs = L.len_corrs, a.len_inputs = L.len_inputs, a.proof = d_proof;
            launch_extract(DevProgram(), a, st);
        }
        for (Segment &S : segs) {
            const rv_circuit *c = S.c;
            const Program &P = c->prog;
            const DevProgram &D = c->dev;
            const uint32_t base_on = (uint32_t)(S.on0 % 1024), base_pre = (uint32_t)(S.pre0 % 1024);
            const bool vm = linear_uses_vm(D);
            if (S.n_in) CU(cudaMemcpyAsync(d_leaf, wit_gf2 + S.wit0, S.n_in, cudaMemcpyHostToDevice, st));
            launch_seg_import(S.d_imp_slot, S.n_imp, d_cell_rows, d_cell_vals, NPI, P.n_prg, d_rows, vm ? d_fresh : nullptr, pitch_fresh, false, d_leaf + S.n_in, st);
            if (P.values_wide) {
                DevProgram Dv = D;
                Dv.input_vid = S.d_leaf_ids;
                Dv.n_inputs = S.n_in + S.n_imp;
                launch_values_wide(Dv, P.wlevel_off.data(), d_leaf, d_vals, st);
            } else {
                launch_values(D.lut_steps, D.n_lut_steps, S.d_leaf_ids, d_leaf, 0, S.n_in + S.n_imp, d_vals, 0, D.n_vals, 1, st);
            }
            launch_mask_gen_tt(d_rk, nslices, P.n_prg, d_rows, vm ? d_fresh : nullptr, pitch_fresh, n_sms, st, 0, 1, false, S.mask0);
            CU(cudaMemsetAsync(d_rows + (size_t)P.zero_row() * NPI, 0, (size_t)NPI * 8, st));
            if (D.n_llevels) launch_linear(D, P.xlevel_off.data(), d_rows, NPI, d_fresh, pitch_fresh, st, nullptr);
            launch_items(D, d_rows, NPI, d_vals, nullptr, d_on, pitch_on, d_pre, pitch_pre, d_bad, st, 1, 0, 0, base_on, base_pre);
            // the bytes of the chunk this segment starts in, left behind by the previous one
            if (base_on) CU(cudaMemcpy2DAsync(d_on, pitch_on, d_carry_on, 1024, base_on, NREPS, cudaMemcpyDeviceToDevice, st));
            if (base_pre) CU(cudaMemcpy2DAsync(d_pre, pitch_pre, d_carry_pre, 1024, base_pre, NREPS, cudaMemcpyDeviceToDevice, st));
            const bool last = &S == &segs.back();
            const uint32_t len_on = base_on + S.n_on, len_pre = base_pre + S.n_pre;
            if (pass == 1) {
                const uint32_t nch_on = last ? (tot_on == 0 ? 1 : (len_on + 1023) / 1024) : len_on / 1024;
                const uint32_t nch_pre = last ? (tot_pre == 0 ? 1 : (len_pre + 1023) / 1024) : len_pre / 1024;
                launch_chunk_cv_window(d_on, pitch_on, last ? len_on : nch_on * 1024, nch_on, (uint32_t)((S.on0 - base_on) / 1024), tot_chunks_on, d_cv_on, d_pre, pitch_pre,
                                       last ? len_pre : nch_pre * 1024, nch_pre, (uint32_t)((S.pre0 - base_pre) / 1024), tot_chunks_pre, d_cv_pre, NREPS, st);
            } else {
                SegExtractArgs x;
                x.on = d_on, x.pre = d_pre, x.pitch_on = pitch_on, x.pitch_pre = pitch_pre, x.base_on = base_on, x.base_pre = base_pre;
                x.recon_pos = D.recon_pos, x.input_pos = D.input_pos, x.n_recon = S.n_recon, x.n_corr = S.n_pre, x.n_inputs = S.n_in;
                x.first_recon = S.recon0, x.first_corr = S.pre0, x.first_input = S.wit0, x.omit_of_rep = d_omit, x.rank_of_rep = d_rank;
                x.len_recons = L.len_recons, x.len_corrs = L.len_corrs, x.len_inputs = L.len_inputs, x.proof = d_proof;
                launch_seg_extract(x, NREPS, st);
            }
            if (!last) {
                const uint32_t keep_on = len_on % 1024, keep_pre = len_pre % 1024;
                if (keep_on) CU(cudaMemcpy2DAsync(d_carry_on, 1024, d_on + (len_on - keep_on), pitch_on, keep_on, NREPS, cudaMemcpyDeviceToDevice, st));
                if (keep_pre) CU(cudaMemcpy2DAsync(d_carry_pre, 1024, d_pre + (len_pre - keep_pre), pitch_pre, keep_pre, NREPS, cudaMemcpyDeviceToDevice, st));
                launch_seg_export(S.d_exp_slot, S.d_exp_row, S.d_exp_vref, S.n_exp, d_rows, d_vals, NPI, d_cell_rows, d_cell_vals, st);
            }
            CU(cudaGetLastError());
        }
    }
    mark("pass 2 (openings)", st);
    // ---- 5. the proof ----
    uint8_t tail[64];
    CU(cudaMemcpyAsync(tail, d_proof + tail_off, 36, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    int bad;
    memcpy(&bad, tail, 4);
    if (const int stt = status_of(bad)) return stt;
    uint8_t *p = plen >= PIN_THRESHOLD ? (uint8_t *)pinned_get(plen) : (uint8_t *)malloc(plen);
    if (!p) return fail(RV_E_NOMEM, "out of memory");
    cudaError_t e = cudaMemcpyAsync(p, d_proof, plen, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        rv_free(p);
        return fail(RV_E_CUDA, std::string("proof copy: ") + cudaGetErrorString(e));
    }
    mark("proof to the host");
    teardown.t0 = std::chrono::steady_clock::now();
    *proof = p;
    *proof_len = plen;
    return RV_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
//  Proof::verify (src/proof/mod.rs:224-307)
// ---------------------------------------------------------------------------------------------------------------------
static int verify_on_session(rv_session *s, const uint8_t *proof, size_t proof_len, const PDomain &g, const PDomain &z, int *okay, int *accept) {
    if (s->n_proofs != 1 || s->npi != RV_PACKED_REPS) return fail(RV_E_ARG, "verification runs on a single-proof, full-shard session");
    const rv_circuit *c = s->c;
    CU(cudaSetDevice(c->device));  // the calling thread may have been left on another device (a group's last member)
    const Program &P = c->prog;
    const DevProgram &D = c->dev;
    constexpr uint32_t NON = RV_ONLINE_REPS, NPRE = RV_PREPROCESSING_REPS, NPI_ON = RV_ONLINE_REPS / 8;
    // ---- staging blob: [seeds 256x16][pkeys 256x128][mode 256][omit 256][VOpen x40][on_given 216x32][z_on_given 216x32][proof bytes] ----
    // Z64 (only when the circuit has Z64 ops): [ZOpen x40][z seeds 256x16][z pkeys 256x128][z omit 256] -- the Z64 openings carry their
    // own keys and unopened player (src/proof/mod.rs:262-280); an honest proof repeats the GF(2) ones
    const size_t o_seeds = 0, o_pkeys = o_seeds + 256 * 16, o_mode = o_pkeys + 256 * 128, o_omit = o_mode + 256, o_opens = o_omit + 256,
                 o_ong = o_opens + NON * sizeof(VOpen), o_zg = o_ong + NPRE * 32, o_zopens = round_up(o_zg + NPRE * 32, 16),
                 o_zseeds = o_zopens + NON * sizeof(ZOpen), o_zpkeys = o_zseeds + 256 * 16, o_zomit = o_zpkeys + 256 * 128,
                 o_proof = round_up(o_zomit + 256, 16);
    const size_t need = o_proof + round_up(proof_len, 16);
    // every buffer is guarded on its own: a failed allocation leaves the session usable (and poolable) for the next call
    if (s->vin_bytes < need) {
        s->vin_bytes = 0;
        if (s->h_vin) cudaFreeHost(s->h_vin);
        s->h_vin = nullptr;
        if (cudaMallocHost(&s->h_vin, need) != cudaSuccess) {
            s->h_vin = nullptr;
            cudaGetLastError();
            return fail(RV_E_NOMEM, "pinned host allocation failed");
        }
        int rc = dalloc(s, &s->d_vin, need);
        if (rc) return rc;
        s->vin_bytes = need;
    }
    {
        s->leaf_pitch = round_up((size_t)P.n_inputs + P.n_pre + P.rand_row.size() + 16, 16);
        s->upitch = round_up((size_t)P.n_uvals + 16, 16);
        int rc;
        if (!s->d_leaf_vals && (rc = dalloc(s, &s->d_leaf_vals, s->leaf_pitch * NON))) return rc;
        if (!s->d_uvals && (rc = dalloc(s, &s->d_uvals, s->upitch * NON))) return rc;
        if (s->has_z) {
            s->zleaf_pitch = P.z.leaf_ids.size() + 1;
            s->zupitch = (size_t)P.z.n_vals + 1;
            if (!s->d_zleaf_v && (rc = dalloc(s, &s->d_zleaf_v, s->zleaf_pitch * NON))) return rc;
            if (!s->d_zuvals && (rc = dalloc(s, &s->d_zuvals, s->zupitch * NON))) return rc;
        }
        if (!s->h_vout && cudaMallocHost(&s->h_vout, RV_TOTAL_REPS * 32 + 16) != cudaSuccess) {
            s->h_vout = nullptr;
            cudaGetLastError();
            return fail(RV_E_NOMEM, "pinned host allocation failed");
        }
    }
    CU(cudaStreamSynchronize(s->st));
    uint8_t *h = s->h_vin;
    memset(h, 0, o_proof);
    VOpen *opens = reinterpret_cast<VOpen *>(h + o_opens);
    for (uint32_t k = 0; k < NON; k++) {  // opened repetitions occupy slots 0..39 in proof order (src/proof/mod.rs:234-246)
        const POnline &o = g.online[k];
        memcpy(h + o_pkeys + (size_t)k * 128, proof + o.keys, 128);
        h[o_mode + k] = 1;
        h[o_omit + k] = o.omit;
        const POnline &first = g.online[k & ~7u];  // the pack's first repetition fixes how many elements are unpacked
        opens[k] = VOpen{(uint32_t)(o_proof + o.recons.off), (uint32_t)(o_proof + o.corrs.off), (uint32_t)(o_proof + o.inputs.off),
                         (uint32_t)first.recons.len, (uint32_t)first.corrs.len, (uint32_t)first.inputs.len, o.omit, 0};
    }
    for (uint32_t k = 0; k < NPRE; k++) {  // then the 216 preprocessing repetitions
        memcpy(h + o_seeds + (size_t)(NON + k) * 16, proof + g.pre[k].seed, 16);
        h[o_omit + NON + k] = RV_PLAYERS;
        memcpy(h + o_ong + (size_t)k * 32, proof + g.pre[k].comm_online, 32);
        memcpy(h + o_zg + (size_t)k * 32, proof + z.pre[k].comm_online, 32);
    }
    bool z_own_keys = false;
    if (s->has_z) {
        ZOpen *zopens = reinterpret_cast<ZOpen *>(h + o_zopens);
        for (uint32_t k = 0; k < NON; k++) {
            const POnline &o = z.online[k], &first = z.online[k & ~7u];
            zopens[k] = ZOpen{o_proof + o.recons.off, o_proof + o.corrs.off, o_proof + o.inputs.off, (uint32_t)(first.recons.len / 8),
                              (uint32_t)(first.corrs.len / 8), (uint32_t)(first.inputs.len / 8), (uint32_t)o.recons.len, (uint32_t)o.corrs.len,
                              (uint32_t)o.inputs.len, o.omit, 0};
            memcpy(h + o_zpkeys + (size_t)k * 128, proof + o.keys, 128);
            h[o_zomit + k] = o.omit;
            if (o.omit != g.online[k].omit || memcmp(proof + o.keys, proof + g.online[k].keys, 128) != 0) z_own_keys = true;
        }
        for (uint32_t k = 0; k < NPRE; k++) {
            memcpy(h + o_zseeds + (size_t)(NON + k) * 16, proof + z.pre[k].seed, 16);
            h[o_zomit + NON + k] = RV_PLAYERS;
            if (memcmp(proof + z.pre[k].seed, proof + g.pre[k].seed, 16) != 0) z_own_keys = true;
        }
    }
    memcpy(h + o_proof, proof, proof_len);
    uint8_t *dv = s->d_vin;
    // Everything from the upload of the staging blob to the copy of the repetition hashes is the same sequence for every proof of
    // this length (what differs travels inside the blob): after one eager run it is captured and replayed as one CUDA graph launch.
    // Z64 circuits keep the eager sequence (the Z64 openings may name their own keys: a data-dependent extra launch).
    const bool graphable = !s->has_z && (s->g_verify.exec == nullptr || s->verify_graph_need == need);
    if (!graphable && s->g_verify.exec) {  // another proof length: the captured copy sizes do not fit
        cudaGraphExecDestroy(s->g_verify.exec);
        s->g_verify = rv_session::GraphSlot();
    }
    s->verify_graph_need = need;
    auto body = [&]() -> int {
    CU(cudaMemcpyAsync(dv, h, need, cudaMemcpyHostToDevice, s->st));
    const uint32_t nslices = 2 * s->npi;
    const VOpen *d_opens = reinterpret_cast<const VOpen *>(dv + o_opens);
    {
        Scope k(s, "v.key_setup", 0);
        launch_key_setup(dv + o_seeds, dv + o_pkeys, dv + o_mode, dv + o_omit, nslices, s->d_pkeys, s->d_rk_plain, s->st, s->d_bad);
    }
    {
        Scope k(s, "v.mask_gen", (uint64_t)P.n_masks * s->npi * 8);
        launch_mask_gen_tt(s->d_rk_plain, nslices, P.n_masks, s->d_rows, s->d_fresh_sm, s->pitch_fresh, c->n_sms, s->st, 0, 1, linear_vm_pairs(D, s->npi));
    }
    if (D.n_llevels) {
        Scope k(s, "v.linear", (uint64_t)P.n_lin * s->npi * 8 * 3, 2);
        launch_linear(D, P.xlevel_off.data(), s->d_rows, s->npi, s->d_fresh_sm, s->pitch_fresh, s->st, nullptr);
    }
    {
        Scope k(s, "v.leaves", 0);
        launch_verify_leaves(D, d_opens, dv, s->d_rows, s->npi, NON, s->d_leaf_vals, s->leaf_pitch, s->st);
    }
    {
        Scope k(s, "v.values", (uint64_t)P.vlut_steps.size() * sizeof(LutInstr));
        if (P.verify_wide)
            launch_uvalues_wide(D, P.vwlevel_off.data(), s->d_leaf_vals, s->leaf_pitch, D.n_inputs + D.n_pre + D.n_rand, s->d_uvals, s->upitch, NON, s->st);
        else
            launch_values(D.vlut_steps, D.n_vlut_steps, D.vleaf_ids, s->d_leaf_vals, s->leaf_pitch, D.n_inputs + D.n_pre + D.n_rand, s->d_uvals, s->upitch, D.n_uvals, NON,
                      s->st);
    }
    {
        Scope k(s, "v.items", 0, 3);
        launch_items_pre_range(D, s->d_rows, s->npi, NPI_ON, s->d_pre, s->pitch_pre, s->st);
        launch_verify_items(D, d_opens, dv, s->d_rows, s->npi, NPI_ON, s->d_uvals, s->upitch, s->d_on, s->pitch_on, s->d_pre, s->pitch_pre, s->d_bad, s->st);
    }
    {
        Scope k(s, "v.chunk_cv", 0);
        launch_chunk_cv2(s->d_on, s->pitch_on, P.n_online, s->d_cv_on, NON, s->d_pre, s->pitch_pre, P.n_pre, s->d_cv_pre, s->nreps, s->st);
    }
    if (s->has_z) {  // the Z64 instances of src/proof/mod.rs:262-280 (online) and :247-260 (preprocessing)
        const DevZProgram &DZ = c->zdev;
        const ZOpen *d_zopens = reinterpret_cast<const ZOpen *>(dv + o_zopens);
        if (z_own_keys) {  // a (dishonest) proof whose Z64 openings name other keys: the reference would use them, so do we
            Scope k(s, "v.z.key_setup", 0);
            launch_key_setup(dv + o_zseeds, dv + o_zpkeys, dv + o_mode, dv + o_zomit, nslices, s->d_pkeys, s->d_rk_plain, s->st);
        }
        {
            Scope k(s, "v.z.mask_gen", (uint64_t)P.z.n_masks * s->zrowlen * 8);
            launch_zmask_gen_tt(s->d_rk_plain, (uint32_t)s->zrowlen, P.z.n_masks, s->d_zrows, c->n_sms, s->st);
        }
        if (DZ.n_llevels) {
            Scope k(s, "v.z.linear", 0, DZ.n_llevels);
            launch_zlinear(DZ, P.z.llevel_off.data(), s->d_zrows, (uint32_t)s->zrowlen, s->st);
        }
        {
            Scope k(s, "v.z.leaves", 0);
            launch_zverify_leaves(DZ, d_zopens, dv, s->d_zrows, s->zrowlen, NON, s->d_zleaf_v, s->zleaf_pitch, d_opens, s->d_uvals, s->upitch, D.b2a_urefs, s->st);
        }
        {
            Scope k(s, "v.z.values", 0);
            launch_zvalues(DZ, s->d_zleaf_v, s->zleaf_pitch, s->d_zuvals, s->zupitch, NON, nullptr, nullptr, s->st);
        }
        {
            Scope k(s, "v.z.items", 0, 3);
            launch_zitems_pre_range(DZ, s->d_zrows, s->zrowlen, NON, s->nreps, s->d_rows, s->d_zpre, s->pitch_zpre, s->st);
            launch_zverify_items(DZ, d_zopens, dv, s->d_zrows, s->zrowlen, NON, s->d_zuvals, s->zupitch, s->d_zon, s->pitch_zon, s->d_zpre, s->pitch_zpre,
                                 s->d_bad, s->st);
        }
        {
            Scope k(s, "v.z.chunk_cv", 0);
            launch_chunk_cv2(s->d_zon, s->pitch_zon, (uint32_t)P.z.on_bytes, s->d_zcv_on, NON, s->d_zpre, s->pitch_zpre, (uint32_t)P.z.pre_bytes, s->d_zcv_pre,
                             s->nreps, s->st);
        }
        {
            Scope k(s, "v.z.rep_hash", 0);
            launch_zrep_hash(s->d_zcv_on, s->n_chunks_zon, s->d_zcv_pre, s->n_chunks_zpre, s->nreps, s->d_zon_hash, s->d_zrep, s->st, NON, dv + o_zg);
        }
    }
    {
        Scope k(s, "v.rep_hash", 0);
        launch_rep_hash(s->d_cv_on, s->n_chunks_on, s->d_cv_pre, s->n_chunks_pre, s->d_zconst, s->nreps, s->d_on_hash, s->d_rep_hash, s->st, NON,
                        dv + o_ong, dv + o_zg, s->has_z ? s->d_zrep : nullptr);
    }
    CU(cudaMemcpyAsync(s->h_vout, s->d_rep_hash, RV_TOTAL_REPS * 32, cudaMemcpyDeviceToHost, s->st));
    CU(cudaMemcpyAsync(s->h_vout + RV_TOTAL_REPS * 32, s->d_bad, 4, cudaMemcpyDeviceToHost, s->st));
    CU(cudaGetLastError());
    return RV_OK;
    };
    {
        const int brc = s->has_z ? body() : run_graphed(s, s->g_verify, body);
        if (brc != RV_OK) return brc;
    }
    CU(cudaStreamSynchronize(s->st));
    s->committed = s->opened = false;
    s->ever_committed = true;
    // re-interleave into original repetition order using the challenge derived from the claimed comm (src/proof/mod.rs:292-302)
    uint8_t omit_of_rep[RV_TOTAL_REPS], ordered[RV_TOTAL_REPS * 32];
    host_challenge(proof, omit_of_rep);
    size_t on = 0, pre = NON;
    for (int i = 0; i < RV_TOTAL_REPS; i++) memcpy(ordered + 32 * i, s->h_vout + 32 * (omit_of_rep[i] < RV_PLAYERS ? on++ : pre++), 32);
    uint32_t comm2[8];
    host_hash(ordered, sizeof ordered, comm2);
    int bad;
    memcpy(&bad, s->h_vout + RV_TOTAL_REPS * 32, 4);
    if (okay) *okay = bad ? 0 : 1;
    *accept = memcmp(comm2, proof, 32) == 0 ? 1 : 0;  // src/proof/mod.rs:305-306
    return RV_OK;
}

extern "C" int rv_verify(const rv_circuit *c, const uint8_t *proof, size_t proof_len, int *okay) {
    if (!c || (!proof && proof_len)) return fail(RV_E_ARG, "NULL argument");
    if (!c->prog.has_verify)
        return fail(RV_E_UNSUPPORTED, "this handle has no verifier tables (compiled with RV_COMPILE_PROVE_ONLY, or more than 2^28 ops)");
    if (proof_len < 32) return fail(RV_E_FORMAT, "proof shorter than its commitment");
    PDomain g, z;
    size_t pos = 32;
    if (!parse_domain(proof, proof_len, pos, g) || !parse_domain(proof, proof_len, pos, z) || pos != proof_len)
        return fail(RV_E_FORMAT, "malformed proof bytes");
    if (okay) *okay = 1;
    // check_format, src/proof/mod.rs:110-114,225-230
    if (g.online.size() != RV_ONLINE_REPS || g.pre.size() != RV_PREPROCESSING_REPS || z.online.size() != RV_ONLINE_REPS || z.pre.size() != RV_PREPROCESSING_REPS)
        return 0;
    for (uint32_t k = 0; k < RV_ONLINE_REPS; k++) {
        if (g.online[k].omit >= RV_PLAYERS || z.online[k].omit >= RV_PLAYERS) return fail(RV_E_FORMAT, "omitted player out of range");
        const POnline &first = g.online[k & ~7u], &o = g.online[k];
        // unpack_selected asserts equal lengths (src/algebra/gf2/share.rs:158-164); ReconGF2::unpack indexes every lane up to the
        // first lane's length (src/algebra/gf2/recon.rs:241-259)
        if (o.recons.len != first.recons.len || o.corrs.len < first.corrs.len || o.inputs.len < first.inputs.len)
            return fail(RV_E_FORMAT, "ragged packed lengths inside a pack of 8 openings");
    }
    rv_session *s = nullptr;
    {
        std::lock_guard<std::mutex> lk(c->pool_mu);
        if (!c->pool.empty()) {
            s = c->pool.back();
            c->pool.pop_back();
        }
    }
    int rc = RV_OK, accept = 0;
    if (!s && (rc = rv_session_create(c, 0, RV_PACKED_REPS, &s))) return rc;
    rc = verify_on_session(s, proof, proof_len, g, z, okay, &accept);
    if (rc == RV_E_CUDA) {
        rv_session_free(s);
        return rc;
    }
    std::lock_guard<std::mutex> lk(c->pool_mu);
    if (c->pool.size() < 32) c->pool.push_back(s);
    else rv_session_free(s);
    return rc == RV_OK ? accept : rc;
}

// ---------------------------------------------------------------------------------------------------------------------
//  One-shot forms with the reference's argument shape (src/proof/mod.rs:119-124,224): the circuit arrives as an op list with
//  every call, so compiled circuits are kept in a small content-addressed cache (128-bit hash of the op bytes + wire counts +
//  device).  Proving needs no verifier tables and compiles without them; a later verification of the same circuit upgrades
//  the entry.  Entries in use are never evicted; idle ones leave in least-recently-used order.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
struct Hash128 {
    uint64_t a, b;
    bool operator==(const Hash128 &o) const { return a == o.a && b == o.b; }
};
inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
inline uint64_t mix64(uint64_t x) {
    x ^= x >> 32;
    x *= 0xd6e8feb86659fd93ull;
    x ^= x >> 32;
    x *= 0xd6e8feb86659fd93ull;
    return x ^ (x >> 32);
}
// four independent multiply-rotate lanes over 32-byte stripes (memory-speed on one core), folded to 128 bits
Hash128 hash_bytes(const void *data, size_t len, uint64_t seed) {
    const uint8_t *p = static_cast<const uint8_t *>(data);
    uint64_t h[4] = {seed ^ 0x9e3779b97f4a7c15ull, seed + 0xc2b2ae3d27d4eb4full, ~seed * 0x165667b19e3779f9ull, seed ^ 0x27d4eb2f165667c5ull};
    size_t n = len / 32;
    for (size_t i = 0; i < n; i++, p += 32) {
        uint64_t w[4];
        memcpy(w, p, 32);
        for (int k = 0; k < 4; k++) h[k] = rotl64(h[k] ^ (w[k] * 0x9fb21c651e98df25ull), 29) * 0xff51afd7ed558ccdull + k;
    }
    uint8_t tail[32] = {0};
    memcpy(tail, p, len % 32);
    uint64_t w[4];
    memcpy(w, tail, 32);
    for (int k = 0; k < 4; k++) h[k] = rotl64(h[k] ^ (w[k] * 0x9fb21c651e98df25ull), 29) * 0xff51afd7ed558ccdull + k;
    Hash128 r;
    r.a = mix64(h[0] ^ rotl64(h[1], 17) ^ len) ^ mix64(h[2] + rotl64(h[3], 41));
    r.b = mix64(h[1] ^ rotl64(h[2], 23) ^ (len * 0x9e3779b97f4a7c15ull)) ^ mix64(h[3] + rotl64(h[0], 37));
    return r;
}
struct CacheEntry {
    Hash128 key;
    size_t n_ops, z64_cells, gf2_cells;
    int device;
    rv_circuit *c;
    int in_use;
    uint64_t last_use;
    bool dead;  // dropped by rv_circuit_cache_clear while in use: invisible to lookups, freed by its last user
};
std::mutex g_cache_mu;
std::vector<CacheEntry> g_cache;
uint64_t g_cache_clock = 0, g_cache_hits = 0, g_cache_misses = 0;
size_t g_cache_max = 8;

void cache_release(rv_circuit *c) {
    bool free_it = false;
    {
        std::lock_guard<std::mutex> g(g_cache_mu);
        for (size_t i = 0; i < g_cache.size(); i++)
            if (g_cache[i].c == c) {
                if (--g_cache[i].in_use == 0 && g_cache[i].dead) {
                    g_cache.erase(g_cache.begin() + i);
                    free_it = true;
                }
                break;
            }
    }
    if (free_it) rv_circuit_free(c);
}

// Out of device memory in a one-shot call: the idle entries of the cache (compiled tables and the pooled sessions of circuits
// nobody is using right now -- a 10^8-gate circuit keeps ~110 GB) are the first thing to give back.  Returns how many went.
size_t cache_drop_idle() {
    std::vector<rv_circuit *> drop;
    {
        std::lock_guard<std::mutex> g(g_cache_mu);
        for (size_t i = 0; i < g_cache.size();) {
            if (g_cache[i].in_use == 0) {
                drop.push_back(g_cache[i].c);
                g_cache.erase(g_cache.begin() + i);
            } else i++;
        }
    }
    for (rv_circuit *d : drop) rv_circuit_free(d);
    return drop.size();
}

// One-shot proofs of big circuits (rv_proof_new, below): keys of circuits seen once and proved in streaming mode.
std::vector<Hash128> g_streamed_once;
size_t g_oneshot_stream_min = (size_t)1 << 24;

// The cache key of a circuit.  One core hashes ~6 GB/s: 0.4 s for the 2.4 GB of a 10^8-gate op list, with every call -- so op
// lists of more than KEY_SLICE bytes are hashed slice by slice on several threads (slice boundaries depend on the length only)
// and the slice hashes are hashed once more.
Hash128 circuit_key(const rv_op *ops, size_t n_ops, size_t z64_cells, size_t gf2_cells) {
    const uint64_t seed = (uint64_t)z64_cells * 0x100000001b3ull ^ gf2_cells;
    const size_t bytes = n_ops * sizeof(rv_op);
    constexpr size_t KEY_SLICE = (size_t)32 << 20;
    if (bytes <= KEY_SLICE) return hash_bytes(ops, bytes, seed);
    const size_t n_slices = (bytes + KEY_SLICE - 1) / KEY_SLICE;
    std::vector<Hash128> part(n_slices);
    std::atomic<size_t> next{0};
    auto work = [&]() {
        for (size_t k; (k = next.fetch_add(1)) < n_slices;)
            part[k] = hash_bytes((const uint8_t *)ops + k * KEY_SLICE, std::min(KEY_SLICE, bytes - k * KEY_SLICE), seed + 0x9e3779b97f4a7c15ull * (k + 1));
    };
    const unsigned nt = (unsigned)std::min<size_t>(n_slices, std::max(1u, std::min(16u, std::thread::hardware_concurrency())));
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; t++) th.emplace_back(work);
    work();
    for (std::thread &t : th) t.join();
    return hash_bytes(part.data(), n_slices * sizeof(Hash128), seed ^ bytes);
}
// true if this circuit is neither cached nor has been through rv_proof_new before (and remembers that it has now)
bool first_sight(const Hash128 &key, size_t n_ops) {
    std::lock_guard<std::mutex> g(g_cache_mu);
    for (const CacheEntry &e : g_cache)
        if (!e.dead && e.key == key && e.n_ops == n_ops) return false;
    for (const Hash128 &k : g_streamed_once)
        if (k == key) return false;
    if (g_streamed_once.size() >= 64) g_streamed_once.erase(g_streamed_once.begin());
    g_streamed_once.push_back(key);
    return true;
}

// Looks the circuit up (compiling it on a miss); the returned handle stays valid until cache_release.
int cache_acquire(const rv_op *ops, size_t n_ops, size_t z64_cells, size_t gf2_cells, bool need_verify, rv_circuit **out, const Hash128 *known_key = nullptr) {
    if (n_ops && !ops) return fail(RV_E_ARG, "ops is NULL");
    const Hash128 key = known_key ? *known_key : circuit_key(ops, n_ops, z64_cells, gf2_cells);
    const int device = rv_device_count() ? g_device : -1;
    {
        std::lock_guard<std::mutex> g(g_cache_mu);
        for (CacheEntry &e : g_cache)
            if (!e.dead && e.key == key && e.n_ops == n_ops && e.z64_cells == z64_cells && e.gf2_cells == gf2_cells && e.device == device &&
                (!need_verify || e.c->prog.has_verify || n_ops > VERIFY_MAX_OPS)) {
                e.in_use++;
                e.last_use = ++g_cache_clock;
                g_cache_hits++;
                *out = e.c;
                return RV_OK;
            }
        g_cache_misses++;
    }
    rv_circuit *c = nullptr;
    const int rc = rv_circuit_compile_ex(ops, n_ops, z64_cells, gf2_cells, need_verify ? 0 : RV_COMPILE_PROVE_ONLY, &c);
    if (rc) return rc;
    std::vector<rv_circuit *> drop;
    {
        std::lock_guard<std::mutex> g(g_cache_mu);
        // a prove-only twin of a circuit now compiled with verifier tables is superseded; then evict idle entries, oldest first
        for (size_t i = 0; i < g_cache.size();) {
            CacheEntry &e = g_cache[i];
            const bool twin = e.key == key && e.n_ops == n_ops && e.z64_cells == z64_cells && e.gf2_cells == gf2_cells && e.device == device;
            if (twin && e.in_use == 0) {
                drop.push_back(e.c);
                g_cache.erase(g_cache.begin() + i);
            } else i++;
        }
        while (g_cache.size() + 1 > g_cache_max) {
            size_t victim = g_cache.size();
            for (size_t i = 0; i < g_cache.size(); i++)
                if (g_cache[i].in_use == 0 && (victim == g_cache.size() || g_cache[i].last_use < g_cache[victim].last_use)) victim = i;
            if (victim == g_cache.size()) break;  // everything is in use: grow for now
            drop.push_back(g_cache[victim].c);
            g_cache.erase(g_cache.begin() + victim);
        }
        g_cache.push_back(CacheEntry{key, n_ops, z64_cells, gf2_cells, device, c, 1, ++g_cache_clock, false});
    }
    for (rv_circuit *d : drop) rv_circuit_free(d);
    *out = c;
    return RV_OK;
}
}  // namespace

extern "C" void rv_circuit_cache_clear(void) {
    std::vector<rv_circuit *> drop;
    {
        std::lock_guard<std::mutex> g(g_cache_mu);
        for (size_t i = 0; i < g_cache.size();) {
            if (g_cache[i].in_use == 0) {
                drop.push_back(g_cache[i].c);
                g_cache.erase(g_cache.begin() + i);
            } else g_cache[i++].dead = true;  // freed by its last user
        }
    }
    for (rv_circuit *d : drop) rv_circuit_free(d);
}
extern "C" void rv_circuit_cache_limit(size_t max_entries) {
    std::vector<rv_circuit *> drop;
    {
        std::lock_guard<std::mutex> g(g_cache_mu);
        g_cache_max = max_entries ? max_entries : 1;
        while (g_cache.size() > g_cache_max) {
            size_t victim = g_cache.size();
            for (size_t i = 0; i < g_cache.size(); i++)
                if (g_cache[i].in_use == 0 && (victim == g_cache.size() || g_cache[i].last_use < g_cache[victim].last_use)) victim = i;
            if (victim == g_cache.size()) break;
            drop.push_back(g_cache[victim].c);
            g_cache.erase(g_cache.begin() + victim);
        }
    }
    for (rv_circuit *d : drop) rv_circuit_free(d);
}
extern "C" void rv_circuit_cache_stats(uint64_t *hits, uint64_t *misses, size_t *entries) {
    std::lock_guard<std::mutex> g(g_cache_mu);
    if (hits) *hits = g_cache_hits;
    if (misses) *misses = g_cache_misses;
    if (entries) *entries = g_cache.size();
}

extern "C" void rv_oneshot_streaming_min(size_t n_ops) {
    std::lock_guard<std::mutex> g(g_cache_mu);
    g_oneshot_stream_min = n_ops;
}

extern "C" int rv_proof_new(const rv_op *ops, size_t n_ops, const uint8_t *wit_gf2, size_t n_gf2, const uint64_t *wit_z64, size_t n_z64,
                            size_t z64_cells, size_t gf2_cells, const uint8_t *seeds, uint8_t **proof, size_t *proof_len) {
    // A big circuit seen for the first time: compiling it for residency is one thread walking the whole op list (10^8 flat gates:
    // 7 s before a 0.2 s proof), the streaming prover compiles its segments on all host cores while it plans and uploads, and
    // the bytes are the same.  The second call with the same circuit compiles it (from then on: cache hits, resident proofs).
    size_t stream_min;
    {
        std::lock_guard<std::mutex> g(g_cache_mu);
        stream_min = g_oneshot_stream_min;
    }
    if (n_ops && !ops) return fail(RV_E_ARG, "ops is NULL");
    const Hash128 key = circuit_key(ops, n_ops, z64_cells, gf2_cells);
    if (stream_min && n_ops >= stream_min && proof && proof_len && rv_device_count() > 0 && first_sight(key, n_ops)) {
        const int src = rv_prove_streaming(ops, n_ops, z64_cells, gf2_cells, wit_gf2, n_gf2, wit_z64, n_z64, seeds, 0, proof, proof_len);
        if (src != RV_E_UNSUPPORTED && src != RV_E_NOMEM) return src;  // (Z64 / Random / B2A, or no room for the window: the resident path decides)
    }
    rv_circuit *c = nullptr;
    int rc = cache_acquire(ops, n_ops, z64_cells, gf2_cells, false, &c, &key);
    if (rc == RV_E_NOMEM && cache_drop_idle()) rc = cache_acquire(ops, n_ops, z64_cells, gf2_cells, false, &c, &key);
    if (rc) return rc;
    rc = rv_prove(c, wit_gf2, n_gf2, wit_z64, n_z64, seeds, proof, proof_len);
    if (rc == RV_E_NOMEM && cache_drop_idle()) rc = rv_prove(c, wit_gf2, n_gf2, wit_z64, n_z64, seeds, proof, proof_len);  // (c itself is in use: kept)
    cache_release(c);
    return rc;
}

extern "C" int rv_proof_verify_ex(const rv_op *ops, size_t n_ops, size_t z64_cells, size_t gf2_cells, const uint8_t *proof, size_t proof_len, int *okay) {
    rv_circuit *c = nullptr;
    int rc = cache_acquire(ops, n_ops, z64_cells, gf2_cells, true, &c);
    if (rc == RV_E_NOMEM && cache_drop_idle()) rc = cache_acquire(ops, n_ops, z64_cells, gf2_cells, true, &c);
    if (rc) return rc;
    rc = rv_verify(c, proof, proof_len, okay);
    if (rc == RV_E_NOMEM && cache_drop_idle()) rc = rv_verify(c, proof, proof_len, okay);
    cache_release(c);
    return rc;
}

extern "C" int rv_proof_verify(const rv_op *ops, size_t n_ops, size_t z64_cells, size_t gf2_cells, const uint8_t *proof, size_t proof_len) {
    int okay = 1;
    const int rc = rv_proof_verify_ex(ops, n_ops, z64_cells, gf2_cells, proof, proof_len, &okay);
    return rc == 1 ? (okay ? 1 : 0) : rc;  // strict: see the header
}
