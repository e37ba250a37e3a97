// tests/hostsim/hostsim.cpp -- TEST INFRASTRUCTURE ONLY.  Replays the device pipeline on the CPU, thread by thread, by
// calling the very same __host__ __device__ bodies the CUDA kernels call (csrc/rv_planes.cuh, rv_aes_bs.cuh,
// rv_blake3.cuh) on the tables produced by the product's circuit compiler.  It lets the CPU-only test-suite pin the
// kernels' arithmetic, bit orders and proof layout against the oracle before any GPU time is spent.  The product never
// links this file.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../reverie_b200/csrc/rv_planes.cuh"
#include "../../reverie_b200/csrc/rv_zplanes.cuh"
#include "../../reverie_b200/csrc/rv_stream_plan.h"

using namespace rv;

static std::string g_err;
extern "C" const char *hs_last_error() { return g_err.c_str(); }

// the pairwise-with-carry tree of k_rep_hash / tree_reduce, serial
static void tree_root(std::vector<uint32_t> &cvs, uint32_t n) {
    while (n > 1) {
        const uint32_t pairs = n / 2, outn = (n + 1) / 2;
        const bool root = n == 2;
        std::vector<uint32_t> nxt(outn * 8);
        for (uint32_t p = 0; p < outn; p++) {
            if (p < pairs) b3_parent_cv(&cvs[2 * p * 8], &cvs[(2 * p + 1) * 8], root, &nxt[p * 8]);
            else memcpy(&nxt[p * 8], &cvs[2 * p * 8], 32);
        }
        cvs.swap(nxt);
        n = outn;
    }
}

static void stream_hash(const uint8_t *data, uint32_t len, uint32_t out[8]) {
    const uint32_t n_chunks = len == 0 ? 1 : (len + 1023) / 1024;
    std::vector<uint32_t> cvs(n_chunks * 8);
    std::vector<uint32_t> buf(256);
    for (uint32_t c = 0; c < n_chunks; c++) {
        const uint32_t off = c * 1024, clen = std::min<uint32_t>(1024, len - off);
        memset(buf.data(), 0, 1024);
        if (clen) memcpy(buf.data(), data + off, clen);
        b3_chunk_cv(buf.data(), clen, c, n_chunks == 1, &cvs[c * 8]);
    }
    tree_root(cvs, n_chunks);
    memcpy(out, cvs.data(), 32);
}

extern "C" void hs_blake3(const uint8_t *data, uint32_t len, uint8_t out[32]) {
    uint32_t h[8];
    stream_hash(data, len, h);
    memcpy(out, h, 32);
}

extern "C" void hs_aes128_encrypt(const uint8_t key[16], const uint8_t in[16], uint8_t out[16]) {
    uint32_t k[4], rk[44], i4[4], o4[4];
    memcpy(k, key, 16);
    memcpy(i4, in, 16);
    aes128_expand_key(k, rk);
    aes128_encrypt_block(rk, i4, o4);
    memcpy(out, o4, 16);
}

// The T-table rounds the GPU mask generators run (tt_aes128_encrypt), with the tables built exactly like the kernels build
// them (te0_entry of the netlist S-box, rotations for Te1..Te3) but without the per-bank replication.
extern "C" void hs_tt_aes128_encrypt(const uint8_t key[16], const uint8_t in[16], uint8_t out[16]) {
    static uint32_t te[4][256];
    static bool built = false;
    if (!built) {
        for (uint32_t x = 0; x < 256; x++) {
            const uint32_t t0 = te0_entry(sub_word(x) & 0xff);
            te[0][x] = t0;
            te[1][x] = (t0 << 8) | (t0 >> 24);
            te[2][x] = (t0 << 16) | (t0 >> 16);
            te[3][x] = (t0 << 24) | (t0 >> 8);
        }
        built = true;
    }
    uint32_t k[4], rk[44], i4[4], o4[4];
    memcpy(k, key, 16);
    memcpy(i4, in, 16);
    aes128_expand_key(k, rk);
    tt_aes128_encrypt(rk, i4[0], i4[1], i4[2], i4[3], [](int t, uint32_t w, int b) { return te[t][(w >> (8 * b)) & 0xff]; }, o4);
    memcpy(out, o4, 16);
}

// rows of the share tensor for `npi` packed instances starting at first_instance: the K1 + K2 kernels
struct SimKeys {
    std::vector<uint32_t> ks, lane_mask;
};
static void gen_masks(const uint8_t *seeds_shard, const uint8_t *pkeys_in, const uint8_t *mode, const uint8_t *omit, uint32_t npi,
                      uint32_t n_masks, std::vector<uint64_t> &rows, std::vector<uint8_t> &pkeys, SimKeys *keep = nullptr) {
    const uint32_t nslices = 2 * npi;
    std::vector<uint32_t> ks((size_t)nslices * 1408, 0), lane_mask(nslices, 0);
    pkeys.assign((size_t)npi * 8 * 128, 0);
    for (uint32_t w = 0; w < nslices; w++)
        for (uint32_t q = 0; q < 32; q++) {  // lane q of the warp
            uint32_t rk[44];
            const bool active = key_setup_stream(slice_rep(w, q), slice_player(q), seeds_shard, pkeys_in, mode, omit, pkeys.data(), rk);
            if (active) lane_mask[w] |= 1u << q;
            for (int qq = 0; qq < 44; qq++)
                for (int i = 0; i < 32; i++)
                    if ((rk[qq] >> i) & 1) ks[(size_t)w * 1408 + qq * 32 + i] |= 1u << q;  // __ballot_sync
        }
    uint32_t *rows32 = reinterpret_cast<uint32_t *>(rows.data());
    const uint64_t n_blocks = ((uint64_t)n_masks + 127) / 128;
    for (uint32_t w = 0; w < nslices; w++)
        for (uint64_t j = 0; j < n_blocks; j++) {
            uint32_t s[128];
            const uint32_t *k = &ks[(size_t)w * 1408];
            bs_aes128_ctr_block(j, [k](int round, int plane) { return k[round * 128 + plane]; }, s);
            for (int p = 0; p < 128; p++) {
                const uint64_t i = plane_to_mask_index(j, p);
                if (i < n_masks) rows32[i * nslices + w] = s[p] & lane_mask[w];
            }
        }
    if (keep) {
        keep->ks.swap(ks);
        keep->lane_mask.swap(lane_mask);
    }
}

// k_zmask_gen + k_zlinear_level
static void gen_zrows(const SimKeys &K, uint32_t npi, const ZProgram &Z, std::vector<uint64_t> &zrows) {
    const uint32_t nslices = 2 * npi;
    const size_t rowlen = (size_t)64 * npi;
    zrows.assign((size_t)Z.n_rows * rowlen, 0);
    const uint64_t n_blocks = ((uint64_t)Z.n_masks + 1) / 2;
    for (uint32_t w = 0; w < nslices; w++)
        for (uint64_t j = 0; j < n_blocks; j++) {
            uint32_t s[128];
            const uint32_t *k = &K.ks[(size_t)w * 1408];
            bs_aes128_ctr_block(j, [k](int round, int plane) { return k[round * 128 + plane]; }, s);
            for (int h = 0; h < 2; h++) {
                if (2 * j + h >= Z.n_masks) break;
                uint32_t lo[32], hi[32];
                planes_to_mask_words(s, K.lane_mask[w], h, lo, hi);
                uint64_t *dst = &zrows[(size_t)(2 * j + h) * rowlen + zrow_index(w, 0)];
                for (int q = 0; q < 32; q++) dst[q] = ((uint64_t)hi[q] << 32) | lo[q];
            }
        }
    for (const ZLin &n : Z.lin)
        for (size_t e = 0; e < rowlen; e++) zrows[(size_t)n.dst * rowlen + e] = n.ca * zrows[(size_t)n.a * rowlen + e] + n.cb * zrows[(size_t)n.b * rowlen + e];
}

static void z_values(const ZProgram &Z, const uint64_t *leaves, uint64_t *v, const uint8_t *gvals, const uint32_t *b2a_vrefs) {
    v[0] = 0;
    for (size_t k = 0; k < Z.leaf_ids.size(); k++) v[Z.leaf_ids[k]] = leaves[k];
    for (const ZInstr &in : Z.vprog) v[in.dst] = z_exec(in, v, gvals, b2a_vrefs);
}

extern "C" void hs_gf2_masks(const uint8_t *seeds8, const uint8_t *omit8, uint64_t *out, uint32_t n) {
    std::vector<uint64_t> rows((size_t)n + 1, 0);
    std::vector<uint8_t> pkeys;
    uint8_t om[8];
    for (int i = 0; i < 8; i++) om[i] = omit8 ? omit8[i] : 8;
    gen_masks(seeds8, nullptr, nullptr, om, 1, n, rows, pkeys);
    memcpy(out, rows.data(), (size_t)n * 8);
}

// Proof::new on the CPU through the kernel bodies.  rep_hashes (optional): 256*32 bytes.
extern "C" int hs_prove(const rv_op *ops, size_t n_ops, size_t z64_cells, size_t gf2_cells, const uint8_t *wit, size_t n_wit,
                        const uint64_t *wit_z, size_t n_wit_z, const uint8_t *seeds, uint8_t **proof, size_t *proof_len, uint8_t *rep_hashes) {
    Program P;
    int rc = compile(ops, n_ops, z64_cells, gf2_cells, P, g_err);
    if (rc) return rc;
    if (n_wit < P.n_inputs || n_wit_z < P.z.n_inputs) return RV_E_WITNESS_SHORT;
    const ZProgram &Z = P.z;
    const uint32_t npi = 32, nreps = 256;
    // K1 + K2
    std::vector<uint64_t> rows((size_t)P.n_rows * npi, 0);
    std::vector<uint8_t> pkeys;
    SimKeys K;
    gen_masks(seeds, nullptr, nullptr, nullptr, npi, P.n_masks, rows, pkeys, &K);
    // K0
    std::vector<uint8_t> vals(P.n_vals, 0);
    for (size_t k = 0; k < P.n_inputs; k++) vals[P.input_vid[k]] = wit[k] & 1;
    for (const VGate &g : P.vgates) {
        const uint32_t a = vals[g.a >> 1] ^ (g.a & 1), b = vals[g.b >> 1] ^ (g.b & 1);
        vals[g.dst] = (uint8_t)((g.op ? (a & b) : (a ^ b)) & 1);
    }
    // K3
    for (const XGate &g : P.xgates)
        for (uint32_t pi = 0; pi < npi; pi++) {
            uint64_t v = 0;
            for (int k = 0; k < 6; k++) v ^= rows[(size_t)g.in[k] * npi + pi];
            rows[(size_t)g.dst * npi + pi] = v;
        }
    // tainted plane (k_tainted): per-repetition plaintext of everything that depends on Random / B2A fresh wires
    std::vector<uint64_t> tvals((size_t)P.n_tvals * npi + 1, 0);
    for (const TGate &g : P.tgates)
        for (uint32_t pi = 0; pi < npi; pi++) tvals[(size_t)g.dst * npi + pi] = tainted_eval(g, rows.data(), npi, pi, vals.data(), tvals.data());
    // Z64: masks, value plane, item plane, stream hashes
    const size_t rowlen = (size_t)64 * npi;
    const size_t pitch_zon = (std::max<size_t>(Z.on_bytes, 1) + 63) / 64 * 64, pitch_zpre = (std::max<size_t>(Z.pre_bytes, 1) + 63) / 64 * 64;
    std::vector<uint8_t> zon, zpre;
    std::vector<uint32_t> zon_hash(nreps * 8), zrep(nreps * 8);
    int zbad = 0;
    if (Z.any()) {
        std::vector<uint64_t> zrows;
        gen_zrows(K, npi, Z, zrows);
        std::vector<uint64_t> leaves(Z.leaf_ids.size(), 0), zvals((size_t)Z.n_vals + 1, 0);
        for (size_t k = 0; k < Z.n_inputs; k++) leaves[k] = wit_z[k];
        z_values(Z, leaves.data(), zvals.data(), vals.data(), P.b2a_vrefs.data());
        zon.assign(pitch_zon * nreps, 0);
        zpre.assign(pitch_zpre * nreps, 0);
        for (uint32_t rep = 0; rep < nreps; rep++) {
            for (const ZItem &it : Z.items)  // k_zitems_online: both streams in one pass
                z_prover_online(it, zrows.data(), rowlen, rep, zvals.data(), &zon[(size_t)rep * pitch_zon], &zbad, &zpre[(size_t)rep * pitch_zpre], rows.data(), npi);
            uint32_t h_pre[8];
            stream_hash(&zon[(size_t)rep * pitch_zon], (uint32_t)Z.on_bytes, &zon_hash[rep * 8]);
            stream_hash(&zpre[(size_t)rep * pitch_zpre], (uint32_t)Z.pre_bytes, h_pre);
            b3_hash64(h_pre, &zon_hash[rep * 8], &zrep[rep * 8]);
        }
    }
    if (zbad) return RV_E_WITNESS_INVALID;
    // K4
    const size_t pitch_on = (std::max<size_t>(P.n_online, 1) + 63) / 64 * 64, pitch_pre = (std::max<size_t>(P.n_pre, 1) + 63) / 64 * 64;
    std::vector<uint8_t> on(pitch_on * nreps, 0), pre(pitch_pre * nreps, 0);
    std::vector<uint32_t> mul_pos;
    for (uint32_t t = 0; t < P.n_online; t++)
        if (P.items[t].kind == ITEM_MUL) mul_pos.push_back(t);
    int bad = 0;
    for (uint32_t pi = 0; pi < npi; pi++) {
        for (uint64_t t0 = 0; t0 < P.n_online; t0 += 8) {
            uint64_t W[8], out[8];
            for (int i = 0; i < 8; i++) W[i] = (t0 + i < P.n_online) ? prover_online_word(P.items[t0 + i], rows.data(), npi, pi, vals.data(), tvals.data(), &bad) : 0;
            words_to_stream_bytes(W, out);
            for (int r = 0; r < 8; r++) memcpy(&on[(size_t)(8 * pi + r) * pitch_on + t0], &out[r], 8);
        }
        for (uint64_t j0 = 0; j0 < P.n_pre; j0 += 8) {
            uint64_t W[8], out[8];
            for (int i = 0; i < 8; i++) W[i] = (j0 + i < P.n_pre) ? pre_word(P.items[mul_pos[j0 + i]], rows.data(), npi, pi) : 0;
            words_to_stream_bytes(W, out);
            for (int r = 0; r < 8; r++) memcpy(&pre[(size_t)(8 * pi + r) * pitch_pre + j0], &out[r], 8);
        }
    }
    if (bad) return RV_E_WITNESS_INVALID;
    // K5
    uint32_t empty[8], zrep0[8];
    b3_chunk_cv(nullptr, 0, 0, true, empty);
    b3_hash64(empty, empty, zrep0);
    std::vector<uint32_t> on_hash(nreps * 8), rep_hash(nreps * 8);
    for (uint32_t r = 0; r < nreps; r++) {
        uint32_t h_pre[8];
        stream_hash(&on[(size_t)r * pitch_on], P.n_online, &on_hash[r * 8]);
        stream_hash(&pre[(size_t)r * pitch_pre], P.n_pre, h_pre);
        rep_join(&on_hash[r * 8], h_pre, Z.any() ? &zrep[r * 8] : zrep0, &rep_hash[r * 8]);
    }
    if (rep_hashes) memcpy(rep_hashes, rep_hash.data(), nreps * 32);
    // K6
    uint32_t comm[8], m[16];
    stream_hash(reinterpret_cast<const uint8_t *>(rep_hash.data()), nreps * 32, comm);
    challenge_block(comm, m);
    uint8_t omit[256];
    uint16_t rank[256];
    memset(omit, RV_PLAYERS, sizeof omit);
    int distinct = 0;
    for (uint64_t t = 0; distinct < RV_ONLINE_REPS; t++) {
        uint32_t o[16];
        challenge_xof_block(m, t, o);
        challenge_consume(o, omit, &distinct);
    }
    uint16_t n_on = 0, n_pre = 0;
    for (int i = 0; i < 256; i++) rank[i] = omit[i] < RV_PLAYERS ? n_on++ : n_pre++;
    // K7
    ProofLayout L{(uint32_t)(P.recon_pos.size() / 8 + 1), P.n_pre / 8 + 1, (uint32_t)(P.n_inputs / 8 + 1)};
    if (Z.any()) {
        L.len_zrecons = (uint32_t)(8 * Z.recon_off.size());
        L.len_zcorrs = (uint32_t)(8 * Z.n_corr);
        L.len_zinputs = (uint32_t)(8 * Z.n_inputs);
    }
    uint8_t *out = (uint8_t *)calloc(L.total(), 1);
    for (uint32_t r = 0; r < nreps && Z.any(); r++) {  // k_zextract, byte by byte
        if (omit[r] >= RV_PLAYERS) continue;
        const uint8_t *zo = &zon[(size_t)r * pitch_zon], *zp = &zpre[(size_t)r * pitch_zpre];
        uint8_t *z = out + L.z_base() + 8 + (size_t)rank[r] * L.sz_on_z();
        const uint64_t nr = 8ull * Z.recon_off.size(), nc = 8ull * Z.n_corr, ni = 8ull * Z.n_inputs;
        for (uint64_t i = 0; i < nr + nc + ni; i++) {
            if (i < nr) z[137 + i] = zo[Z.recon_off[i >> 3] + 8 * omit[r] + (i & 7)];
            else if (i < nr + nc) z[145 + i] = zp[i - nr];
            else z[153 + i] = zo[Z.input_off[(i - nr - nc) >> 3] + ((i - nr - nc) & 7)];
        }
    }
    for (uint32_t r = 0; r < nreps; r++) {
        ExtractView v;
        v.on = &on[(size_t)r * pitch_on];
        v.pre = &pre[(size_t)r * pitch_pre];
        v.on_hash = reinterpret_cast<const uint8_t *>(&on_hash[r * 8]);
        v.pkeys = &pkeys[(size_t)r * 128];
        v.seed = seeds + (size_t)r * 16;
        v.comm = reinterpret_cast<const uint8_t *>(comm);
        v.z64_empty_hash = empty;
        v.z_on_hash = Z.any() ? reinterpret_cast<const uint8_t *>(&zon_hash[r * 8]) : nullptr;
        v.recon_pos = P.recon_pos.data();
        v.input_pos = P.input_pos.data();
        v.n_recon = (uint32_t)P.recon_pos.size();
        v.n_pre = P.n_pre;
        v.n_inputs = (uint32_t)P.n_inputs;
        for (uint32_t tid = 0; tid < 4; tid++) extract_entry(L, v, r, omit[r], rank[r], tid, 4, out);
    }
    *proof = out;
    *proof_len = L.total();
    return RV_OK;
}

// Proof::new in streaming mode on the CPU: the product's planner (rv_stream_plan.h) and its compiler's segment mode (SegmentIO),
// with the carried wires kept in a cell file exactly as rv_prove_streaming does -- imports are leaves of both planes, exports are
// written back after the segment.  The hash streams are simply concatenated here (the chunk carry is kernel-level logic that the
// GPU tests cover); what this pins on the CPU is segmentation, liveness, slot recycling and the compiler's import / export rows.
extern "C" int hs_prove_streaming(const rv_op *ops, size_t n_ops, size_t gf2_cells, const uint8_t *wit, size_t n_wit, const uint8_t *seeds,
                                  size_t window_ops, uint8_t **proof, size_t *proof_len) {
    StreamPlan plan;
    int rc = plan_stream(ops, n_ops, gf2_cells, std::max<size_t>(window_ops, 64), plan, g_err);
    if (rc) return rc;
    if (n_wit < plan.tot_inputs) return RV_E_WITNESS_SHORT;
    const uint32_t npi = 32, nreps = 256;
    std::vector<uint64_t> all_rows(((size_t)plan.masks + 1) * npi, 0);
    std::vector<uint8_t> pkeys;
    gen_masks(seeds, nullptr, nullptr, nullptr, npi, (uint32_t)plan.masks, all_rows, pkeys);
    std::vector<uint64_t> cell_rows((size_t)std::max(plan.n_slots, 1u) * npi, 0);
    std::vector<uint8_t> cell_vals(std::max(plan.n_slots, 1u), 0);
    const size_t pitch_on = (std::max<size_t>(plan.tot_on, 1) + 63) / 64 * 64, pitch_pre = (std::max<size_t>(plan.tot_pre, 1) + 63) / 64 * 64;
    std::vector<uint8_t> on(pitch_on * nreps, 0), pre(pitch_pre * nreps, 0);
    std::vector<uint32_t> recon_pos, input_pos;
    std::vector<std::vector<uint32_t>> seg_recon_pos, seg_input_pos;
    int bad = 0;
    for (Segment &S : plan.segs) {
        Program P;
        rc = compile(S.ops.data(), S.ops.size(), 0, S.n_local, P, g_err, COMPILE_PROVE_ONLY, &S.io);
        if (rc) return rc;
        if (P.n_tvals || P.z.any()) return RV_E_UNSUPPORTED;
        const uint32_t n_imp = (uint32_t)S.io.import_cells.size();
        if (P.n_masks != P.n_prg + n_imp) { g_err = "segment rows"; return -301; }
        std::vector<uint64_t> rows((size_t)P.n_rows * npi, 0);
        memcpy(rows.data(), &all_rows[(size_t)S.mask0 * npi], (size_t)P.n_prg * npi * 8);                                  // k_mask_gen_tt with mask_base
        for (uint32_t j = 0; j < n_imp; j++) memcpy(&rows[(size_t)(P.n_prg + j) * npi], &cell_rows[(size_t)S.import_slot[j] * npi], npi * 8);  // k_seg_import
        std::vector<uint8_t> vals(P.n_vals, 0);
        for (size_t k = 0; k < P.n_inputs; k++) vals[P.input_vid[k]] = wit[S.wit0 + k] & 1;
        for (uint32_t j = 0; j < n_imp; j++) vals[S.io.import_vid[j]] = cell_vals[S.import_slot[j]];
        for (const VGate &g : P.vgates) {
            const uint32_t a = vals[g.a >> 1] ^ (g.a & 1), b = vals[g.b >> 1] ^ (g.b & 1);
            vals[g.dst] = (uint8_t)((g.op ? (a & b) : (a ^ b)) & 1);
        }
        for (const XGate &g : P.xgates)
            for (uint32_t pi = 0; pi < npi; pi++) {
                uint64_t v = 0;
                for (int k = 0; k < 6; k++) v ^= rows[(size_t)g.in[k] * npi + pi];
                rows[(size_t)g.dst * npi + pi] = v;
            }
        uint32_t j = 0;
        for (uint32_t t = 0; t < P.n_online; t++) {
            const Item &it = P.items[t];
            for (uint32_t pi = 0; pi < npi; pi++) {
                const uint64_t w = prover_online_word(it, rows.data(), npi, pi, vals.data(), nullptr, &bad);
                for (int r = 0; r < 8; r++) on[(size_t)(8 * pi + r) * pitch_on + S.on0 + t] = (uint8_t)(w >> (8 * (7 - r)));  // rep r = big-endian byte r
                if (it.kind == ITEM_MUL) {
                    const uint64_t d = pre_word(it, rows.data(), npi, pi);
                    for (int r = 0; r < 8; r++) pre[(size_t)(8 * pi + r) * pitch_pre + S.pre0 + j] = (uint8_t)(d >> (8 * (7 - r)));
                }
            }
            if (it.kind == ITEM_MUL) j++;
        }
        for (uint32_t k : P.recon_pos) recon_pos.push_back((uint32_t)(S.on0 + k));
        for (uint32_t k : P.input_pos) input_pos.push_back((uint32_t)(S.on0 + k));
        seg_recon_pos.push_back(P.recon_pos);  // kept per segment for the segment-wise packing of the openings below
        seg_input_pos.push_back(P.input_pos);
        for (size_t k = 0; k < S.io.export_cells.size(); k++) {  // k_seg_export
            memcpy(&cell_rows[(size_t)S.export_slot[k] * npi], &rows[(size_t)S.io.export_row[k] * npi], npi * 8);
            cell_vals[S.export_slot[k]] = (uint8_t)((vals[S.io.export_vref[k] >> 1] ^ S.io.export_vref[k]) & 1);
        }
    }
    if (bad) return RV_E_WITNESS_INVALID;
    // K5 .. K7 on the concatenated streams (GF(2) only: the Z64 transcript of every repetition is the empty one)
    uint32_t empty[8], zrep0[8];
    b3_chunk_cv(nullptr, 0, 0, true, empty);
    b3_hash64(empty, empty, zrep0);
    std::vector<uint32_t> on_hash(nreps * 8), rep_hash(nreps * 8);
    for (uint32_t r = 0; r < nreps; r++) {
        uint32_t h_pre[8];
        stream_hash(&on[(size_t)r * pitch_on], (uint32_t)plan.tot_on, &on_hash[r * 8]);
        stream_hash(&pre[(size_t)r * pitch_pre], (uint32_t)plan.tot_pre, h_pre);
        rep_join(&on_hash[r * 8], h_pre, zrep0, &rep_hash[r * 8]);
    }
    uint32_t comm[8], m[16];
    stream_hash(reinterpret_cast<const uint8_t *>(rep_hash.data()), nreps * 32, comm);
    challenge_block(comm, m);
    uint8_t omit[256];
    uint16_t rank[256];
    memset(omit, RV_PLAYERS, sizeof omit);
    int distinct = 0;
    for (uint64_t t = 0; distinct < RV_ONLINE_REPS; t++) {
        uint32_t o[16];
        challenge_xof_block(m, t, o);
        challenge_consume(o, omit, &distinct);
    }
    uint16_t n_on = 0, n_pre = 0;
    for (int i = 0; i < 256; i++) rank[i] = omit[i] < RV_PLAYERS ? n_on++ : n_pre++;
    const ProofLayout L{(uint32_t)(recon_pos.size() / 8 + 1), (uint32_t)(plan.tot_pre / 8 + 1), (uint32_t)(plan.tot_inputs / 8 + 1)};
    uint8_t *out = (uint8_t *)calloc(L.total(), 1);
    for (uint32_t r = 0; r < nreps; r++) {
        ExtractView v;
        v.on = &on[(size_t)r * pitch_on];
        v.pre = &pre[(size_t)r * pitch_pre];
        v.on_hash = reinterpret_cast<const uint8_t *>(&on_hash[r * 8]);
        v.pkeys = &pkeys[(size_t)r * 128];
        v.seed = seeds + (size_t)r * 16;
        v.comm = reinterpret_cast<const uint8_t *>(comm);
        v.z64_empty_hash = empty;
        v.z_on_hash = nullptr;
        v.recon_pos = recon_pos.data();
        v.input_pos = input_pos.data();
        v.n_recon = (uint32_t)recon_pos.size();
        v.n_pre = (uint32_t)plan.tot_pre;
        v.n_inputs = (uint32_t)plan.tot_inputs;
        // headers, keys, hashes and zeroed vectors first (k_extract with empty tables), then every segment ORs in its bits (k_seg_extract)
        v.n_recon = v.n_pre = v.n_inputs = 0;
        for (uint32_t tid = 0; tid < 4; tid++) extract_entry(L, v, r, omit[r], rank[r], tid, 4, out);
        if (omit[r] >= RV_PLAYERS) continue;
        uint8_t *e = out + L.g_base() + 8 + (size_t)rank[r] * L.sz_on_g();
        for (size_t k = 0; k < plan.segs.size(); k++) {
            const Segment &S = plan.segs[k];
            const uint8_t *son = &on[(size_t)r * pitch_on + S.on0], *spre = &pre[(size_t)r * pitch_pre + S.pre0];
            const uint32_t n_rec = (uint32_t)seg_recon_pos[k].size(), n_in = (uint32_t)seg_input_pos[k].size();
            const uint32_t n_cor = (uint32_t)((k + 1 < plan.segs.size() ? plan.segs[k + 1].pre0 : plan.tot_pre) - S.pre0);
            auto gather = [&](uint8_t *dst, uint64_t first, uint32_t n, const uint8_t *stream, const uint32_t *pos, uint32_t bit) {
                if (!n) return;
                for (uint64_t g = first / 8; g <= (first + n - 1) / 8; g++) dst[g] |= seg_pack_byte(stream, pos, first, n, g, bit);
            };
            gather(e + 137, S.recon0, n_rec, son, seg_recon_pos[k].data(), 7 - omit[r]);
            gather(e + 145 + L.len_recons, S.pre0, n_cor, spre, nullptr, 0);
            gather(e + 153 + L.len_recons + L.len_corrs, S.wit0, n_in, son, seg_input_pos[k].data(), 0);
        }
    }
    *proof = out;
    *proof_len = L.total();
    return RV_OK;
}

extern "C" void hs_free(void *p) { free(p); }

// The threaded planner against the serial one (rv_stream_plan.h): 0 = same return code, same error text, same plan in every field.
extern "C" int hs_plan_compare(const rv_op *ops, size_t n_ops, size_t gf2_cells, size_t window_ops, unsigned n_threads) {
    StreamPlan A, B;
    std::string ea, eb;
    const int ra = plan_stream_serial(ops, n_ops, gf2_cells, window_ops, A, ea);
    const int rb = plan_stream(ops, n_ops, gf2_cells, window_ops, B, eb, nullptr, n_threads);
    if (ra != rb || ea != eb) { g_err = "rc / error differ: " + std::to_string(ra) + " '" + ea + "' vs " + std::to_string(rb) + " '" + eb + "'"; return 1; }
    if (ra != RV_OK) return 0;
    if (A.n_slots != B.n_slots || A.masks != B.masks || A.tot_on != B.tot_on || A.tot_pre != B.tot_pre || A.tot_inputs != B.tot_inputs || A.tot_recon != B.tot_recon ||
        A.gf2_cells != B.gf2_cells || A.segs.size() != B.segs.size()) { g_err = "totals differ"; return 1; }
    for (size_t k = 0; k < A.segs.size(); k++) {
        const Segment &x = A.segs[k], &y = B.segs[k];
        const bool same = x.a == y.a && x.b == y.b && x.n_local == y.n_local && x.ops.size() == y.ops.size() &&
                          (x.ops.empty() || !memcmp(x.ops.data(), y.ops.data(), x.ops.size() * sizeof(rv_op))) && x.io.import_cells == y.io.import_cells &&
                          x.io.export_cells == y.io.export_cells && x.import_slot == y.import_slot && x.export_slot == y.export_slot && x.import_global == y.import_global &&
                          x.export_global == y.export_global && x.mask0 == y.mask0 && x.on0 == y.on0 && x.pre0 == y.pre0 && x.wit0 == y.wit0 && x.recon0 == y.recon0;
        if (!same) { g_err = "segment " + std::to_string(k) + " differs"; return 1; }
    }
    return 0;
}

// 64-bit FNV-1a over every table of the compiled program that reaches the device (tests: two compiles that must agree).
extern "C" int hs_program_digest(const rv_op *ops, size_t n_ops, size_t z64_cells, size_t gf2_cells, uint32_t flags, uint64_t *digest) {
    Program P;
    int rc = compile(ops, n_ops, z64_cells, gf2_cells, P, g_err, flags);
    if (rc) return rc;
    uint64_t h = 1469598103934665603ull;
    auto eat = [&](const void *p, size_t n) {
        const uint8_t *b = (const uint8_t *)p;
        for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 1099511628211ull;
        h = (h ^ (n & 0xFF)) * 1099511628211ull;
    };
    auto vec = [&](const auto &v) { eat(v.data(), v.size() * sizeof(v[0])); };
    vec(P.items), vec(P.recon_pos), vec(P.input_pos), vec(P.input_vid), vec(P.xgates), vec(P.xlevel_off), vec(P.vm_steps), vec(P.lut_steps);
    vec(P.wgates), vec(P.wlevel_off), vec(P.vlut_steps), vec(P.vwgates), vec(P.vwlevel_off), vec(P.input_uid), vec(P.kappa_uid);
    vec(P.item_ua), vec(P.item_ub), vec(P.tgates), vec(P.tlevel_off), vec(P.rand_row), vec(P.rand_uid), vec(P.b2a_vrefs), vec(P.b2a_urefs);
    vec(P.z.vprog), vec(P.z.lin), vec(P.z.items), vec(P.z.leaf_ids), vec(P.z.recon_off), vec(P.z.input_off), vec(P.z.mul_pos);
    const uint64_t scalars[] = {P.n_and, P.n_inputs, P.n_assert, P.n_masks, P.n_lin, P.n_rows, P.n_vals, P.n_online, P.n_uvals, P.n_vm_steps, P.n_lut_steps,
                                P.n_vlut_steps, P.vm_cells, P.algorithmic_bytes, (uint64_t)P.values_wide, (uint64_t)P.verify_wide};
    eat(scalars, sizeof(scalars));
    *digest = h;
    return RV_OK;
}

// Replays the padded device step streams of both planes (what the kernels actually execute) against the plain networks:
// the mask VM on random fresh rows vs. the unmapped... mapped XOR gates, the LUT stream on random witnesses vs. the circuit's
// own 2-input gates.  stats: [n_vm_steps, n_lut_steps, vm_cells, n_luts].
extern "C" int hs_check_steps(const rv_op *ops, size_t n_ops, size_t z64_cells, size_t gf2_cells, uint32_t *stats) {
    Program P;
    int rc = compile(ops, n_ops, z64_cells, gf2_cells, P, g_err);
    if (rc) return rc;
    stats[0] = P.n_vm_steps;
    stats[1] = P.n_lut_steps;
    stats[2] = P.vm_cells;
    stats[3] = (uint32_t)(P.values_wide ? P.wgates.size() : P.luts.size());
    stats[4] = (uint32_t)P.xlevel_off.size() - 1;
    stats[5] = (uint32_t)(P.values_wide ? P.wlevel_off.size() : P.lut_level_off.size()) - 1;
    stats[6] = P.n_lin;
    if (P.vm_steps.size() != (size_t)P.n_vm_steps * VM_STEP || P.lut_steps.size() != (size_t)P.n_lut_steps * LUT_STEP) { g_err = "stream size"; return -300; }
    // --- mask VM: once with every LOAD landing immediately, once landing as late as the cp.async group wait allows ---
    for (int late = 0; late < 2; late++) {
        std::vector<uint32_t> rows(P.n_rows, 0), ref;
        uint32_t x = 99;
        for (uint32_t r = 0; r < P.n_masks; r++) rows[r] = (x = x * 1664525u + 1013904223u);
        ref = rows;
        for (const XGate &g : P.xgates) {
            uint32_t v = 0;
            for (int k = 0; k < 6; k++) v ^= ref[g.in[k]];
            ref[g.dst] = v;
        }
        std::vector<uint32_t> cells(P.vm_cells + 1, 0xDEADBEEF);
        cells[0] = 0;
        std::vector<std::pair<uint32_t, uint32_t>> writes;
        std::vector<std::pair<uint32_t, uint32_t>> pending[2];  // [0]: issued in the current level, [1]: in the previous one
        // The device orders the steps of one level only loosely (a barrier at the level's end and at chunk ends), and a LOAD lands
        // any time between its issue and the group wait: inside a level no cell may be both read and written or written twice, and
        // a cell with a LOAD in flight (this level and the next) is off limits for everybody else.  Stamps: level of the last read /
        // write / LOAD issue of every cell (the scratch cell takes any number of writes and is never read by a real instruction).
        std::vector<uint32_t> rd_level(P.vm_cells + 1, 0), wr_level(P.vm_cells + 1, 0), ld_level(P.vm_cells + 1, 0);
        uint32_t level_no = 2;  // (0 = never; LOADs of level L block their cell during L and L + 1)
        for (uint32_t st = 0; st < P.n_vm_steps; st++) {
            writes.clear();
            bool bar = false, level_end = false;
            for (uint32_t t = 0; t < VM_STEP; t++) {
                const VmInstr &in = P.vm_steps[(size_t)st * VM_STEP + t];
                bar = (in.flags & VM_F_BAR) != 0;
                level_end = (in.flags & VM_F_LEVEL_END) != 0;
                if (((P.vm_steps[(size_t)st * VM_STEP].flags & VM_F_BAR) != 0) != bar) { g_err = "non-uniform barrier flag"; return -301; }
                if ((in.flags & VM_F_LEVEL_END) && !bar) { g_err = "level end without barrier"; return -305; }
                if (in.dst > P.vm_cells) { g_err = "cell id out of range"; return -306; }
                if (in.flags & VM_F_LOAD) {
                    if (in.dst == 0 || in.dst >= P.vm_cells) { g_err = "LOAD into the zero or scratch cell"; return -307; }
                    if (rd_level[in.dst] == level_no || wr_level[in.dst] == level_no || ld_level[in.dst] + 1 >= level_no) { g_err = "LOAD into a cell in use in its level (step " + std::to_string(st) + ")"; return -308; }
                    ld_level[in.dst] = level_no;
                    if (late) pending[0].push_back({in.dst, rows[in.row]});
                    else writes.push_back({in.dst, rows[in.row]});
                } else {
                    uint32_t v = 0;
                    for (int k = 0; k < 6; k++) {
                        const uint32_t c = in.in[k];
                        if (c != 0 && (wr_level[c] == level_no || ld_level[c] + 1 >= level_no)) { g_err = "XOR reads a cell written or loaded in its own level (step " + std::to_string(st) + ")"; return -309; }
                        rd_level[c] = level_no;
                        v ^= cells[c];
                    }
                    if (in.dst != P.vm_cells) {  // (not the scratch cell)
                        if (in.dst == 0) { g_err = "XOR writes the zero cell"; return -310; }
                        if (rd_level[in.dst] == level_no || wr_level[in.dst] == level_no || ld_level[in.dst] + 1 >= level_no) { g_err = "XOR writes a cell in use in its level (step " + std::to_string(st) + ")"; return -311; }
                        wr_level[in.dst] = level_no;
                    }
                    writes.push_back({in.dst, v});
                    if (in.row != VM_ROW_NONE) rows[in.row] = v;
                }
            }
            if ((st + 1) % VM_STEPS_PER_CHUNK == 0 && !bar) { g_err = "missing barrier at chunk end"; return -302; }
            for (auto &w : writes) cells[w.first] = w.second;
            if (level_end) {  // commit; wait_prior(VM_DELTA - 1): everything but the newest group has landed
                for (auto &w : pending[1]) cells[w.first] = w.second;
                pending[1].swap(pending[0]);
                pending[0].clear();
                level_no++;
            }
        }
        for (const Item &it : P.items) {
            const uint32_t rr[2] = {it.ra, it.kind == ITEM_MUL ? it.rb : it.ra};
            for (uint32_t r : rr)
                if (rows[r] != ref[r]) { g_err = "VM step stream: row mismatch at row " + std::to_string(r) + (late ? " (late LOADs)" : ""); return -303; }
        }
    }
    // --- the verifier's u-plane stream: the same barrier discipline (its values are checked by hs_verify) ---
    {
        std::vector<uint32_t> wr_region(P.n_uvals + 1, 0);
        uint32_t region = 1;
        if (P.vlut_steps.size() != (size_t)P.n_vlut_steps * LUT_STEP) { g_err = "u-plane stream size"; return -315; }
        for (uint32_t st = 0; st < P.n_vlut_steps; st++) {
            bool bar = false;
            for (uint32_t t = 0; t < LUT_STEP; t++) {
                const LutInstr &li = P.vlut_steps[(size_t)st * LUT_STEP + t];
                bar |= (li.pad & LUT_F_BAR) != 0;
                for (int k = 0; k < 6; k++) {
                    if (li.in[k] > P.n_uvals) { g_err = "u-plane stream: value id out of range"; return -316; }
                    if (wr_region[li.in[k]] == region) { g_err = "u-plane stream: step " + std::to_string(st) + " reads a value written since the last barrier"; return -317; }
                }
                if (li.dst != P.n_uvals) {
                    if (li.dst > P.n_uvals || wr_region[li.dst] == region) { g_err = "u-plane stream: value written twice between barriers"; return -318; }
                    wr_region[li.dst] = region;
                }
            }
            if (bar) region++;
        }
    }
    // --- LUT stream ---
    uint32_t x = 4242;
    for (int trial = 0; trial < 3; trial++) {
        std::vector<uint8_t> a(P.n_vals + 1, 0), b(P.n_vals + 1, 0);
        for (size_t k = 0; k < P.n_inputs; k++) {
            x = x * 1664525u + 1013904223u;
            a[P.input_vid[k]] = b[P.input_vid[k]] = (x >> 16) & 1;
        }
        for (const VGate &g : P.vgates) {
            const uint32_t u = a[g.a >> 1] ^ (g.a & 1), v = a[g.b >> 1] ^ (g.b & 1);
            a[g.dst] = (uint8_t)((g.op ? (u & v) : (u ^ v)) & 1);
        }
        std::vector<std::pair<uint32_t, uint8_t>> w;
        if (P.values_wide)  // k_values_level: one launch per level over the level-sorted 2-input gates
            for (const VGate &g : P.wgates) {
                const uint32_t u = b[g.a >> 1] ^ (g.a & 1), v = b[g.b >> 1] ^ (g.b & 1);
                b[g.dst] = (uint8_t)((g.op ? (u & v) : (u ^ v)) & 1);
            }
        // steps between two barriers run in no particular order on the device: a value written in such a region must not be read
        // (or written again) inside it.  wr_region: the region (1-based) of every value's write.
        std::vector<uint32_t> wr_region(P.n_vals + 1, 0);
        uint32_t region = 1;
        for (uint32_t st = 0; st < P.n_lut_steps; st++) {
            w.clear();
            bool bar = false;
            for (uint32_t t = 0; t < LUT_STEP; t++) {
                const LutInstr &li = P.lut_steps[(size_t)st * LUT_STEP + t];
                bar |= (li.pad & LUT_F_BAR) != 0;
                uint32_t idx = 0;
                for (int k = 0; k < 6; k++) {
                    if (li.in[k] > P.n_vals) { g_err = "LUT step stream: value id out of range"; return -312; }
                    if (wr_region[li.in[k]] == region) { g_err = "LUT step stream: step " + std::to_string(st) + " reads a value written since the last barrier"; return -313; }
                    idx |= (uint32_t)b[li.in[k]] << k;
                }
                if (li.dst != P.n_vals) {  // (not the scratch slot of the padding)
                    if (wr_region[li.dst] == region) { g_err = "LUT step stream: value written twice between barriers"; return -314; }
                    wr_region[li.dst] = region;
                }
                w.push_back({li.dst, (uint8_t)((li.tt >> idx) & 1)});
            }
            if (bar) region++;
            // without a barrier the next step may or may not see these writes; with the level structure it must not matter,
            // so apply them only at barriers for non-barrier steps' sake: here we apply immediately (reads of a later step
            // of the same level never touch this level's outputs)
            for (auto &p : w) b[p.first] = p.second;
        }
        for (const Item &it : P.items) {
            const uint32_t vv[2] = {it.va >> 1, it.kind == ITEM_MUL ? it.vb >> 1 : it.va >> 1};
            for (uint32_t v : vv)
                if (a[v] != b[v]) { g_err = "LUT step stream: value mismatch at vid " + std::to_string(v); return -304; }
        }
    }
    return 0;
}

// Proof::verify on the CPU through the kernel bodies (mirrors verify_on_session in csrc/rv_api.cu).
// Returns 1 accept / 0 reject / <0 error; rep_hashes (optional, 256*32) receives the hashes in ORIGINAL repetition order.
#include "../../reverie_b200/csrc/rv_bincode.h"
extern "C" int hs_verify(const rv_op *ops, size_t n_ops, size_t z64_cells, size_t gf2_cells, const uint8_t *proof, size_t proof_len, int *okay,
                         uint8_t *rep_hashes) {
    Program P;
    int rc = compile(ops, n_ops, z64_cells, gf2_cells, P, g_err);
    if (rc) return rc;
    if (proof_len < 32) return RV_E_FORMAT;
    PDomain g, z;
    size_t pos = 32;
    if (!parse_domain(proof, proof_len, pos, g) || !parse_domain(proof, proof_len, pos, z) || pos != proof_len) return RV_E_FORMAT;
    if (g.online.size() != 40 || g.pre.size() != 216 || z.online.size() != 40 || z.pre.size() != 216) return 0;
    const uint32_t npi = 32, NON = 40;
    std::vector<uint8_t> seeds(256 * 16, 0), pkeys_in(256 * 128, 0), mode(256, 0), omit(256, 8);
    std::vector<VOpen> opens(NON);
    for (uint32_t k = 0; k < NON; k++) {
        const POnline &o = g.online[k], &first = g.online[k & ~7u];
        if (o.omit >= 8 || z.online[k].omit >= 8) return RV_E_FORMAT;
        if (o.recons.len != first.recons.len || o.corrs.len < first.corrs.len || o.inputs.len < first.inputs.len) return RV_E_FORMAT;
        memcpy(&pkeys_in[k * 128], proof + o.keys, 128);
        mode[k] = 1;
        omit[k] = o.omit;
        opens[k] = VOpen{(uint32_t)o.recons.off, (uint32_t)o.corrs.off, (uint32_t)o.inputs.off, (uint32_t)first.recons.len,
                         (uint32_t)first.corrs.len, (uint32_t)first.inputs.len, o.omit, 0};
    }
    for (uint32_t k = 0; k < 216; k++) memcpy(&seeds[(NON + k) * 16], proof + g.pre[k].seed, 16);
    std::vector<uint64_t> rows((size_t)P.n_rows * npi, 0);
    std::vector<uint8_t> pkeys;
    SimKeys K;
    gen_masks(seeds.data(), pkeys_in.data(), mode.data(), omit.data(), npi, P.n_masks, rows, pkeys, &K);
    for (const XGate &x : P.xgates)
        for (uint32_t pi = 0; pi < npi; pi++) {
            uint64_t v = 0;
            for (int k = 0; k < 6; k++) v ^= rows[(size_t)x.in[k] * npi + pi];
            rows[(size_t)x.dst * npi + pi] = v;
        }
    std::vector<uint32_t> mul_pos, recon_idx(P.n_online, 0);
    for (uint32_t t = 0; t < P.n_online; t++)
        if (P.items[t].kind == ITEM_MUL) mul_pos.push_back(t);
    for (uint32_t k = 0; k < P.recon_pos.size(); k++) recon_idx[P.recon_pos[k]] = k;
    // u-plane of the 40 opened repetitions
    const size_t upitch = (size_t)P.n_uvals + 1;
    std::vector<uint8_t> uvals(upitch * NON, 0);
    for (uint32_t s = 0; s < NON; s++) {
        uint8_t *uv = &uvals[s * upitch];
        for (uint32_t k = 0; k < P.n_inputs; k++) uv[P.input_uid[k]] = verify_leaf_input(P.items[P.input_pos[k]], k, opens[s], proof, rows.data(), npi, s);
        for (uint32_t j = 0; j < P.n_pre; j++) uv[P.kappa_uid[j]] = verify_leaf_kappa(P.items[mul_pos[j]], recon_idx[mul_pos[j]], opens[s], proof, rows.data(), npi, s);
        for (uint32_t k = 0; k < P.rand_uid.size(); k++) uv[P.rand_uid[k]] = verify_leaf_random(P.rand_row[k], rows.data(), npi, s);
        if (P.verify_wide)  // k_uvalues_level: one launch per level over the level-sorted 2-input gates
            for (const VGate &g : P.vwgates) {
                const uint32_t x = uv[g.a >> 1] ^ (g.a & 1), y = uv[g.b >> 1] ^ (g.b & 1);
                uv[g.dst] = (uint8_t)((g.op ? (x & y) : (x ^ y)) & 1);
            }
        for (uint32_t st = 0; st < P.n_vlut_steps; st++)
            for (uint32_t t = 0; t < LUT_STEP; t++) {
                const LutInstr &li = P.vlut_steps[(size_t)st * LUT_STEP + t];
                uint32_t idx = 0;
                for (int k = 0; k < 6; k++) idx |= (uint32_t)uv[li.in[k]] << k;
                uv[li.dst] = (uint8_t)((li.tt >> idx) & 1);
            }
    }
    // ---- Z64 instances (mirrors the has_z block of verify_on_session) ----
    const ZProgram &Z = P.z;
    std::vector<uint32_t> zrep_slot(256 * 8, 0);
    int z_not_okay = 0;
    if (Z.any()) {
        std::vector<ZOpen> zopens(NON);
        std::vector<uint8_t> zseeds(256 * 16, 0), zpkeys_in(256 * 128, 0), zomit(256, 8);
        bool own = false;
        for (uint32_t k = 0; k < NON; k++) {
            const POnline &o = z.online[k], &first = z.online[k & ~7u];
            zopens[k] = ZOpen{o.recons.off, o.corrs.off, o.inputs.off, (uint32_t)(first.recons.len / 8), (uint32_t)(first.corrs.len / 8),
                              (uint32_t)(first.inputs.len / 8), (uint32_t)o.recons.len, (uint32_t)o.corrs.len, (uint32_t)o.inputs.len, o.omit, 0};
            memcpy(&zpkeys_in[k * 128], proof + o.keys, 128);
            zomit[k] = o.omit;
            if (o.omit != g.online[k].omit || memcmp(proof + o.keys, proof + g.online[k].keys, 128) != 0) own = true;
        }
        for (uint32_t k = 0; k < 216; k++) {
            memcpy(&zseeds[(NON + k) * 16], proof + z.pre[k].seed, 16);
            if (memcmp(proof + z.pre[k].seed, proof + g.pre[k].seed, 16) != 0) own = true;
        }
        SimKeys KZ;
        if (own) {
            std::vector<uint64_t> dummy(1, 0);
            std::vector<uint8_t> pk2;
            gen_masks(zseeds.data(), zpkeys_in.data(), mode.data(), zomit.data(), npi, 0, dummy, pk2, &KZ);
        }
        const size_t rowlen = (size_t)64 * npi;
        std::vector<uint64_t> zrows;
        gen_zrows(own ? KZ : K, npi, Z, zrows);
        const size_t pitch_zon = (std::max<size_t>(Z.on_bytes, 1) + 63) / 64 * 64, pitch_zpre = (std::max<size_t>(Z.pre_bytes, 1) + 63) / 64 * 64;
        std::vector<uint8_t> zon(pitch_zon * 256, 0), zpre(pitch_zpre * 256, 0);
        std::vector<uint64_t> leaves(Z.leaf_ids.size() + 1), uv((size_t)Z.n_vals + 1);
        std::vector<uint32_t> input_item;
        for (uint32_t t = 0; t < Z.items.size(); t++)
            if (Z.items[t].kind == ITEM_INPUT) input_item.push_back(t);
        for (uint32_t slot = 0; slot < 256; slot++) {
            uint32_t h_on[8], h_pre[8];
            if (slot < NON) {
                for (uint32_t k = 0; k < Z.n_inputs; k++) leaves[k] = z_verify_leaf_input(Z.items[input_item[k]], k, zopens[slot], proof, zrows.data(), rowlen, slot);
                for (uint32_t j = 0; j < Z.n_corr; j++) {
                    const ZItem &it = Z.items[Z.mul_pos[j]];
                    leaves[Z.n_inputs + j] = it.kind == ITEM_B2A ? z_verify_leaf_b2a(it, zopens[slot], opens[slot], proof, zrows.data(), rowlen, slot,
                                                                                     &uvals[slot * upitch], P.b2a_urefs.data())
                                                                 : z_verify_leaf_kappa(it, Z.recon_idx[Z.mul_pos[j]], zopens[slot], proof, zrows.data(), rowlen, slot);
                }
                z_values(Z, leaves.data(), uv.data(), nullptr, nullptr);
                for (uint32_t t = 0; t < Z.items.size(); t++)
                    z_verify_online(Z.items[t], Z.recon_idx[t], zopens[slot], proof, zrows.data(), rowlen, slot, uv.data(), &zon[(size_t)slot * pitch_zon], &z_not_okay);
                for (uint32_t j = 0; j < Z.n_corr; j++)
                    put64(&zpre[(size_t)slot * pitch_zpre + 8ull * j], z_packed(proof, zopens[slot].off_corrs, zopens[slot].n_corrs, zopens[slot].len_corrs, j));
                stream_hash(&zon[(size_t)slot * pitch_zon], (uint32_t)Z.on_bytes, h_on);
            } else {
                for (uint32_t j = 0; j < Z.n_corr; j++)
                    put64(&zpre[(size_t)slot * pitch_zpre + 8ull * j], z_pre_word(Z.items[Z.mul_pos[j]], zrows.data(), rowlen, slot, rows.data(), npi));
                memcpy(h_on, proof + z.pre[slot - NON].comm_online, 32);
            }
            stream_hash(&zpre[(size_t)slot * pitch_zpre], (uint32_t)Z.pre_bytes, h_pre);
            b3_hash64(h_pre, h_on, &zrep_slot[slot * 8]);
        }
    }
    const size_t pitch_on = (std::max<size_t>(P.n_online, 1) + 63) / 64 * 64, pitch_pre = (std::max<size_t>(P.n_pre, 1) + 63) / 64 * 64;
    std::vector<uint8_t> on(pitch_on * 256, 0), pre(pitch_pre * 256, 0);
    int not_okay = 0;
    for (uint32_t pi = 0; pi < npi; pi++) {
        if (pi < NON / 8)
            for (uint64_t t0 = 0; t0 < P.n_online; t0 += 8) {
                uint64_t W[8], out[8];
                for (int i = 0; i < 8; i++) {
                    const uint32_t t = (uint32_t)t0 + i;
                    W[i] = t < P.n_online ? verify_online_word(P.items[t], t, P.item_ua[t], P.item_ub[t], recon_idx[t], opens.data(), proof, rows.data(), npi, pi,
                                                               uvals.data(), upitch, &not_okay)
                                          : 0;
                }
                words_to_stream_bytes(W, out);
                for (int r = 0; r < 8; r++) memcpy(&on[(size_t)(8 * pi + r) * pitch_on + t0], &out[r], 8);
            }
        for (uint64_t j0 = 0; j0 < P.n_pre; j0 += 8) {
            uint64_t W[8], out[8];
            for (int i = 0; i < 8; i++) {
                const uint32_t j = (uint32_t)j0 + i;
                W[i] = j >= P.n_pre ? 0 : (pi < NON / 8 ? verify_pre_word(j, opens.data(), proof, pi) : pre_word(P.items[mul_pos[j]], rows.data(), npi, pi));
            }
            words_to_stream_bytes(W, out);
            for (int r = 0; r < 8; r++) memcpy(&pre[(size_t)(8 * pi + r) * pitch_pre + j0], &out[r], 8);
        }
    }
    uint32_t empty[8], zrep[8];
    b3_chunk_cv(nullptr, 0, 0, true, empty);
    b3_hash64(empty, empty, zrep);
    std::vector<uint8_t> slot_hash(256 * 32);
    for (uint32_t s = 0; s < 256; s++) {
        uint32_t h_on[8], h_pre[8], zz[8], out[8];
        stream_hash(&pre[(size_t)s * pitch_pre], P.n_pre, h_pre);
        if (s < NON) {
            stream_hash(&on[(size_t)s * pitch_on], P.n_online, h_on);
            memcpy(zz, zrep, 32);
        } else {
            uint32_t zon[8];
            memcpy(h_on, proof + g.pre[s - NON].comm_online, 32);
            memcpy(zon, proof + z.pre[s - NON].comm_online, 32);
            b3_hash64(empty, zon, zz);
        }
        if (Z.any()) memcpy(zz, &zrep_slot[s * 8], 32);
        rep_join(h_on, h_pre, zz, out);
        memcpy(&slot_hash[s * 32], out, 32);
    }
    uint8_t omit_of_rep[256], ordered[256 * 32];
    host_challenge(proof, omit_of_rep);
    size_t a = 0, b = NON;
    for (int i = 0; i < 256; i++) memcpy(ordered + 32 * i, &slot_hash[32 * (omit_of_rep[i] < 8 ? a++ : b++)], 32);
    if (rep_hashes) memcpy(rep_hashes, ordered, sizeof ordered);
    uint32_t comm2[8];
    host_hash(ordered, sizeof ordered, comm2);
    if (okay) *okay = (not_okay || z_not_okay) ? 0 : 1;
    return memcmp(comm2, proof, 32) == 0 ? 1 : 0;
}
