// rv_z64.cu -- sm_100a kernels of the Z64 domain (src/algebra/z64/*, the Z64 instance of src/interpreter/single.rs and of
// the three transcripts).  Integer work only.  Layouts: rv_zplanes.cuh.
#include <algorithm>

#include "rv_kernels.cuh"
#include "rv_zplanes.cuh"

namespace rv {

// =====================================================================================================================
//  ZK2  Z64 mask generation (src/algebra/z64/batch.rs:25-30, src/algebra/z64/domain.rs:64-83): mask i of a stream is the
//       little-endian u64 at keystream byte 8 i, i.e. AES block j = masks 2j, 2j+1 in natural byte order.
// =====================================================================================================================
//      T-table AES: thread = one PRG stream (its 44 round-key words in registers), looping over counter blocks; a warp =
//      32 consecutive streams, so every mask leaves as one 256-byte row segment.  The four 1 KB Te tables are replicated
//      once per shared-memory bank (128 KB): lane l only ever reads words = l (mod 32), so the 32 data-dependent lookups
//      of a warp are conflict-free and the generator is bound by the LDS pipe (160 lookups per block) instead of the
//      ~640 LOP3 per block of the bitsliced form (whose planes would need transposing back for Z64 anyway).
constexpr int ZT_THREADS = 512, ZT_BLOCKS_PER_TASK = 64;
// Shared-memory image: entry x of table t at byte (t >> 1) * 65536 + x * 256 + (t & 1) * 128 + 4 * lane.  The 256-byte entry
// stride makes the data-dependent part of the address a single PRMT: (byte b of w) << 8 | 4 * lane.
struct SmemTe {
    const uint8_t *base;  // the table image (uniform)
    uint32_t lane4;       // 4 * lane
    __device__ __forceinline__ uint32_t operator()(int t, uint32_t w, int b) const {
        const uint32_t off = __byte_perm(w, lane4, 0x7604 | (b << 4));
        return *reinterpret_cast<const uint32_t *>(base + (t >> 1) * 65536 + (t & 1) * 128 + off);
    }
};

__global__ void __launch_bounds__(ZT_THREADS, 1) k_zmask_gen_tt(const uint32_t *__restrict__ rk_plain, uint32_t nstreams, uint32_t n_masks,
                                                                uint64_t *__restrict__ zrows) {
    extern __shared__ __align__(16) uint32_t te[];  // [4][256][32]
    __shared__ uint32_t sbox32[64];
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    if (tid < 64) {
        const uint32_t b = 4 * tid;
        sbox32[tid] = sub_word(b | ((b + 1) << 8) | ((b + 2) << 16) | ((b + 3) << 24));
    }
    __syncthreads();
    for (uint32_t e = tid; e < 256 * 32; e += ZT_THREADS) {
        const uint32_t x = e >> 5, l = e & 31, t0 = te0_entry(reinterpret_cast<const uint8_t *>(sbox32)[x]);
        te[x * 64 + l] = t0;
        te[x * 64 + 32 + l] = (t0 << 8) | (t0 >> 24);
        te[16384 + x * 64 + l] = (t0 << 16) | (t0 >> 16);
        te[16384 + x * 64 + 32 + l] = (t0 << 24) | (t0 >> 8);
    }
    __syncthreads();
    const SmemTe tab{reinterpret_cast<const uint8_t *>(te), 4 * lane};
    const uint32_t n_sg = nstreams / 32, n_blocks = (n_masks + 1) / 2;
    const uint32_t n_ranges = (n_blocks + ZT_BLOCKS_PER_TASK - 1) / ZT_BLOCKS_PER_TASK;
    const uint64_t n_tasks = (uint64_t)n_sg * n_ranges;
    constexpr uint32_t WARPS = ZT_THREADS / 32;
    for (uint64_t task = (uint64_t)blockIdx.x * WARPS + (tid >> 5); task < n_tasks; task += (uint64_t)gridDim.x * WARPS) {
        const uint32_t sidx = (uint32_t)(task % n_sg) * 32 + lane, range = (uint32_t)(task / n_sg);
        uint32_t rk[44];
#pragma unroll
        for (int q = 0; q < 44; q++) rk[q] = rk_plain[(size_t)q * nstreams + sidx];
        const uint32_t act = rk_plain[(size_t)44 * nstreams + sidx];
        const uint32_t j_end = min(n_blocks, (range + 1) * ZT_BLOCKS_PER_TASK);
#pragma unroll 1
        for (uint32_t j = range * ZT_BLOCKS_PER_TASK; j < j_end; j++) {
            uint32_t in[4], o[4];
            ctr_block_words(j, in);
            tt_aes128_encrypt(rk, in[0], in[1], in[2], in[3], tab, o);
            uint64_t *dst = zrows + (size_t)(2 * j) * nstreams + sidx;
            dst[0] = ((uint64_t)(o[1] & act) << 32) | (o[0] & act);
            if (2 * j + 1 < n_masks) dst[nstreams] = ((uint64_t)(o[3] & act) << 32) | (o[2] & act);
        }
    }
}

// called by configure_kernels (rv_kernels.cu) once per device, with that device current
// defined at the end of this file (it names every Z64 kernel)


void launch_zmask_gen_tt(const uint32_t *rk_plain, uint32_t nstreams, uint32_t n_masks, uint64_t *zrows, int n_sms, cudaStream_t st) {
    if (n_masks == 0) return;
    const uint32_t n_blocks = (n_masks + 1) / 2, n_ranges = (n_blocks + ZT_BLOCKS_PER_TASK - 1) / ZT_BLOCKS_PER_TASK;
    const uint64_t n_tasks = (uint64_t)(nstreams / 32) * n_ranges;
    const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)n_sms, (n_tasks + ZT_THREADS / 32 - 1) / (ZT_THREADS / 32));
    k_zmask_gen_tt<<<grid, ZT_THREADS, 4 * 256 * 32 * 4, st>>>(rk_plain, nstreams, n_masks, zrows);
}

// =====================================================================================================================
//  ZK3  mask plane: zrow[dst] = ca * zrow[a] + cb * zrow[b], one launch per level, thread = (node, element)
// =====================================================================================================================
__global__ void __launch_bounds__(256) k_zlinear_level(const ZLin *__restrict__ lin, uint32_t n, uint64_t *zrows, uint32_t rowlen) {
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t g = gid / rowlen;
    const uint32_t e = (uint32_t)(gid % rowlen);
    if (g >= n) return;
    const ZLin nd = lin[g];
    zrows[(size_t)nd.dst * rowlen + e] = nd.ca * zrows[(size_t)nd.a * rowlen + e] + nd.cb * zrows[(size_t)nd.b * rowlen + e];
}

int launch_zlinear(const DevZProgram &Z, const uint32_t *llevel_off_host, uint64_t *zrows, uint32_t rowlen, cudaStream_t st) {
    for (uint32_t l = 0; l < Z.n_llevels; l++) {
        const uint32_t n = llevel_off_host[l + 1] - llevel_off_host[l];
        const uint64_t threads = (uint64_t)n * rowlen;
        k_zlinear_level<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(Z.lin + llevel_off_host[l], n, zrows, rowlen);
    }
    return (int)Z.n_llevels;
}

// =====================================================================================================================
//  ZK0  value plane: level-synchronous walk of the u64 program by one CTA per instance (prover: 1; verifier: one per
//       opened repetition).  The next level's instructions are fetched before the barrier that ends the current one.
// =====================================================================================================================
constexpr int ZV_THREADS = 256;
__global__ void __launch_bounds__(ZV_THREADS) k_zvalues(const ZInstr *__restrict__ prog, const uint32_t *__restrict__ level_off, uint32_t n_levels,
                                                        const uint32_t *__restrict__ leaf_ids, const uint64_t *__restrict__ leaf_vals,
                                                        size_t leaf_pitch, uint32_t n_leaves, uint64_t *vals_out, size_t vals_pitch,
                                                        const uint8_t *__restrict__ gvals, const uint32_t *__restrict__ b2a_vrefs) {
    uint64_t *v = vals_out + (size_t)blockIdx.x * vals_pitch;
    const uint64_t *lv = leaf_vals + (size_t)blockIdx.x * leaf_pitch;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) v[0] = 0;
    for (uint32_t k = tid; k < n_leaves; k += ZV_THREADS) v[leaf_ids[k]] = lv[k];
    if (n_levels == 0) return;
    uint32_t s = level_off[0], e = level_off[1];
    ZInstr nx;
    bool have = s + tid < e;
    if (have) nx = prog[s + tid];
    __syncthreads();
    for (uint32_t l = 0; l < n_levels; l++) {
        const uint32_t e_next = l + 1 < n_levels ? level_off[l + 2] : e;
        uint32_t g = s + tid;
        ZInstr in = nx;
        const bool had = have;
        have = e + tid < e_next;
        if (have) nx = prog[e + tid];  // prefetch: does not depend on this level's values
        if (had) {
            v[in.dst] = z_exec(in, v, gvals, b2a_vrefs);
            for (g += ZV_THREADS; g < e; g += ZV_THREADS) {
                in = prog[g];
                v[in.dst] = z_exec(in, v, gvals, b2a_vrefs);
            }
        }
        __syncthreads();
        s = e;
        e = e_next;
    }
}

void launch_zvalues(const DevZProgram &Z, const uint64_t *leaf_vals, size_t leaf_pitch, uint64_t *vals, size_t vals_pitch, uint32_t n_instances,
                    const uint8_t *gvals, const uint32_t *b2a_vrefs, cudaStream_t st) {
    k_zvalues<<<n_instances, ZV_THREADS, 0, st>>>(Z.vprog, Z.vlevel_off, Z.n_vlevels, Z.leaf_ids, leaf_vals, leaf_pitch, Z.n_leaves, vals, vals_pitch, gvals,
                                                  b2a_vrefs);
}

// =====================================================================================================================
//  ZK4  item plane.  thread = (item, repetition); a CTA = 32 consecutive items x the 8 repetitions of one packed instance
//       (repetition fastest), so a warp reads each operand row as 512 contiguous bytes and appends 256 contiguous bytes
//       to each of its 8 streams.
// =====================================================================================================================
__global__ void __launch_bounds__(256) k_zitems_online(const ZItem *__restrict__ items, uint32_t n_items, const uint64_t *__restrict__ zrows,
                                                       size_t rowlen, const uint64_t *__restrict__ vals, const uint64_t *__restrict__ grows,
                                                       uint8_t *__restrict__ on, size_t pitch, uint8_t *__restrict__ pre, size_t pitch_pre, int *bad) {
    const uint32_t t = blockIdx.x * 32 + (threadIdx.x >> 3), rep = 8 * blockIdx.y + (threadIdx.x & 7);
    if (t >= n_items) return;
    int flag = 0;
    z_prover_online(items[t], zrows, rowlen, rep, vals, on + (size_t)rep * pitch, &flag, pre + (size_t)rep * pitch_pre, grows, (uint32_t)(rowlen / 64));
    if (flag && rep == 0) atomicOr(bad, 1);
}

__global__ void __launch_bounds__(256) k_zitems_pre(const ZItem *__restrict__ items, const uint32_t *__restrict__ mul_pos, uint32_t n_mul,
                                                    const uint64_t *__restrict__ zrows, size_t rowlen, const uint64_t *__restrict__ grows,
                                                    uint8_t *__restrict__ pre, size_t pitch, uint32_t first_rep) {
    const uint32_t j = blockIdx.x * 32 + (threadIdx.x >> 3), rep = first_rep + 8 * blockIdx.y + (threadIdx.x & 7);
    if (j >= n_mul) return;
    put64(pre + (size_t)rep * pitch + 8ull * j, z_pre_word(items[mul_pos[j]], zrows, rowlen, rep, grows, (uint32_t)(rowlen / 64)));
}

void launch_zitems(const DevZProgram &Z, const uint64_t *zrows, size_t rowlen, uint32_t nreps, const uint64_t *vals, const uint64_t *grows, uint8_t *on,
                   size_t pitch_on, uint8_t *pre, size_t pitch_pre, int *bad, cudaStream_t st) {
    // one pass fills both streams: a Mul's preprocessing word needs the operand segments its online word has just loaded
    if (Z.n_items)
        k_zitems_online<<<dim3((Z.n_items + 31) / 32, nreps / 8), 256, 0, st>>>(Z.items, Z.n_items, zrows, rowlen, vals, grows, on, pitch_on, pre, pitch_pre, bad);
}

void launch_zitems_pre_range(const DevZProgram &Z, const uint64_t *zrows, size_t rowlen, uint32_t first_rep, uint32_t nreps, const uint64_t *grows,
                             uint8_t *pre, size_t pitch_pre, cudaStream_t st) {
    if (!Z.n_corr || first_rep >= nreps) return;
    k_zitems_pre<<<dim3((Z.n_corr + 31) / 32, (nreps - first_rep) / 8), 256, 0, st>>>(Z.items, Z.mul_pos, Z.n_corr, zrows, rowlen, grows, pre, pitch_pre,
                                                                                       first_rep);
}

// =====================================================================================================================
//  online verifier
// =====================================================================================================================
__global__ void __launch_bounds__(256) k_zverify_leaves(const ZItem *__restrict__ items, const uint32_t *__restrict__ input_item,
                                                        const uint32_t *__restrict__ mul_pos, const uint32_t *__restrict__ recon_idx,
                                                        uint32_t n_inputs, uint32_t n_mul, const ZOpen *__restrict__ opens,
                                                        const uint8_t *__restrict__ proof, const uint64_t *__restrict__ zrows, size_t rowlen,
                                                        uint64_t *__restrict__ leaf_vals, size_t leaf_pitch, const VOpen *__restrict__ gopens,
                                                        const uint8_t *__restrict__ guvals, size_t gupitch, const uint32_t *__restrict__ b2a_urefs) {
    const uint32_t leaf = blockIdx.x * 32 + (threadIdx.x >> 3), slot = 8 * blockIdx.y + (threadIdx.x & 7);
    if (leaf >= n_inputs + n_mul) return;
    uint64_t v;
    if (leaf < n_inputs) v = z_verify_leaf_input(items[input_item[leaf]], leaf, opens[slot], proof, zrows, rowlen, slot);
    else {
        const uint32_t t = mul_pos[leaf - n_inputs];
        if (items[t].kind == ITEM_B2A)
            v = z_verify_leaf_b2a(items[t], opens[slot], gopens[slot], proof, zrows, rowlen, slot, guvals + (size_t)slot * gupitch, b2a_urefs);
        else v = z_verify_leaf_kappa(items[t], recon_idx[t], opens[slot], proof, zrows, rowlen, slot);
    }
    leaf_vals[(size_t)slot * leaf_pitch + leaf] = v;
}

__global__ void __launch_bounds__(256) k_zverify_items_online(const ZItem *__restrict__ items, const uint32_t *__restrict__ recon_idx, uint32_t n_items,
                                                              const ZOpen *__restrict__ opens, const uint8_t *__restrict__ proof,
                                                              const uint64_t *__restrict__ zrows, size_t rowlen, const uint64_t *__restrict__ uvals,
                                                              size_t upitch, uint8_t *__restrict__ on, size_t pitch, int *not_okay) {
    const uint32_t t = blockIdx.x * 32 + (threadIdx.x >> 3), slot = 8 * blockIdx.y + (threadIdx.x & 7);
    if (t >= n_items) return;
    int flag = 0;
    z_verify_online(items[t], recon_idx[t], opens[slot], proof, zrows, rowlen, slot, uvals + (size_t)slot * upitch, on + (size_t)slot * pitch, &flag);
    if (flag) atomicOr(not_okay, 1);
}

// preprocessing stream of an opened repetition = the proof's corrections (online.rs:169-174)
__global__ void __launch_bounds__(256) k_zverify_items_pre(uint32_t n_mul, const ZOpen *__restrict__ opens, const uint8_t *__restrict__ proof,
                                                           uint8_t *__restrict__ pre, size_t pitch) {
    const uint32_t j = blockIdx.x * 32 + (threadIdx.x >> 3), slot = 8 * blockIdx.y + (threadIdx.x & 7);
    if (j >= n_mul) return;
    const ZOpen &o = opens[slot];
    put64(pre + (size_t)slot * pitch + 8ull * j, z_packed(proof, o.off_corrs, o.n_corrs, o.len_corrs, j));
}

void launch_zverify_leaves(const DevZProgram &Z, const ZOpen *opens, const uint8_t *proof, const uint64_t *zrows, size_t rowlen, uint32_t n_slots,
                           uint64_t *leaf_vals, size_t leaf_pitch, const VOpen *gopens, const uint8_t *guvals, size_t gupitch,
                           const uint32_t *b2a_urefs, cudaStream_t st) {
    if (!Z.n_leaves) return;
    k_zverify_leaves<<<dim3((Z.n_leaves + 31) / 32, n_slots / 8), 256, 0, st>>>(Z.items, Z.input_item, Z.mul_pos, Z.recon_idx, Z.n_inputs, Z.n_corr, opens, proof,
                                                                             zrows, rowlen, leaf_vals, leaf_pitch, gopens, guvals, gupitch, b2a_urefs);
}

void launch_zverify_items(const DevZProgram &Z, const ZOpen *opens, const uint8_t *proof, const uint64_t *zrows, size_t rowlen, uint32_t n_slots,
                          const uint64_t *uvals, size_t upitch, uint8_t *on, size_t pitch_on, uint8_t *pre, size_t pitch_pre, int *not_okay,
                          cudaStream_t st) {
    if (Z.n_items)
        k_zverify_items_online<<<dim3((Z.n_items + 31) / 32, n_slots / 8), 256, 0, st>>>(Z.items, Z.recon_idx, Z.n_items, opens, proof, zrows, rowlen, uvals,
                                                                                      upitch, on, pitch_on, not_okay);
    if (Z.n_corr) k_zverify_items_pre<<<dim3((Z.n_corr + 31) / 32, n_slots / 8), 256, 0, st>>>(Z.n_corr, opens, proof, pre, pitch_pre);
}

// =====================================================================================================================
//  ZK7  extraction of the Z64 vectors of the opened repetitions (src/transcript/prover.rs:57-175 with
//       src/algebra/z64/share.rs:37-49 and recon.rs:46-66).  The vectors sit at arbitrary byte alignment inside the bincode
//       proof, so a thread writes one ALIGNED 8-byte word of the destination, spliced from the two elements that straddle
//       it; only the first and last word of a vector fall back to byte stores (their other bytes belong to neighbours).
//       grid.y = repetition of the shard, grid.x covers the n + 1 words of each of the three vectors.
// =====================================================================================================================
template <typename F>
__device__ __forceinline__ void put_spliced(uint8_t *D, uint64_t n, uint64_t w, F src) {  // w in [0, n]
    if (n == 0) return;
    const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(D) & 7);
    if (a == 0) {
        if (w < n) *reinterpret_cast<uint64_t *>(D + 8 * w) = src(w);
        return;
    }
    uint8_t *D0 = D - a;
    if (w == 0) {
        const uint64_t v = src(0);
        for (uint32_t i = a; i < 8; i++) D0[i] = (uint8_t)(v >> (8 * (i - a)));
    } else if (w == n) {
        const uint64_t v = src(n - 1) >> (8 * (8 - a));
        for (uint32_t i = 0; i < a; i++) D0[8 * n + i] = (uint8_t)(v >> (8 * i));
    } else {
        *reinterpret_cast<uint64_t *>(D0 + 8 * w) = (src(w - 1) >> (8 * (8 - a))) | (src(w) << (8 * a));
    }
}

__global__ void __launch_bounds__(256) k_zextract(const uint32_t *__restrict__ recon_off, const uint32_t *__restrict__ input_off, uint32_t n_recon,
                                                  uint32_t n_corr, uint32_t n_inputs, ZExtractArgs a) {
    const uint32_t lrep = blockIdx.y, rep = a.first_rep + lrep;
    const uint32_t omit = a.omit_of_rep[rep];
    if (omit >= RV_PLAYERS) return;
    uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint8_t *on = a.on + (size_t)lrep * a.pitch_on, *pre = a.pre + (size_t)lrep * a.pitch_pre;
    uint8_t *z = a.proof + a.z_base + 8 + (size_t)a.rank_of_rep[rep] * a.sz_on_z;
    const uint64_t nr = n_recon, nc = n_corr, ni = n_inputs;
    if (w <= nr) {  // recons: the unopened player's u64 of every broadcast share
        put_spliced(z + 137, nr, w, [&](uint64_t e) { return *reinterpret_cast<const uint64_t *>(on + recon_off[e] + 8 * omit); });
        return;
    }
    w -= nr + 1;
    if (w <= nc) {  // corrs: the preprocessing stream itself
        put_spliced(z + 145 + 8 * nr, nc, w, [&](uint64_t e) { return *reinterpret_cast<const uint64_t *>(pre + 8 * e); });
        return;
    }
    w -= nc + 1;
    if (w <= ni) put_spliced(z + 153 + 8 * (nr + nc), ni, w, [&](uint64_t e) { return *reinterpret_cast<const uint64_t *>(on + input_off[e]); });
}

void launch_zextract(const DevZProgram &Z, const ZExtractArgs &a, cudaStream_t st) {
    const uint64_t words = (uint64_t)Z.n_recon + Z.n_corr + Z.n_inputs + 3;
    if (words == 3) return;
    k_zextract<<<dim3((unsigned)((words + 255) / 256), a.nreps), 256, 0, st>>>(Z.recon_off, Z.input_off, Z.n_recon, Z.n_corr, Z.n_inputs, a);
}

// Opt-in shared memory of the Z64 mask generator, and every Z64 kernel loaded up front (see configure_kernels).
int configure_zkernels() {
    cudaError_t e = cudaFuncSetAttribute(k_zmask_gen_tt, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 256 * 32 * 4);
    cudaFuncAttributes at;
    for (const void *fn : {(const void *)k_zlinear_level, (const void *)k_zvalues, (const void *)k_zitems_online, (const void *)k_zitems_pre,
                           (const void *)k_zverify_leaves, (const void *)k_zverify_items_online, (const void *)k_zverify_items_pre, (const void *)k_zextract})
        if (e == cudaSuccess) e = cudaFuncGetAttributes(&at, fn);
    return (int)e;
}

}  // namespace rv
