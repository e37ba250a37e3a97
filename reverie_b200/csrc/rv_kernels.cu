// rv_kernels.cu -- the sm_100a kernels of the KKW prover.  Compiled with -gencode arch=compute_100a,code=sm_100a.
// Integer / bitwise work only: no tensor cores.  See DESIGN.md section 5 for the roofline of each kernel.
#include <cuda_pipeline.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "rv_kernels.cuh"
#include "rv_planes.cuh"

namespace rv {

// =====================================================================================================================
//  K1  key setup: one warp per slice (32 PRG streams = 4 repetitions x 8 players = one u32 half of a share word)
// =====================================================================================================================
// Lane q owns the stream that lives at bit q of the slice word: stream index 31-q = 8*rep_in_slice + player.
// Slice w: packed instance w/2; odd w = high u32 (repetitions 0..3), even w = low u32 (repetitions 4..7).
// Output: rk_plain [45][32 * nslices] -- the 44 round-key words of every stream (stream = 8 * rep + player) and a row of
// all-ones / zero "stream is active" words (the verifier's unopened player stays zero, src/generator/batch.rs:31-34).
__global__ void __launch_bounds__(128) k_key_setup(const uint8_t *__restrict__ seeds, const uint8_t *__restrict__ pkeys_in,
                                                   const uint8_t *__restrict__ mode, const uint8_t *__restrict__ omit, uint32_t nslices,
                                                   uint8_t *__restrict__ pkeys_out, uint32_t *__restrict__ rk_plain, int *__restrict__ bad,
                                                   uint32_t n_flags, size_t flag_stride) {
    // first kernel of every proof / verification: clear the "an AssertZero failed" flag of each proof of the session
    if (bad != nullptr && blockIdx.x == 0 && threadIdx.x < n_flags) *reinterpret_cast<int *>(reinterpret_cast<uint8_t *>(bad) + threadIdx.x * flag_stride) = 0;
    __shared__ uint32_t sbox32[64];  // the S-box as a byte table, built from the netlist (4 entries per thread)
    if (threadIdx.x < 64) {
        const uint32_t b = 4 * threadIdx.x;
        sbox32[threadIdx.x] = sub_word(b | ((b + 1) << 8) | ((b + 2) << 16) | ((b + 3) << 24));
    }
    __syncthreads();
    const TableSubWord sw{reinterpret_cast<const uint8_t *>(sbox32)};
    const uint32_t lane = threadIdx.x & 31, w = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (w >= nslices) return;
    const uint32_t rep = slice_rep(w, lane), p = slice_player(lane);
    uint32_t rk[44];
    const bool active = key_setup_stream(rep, p, seeds, pkeys_in, mode, omit, pkeys_out, rk, sw);
    const uint32_t ns = nslices * 32, sidx = 8 * rep + p;
#pragma unroll
    for (int q = 0; q < 44; q++) rk_plain[(size_t)q * ns + sidx] = rk[q];
    rk_plain[(size_t)44 * ns + sidx] = active ? 0xFFFFFFFFu : 0u;
}

void launch_key_setup(const uint8_t *seeds, const uint8_t *pkeys_in, const uint8_t *mode, const uint8_t *omit, uint32_t nslices,
                      uint8_t *pkeys_out, uint32_t *rk_plain, cudaStream_t st, int *bad, uint32_t n_flags, size_t flag_stride) {
    k_key_setup<<<(nslices + 3) / 4, 128, 0, st>>>(seeds, pkeys_in, mode, omit, nslices, pkeys_out, rk_plain, bad, n_flags, flag_stride);
}

// =====================================================================================================================
//  K2  mask generation: AES-128-CTR of every PRG stream, transposed into packed share words
// =====================================================================================================================
// T-table AES.  A warp = one slice: lane q runs AES-128 for the PRG stream that lives at bit q of the
//      slice word (its 44 round-key words in registers), one counter block at a time, with Te0 / Te2 replicated once per
//      shared-memory bank (64 KB; Te1 / Te3 are byte rotations), so the 160 data-dependent lookups per block never conflict.
//      Five shuffle-exchange stages per keystream word then transpose the warp's 32 x 128 keystream bits into the 128 share
//      words of the slice (what the reference's AVX2 movemask transpose produces, src/algebra/gf2/domain.rs:66-173), and a
//      small shared-memory tile regroups the CTA's 16 slices so every mask row leaves as one 64-byte segment.
//      Cost per block and lane: ~190 LDS/SHFL + ~380 ALU-pipe instructions, against ~640 LOP3 bitsliced.
// A CTA = 16 warps = NS slices x NB counter blocks (NS * NB = 16): full shards give every warp its own slice (NS = 16); the small
// shards of a proof spread over 8 or 16 GPUs (8 or 4 slices) let the spare warps take the next counter blocks instead of idling.
constexpr int GT_WARPS = 16, GT_THREADS = 32 * GT_WARPS;
constexpr size_t GT_TILE_BYTES = 512 * (4 + 4) * 4;  // the largest tile: NS = 4 -> 4 blocks x 128 masks, pitch NS + 4 words
constexpr size_t GT_SMEM2 = 2 * 256 * 32 * 4 + GT_TILE_BYTES, GT_SMEM4 = 4 * 256 * 32 * 4 + GT_TILE_BYTES;
// FOUR = false: Te0 / Te2 in 64 KB (Te1 / Te3 by PRMT rotation): leaves room for a mask-VM CTA on the same SM -- small proofs
// in flight.  FOUR = true: all four tables (128 KB), 72 fewer ALU instructions per block -- circuits big enough to own the chip.
template <bool FOUR>
struct SmemTe {
    const uint8_t *base;  // entry x at byte 256 x: [0, 128) = Te0[x] once per lane, [128, 256) = Te2[x]; FOUR: Te1 / Te3 64 KB further
    uint32_t lane4;
    __device__ __forceinline__ uint32_t operator()(int t, uint32_t w, int b) const {
        const uint32_t off = __byte_perm(w, lane4, 0x7604 | (b << 4));  // (byte b of w) << 8 | 4 * lane
        if (FOUR) return *reinterpret_cast<const uint32_t *>(base + (t & 1) * 65536 + (t >> 1) * 128 + off);
        const uint32_t v = *reinterpret_cast<const uint32_t *>(base + (t >> 1) * 128 + off);
        return (t & 1) ? __byte_perm(v, v, 0x2103) : v;  // Te1 = rotl(Te0, 8), Te3 = rotl(Te2, 8)
    }
};
// 32 x 32 bit transpose across the lanes of a warp: result bit q of lane l = input bit l of lane q.
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, uint32_t lane) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const uint32_t m = d == 16 ? 0x0000FFFFu : d == 8 ? 0x00FF00FFu : d == 4 ? 0x0F0F0F0Fu : d == 2 ? 0x33333333u : 0x55555555u;
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, d);
        x = (lane & d) ? ((x & ~m) | ((y >> d) & m)) : ((x & m) | ((y << d) & ~m));
    }
    return x;
}

// mask_base: PRG index of row 0 (streaming segments continue the streams where the previous segment stopped; 0 otherwise).
// groups_per_cta: a CTA walks that many groups of NB consecutive counter blocks.
template <bool FOUR, int NS>
__global__ void __launch_bounds__(GT_THREADS, 1) k_mask_gen_tt(const uint32_t *__restrict__ rk_plain, uint32_t nslices, uint32_t n_masks,
                                                               uint32_t groups_per_cta, uint32_t *__restrict__ rows32,
                                                               uint64_t *__restrict__ fresh_pm, size_t pitch_pm, bool pm_pairs, uint64_t mask_base) {
    constexpr int NB = GT_WARPS / NS, PITCH = NS + 4, QUADS = NS / 4, ROWS = NB * 128;  // tile row = the CTA's slice words + pad (16-byte aligned rows)
    static_assert(NS == 16 || NS == 8 || NS == 4, "slices per CTA");
    static_assert((size_t)ROWS * PITCH * 4 <= GT_TILE_BYTES, "tile size");
    extern __shared__ __align__(16) uint32_t gt_smem[];
    uint32_t *te = gt_smem, *tile = gt_smem + (FOUR ? 4 : 2) * 256 * 32;
    __shared__ uint32_t sbox32[64];
    const uint32_t tid = threadIdx.x, lane = tid & 31, wv = tid >> 5, sl = wv % NS, bl = wv / NS;
    if (tid < 64) {
        const uint32_t b = 4 * tid;
        sbox32[tid] = sub_word(b | ((b + 1) << 8) | ((b + 2) << 16) | ((b + 3) << 24));
    }
    __syncthreads();
    for (uint32_t e = tid; e < 256 * 32; e += GT_THREADS) {
        const uint32_t x = e >> 5, l = e & 31, t0 = te0_entry(reinterpret_cast<const uint8_t *>(sbox32)[x]);
        te[x * 64 + l] = t0;
        te[x * 64 + 32 + l] = (t0 << 16) | (t0 >> 16);
        if (FOUR) {
            te[16384 + x * 64 + l] = (t0 << 8) | (t0 >> 24);
            te[16384 + x * 64 + 32 + l] = (t0 << 24) | (t0 >> 8);
        }
    }
    const uint32_t w0 = blockIdx.y * NS, w = w0 + sl, nstreams = nslices * 32;
    const bool live = w < nslices;
    uint32_t rk[44], act = 0;
    if (live) {
        const uint32_t sidx = 8 * slice_rep(w, lane) + slice_player(lane);
#pragma unroll
        for (int q = 0; q < 44; q++) rk[q] = rk_plain[(size_t)q * nstreams + sidx];
        act = rk_plain[(size_t)44 * nstreams + sidx];
    }
    __syncthreads();
    const SmemTe<FOUR> tab{reinterpret_cast<const uint8_t *>(te), 4 * lane};
    const uint64_t jb0 = mask_base / 128;  // first counter block that holds a mask of this launch
    const uint32_t n_blocks = (uint32_t)((mask_base + n_masks + 127) / 128 - jb0), n_groups = (n_blocks + NB - 1) / NB;
    const uint32_t g_end = min(n_groups, (blockIdx.x + 1) * groups_per_cta);
    const int64_t shift = (int64_t)(jb0 * 128) - (int64_t)mask_base;  // row of mask m of local block j = j * 128 + m + shift (may be < 0: a mask of the previous segment)
#pragma unroll 1
    for (uint32_t g = blockIdx.x * groups_per_cta; g < g_end; g++) {
        const uint32_t j = g * NB + bl;  // this warp's counter block
        if (live && j < n_blocks) {
            uint32_t in[4], o[4];
            ctr_block_words((uint32_t)(jb0 + j), in);
            tt_aes128_encrypt(rk, in[0], in[1], in[2], in[3], tab, o);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                // lane now holds plane k = 32 q + lane of the slice = keystream byte B = k / 8, bit b = k % 8 -> mask 8 B + 7 - b of the block
                const uint32_t t = warp_transpose32(o[q] & act, lane);
                const uint32_t k = 32 * q + lane, m = (k & ~7u) | (7 - (k & 7));
                tile[(bl * 128 + m) * PITCH + sl] = t;
            }
        }
        __syncthreads();
        const int64_t i0 = (int64_t)g * ROWS + shift;  // row of tile row 0 (rows of blocks past the end fall behind n_masks)
        {  // row-major share tensor: thread = (mask, 4 slices) -> one 16-byte store; a row's NS slices are contiguous
            for (uint32_t e = tid; e < ROWS * QUADS; e += GT_THREADS) {
                const uint32_t m = e / QUADS, q4 = 4 * (e % QUADS);
                const int64_t i = i0 + m;
                if (i >= 0 && i < (int64_t)n_masks && w0 + q4 < nslices) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(tile + m * PITCH + q4);
                    if (nslices >= 4) *reinterpret_cast<uint4 *>(rows32 + i * nslices + w0 + q4) = v;
                    else *reinterpret_cast<uint2 *>(rows32 + i * nslices) = make_uint2(v.x, v.y);  // a shard of one packed instance: rows of 8 bytes
                }
            }
        }
        if (fresh_pm != nullptr) {  // instance-major copy for the mask VM: NS / 2 instances x the group's masks, u64 each
            if (pm_pairs) {  // two instances interleaved ([instance pair][mask][2]): a VM CTA that runs two columns loads 16 bytes at once
                for (uint32_t e = tid; e < (NS / 4) * ROWS; e += GT_THREADS) {
                    const uint32_t p = e / ROWS, m = e % ROWS;
                    const int64_t i = i0 + m;
                    if (i >= 0 && i < (int64_t)n_masks && w0 + 4 * p < nslices) {
                        const uint4 v = *reinterpret_cast<const uint4 *>(tile + m * PITCH + 4 * p);
                        *reinterpret_cast<uint4 *>(fresh_pm + ((size_t)((w0 >> 2) + p) * pitch_pm + (uint64_t)i) * 2) = v;
                    }
                }
            } else {
                for (uint32_t e = tid; e < (NS / 2) * ROWS; e += GT_THREADS) {
                    const uint32_t p = e / ROWS, m = e % ROWS;
                    const int64_t i = i0 + m;
                    if (i >= 0 && i < (int64_t)n_masks && w0 + 2 * p < nslices) {
                        const uint2 v = *reinterpret_cast<const uint2 *>(tile + m * PITCH + 2 * p);
                        fresh_pm[(size_t)((w0 >> 1) + p) * pitch_pm + (uint64_t)i] = ((uint64_t)v.y << 32) | v.x;
                    }
                }
            }
        }
        __syncthreads();
    }
}

template <bool FOUR, int NS>
static void mask_gen_launch(dim3 grid, cudaStream_t st, const uint32_t *rk_plain, uint32_t nslices, uint32_t n_masks, uint32_t per, uint32_t *rows32,
                            uint64_t *fresh_pm, size_t pitch_pm, bool pm_pairs, uint64_t mask_base) {
    k_mask_gen_tt<FOUR, NS><<<grid, GT_THREADS, FOUR ? GT_SMEM4 : GT_SMEM2, st>>>(rk_plain, nslices, n_masks, per, rows32, fresh_pm, pitch_pm, pm_pairs, mask_base);
}

void launch_mask_gen_tt(const uint32_t *rk_plain, uint32_t nslices, uint32_t n_masks, uint64_t *rows, uint64_t *fresh_pm, size_t pitch_pm, int n_sms,
                        cudaStream_t st, uint32_t busy_sms, uint32_t share, bool pm_pairs, uint64_t mask_base) {
    if (n_masks == 0) return;
    const uint32_t ns = nslices <= 4 ? 4 : nslices <= 8 ? 8 : 16, nb = GT_WARPS / ns;
    const uint32_t n_blocks = (uint32_t)((mask_base + n_masks + 127) / 128 - mask_base / 128), n_groups = (n_blocks + nb - 1) / nb, gy = (nslices + ns - 1) / ns;
    // One CTA per SM (104 registers x 512 threads), so the grid is sized to finish in ONE wave over the SMs this launch can count
    // on: the value plane's CTAs (busy_sms, one SM each for the whole mask pipeline) and the other sessions of the batch (share)
    // take theirs -- a grid a few CTAs larger than the free SMs would run a second, almost empty wave and double the kernel.
    share = std::max(1u, share);
    const uint32_t busy_all = busy_sms * share;  // every session of the batch runs its own value plane
    const uint32_t avail = std::max(8u, ((uint32_t)n_sms > busy_all ? (uint32_t)n_sms - busy_all : 0u) / share);
    // up to 96 counter blocks per warp: an 8-proof session (32 CTA rows) then fits one wave of 128 CTAs (measured alone: 286 -> 202 us;
    // the 4-session step is throughput-bound and does not care)
    const uint32_t want_x = std::max(1u, avail / gy);
    const uint32_t per = std::min(96u, std::max(4u, (n_groups + want_x - 1) / want_x));
    dim3 grid((n_groups + per - 1) / per, gy);
    uint32_t *rows32 = reinterpret_cast<uint32_t *>(rows);
    const bool four = (uint64_t)n_blocks * gy >= 64ull * n_sms;  // enough work for many waves: the mask generator owns the chip
#define RV_MG(F, N) mask_gen_launch<F, N>(grid, st, rk_plain, nslices, n_masks, per, rows32, fresh_pm, pitch_pm, pm_pairs, mask_base)
    if (ns == 16) four ? RV_MG(true, 16) : RV_MG(false, 16);
    else if (ns == 8) four ? RV_MG(true, 8) : RV_MG(false, 8);
    else four ? RV_MG(true, 4) : RV_MG(false, 4);
#undef RV_MG
}

// =====================================================================================================================
//  Instruction ring: a level-sorted program streamed global -> shared memory with 1-D bulk async copies (TMA,
//  cp.async.bulk + mbarrier complete_tx), RING_NB chunks in flight.  Thread 0 issues; every thread consumes.
// =====================================================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

constexpr size_t SMEM_DYN_CAP = 226 * 1024;  // of the 227 KB a CTA may opt into; the rest covers static __shared__

// CHUNK_BYTES of program per chunk.  Every thread calls begin_chunk(c) for c = 0, 1, 2, ... in order (uniformly), and a
// CTA barrier separates the last read of chunk c-1 from begin_chunk(c): the host compiler forces one at every chunk end.
// All bookkeeping is 32-bit and precomputed: the per-chunk cost sits on the critical path of the level-synchronous kernels.
template <uint32_t CHUNK_BYTES, int RING_NB>  // RING_NB chunks in flight
struct ChunkStream {
    static constexpr size_t BYTES = (size_t)RING_NB * CHUNK_BYTES + 64;
    uint32_t buf_s, bars_s;  // shared-space addresses
    const uint8_t *src;
    uint32_t n_chunks, last_bytes;

    __device__ __forceinline__ void issue(uint32_t c) {  // c < n_chunks
        const uint32_t bytes = c + 1 == n_chunks ? last_bytes : CHUNK_BYTES;
        const uint32_t slot = c % RING_NB, bar = bars_s + 8 * slot;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(buf_s + slot * CHUNK_BYTES),
                     "l"(src + (size_t)c * CHUNK_BYTES), "r"(bytes), "r"(bar)
                     : "memory");
    }
    __device__ void init(uint8_t *smem, const void *program, uint32_t chunks, uint32_t last_chunk_bytes) {  // all threads; has a CTA barrier
        buf_s = smem_u32(smem);
        bars_s = buf_s + RING_NB * CHUNK_BYTES;
        src = reinterpret_cast<const uint8_t *>(program);
        n_chunks = chunks;
        last_bytes = last_chunk_bytes;
        if (threadIdx.x == 0) {
            for (int i = 0; i < RING_NB; i++) mbar_init(reinterpret_cast<uint64_t *>(smem + RING_NB * CHUNK_BYTES) + i, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0)
            for (uint32_t c = 0; c < RING_NB && c < n_chunks; c++) issue(c);
    }
    // Returns the shared-space address of the chunk's image.  Thread 0 also re-arms the buffer that chunk c-1 occupied.
    __device__ __forceinline__ uint32_t begin_chunk(uint32_t c) {
        if (threadIdx.x == 0 && c >= 1 && c + RING_NB - 1 < n_chunks) issue(c + RING_NB - 1);
        const uint32_t slot = c % RING_NB, bar = bars_s + 8 * slot, parity = (c / RING_NB) & 1;
        uint32_t ok;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok)
                         : "r"(bar), "r"(parity)
                         : "memory");
        } while (!ok);
        return buf_s + slot * CHUNK_BYTES;
    }
};
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}

// =====================================================================================================================
//  K0  value plane: the mapped LUT program as a step stream (rv_compile.h), one CTA of LUT_STEP threads, one slot per
//      thread per step.  Values live in shared memory (or global when they do not fit).
// =====================================================================================================================
constexpr int VP_THREADS = LUT_STEP;
template <int SPC>  // steps per chunk: 4 or 8 (the stream has a barrier at least every 4 steps)
using LutStream = ChunkStream<SPC * LUT_STEP * (uint32_t)sizeof(LutInstr), 2>;
static_assert(LUT_STEPS_PER_CHUNK_MAX % LUT_STEPS_PER_CHUNK == 0, "chunk sizes must be multiples of the barrier period");

// CTA b evaluates instance b: leaves leaf_vals[b * leaf_pitch + k] -> value id leaf_ids[k]; results to vals_g + b * vals_pitch.
// Prover: one instance, leaves = the witness bits.  Online verifier: one instance per opened repetition (u-plane).
template <bool SMEM_VALS, int SPC>
__global__ void __launch_bounds__(VP_THREADS) k_values(const LutInstr *__restrict__ prog, uint32_t n_steps, const uint32_t *__restrict__ leaf_ids,
                                                       const uint8_t *__restrict__ leaf_vals, size_t leaf_pitch, uint32_t n_leaves,
                                                       uint8_t *__restrict__ vals_out, size_t vals_pitch, uint32_t n_vals) {
    extern __shared__ __align__(128) uint8_t smem[];
    using Stream = LutStream<SPC>;
    const uint32_t n_chunks = (n_steps + SPC - 1) / SPC;
    Stream stream;
    stream.init(smem, prog, n_chunks, (n_steps - (n_chunks ? n_chunks - 1 : 0) * SPC) * LUT_STEP * (uint32_t)sizeof(LutInstr));
    uint8_t *vals_g = vals_out + (size_t)blockIdx.x * vals_pitch;
    const uint8_t *wit = leaf_vals + (size_t)blockIdx.x * leaf_pitch;
    // values: slot n_vals is the scratch target of empty slots.  In shared memory they are addressed as smem[BYTES + id] so
    // that every access is a plain LDS/STS with an immediate offset (no generic-address arithmetic in the hot loop).
    auto ld = [&](uint32_t id) -> uint32_t { return SMEM_VALS ? (uint32_t)smem[Stream::BYTES + id] : (uint32_t)vals_g[id]; };
    auto st = [&](uint32_t id, uint32_t v) {
        if (SMEM_VALS) smem[Stream::BYTES + id] = (uint8_t)v;
        else vals_g[id] = (uint8_t)v;
    };
    const uint32_t tid = threadIdx.x;
    if (tid == 0) st(0, 0);
    for (uint32_t k = tid; k < n_leaves; k += VP_THREADS) st(leaf_ids[k], wit[k] & 1);
    __syncthreads();
    for (uint32_t c = 0; c < n_chunks; c++) {
        const uint32_t img = stream.begin_chunk(c) + tid * (uint32_t)sizeof(LutInstr);
        const uint32_t nst = min((uint32_t)SPC, n_steps - c * SPC);
        uint4 u0[SPC], u1[SPC];
        uint2 u2[SPC];
#pragma unroll
        for (int k = 0; k < SPC; k++)
            if (k < (int)nst) {
                const uint32_t p = img + k * LUT_STEP * (uint32_t)sizeof(LutInstr);
                u0[k] = lds128(p);       // {dst, in0, in1, in2}
                u1[k] = lds128(p + 16);  // {in3, in4, in5, flags}
                u2[k] = lds64(p + 32);   // {tt lo, tt hi}
            }
#pragma unroll
        for (int k = 0; k < SPC; k++)
            if (k < (int)nst) {
                const uint32_t idx = ld(u0[k].y) | (ld(u0[k].z) << 1) | (ld(u0[k].w) << 2) | (ld(u1[k].x) << 3) | (ld(u1[k].y) << 4) | (ld(u1[k].z) << 5);
                const uint64_t tt = ((uint64_t)u2[k].y << 32) | u2[k].x;
                st(u0[k].x, (uint32_t)(tt >> idx) & 1u);
                if (u1[k].w & LUT_F_BAR) __syncthreads();
            }
    }
    __syncthreads();
    if (SMEM_VALS) {
        for (uint32_t i = tid; i < n_vals; i += VP_THREADS) vals_g[i] = smem[Stream::BYTES + i];
    }
}

size_t launch_values(const LutInstr *steps, uint32_t n_steps, const uint32_t *leaf_ids, const uint8_t *leaf_vals, size_t leaf_pitch,
                     uint32_t n_leaves, uint8_t *vals, size_t vals_pitch, uint32_t n_vals, uint32_t n_instances, cudaStream_t st) {
    const size_t cap = SMEM_DYN_CAP, vbytes = ((size_t)n_vals + 1 + 15) & ~(size_t)15;
    constexpr int S8 = (int)LUT_STEPS_PER_CHUNK_MAX, S4 = (int)LUT_STEPS_PER_CHUNK;
    if (LutStream<S8>::BYTES + vbytes <= cap) {
        k_values<true, S8><<<n_instances, VP_THREADS, LutStream<S8>::BYTES + vbytes, st>>>(steps, n_steps, leaf_ids, leaf_vals, leaf_pitch, n_leaves, vals, vals_pitch, n_vals);
        return LutStream<S8>::BYTES + vbytes;
    }
    if (LutStream<S4>::BYTES + vbytes <= cap) {  // a smaller instruction ring leaves room for the values (the verifier's u-plane of SHA-256)
        k_values<true, S4><<<n_instances, VP_THREADS, LutStream<S4>::BYTES + vbytes, st>>>(steps, n_steps, leaf_ids, leaf_vals, leaf_pitch, n_leaves, vals, vals_pitch, n_vals);
        return LutStream<S4>::BYTES + vbytes;
    }
    k_values<false, S8><<<n_instances, VP_THREADS, LutStream<S8>::BYTES, st>>>(steps, n_steps, leaf_ids, leaf_vals, leaf_pitch, n_leaves, vals, vals_pitch, n_vals);
    return LutStream<S8>::BYTES;
}

// Wide circuits: thread = one LUT of the level; values in global memory (L2-resident between the per-level launches).
__global__ void __launch_bounds__(256) k_values_leaves(const uint32_t *__restrict__ leaf_ids, const uint8_t *__restrict__ wit, uint32_t n_leaves,
                                                       uint8_t *__restrict__ vals) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) vals[0] = 0;
    if (k < n_leaves) vals[leaf_ids[k]] = wit[k] & 1;
}
// 2-input gate on byte values: v[dst] = (v[a >> 1] ^ (a & 1)) OP (v[b >> 1] ^ (b & 1))
__device__ __forceinline__ void eval_vgate(const VGate g, uint8_t *v) {
    const uint32_t x = (uint32_t)v[g.a >> 1] ^ (g.a & 1), y = (uint32_t)v[g.b >> 1] ^ (g.b & 1);
    v[g.dst] = (uint8_t)((g.op ? (x & y) : (x ^ y)) & 1);
}
__global__ void __launch_bounds__(256) k_values_level(const VGate *__restrict__ gates, uint32_t n, uint8_t *vals) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n) eval_vgate(gates[g], vals);
}

int launch_values_wide(const DevProgram &P, const uint32_t *off_host, const uint8_t *wit, uint8_t *vals, cudaStream_t st) {
    k_values_leaves<<<(std::max(P.n_inputs, 1u) + 255) / 256, 256, 0, st>>>(P.input_vid, wit, P.n_inputs, vals);
    for (uint32_t l = 0; l < P.n_wlevels; l++) {
        const uint32_t n = off_host[l + 1] - off_host[l];
        if (n) k_values_level<<<(n + 255) / 256, 256, 0, st>>>(P.wgates + off_host[l], n, vals);
    }
    return 1 + (int)P.n_wlevels;
}

__global__ void __launch_bounds__(256) k_uvalues_leaves(const uint32_t *__restrict__ leaf_ids, const uint8_t *__restrict__ leaf_vals, size_t leaf_pitch,
                                                        uint32_t n_leaves, uint8_t *__restrict__ uvals, size_t upitch) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    uint8_t *v = uvals + (size_t)blockIdx.y * upitch;
    if (k == 0) v[0] = 0;
    if (k < n_leaves) v[leaf_ids[k]] = leaf_vals[(size_t)blockIdx.y * leaf_pitch + k] & 1;
}
__global__ void __launch_bounds__(256) k_uvalues_level(const VGate *__restrict__ gates, uint32_t n, uint8_t *uvals, size_t upitch) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n) eval_vgate(gates[g], uvals + (size_t)blockIdx.y * upitch);
}

int launch_uvalues_wide(const DevProgram &P, const uint32_t *off_host, const uint8_t *leaf_vals, size_t leaf_pitch, uint32_t n_leaves, uint8_t *uvals,
                        size_t upitch, uint32_t n_instances, cudaStream_t st) {
    k_uvalues_leaves<<<dim3((std::max(n_leaves, 1u) + 255) / 256, n_instances), 256, 0, st>>>(P.vleaf_ids, leaf_vals, leaf_pitch, n_leaves, uvals, upitch);
    for (uint32_t l = 0; l < P.n_vwlevels; l++) {
        const uint32_t n = off_host[l + 1] - off_host[l];
        if (n) k_uvalues_level<<<dim3((n + 255) / 256, n_instances), 256, 0, st>>>(P.vwgates + off_host[l], n, uvals, upitch);
    }
    return 1 + (int)P.n_vwlevels;
}

// =====================================================================================================================
//  K3  mask plane: row[dst] = row[a] ^ row[b], level by level
// =====================================================================================================================
constexpr int LIN_THREADS = 256;
constexpr int VM_THREADS = VM_STEP;

// (a) VM over shared-memory cells: one CTA per packed instance (cell = the u64 share word: 8 repetitions x 8 players), one
//     slot per thread per step.  The dependent chain per level is LDS -> XOR -> STS -> barrier; fresh rows arrive through
//     cp.async LOADs issued VM_DELTA levels early; only rows the item plane needs are written back.  Every CTA streams the
//     whole program from L2, so the 20-byte instruction (16-bit cell ids) is what sets the pace.
constexpr uint32_t VM_CHUNK_BYTES = VM_STEPS_PER_CHUNK * VM_STEP * (uint32_t)sizeof(VmInstr);
using VmStream = ChunkStream<VM_CHUNK_BYTES, 2>;
static_assert(VM_THREADS == (int)VM_STEP, "one slot per thread");
static_assert(VM_CHUNK_BYTES % 16 == 0 && (VM_STEP * sizeof(VmInstr)) % 16 == 0, "bulk copies move multiples of 16 bytes");
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// COLS = columns (packed instances) one CTA runs: 1 (u64 cells) or 2 adjacent ones (16-byte cells: the same decoded instruction
// moves twice the payload, so the program is streamed from L2 half as often per proof and the cell traffic uses 128-bit LDS/STS).
template <int COLS>
struct VmCell;
template <>
struct VmCell<1> {
    using T = uint64_t;
    static __device__ __forceinline__ T zero() { return 0; }
    static __device__ __forceinline__ T x(T a, T b) { return a ^ b; }
};
template <>
struct VmCell<2> {
    using T = ulonglong2;
    static __device__ __forceinline__ T zero() { return make_ulonglong2(0, 0); }
    static __device__ __forceinline__ T x(T a, T b) { return make_ulonglong2(a.x ^ b.x, a.y ^ b.y); }
};

template <int COLS>
__global__ void __launch_bounds__(VM_THREADS) k_mask_vm(const VmInstr *__restrict__ prog, uint32_t n_steps, const uint64_t *__restrict__ fresh_pm,
                                                        size_t pitch_fresh, uint64_t *__restrict__ rows, uint32_t npi, uint32_t scratch) {
    using Cell = typename VmCell<COLS>::T;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t n_chunks = (n_steps + VM_STEPS_PER_CHUNK - 1) / VM_STEPS_PER_CHUNK;
    VmStream stream;
    stream.init(smem, prog, n_chunks, (n_steps - (n_chunks ? n_chunks - 1 : 0) * VM_STEPS_PER_CHUNK) * VM_STEP * (uint32_t)sizeof(VmInstr));
    Cell *cells = reinterpret_cast<Cell *>(smem + VmStream::BYTES);
    const uint32_t tid = threadIdx.x, q = blockIdx.x;
    // LOADs read the instance-major copy of the fresh masks (COLS = 2: the pair-interleaved one), so a warp's 32 requests (sorted
    // by row inside a level) fall into a few 128-byte lines.  Exported rows go straight into the row-major share tensor: the
    // tensor of a VM-sized circuit lives in L2, which merges the instances' pieces of a row before the item plane reads it.
    const Cell *src = reinterpret_cast<const Cell *>(fresh_pm) + (size_t)q * pitch_fresh;
    uint64_t *dst = rows + (size_t)COLS * q;
    if (tid == 0) cells[0] = VmCell<COLS>::zero();  // cell 0 is the constant zero
    __syncthreads();                                  // (empty slots of the first steps already read it)
    for (uint32_t c = 0; c < n_chunks; c++) {
        const uint32_t img = stream.begin_chunk(c) + tid * (uint32_t)sizeof(VmInstr);  // 5-word stride: conflict-free LDS.32
        const uint32_t nst = min((uint32_t)VM_STEPS_PER_CHUNK, n_steps - c * VM_STEPS_PER_CHUNK);
        uint32_t u[VM_STEPS_PER_CHUNK][5];
#pragma unroll
        for (int k = 0; k < (int)VM_STEPS_PER_CHUNK; k++)
            if (k < (int)nst) {
                const uint32_t p = img + k * VM_STEP * (uint32_t)sizeof(VmInstr);
#pragma unroll
                for (int w = 0; w < 5; w++) u[k][w] = lds32(p + 4 * w);  // {row, dst | flags << 16, in0 | in1 << 16, in2 | in3 << 16, in4 | in5 << 16}
            }
#pragma unroll
        for (int k = 0; k < (int)VM_STEPS_PER_CHUNK; k++)
            if (k < (int)nst) {
                const uint32_t row = u[k][0], d = u[k][1] & 0xFFFFu, fl = u[k][1] >> 16;
                if (fl & VM_F_LOAD) {
                    __pipeline_memcpy_async(cells + d, src + row, sizeof(Cell));
                } else {
                    typedef VmCell<COLS> V;
                    const Cell v = V::x(V::x(V::x(cells[u[k][2] & 0xFFFFu], cells[u[k][2] >> 16]), V::x(cells[u[k][3] & 0xFFFFu], cells[u[k][3] >> 16])),
                                        V::x(cells[u[k][4] & 0xFFFFu], cells[u[k][4] >> 16]));
                    if (d != scratch) cells[d] = v;  // (empty slots and export-only XORs name the scratch cell: nobody reads it)
                    if (row != VM_ROW_NONE) *reinterpret_cast<Cell *>(dst + (size_t)row * npi) = v;
                }
                if (fl & VM_F_LEVEL_END) {  // one cp.async group per level; LOADs of level L-2 are complete before level L starts
                    __pipeline_commit();
                    __pipeline_wait_prior(VM_DELTA - 1);
                }
                if (fl & VM_F_BAR) __syncthreads();
            }
    }
}

// (b) fallback for networks whose live set does not fit in shared memory: one CTA per packed instance walks all levels
//     over the share tensor itself (CTA barrier only; columns are independent)
__device__ __forceinline__ uint64_t xor6(const uint64_t *rows, const XGate &g, uint32_t npi, uint32_t pi) {
    uint64_t v = 0;
#pragma unroll
    for (int k = 0; k < 6; k++) v ^= rows[(size_t)g.in[k] * npi + pi];
    return v;
}

__global__ void __launch_bounds__(LIN_THREADS) k_linear_cta(const XGate *__restrict__ gates, const uint32_t *__restrict__ level_off,
                                                            uint32_t n_levels, uint64_t *rows, uint32_t npi) {
    const uint32_t pi = blockIdx.x, tid = threadIdx.x;
    for (uint32_t l = 0; l < n_levels; l++) {
        const uint32_t s = level_off[l], e = level_off[l + 1];
        for (uint32_t g = s + tid; g < e; g += LIN_THREADS) {
            const XGate gt = gates[g];
            rows[(size_t)gt.dst * npi + pi] = xor6(rows, gt, npi, pi);
        }
        __syncthreads();
    }
}

// (c) wide levels: one launch per level, thread = (gate, packed instance)
__global__ void __launch_bounds__(256) k_linear_level(const XGate *__restrict__ gates, uint32_t n, uint64_t *rows, uint32_t npi) {
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t g = gid / npi;
    const uint32_t pi = (uint32_t)(gid % npi);
    if (g >= n) return;
    const XGate gt = gates[g];
    rows[(size_t)gt.dst * npi + pi] = xor6(rows, gt, npi, pi);
}

static size_t vm_smem_bytes(const DevProgram &P, int cols = 1) { return VmStream::BYTES + ((size_t)P.vm_cells + 1) * 8 * cols; }
static_assert(VmStream::BYTES == VM_STREAM_SMEM && SMEM_DYN_CAP == VM_SMEM_CAP, "rv_compile.h mirrors the VM's shared-memory budget (vm_columns)");
bool linear_uses_vm(const DevProgram &P) {
    if (P.n_llevels == 0 || (double)P.n_xgates / P.n_llevels >= 4096.0) return false;
    return P.n_vm_steps && P.vm_cells < VM_CELL_MASK && vm_smem_bytes(P) <= SMEM_DYN_CAP;
}
// Two columns per VM CTA: when the wider cells still fit and the tensor has an even number of columns (a lone proof stays on one
// column per CTA: it is latency-bound and wants all its instances on their own SMs).
bool linear_vm_pairs(const DevProgram &P, uint32_t npi) { return linear_uses_vm(P) && npi % 2 == 0 && npi > 32 && vm_smem_bytes(P, 2) <= SMEM_DYN_CAP; }

int launch_linear(const DevProgram &P, const uint32_t *off_host, uint64_t *rows, uint32_t npi, const uint64_t *fresh_sm, size_t pitch_fresh,
                  cudaStream_t st, int *which) {
    if (which) *which = -1;
    if (P.n_llevels == 0) return 0;
    const double avg_width = (double)P.n_xgates / P.n_llevels;
    if (avg_width >= 4096.0) {  // wide: per-level launches (~3 us each) keep every SM busy
        if (which) *which = 2;
        for (uint32_t l = 0; l < P.n_llevels; l++) {
            const uint32_t n = off_host[l + 1] - off_host[l];
            const uint64_t threads = (uint64_t)n * npi;
            k_linear_level<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(P.xgates + off_host[l], n, rows, npi);
        }
        return (int)P.n_llevels;
    }
    if (linear_uses_vm(P) && fresh_sm) {
        if (which) *which = 0;
        if (linear_vm_pairs(P, npi)) k_mask_vm<2><<<npi / 2, VM_THREADS, vm_smem_bytes(P, 2), st>>>(P.vm_steps, P.n_vm_steps, fresh_sm, pitch_fresh, rows, npi, P.vm_cells);
        else k_mask_vm<1><<<npi, VM_THREADS, vm_smem_bytes(P), st>>>(P.vm_steps, P.n_vm_steps, fresh_sm, pitch_fresh, rows, npi, P.vm_cells);
        return 1;
    }
    if (which) *which = 1;
    k_linear_cta<<<npi, LIN_THREADS, 0, st>>>(P.xgates, P.xlevel_off, P.n_llevels, rows, npi);
    return 1;
}

// =====================================================================================================================
//  K4  item plane: thread = (8 consecutive stream positions, packed instance); 8x8 byte transpose in registers so each
//      repetition's 8 stream bytes leave as one aligned 64-bit store
// =====================================================================================================================
__global__ void __launch_bounds__(256) k_items_pre(const Item *__restrict__ items, const uint32_t *__restrict__ mul_pos, uint32_t n_pre,
                                                   const uint64_t *__restrict__ rows, uint32_t npi, uint8_t *__restrict__ pre, size_t pitch,
                                                   uint32_t first_pi) {
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t pi = first_pi + (uint32_t)(gid % (npi - first_pi));
    const uint64_t j0 = (gid / (npi - first_pi)) * 8;
    if (j0 >= n_pre) return;
    uint64_t W[8], out[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        W[i] = 0;
        if (j0 + i < n_pre) W[i] = pre_word(items[mul_pos[j0 + i]], rows, npi, pi);
    }
    words_to_stream_bytes(W, out);
#pragma unroll
    for (int r = 0; r < 8; r++) *reinterpret_cast<uint64_t *>(pre + (size_t)(8 * pi + r) * pitch + j0) = out[r];
}

// Tiled item plane (prover).  A CTA computes a tile of T consecutive stream positions for every repetition of the shard:
// lane = (packed instance, position group of 8) so that each mask row is read as one contiguous run of npi * 8 bytes; the
// 8x8 byte transposes happen in registers (words_to_stream_bytes); the tile is staged in shared memory (row = repetition,
// pitch T + 8 bytes: 64-bit stores of consecutive instances fall into distinct banks) and leaves as whole 128-byte lines
// of each repetition's stream.  PRE selects the preprocessing stream (one byte per Mul) instead of the online stream.
constexpr int IT_THREADS = 256;
// npi = packed instances of ONE proof (the tile's geometry); the session may hold several proofs side by side: col0 = first
// column of this proof in the share tensor, row_stride = columns of the whole tensor.
// base: buffer position of item 0 (streaming segments start inside a BLAKE3 chunk whose first bytes the previous segment left
// behind; positions below `base` are written as zero and filled in by the caller afterwards).  0 otherwise.
template <bool PRE>
__device__ __forceinline__ void items_tile_body(uint32_t tile_idx, const Item *__restrict__ items, const uint32_t *__restrict__ mul_pos, uint32_t n,
                                                const uint64_t *__restrict__ rows, uint32_t npi, uint32_t col0, uint32_t row_stride,
                                                const uint8_t *__restrict__ vals, const uint64_t *__restrict__ tvals, uint8_t *__restrict__ out,
                                                size_t pitch, uint32_t T, int *bad, uint32_t base) {
    extern __shared__ __align__(16) uint8_t tile[];
    const uint32_t tid = threadIdx.x, pi = tid % npi, pg0 = tid / npi, pg_step = IT_THREADS / npi, col = col0 + pi;
    const uint32_t tp = T + 8;  // tile pitch in bytes
    const uint64_t T0 = (uint64_t)tile_idx * T;
    int flag = 0;
    for (uint32_t g = pg0; g < T / 8; g += pg_step) {
        const uint64_t t0 = T0 + 8ull * g;
        uint64_t W[8], o[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            W[i] = 0;
            const uint64_t t = t0 + i - base;  // (wraps for positions below base)
            if (t0 + i >= base && t < n)
                W[i] = PRE ? pre_word(items[mul_pos[t]], rows, row_stride, col) : prover_online_word(items[t], rows, row_stride, col, vals, tvals, &flag);
        }
        words_to_stream_bytes(W, o);
#pragma unroll
        for (int r = 0; r < 8; r++) *reinterpret_cast<uint64_t *>(tile + (size_t)(r * npi + pi) * tp + 8 * g) = o[r];
    }
    if (!PRE && flag) atomicOr(bad, 1);
    __syncthreads();
    // write-out: one warp per tile row, 128 bytes (32 lanes x u32) per step
    const uint32_t lane = tid & 31, wrp = tid >> 5, nrows = 8 * npi;
    for (uint32_t row = wrp; row < nrows; row += IT_THREADS / 32) {
        const uint32_t r = row / npi, p = row % npi;
        uint8_t *dst = out + (size_t)(8 * (col0 + p) + r) * pitch + T0;
        const uint8_t *src = tile + (size_t)row * tp;
        for (uint32_t c = 4 * lane; c < T; c += 128) *reinterpret_cast<uint32_t *>(dst + c) = *reinterpret_cast<const uint32_t *>(src + c);
    }
}

// One launch covers both streams: CTAs [0, tiles_on) take tiles of the online stream, the rest tiles of the preprocessing
// stream (small proofs are bound by the number of kernels in flight, not by their work).
// grid.y = proof of the session (its witness / value plane at vals + y * vals_pitch, its flag at bad + y * flag_stride bytes).
__global__ void __launch_bounds__(IT_THREADS) k_items(const Item *__restrict__ items, const uint32_t *__restrict__ mul_pos, uint32_t n_online,
                                                     uint32_t n_pre, uint32_t tiles_on, const uint64_t *__restrict__ rows, uint32_t npi,
                                                     const uint8_t *__restrict__ vals, size_t vals_pitch, const uint64_t *__restrict__ tvals,
                                                     uint8_t *__restrict__ on, size_t pitch_on, uint8_t *__restrict__ pre, size_t pitch_pre, uint32_t T,
                                                     int *bad, size_t flag_stride, uint32_t base_on, uint32_t base_pre) {
    const uint32_t col0 = blockIdx.y * npi, row_stride = gridDim.y * npi;
    if (blockIdx.x < tiles_on)
        items_tile_body<false>(blockIdx.x, items, nullptr, n_online, rows, npi, col0, row_stride, vals + blockIdx.y * vals_pitch, tvals, on, pitch_on, T,
                               reinterpret_cast<int *>(reinterpret_cast<uint8_t *>(bad) + blockIdx.y * flag_stride), base_on);
    else items_tile_body<true>(blockIdx.x - tiles_on, items, mul_pos, n_pre, rows, npi, col0, row_stride, nullptr, nullptr, pre, pitch_pre, T, nullptr, base_pre);
}

static uint32_t items_tile(uint32_t npi) { return std::max(128u, 8u * IT_THREADS / npi); }

void launch_items(const DevProgram &P, const uint64_t *rows, uint32_t npi, const uint8_t *vals, const uint64_t *tvals, uint8_t *on, size_t pitch_on,
                  uint8_t *pre, size_t pitch_pre, int *bad, cudaStream_t st, uint32_t n_proofs, size_t vals_pitch, size_t flag_stride, uint32_t base_on,
                  uint32_t base_pre) {
    const uint32_t T = items_tile(npi);
    const size_t smem = (size_t)8 * npi * (T + 8);
    const uint32_t tiles_on = P.n_online ? (base_on + P.n_online + T - 1) / T : 0, tiles_pre = P.n_pre ? (base_pre + P.n_pre + T - 1) / T : 0;
    if (tiles_on + tiles_pre)
        k_items<<<dim3(tiles_on + tiles_pre, n_proofs), IT_THREADS, smem, st>>>(P.items, P.mul_pos, P.n_online, P.n_pre, tiles_on, rows, npi, vals, vals_pitch,
                                                                                 tvals, on, pitch_on, pre, pitch_pre, T, bad, flag_stride, base_on, base_pre);
}

// Tainted plane: CTA = one packed instance (columns are independent), level-synchronous with CTA barriers only.
__global__ void __launch_bounds__(256) k_tainted(const TGate *__restrict__ gates, const uint32_t *__restrict__ level_off, uint32_t n_levels,
                                                 const uint64_t *__restrict__ rows, uint32_t npi, const uint8_t *__restrict__ vals, uint64_t *tvals) {
    const uint32_t pi = blockIdx.x;
    for (uint32_t l = 0; l < n_levels; l++) {
        const uint32_t e = level_off[l + 1];
        for (uint32_t g = level_off[l] + threadIdx.x; g < e; g += blockDim.x) {
            const TGate gt = gates[g];
            tvals[(size_t)gt.dst * npi + pi] = tainted_eval(gt, rows, npi, pi, vals, tvals);
        }
        __syncthreads();
    }
}

void launch_tainted(const DevProgram &P, const uint64_t *rows, uint32_t npi, const uint8_t *vals, uint64_t *tvals, cudaStream_t st) {
    if (P.n_tlevels) k_tainted<<<npi, 256, 0, st>>>(P.tgates, P.tlevel_off, P.n_tlevels, rows, npi, vals, tvals);
}

// verifier, preprocessing repetitions: recompute the corrections from the seeds (src/transcript/verifier/preprocess.rs:66-69)
void launch_items_pre_range(const DevProgram &P, const uint64_t *rows, uint32_t npi, uint32_t first_pi, uint8_t *pre, size_t pitch_pre, cudaStream_t st) {
    if (!P.n_pre || first_pi >= npi) return;
    const uint64_t threads = (uint64_t)((P.n_pre + 7) / 8) * (npi - first_pi);
    k_items_pre<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(P.items, P.mul_pos, P.n_pre, rows, npi, pre, pitch_pre, first_pi);
}

// =====================================================================================================================
//  K5  transcript hashing
// =====================================================================================================================
// thread = (stream, repetition, chunk): chaining value of one 1 KiB chunk.  Full 64-byte blocks are fetched as four
// 16-byte loads, one block ahead of the compression that consumes them.
struct ChunkJob {
    const uint8_t *stream;
    size_t pitch;
    uint32_t len, n_chunks, nreps;  // bytes / chunks of the buffer handed to this launch
    uint32_t *cvs;
    // streaming: the buffer holds chunks [chunk0, chunk0 + n_chunks) of streams of cv_stride chunks; single = the whole stream is one chunk
    uint32_t chunk0, cv_stride;
    bool single;
};

__device__ __forceinline__ void load_block(const uint4 *p, uint32_t m[16]) {
    const uint4 a = p[0], b = p[1], c = p[2], d = p[3];
    m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w;
    m[4] = b.x; m[5] = b.y; m[6] = b.z; m[7] = b.w;
    m[8] = c.x; m[9] = c.y; m[10] = c.z; m[11] = c.w;
    m[12] = d.x; m[13] = d.y; m[14] = d.z; m[15] = d.w;
}

__global__ void __launch_bounds__(64) k_chunk_cv(ChunkJob j0, ChunkJob j1) {
    const ChunkJob &J = blockIdx.y == 0 ? j0 : j1;
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t chunk = (uint32_t)(gid % J.n_chunks), rep = (uint32_t)(gid / J.n_chunks);
    if (rep >= J.nreps) return;
    const uint32_t off = chunk * 1024u;
    const uint32_t clen = min(1024u, J.len - off);
    const bool root = J.single;
    const uint8_t *base = J.stream + (size_t)rep * J.pitch + off;
    uint32_t cv[8];
    b3_iv(cv);
    const uint32_t n_blocks = clen == 0 ? 1 : (clen + 63) / 64;
    const uint32_t n_full = clen / 64;  // blocks read with plain vector loads (the pitch is a multiple of 64)
    uint32_t m[16], nx[16];
    if (n_full) load_block(reinterpret_cast<const uint4 *>(base), nx);
    for (uint32_t b = 0; b < n_blocks; b++) {
        uint32_t blen = 64;
        if (b < n_full) {
#pragma unroll
            for (int i = 0; i < 16; i++) m[i] = nx[i];
            if (b + 1 < n_full) load_block(reinterpret_cast<const uint4 *>(base + 64 * (b + 1)), nx);
        } else {  // the ragged tail block (or the single empty block)
            blen = clen - 64 * b;
            const uint32_t *w = reinterpret_cast<const uint32_t *>(base + 64 * b);
#pragma unroll
            for (int i = 0; i < 16; i++) {
                uint32_t v = 0;
                if (4u * i < blen) {
                    v = w[i];
                    if (blen - 4 * i < 4) v &= (1u << (8 * (blen - 4 * i))) - 1u;
                }
                m[i] = v;
            }
        }
        uint32_t flags = (b == 0 ? B3_CHUNK_START : 0) | (b + 1 == n_blocks ? B3_CHUNK_END : 0);
        if (root && b + 1 == n_blocks) flags |= B3_ROOT;
        b3_compress_cv(cv, m, J.chunk0 + chunk, blen, flags);
    }
    uint4 *dst = reinterpret_cast<uint4 *>(J.cvs + ((size_t)rep * J.cv_stride + J.chunk0 + chunk) * 8);
    dst[0] = make_uint4(cv[0], cv[1], cv[2], cv[3]);
    dst[1] = make_uint4(cv[4], cv[5], cv[6], cv[7]);
}

static uint32_t n_chunks_of(uint32_t len) { return len == 0 ? 1 : (len + 1023) / 1024; }

void launch_chunk_cv2(const uint8_t *on, size_t pitch_on, uint32_t len_on, uint32_t *cv_on, uint32_t nreps_on, const uint8_t *pre, size_t pitch_pre,
                      uint32_t len_pre, uint32_t *cv_pre, uint32_t nreps_pre, cudaStream_t st) {
    ChunkJob j0{on, pitch_on, len_on, n_chunks_of(len_on), nreps_on, cv_on, 0, n_chunks_of(len_on), n_chunks_of(len_on) == 1},
        j1{pre, pitch_pre, len_pre, n_chunks_of(len_pre), nreps_pre, cv_pre, 0, n_chunks_of(len_pre), n_chunks_of(len_pre) == 1};
    const uint64_t threads = std::max((uint64_t)j0.n_chunks * nreps_on, (uint64_t)j1.n_chunks * nreps_pre);
    if (threads == 0) return;
    dim3 grid((unsigned)((threads + 63) / 64), 2);
    k_chunk_cv<<<grid, 64, 0, st>>>(j0, j1);
}

// Streaming: chunks [chunk0, chunk0 + n_chunks) of both streams from window buffers (n_chunks may be 0 for a stream); `total_*` are
// the whole streams' chunk counts (the pitch of the CV arrays; 1 = the stream is a single, root chunk).
void launch_chunk_cv_window(const uint8_t *on, size_t pitch_on, uint32_t len_on, uint32_t nchunks_on, uint32_t chunk0_on, uint32_t total_on, uint32_t *cv_on,
                            const uint8_t *pre, size_t pitch_pre, uint32_t len_pre, uint32_t nchunks_pre, uint32_t chunk0_pre, uint32_t total_pre,
                            uint32_t *cv_pre, uint32_t nreps, cudaStream_t st) {
    ChunkJob j0{on, pitch_on, len_on, nchunks_on, nreps, cv_on, chunk0_on, total_on, total_on == 1},
        j1{pre, pitch_pre, len_pre, nchunks_pre, nreps, cv_pre, chunk0_pre, total_pre, total_pre == 1};
    const uint64_t threads = std::max((uint64_t)nchunks_on, (uint64_t)nchunks_pre) * nreps;
    if (threads == 0) return;
    dim3 grid((unsigned)((threads + 63) / 64), 2);
    k_chunk_cv<<<grid, 64, 0, st>>>(j0, j1);
}

// =====================================================================================================================
//  Streaming segments: carried wire state in and out, and the segment's share of the packed openings
// =====================================================================================================================
// Imports: row n_prg + j of the segment's share tensor <- slot[j] of the cell file (and its instance-major copy for the mask VM);
// plaintext leaf n_inputs + j <- the cell's value.
__global__ void __launch_bounds__(256) k_seg_import(const uint32_t *__restrict__ slot, uint32_t n_imports, const uint64_t *__restrict__ cell_rows,
                                                    const uint8_t *__restrict__ cell_vals, uint32_t npi, uint32_t n_prg, uint64_t *__restrict__ rows,
                                                    uint64_t *__restrict__ fresh_pm, size_t pitch_pm, bool pm_pairs, uint8_t *__restrict__ leaf_vals) {
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t pi = (uint32_t)(gid % npi);
    const uint64_t j = gid / npi;
    if (j >= n_imports) return;
    const uint64_t v = cell_rows[(size_t)slot[j] * npi + pi];
    rows[(size_t)(n_prg + j) * npi + pi] = v;
    if (fresh_pm != nullptr) {
        if (pm_pairs) fresh_pm[((size_t)(pi >> 1) * pitch_pm + n_prg + j) * 2 + (pi & 1)] = v;
        else fresh_pm[(size_t)pi * pitch_pm + n_prg + j] = v;
    }
    if (pi == 0) leaf_vals[j] = cell_vals[slot[j]];
}
// Exports: the final state of the wires later segments read.
__global__ void __launch_bounds__(256) k_seg_export(const uint32_t *__restrict__ slot, const uint32_t *__restrict__ row, const uint32_t *__restrict__ vref,
                                                    uint32_t n_exports, const uint64_t *__restrict__ rows, const uint8_t *__restrict__ vals, uint32_t npi,
                                                    uint64_t *__restrict__ cell_rows, uint8_t *__restrict__ cell_vals) {
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t pi = (uint32_t)(gid % npi);
    const uint64_t k = gid / npi;
    if (k >= n_exports) return;
    cell_rows[(size_t)slot[k] * npi + pi] = rows[(size_t)row[k] * npi + pi];
    if (pi == 0) cell_vals[slot[k]] = (uint8_t)((vals[vref[k] >> 1] ^ vref[k]) & 1);
}
void launch_seg_import(const uint32_t *slot, uint32_t n_imports, const uint64_t *cell_rows, const uint8_t *cell_vals, uint32_t npi, uint32_t n_prg, uint64_t *rows,
                       uint64_t *fresh_pm, size_t pitch_pm, bool pm_pairs, uint8_t *leaf_vals, cudaStream_t st) {
    if (!n_imports) return;
    const uint64_t threads = (uint64_t)n_imports * npi;
    k_seg_import<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(slot, n_imports, cell_rows, cell_vals, npi, n_prg, rows, fresh_pm, pitch_pm, pm_pairs, leaf_vals);
}
void launch_seg_export(const uint32_t *slot, const uint32_t *row, const uint32_t *vref, uint32_t n_exports, const uint64_t *rows, const uint8_t *vals, uint32_t npi,
                       uint64_t *cell_rows, uint8_t *cell_vals, cudaStream_t st) {
    if (!n_exports) return;
    const uint64_t threads = (uint64_t)n_exports * npi;
    k_seg_export<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(slot, row, vref, n_exports, rows, vals, npi, cell_rows, cell_vals);
}

// The segment's bits of the opened repetitions' packed vectors (src/transcript/prover.rs:57-175), OR-ed into a proof whose headers
// and (zeroed) vectors k_extract has already written.  CTA = repetition; thread = output byte.  A byte that straddles two segments
// is completed by the later one (segments run one after the other).
__global__ void __launch_bounds__(256) k_seg_extract(SegExtractArgs a) {
    const uint32_t rep = blockIdx.x, omit = a.omit_of_rep[rep];
    if (omit >= RV_PLAYERS) return;
    const ProofLayout L{a.len_recons, a.len_corrs, a.len_inputs, 0, 0, 0};
    uint8_t *e = a.proof + L.g_base() + 8 + a.rank_of_rep[rep] * L.sz_on_g();
    const uint8_t *on = a.on + (size_t)rep * a.pitch_on + a.base_on, *pre = a.pre + (size_t)rep * a.pitch_pre + a.base_pre;
    const uint32_t tid = threadIdx.x + blockDim.x * blockIdx.y, nt = blockDim.x * gridDim.y;
    auto gather = [&](uint8_t *dst, uint64_t first, uint32_t n, const uint8_t *stream, const uint32_t *pos, uint32_t bit) {
        if (n == 0) return;
        const uint64_t g0 = first / 8, g1 = (first + n - 1) / 8;
        for (uint64_t g = g0 + tid; g <= g1; g += nt) dst[g] |= seg_pack_byte(stream, pos, first, n, g, bit);
    };
    gather(e + 137, a.first_recon, a.n_recon, on, a.recon_pos, 7 - omit);
    gather(e + 145 + L.len_recons, a.first_corr, a.n_corr, pre, nullptr, 0);
    gather(e + 153 + L.len_recons + L.len_corrs, a.first_input, a.n_inputs, on, a.input_pos, 0);
}
void launch_seg_extract(const SegExtractArgs &a, uint32_t nreps, cudaStream_t st) {
    const uint32_t bytes = (a.n_recon + a.n_corr + a.n_inputs) / 8 + 1;
    const uint32_t ny = std::min(64u, std::max(1u, bytes / 4096));
    k_seg_extract<<<dim3(nreps, ny), 256, 0, st>>>(a);
}

// BLAKE3 tree over n chunk CVs, in place: adjacent pairs merge, an odd tail is carried up unchanged (this reproduces
// the spec's left-heavy tree).  The root lands in cvs[0..8).  Block-cooperative.
__device__ void tree_reduce(uint32_t *cvs, uint32_t n) {
    const uint32_t T = blockDim.x, tid = threadIdx.x;
    while (n > 1) {
        const uint32_t pairs = n / 2, outn = (n + 1) / 2;
        const bool root = (n == 2);
        for (uint32_t base = 0; base < outn; base += T) {
            const uint32_t p = base + tid;
            uint32_t res[8];
            bool have = false;
            if (p < pairs) {
                uint32_t l[8], r[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    l[i] = cvs[(size_t)(2 * p) * 8 + i];
                    r[i] = cvs[(size_t)(2 * p + 1) * 8 + i];
                }
                b3_parent_cv(l, r, root, res);
                have = true;
            } else if (p < outn) {
#pragma unroll
                for (int i = 0; i < 8; i++) res[i] = cvs[(size_t)(2 * p) * 8 + i];
                have = true;
            }
            __syncthreads();
            if (have) {
#pragma unroll
                for (int i = 0; i < 8; i++) cvs[(size_t)p * 8 + i] = res[i];
            }
            __syncthreads();
        }
        n = outn;
    }
}

// Long streams (10^8 gates = 10^5 chunks per repetition and stream): the lower levels of the tree are wide enough for the whole
// grid, and one CTA per repetition would walk them alone (3 ms whatever the shard).  One launch per level, out of place (a thread
// writing slot p would race with the reader of slots 2p', 2p'+1 of another CTA): thread = (repetition, pair).
__global__ void __launch_bounds__(256) k_cv_tree_level(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint32_t n, uint32_t in_stride,
                                                       uint32_t out_stride, uint32_t nreps) {
    const uint32_t outn = (n + 1) / 2;
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t p = (uint32_t)(gid % outn), rep = (uint32_t)(gid / outn);
    if (rep >= nreps) return;
    const uint4 *src = reinterpret_cast<const uint4 *>(in + ((size_t)rep * in_stride + 2 * (size_t)p) * 8);
    uint4 *dst = reinterpret_cast<uint4 *>(out + ((size_t)rep * out_stride + p) * 8);
    if (2 * p + 1 < n) {
        const uint4 a0 = src[0], a1 = src[1], b0 = src[2], b1 = src[3];
        const uint32_t l[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, r[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        uint32_t res[8];
        b3_parent_cv(l, r, false, res);  // never the root: the per-repetition kernel finishes the last levels
        dst[0] = make_uint4(res[0], res[1], res[2], res[3]);
        dst[1] = make_uint4(res[4], res[5], res[6], res[7]);
    } else {  // the odd tail is carried up unchanged (the spec's left-heavy tree)
        dst[0] = src[0];
        dst[1] = src[1];
    }
}

// Reduces the CV arrays [nreps][stride] of n chunks level by level while a level is still wide; returns the remaining count and
// leaves the survivors at the front of `cvs` rows (same stride).  `scratch` holds nreps * ceil(n / 2) CVs.
uint32_t launch_cv_tree_wide(uint32_t *cvs, uint32_t n, uint32_t stride, uint32_t nreps, uint32_t *scratch, cudaStream_t st) {
    constexpr uint32_t NARROW = 2048;  // below this a CTA per repetition does the rest
    if (n <= NARROW || scratch == nullptr) return n;
    const uint32_t sstride = (n + 1) / 2;
    uint32_t *bufs[2] = {cvs, scratch};
    uint32_t strides[2] = {stride, sstride};
    int cur = 0;
    while (n > NARROW) {
        const uint32_t outn = (n + 1) / 2;
        const uint64_t threads = (uint64_t)outn * nreps;
        k_cv_tree_level<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(bufs[cur], bufs[cur ^ 1], n, strides[cur], strides[cur ^ 1], nreps);
        n = outn;
        cur ^= 1;
    }
    if (cur == 1)  // survivors are in the scratch buffer: move them to the front of the rows of `cvs`
        cudaMemcpy2DAsync(cvs, (size_t)stride * 32, scratch, (size_t)sstride * 32, (size_t)n * 32, nreps, cudaMemcpyDeviceToDevice, st);
    return n;
}

// CTA = one repetition: roots of both streams, then the joins of Transcript::hash (src/transcript/mod.rs:77-96) and
// CombineInstance::hash (src/interpreter/combine.rs:104-118).
// Verifier: repetitions >= first_pre take their online hash from the proof (VerifierTranscriptPreprocess::online_hash,
// src/transcript/verifier/preprocess.rs:54-56) -- `on_given` / `z_on_given` hold those 32-byte values per such repetition.
__global__ void __launch_bounds__(128) k_rep_hash(uint32_t *cv_on, uint32_t n_chunks_on, uint32_t *cv_pre, uint32_t n_chunks_pre,
                                                  const uint32_t *__restrict__ zconst, uint8_t *__restrict__ on_hash,
                                                  uint8_t *__restrict__ rep_hash, uint32_t first_pre, const uint8_t *__restrict__ on_given,
                                                  const uint8_t *__restrict__ z_on_given, const uint32_t *__restrict__ zrep, uint32_t stride_on,
                                                  uint32_t stride_pre) {
    const uint32_t rep = blockIdx.x;
    const bool given = rep >= first_pre;
    uint32_t *on = cv_on + (size_t)rep * stride_on * 8, *pre = cv_pre + (size_t)rep * stride_pre * 8;
    if (!given) tree_reduce(on, n_chunks_on);
    tree_reduce(pre, n_chunks_pre);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t h_on[8], h_pre[8], z[8], out[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            h_pre[i] = pre[i];
            z[i] = zconst[8 + i];  // H(B3("") || B3("")): the empty Z64 transcript of a prover / online-verifier repetition
        }
        if (given) {
            const uint8_t *g = on_given + (size_t)(rep - first_pre) * 32, *zg = z_on_given + (size_t)(rep - first_pre) * 32;
            uint32_t zon[8], e[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                h_on[i] = (uint32_t)g[4 * i] | ((uint32_t)g[4 * i + 1] << 8) | ((uint32_t)g[4 * i + 2] << 16) | ((uint32_t)g[4 * i + 3] << 24);
                zon[i] = (uint32_t)zg[4 * i] | ((uint32_t)zg[4 * i + 1] << 8) | ((uint32_t)zg[4 * i + 2] << 16) | ((uint32_t)zg[4 * i + 3] << 24);
                e[i] = zconst[i];  // B3(""): the (empty) Z64 preprocessing stream
            }
            if (zrep == nullptr) b3_hash64(e, zon, z);
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) h_on[i] = on[i];
        }
        if (zrep != nullptr) {  // the circuit has Z64 ops: k_zrep_hash computed this repetition's Z64 transcript hash
#pragma unroll
            for (int i = 0; i < 8; i++) z[i] = zrep[(size_t)rep * 8 + i];
        }
        rep_join(h_on, h_pre, z, out);
        uint32_t *d0 = reinterpret_cast<uint32_t *>(on_hash + (size_t)rep * 32), *d1 = reinterpret_cast<uint32_t *>(rep_hash + (size_t)rep * 32);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            d0[i] = h_on[i];
            d1[i] = out[i];
        }
    }
}

void launch_rep_hash(uint32_t *cv_on, uint32_t n_chunks_on, uint32_t *cv_pre, uint32_t n_chunks_pre, const uint32_t *zconst, uint32_t nreps,
                     uint8_t *on_hash, uint8_t *rep_hash, cudaStream_t st, uint32_t first_pre, const uint8_t *on_given, const uint8_t *z_on_given,
                     const uint32_t *zrep, uint32_t *scratch) {
    // prover with long streams: the wide lower levels of both trees grid-wide first (every repetition computes both roots there)
    uint32_t n_on = n_chunks_on, n_pre = n_chunks_pre;
    if (scratch != nullptr && first_pre == 0xFFFFFFFFu) {
        n_on = launch_cv_tree_wide(cv_on, n_chunks_on, n_chunks_on, nreps, scratch, st);
        n_pre = launch_cv_tree_wide(cv_pre, n_chunks_pre, n_chunks_pre, nreps, scratch, st);
    }
    k_rep_hash<<<nreps, 128, 0, st>>>(cv_on, n_on, cv_pre, n_pre, zconst, on_hash, rep_hash, first_pre, on_given, z_on_given, zrep, n_chunks_on, n_chunks_pre);
}

// Z64 transcript of one repetition (CTA): Transcript::hash = H(B3(pre) || B3(on)), src/transcript/mod.rs:77-96
__global__ void __launch_bounds__(256) k_zrep_hash(uint32_t *cv_on, uint32_t n_chunks_on, uint32_t *cv_pre, uint32_t n_chunks_pre,
                                                   uint8_t *__restrict__ zon_hash, uint32_t *__restrict__ zrep, uint32_t first_pre,
                                                   const uint8_t *__restrict__ z_on_given) {
    const uint32_t rep = blockIdx.x;
    const bool given = rep >= first_pre;
    uint32_t *on = cv_on + (size_t)rep * n_chunks_on * 8, *pre = cv_pre + (size_t)rep * n_chunks_pre * 8;
    if (!given) tree_reduce(on, n_chunks_on);
    tree_reduce(pre, n_chunks_pre);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t h_on[8], h_pre[8], out[8];
#pragma unroll
        for (int i = 0; i < 8; i++) h_pre[i] = pre[i];
        if (given) {
            const uint8_t *g = z_on_given + (size_t)(rep - first_pre) * 32;
#pragma unroll
            for (int i = 0; i < 8; i++) h_on[i] = (uint32_t)g[4 * i] | ((uint32_t)g[4 * i + 1] << 8) | ((uint32_t)g[4 * i + 2] << 16) | ((uint32_t)g[4 * i + 3] << 24);
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) h_on[i] = on[i];
        }
        b3_hash64(h_pre, h_on, out);
        uint32_t *d0 = reinterpret_cast<uint32_t *>(zon_hash + (size_t)rep * 32);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            d0[i] = h_on[i];
            zrep[(size_t)rep * 8 + i] = out[i];
        }
    }
}

void launch_zrep_hash(uint32_t *cv_on, uint32_t n_chunks_on, uint32_t *cv_pre, uint32_t n_chunks_pre, uint32_t nreps, uint8_t *zon_hash,
                      uint32_t *zrep, cudaStream_t st, uint32_t first_pre, const uint8_t *z_on_given) {
    k_zrep_hash<<<nreps, 256, 0, st>>>(cv_on, n_chunks_on, cv_pre, n_chunks_pre, zon_hash, zrep, first_pre, z_on_given);
}

// =====================================================================================================================
//  Online verifier kernels (src/transcript/verifier/online.rs): u-plane leaves, the item plane with the proof's data
// =====================================================================================================================
// thread = (leaf, opened repetition).  Leaves 0..n_inputs-1 are the inputs, n_inputs.. the kappa of every Mul.
__global__ void __launch_bounds__(256) k_verify_leaves(const Item *__restrict__ items, const uint32_t *__restrict__ input_pos,
                                                       const uint32_t *__restrict__ mul_pos, const uint32_t *__restrict__ recon_idx,
                                                       uint32_t n_inputs, uint32_t n_and, const uint32_t *__restrict__ rand_row, uint32_t n_rand,
                                                       const VOpen *__restrict__ opens, const uint8_t *__restrict__ proof,
                                                       const uint64_t *__restrict__ rows, uint32_t npi, uint32_t n_slots,
                                                       uint8_t *__restrict__ leaf_vals, size_t leaf_pitch) {
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_leaves = n_inputs + n_and + n_rand;
    const uint32_t leaf = (uint32_t)(gid % n_leaves), slot = (uint32_t)(gid / n_leaves);
    if (slot >= n_slots) return;
    uint8_t v;
    if (leaf < n_inputs) v = verify_leaf_input(items[input_pos[leaf]], leaf, opens[slot], proof, rows, npi, slot);
    else if (leaf >= n_inputs + n_and) v = verify_leaf_random(rand_row[leaf - n_inputs - n_and], rows, npi, slot);
    else {
        const uint32_t t = mul_pos[leaf - n_inputs];
        v = verify_leaf_kappa(items[t], recon_idx[t], opens[slot], proof, rows, npi, slot);
    }
    leaf_vals[(size_t)slot * leaf_pitch + leaf] = v;
}

// thread = (8 consecutive online positions, packed instance of opened repetitions)
__global__ void __launch_bounds__(256) k_verify_items_online(const Item *__restrict__ items, const uint32_t *__restrict__ item_ua,
                                                             const uint32_t *__restrict__ item_ub, const uint32_t *__restrict__ recon_idx,
                                                             uint32_t n_online, const VOpen *__restrict__ opens, const uint8_t *__restrict__ proof,
                                                             const uint64_t *__restrict__ rows, uint32_t npi, uint32_t npi_online,
                                                             const uint8_t *__restrict__ uvals, size_t upitch, uint8_t *__restrict__ on, size_t pitch,
                                                             int *not_okay) {
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t pi = (uint32_t)(gid % npi_online);
    const uint64_t t0 = (gid / npi_online) * 8;
    if (t0 >= n_online) return;
    uint64_t W[8], out[8];
    int flag = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        W[i] = 0;
        const uint32_t t = (uint32_t)t0 + i;
        if (t < n_online) W[i] = verify_online_word(items[t], t, item_ua[t], item_ub[t], recon_idx[t], opens, proof, rows, npi, pi, uvals, upitch, &flag);
    }
    if (flag) atomicOr(not_okay, 1);
    words_to_stream_bytes(W, out);
#pragma unroll
    for (int r = 0; r < 8; r++) *reinterpret_cast<uint64_t *>(on + (size_t)(8 * pi + r) * pitch + t0) = out[r];
}

// thread = (8 consecutive Mul indices, packed instance of opened repetitions): the proof's corrections as stream bytes
__global__ void __launch_bounds__(256) k_verify_items_pre(uint32_t n_pre, const VOpen *__restrict__ opens, const uint8_t *__restrict__ proof,
                                                          uint32_t npi_online, uint8_t *__restrict__ pre, size_t pitch) {
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t pi = (uint32_t)(gid % npi_online);
    const uint64_t j0 = (gid / npi_online) * 8;
    if (j0 >= n_pre) return;
    uint64_t W[8], out[8];
#pragma unroll
    for (int i = 0; i < 8; i++) W[i] = (j0 + i < n_pre) ? verify_pre_word((uint32_t)j0 + i, opens, proof, pi) : 0;
    words_to_stream_bytes(W, out);
#pragma unroll
    for (int r = 0; r < 8; r++) *reinterpret_cast<uint64_t *>(pre + (size_t)(8 * pi + r) * pitch + j0) = out[r];
}

void launch_verify_leaves(const DevProgram &P, const VOpen *opens, const uint8_t *proof, const uint64_t *rows, uint32_t npi, uint32_t n_slots,
                          uint8_t *leaf_vals, size_t leaf_pitch, cudaStream_t st) {
    const uint64_t threads = (uint64_t)(P.n_inputs + P.n_pre + P.n_rand) * n_slots;
    if (!threads) return;
    k_verify_leaves<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(P.items, P.input_pos, P.mul_pos, P.recon_idx, P.n_inputs, P.n_pre, P.rand_row, P.n_rand,
                                                                        opens, proof, rows, npi, n_slots, leaf_vals, leaf_pitch);
}

void launch_verify_items(const DevProgram &P, const VOpen *opens, const uint8_t *proof, const uint64_t *rows, uint32_t npi, uint32_t npi_online,
                         const uint8_t *uvals, size_t upitch, uint8_t *on, size_t pitch_on, uint8_t *pre, size_t pitch_pre, int *not_okay,
                         cudaStream_t st) {
    if (P.n_online) {
        const uint64_t threads = (uint64_t)((P.n_online + 7) / 8) * npi_online;
        k_verify_items_online<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(P.items, P.item_ua, P.item_ub, P.recon_idx, P.n_online, opens, proof, rows, npi,
                                                                                  npi_online, uvals, upitch, on, pitch_on, not_okay);
    }
    if (P.n_pre) {
        const uint64_t threads = (uint64_t)((P.n_pre + 7) / 8) * npi_online;
        k_verify_items_pre<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(P.n_pre, opens, proof, npi_online, pre, pitch_pre);
    }
}

// =====================================================================================================================
//  K6  comm + Fiat-Shamir challenge: one warp
// =====================================================================================================================
// grid.x = proof of the session.  The 256 hashes of proof b arrive as n_seg segments of seg_bytes (one per rank of the
// all-gather, rank-major over the session's proofs): segment r of proof b starts at all_hashes + (r * n_proofs + b) * seg_bytes.
//
// Linked sessions (x.world > 1; rv_session_peer_link): the all-gather of src/proof/mod.rs:160-171 happens HERE, over peer memory.
// The warp of proof b first stores this rank's segment into every rank's receive buffer (NVLink peer stores; double-buffered by
// the parity of the proof's step counter), publishes it with a system-scope release of the step number into each rank's flag
// word, then waits (system-scope acquire) until the flags of all ranks carry this step: commit -> gather -> challenge is one
// kernel, no host pacing, no collective library on the data path.
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Spins until *flag == want (lanes with active == false return at once); false on timeout.
__device__ __forceinline__ bool wait_flag(const uint32_t *flag, uint32_t want, bool active, uint64_t timeout_ns) {
    if (!active) return true;
    const uint64_t t0 = globaltimer_ns();
    uint32_t spins = 0;
    while (ld_acquire_sys(flag) != want) {
        if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > timeout_ns) return false;
        __nanosleep(64);
    }
    return true;
}

__global__ void __launch_bounds__(32) k_challenge(const uint8_t *all_hashes_base, uint32_t seg_bytes, uint8_t *__restrict__ comm_base,
                                                  size_t comm_stride, uint8_t *__restrict__ omit_base, uint16_t *__restrict__ rank_base, XchgArgs x,
                                                  const uint8_t *__restrict__ own_hashes) {
    const uint32_t pb = blockIdx.x, n_proofs = gridDim.x;
    uint8_t *comm = comm_base + pb * comm_stride, *omit_of_rep = omit_base + pb * RV_TOTAL_REPS;
    uint16_t *rank_of_rep = rank_base + pb * RV_TOTAL_REPS;
    __shared__ uint32_t cv[8][8];
    __shared__ uint32_t xof[32][16];
    __shared__ uint8_t omit[RV_TOTAL_REPS];
    __shared__ __align__(16) uint32_t hbuf[RV_TOTAL_REPS * 8];
    const uint32_t lane = threadIdx.x;
    if (x.world > 1) {
        const XchgLayout L{n_proofs};
        uint32_t *step = reinterpret_cast<uint32_t *>(x.peer[x.rank] + L.off_step()) + pb;
        const uint32_t e = *step + 1, par = e & 1;
        const uint4 *src = reinterpret_cast<const uint4 *>(own_hashes + (size_t)pb * seg_bytes);
        for (uint32_t r = 0; r < x.world; r++) {
            uint4 *dst = reinterpret_cast<uint4 *>(x.peer[r] + L.off_hash(par) + ((size_t)x.rank * n_proofs + pb) * seg_bytes);
            for (uint32_t i = lane; i < seg_bytes / 16; i += 32) dst[i] = src[i];
        }
        __threadfence_system();
        __syncwarp();
        if (lane < x.world) st_release_sys(reinterpret_cast<uint32_t *>(x.peer[lane] + L.off_flag(par, x.rank)) + pb, e);
        const bool ok = wait_flag(reinterpret_cast<const uint32_t *>(x.peer[x.rank] + L.off_flag(par, lane)) + pb, e, lane < x.world, x.timeout_ns);
        if (!__all_sync(0xffffffffu, ok) && lane == 0) atomicOr(reinterpret_cast<int *>(comm - 4), RV_BAD_PEER_TIMEOUT);  // the flag word sits right before comm
        __threadfence_system();
        if (lane == 0) *step = e;
        all_hashes_base = x.peer[x.rank] + L.off_hash(par);
    }
    // combine_hashes (src/proof/mod.rs:102-108): BLAKE3 of 256 x 32 B = 8 chunks
    {   // the hashes may have been written by other GPUs a moment ago: fetch them with coherent (volatile) loads into shared memory
        for (uint32_t i = lane; i < RV_TOTAL_REPS * 2; i += 32) {  // 16-byte pieces; a 1 KiB chunk (32 hashes) never straddles segments (>= 32 repetitions each)
            const uint32_t byte0 = 16 * i, seg = byte0 / seg_bytes;
            const uint4 *p = reinterpret_cast<const uint4 *>(all_hashes_base + ((size_t)seg * n_proofs + pb) * seg_bytes + (byte0 - seg * seg_bytes));
            uint4 v;
            asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
            reinterpret_cast<uint4 *>(hbuf)[i] = v;
        }
        __syncwarp();
    }
    if (lane < 8) {
        uint32_t c[8];
        b3_chunk_cv(hbuf + 256 * lane, 1024, lane, false, c);
        for (int i = 0; i < 8; i++) cv[lane][i] = c[i];
    }
    __syncwarp();
    for (uint32_t n = 8; n > 1; n >>= 1) {
        uint32_t res[8];
        if (lane < n / 2) b3_parent_cv(cv[2 * lane], cv[2 * lane + 1], n == 2, res);
        __syncwarp();
        if (lane < n / 2)
            for (int i = 0; i < 8; i++) cv[lane][i] = res[i];
        __syncwarp();
    }
    if (lane < 8) reinterpret_cast<uint32_t *>(comm)[lane] = cv[0][lane];
    for (uint32_t i = lane; i < RV_TOTAL_REPS; i += 32) omit[i] = RV_PLAYERS;
    uint32_t m[16];
    challenge_block(cv[0], m);
    __shared__ int distinct;
    if (lane == 0) distinct = 0;
    __syncwarp();
    for (uint32_t round = 0; round < 64; round++) {
        uint32_t o[16];
        challenge_xof_block(m, (uint64_t)round * 32 + lane, o);
        for (int i = 0; i < 16; i++) xof[lane][i] = o[i];
        __syncwarp();
        if (lane == 0)
            for (int b = 0; b < 32 && distinct < RV_ONLINE_REPS; b++) challenge_consume(xof[b], omit, &distinct);
        __syncwarp();
        if (distinct >= RV_ONLINE_REPS) break;
    }
    __syncwarp();
    if (lane == 0) {
        uint16_t n_on = 0, n_pre = 0;
        for (int i = 0; i < RV_TOTAL_REPS; i++) {
            omit_of_rep[i] = omit[i];
            rank_of_rep[i] = (omit[i] < RV_PLAYERS) ? n_on++ : n_pre++;
        }
    }
}

void launch_challenge(const uint8_t *all_hashes, uint32_t seg_bytes, uint8_t *comm, size_t comm_stride, uint8_t *omit_of_rep, uint16_t *rank_of_rep,
                      uint32_t n_proofs, cudaStream_t st, const XchgArgs *x, const uint8_t *own_hashes) {
    XchgArgs none;
    k_challenge<<<n_proofs, 32, 0, st>>>(all_hashes, seg_bytes, comm, comm_stride, omit_of_rep, rank_of_rep, x ? *x : none, own_hashes);
}

// Last kernel of a linked session's open phase (one warp).  Every rank's extraction wrote its entries straight into the assembling
// rank's proof buffer; this publishes "rank r is done with step e" there, and on the assembling rank waits for all ranks before the
// device-to-host copy that follows in stream order: the assembly of src/proof/mod.rs:200-221 with no reduce and no host hop.
__global__ void __launch_bounds__(32) k_xfinish(XchgArgs x, uint32_t n_proofs, int *bad, size_t flag_stride) {
    const XchgLayout L{n_proofs};
    const uint32_t lane = threadIdx.x;
    uint32_t *step = reinterpret_cast<uint32_t *>(x.peer[x.rank] + L.off_done_step());
    const uint32_t e = *step + 1;
    __threadfence_system();
    if (lane == 0) st_release_sys(reinterpret_cast<uint32_t *>(x.peer[x.dst] + L.off_done()) + x.rank, e);
    if (x.rank == x.dst) {
        const bool ok = wait_flag(reinterpret_cast<const uint32_t *>(x.peer[x.rank] + L.off_done()) + lane, e, lane < x.world, x.timeout_ns);
        if (!__all_sync(0xffffffffu, ok))
            for (uint32_t b = lane; b < n_proofs; b += 32) atomicOr(reinterpret_cast<int *>(reinterpret_cast<uint8_t *>(bad) + b * flag_stride), RV_BAD_PEER_TIMEOUT);
        __threadfence_system();
    }
    __syncwarp();
    if (lane == 0) *step = e;
}

void launch_xfinish(const XchgArgs &x, uint32_t n_proofs, int *bad, size_t flag_stride, cudaStream_t st) { k_xfinish<<<1, 32, 0, st>>>(x, n_proofs, bad, flag_stride); }

// =====================================================================================================================
//  K7  extraction: CTA = one repetition of the shard; writes its entry of the bincode `Proof` in place
// =====================================================================================================================
__global__ void __launch_bounds__(256) k_extract(const uint32_t *__restrict__ recon_pos, const uint32_t *__restrict__ input_pos,
                                                 uint32_t n_recon, uint32_t n_pre, uint32_t n_inputs, ExtractArgs a) {
    const uint32_t pb = blockIdx.x / a.nreps, lrep = blockIdx.x, rep = a.first_rep + blockIdx.x % a.nreps;  // lrep indexes the session's streams
    ProofLayout L{a.len_recons, a.len_corrs, a.len_inputs, a.len_zrecons, a.len_zcorrs, a.len_zinputs};
    ExtractView v;
    v.on = a.on + (size_t)lrep * a.pitch_on;
    v.pre = a.pre + (size_t)lrep * a.pitch_pre;
    v.on_hash = a.on_hash + (size_t)lrep * 32;
    v.pkeys = a.pkeys + (size_t)lrep * 128;
    v.seed = a.seeds + (size_t)lrep * 16;
    v.comm = a.comm + pb * a.proof_stride;
    v.z64_empty_hash = a.z64_empty_hash;
    v.z_on_hash = a.z_on_hash ? a.z_on_hash + (size_t)lrep * 32 : nullptr;
    v.recon_pos = recon_pos;
    v.input_pos = input_pos;
    v.n_recon = n_recon;
    v.n_pre = n_pre;
    v.n_inputs = n_inputs;
    // grid.y CTAs share one repetition: the packed vectors of a big circuit are megabytes per opened repetition
    const uint32_t omit = a.omit_of_rep[pb * RV_TOTAL_REPS + rep];
    if (omit >= RV_PLAYERS && blockIdx.y != 0) return;
    extract_entry(L, v, rep, omit, a.rank_of_rep[pb * RV_TOTAL_REPS + rep], threadIdx.x + blockDim.x * blockIdx.y, blockDim.x * gridDim.y,
                  a.proof + pb * a.proof_stride);
}

void launch_extract(const DevProgram &P, const ExtractArgs &a, cudaStream_t st) {
    const uint32_t bytes = a.len_recons + a.len_corrs + a.len_inputs;
    const uint32_t ny = std::min(64u, std::max(1u, bytes / 4096));
    k_extract<<<dim3(a.nreps * a.n_proofs, ny), 256, 0, st>>>(P.recon_pos, P.input_pos, P.n_recon, P.n_pre, P.n_inputs, a);
}

// Per-device kernel attributes (opt-in dynamic shared memory above 48 KB, shared-memory carveout).  The attributes belong to the
// (function, device) pair, so rv_session_create calls this for the session's device; the first call per device does the work.
int configure_kernels(int device) {
    static std::mutex mu;
    static uint64_t done[4] = {0, 0, 0, 0};
    if (device < 0 || device >= 256) return (int)cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> g(mu);
    if (done[device >> 6] >> (device & 63) & 1) return 0;
    cudaError_t e = cudaSuccess;
    auto set = [&](const void *fn, int dyn_bytes, bool carveout) {
        if (e == cudaSuccess && dyn_bytes) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_bytes);
        // the full shared-memory carveout: with the default split a second CTA (of this or of another kernel) does not fit
        if (e == cudaSuccess && carveout) e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    };
    set((const void *)k_mask_gen_tt<false, 16>, (int)GT_SMEM2, true);
    set((const void *)k_mask_gen_tt<true, 16>, (int)GT_SMEM4, true);
    set((const void *)k_mask_gen_tt<false, 8>, (int)GT_SMEM2, true);
    set((const void *)k_mask_gen_tt<true, 8>, (int)GT_SMEM4, true);
    set((const void *)k_mask_gen_tt<false, 4>, (int)GT_SMEM2, true);
    set((const void *)k_mask_gen_tt<true, 4>, (int)GT_SMEM4, true);
    set((const void *)k_values<true, (int)LUT_STEPS_PER_CHUNK_MAX>, (int)SMEM_DYN_CAP, false);
    set((const void *)k_values<true, (int)LUT_STEPS_PER_CHUNK>, (int)SMEM_DYN_CAP, false);
    set((const void *)k_values<false, (int)LUT_STEPS_PER_CHUNK_MAX>, (int)SMEM_DYN_CAP, false);
    set((const void *)k_mask_vm<1>, (int)SMEM_DYN_CAP, true);  // two VM CTAs (~110 KB each for SHA-256) per SM need the full carveout
    set((const void *)k_mask_vm<2>, (int)SMEM_DYN_CAP, true);
    set((const void *)k_items, 0, true);                    // ~35 KB tiles: six CTAs share an SM
    if (e == cudaSuccess) e = (cudaError_t)configure_zkernels();
    // Load every kernel now.  With CUDA's lazy module loading the first launch of a function may have to wait for the device to go
    // idle; the challenge kernel of a linked session waits for the other ranks on the device, so a first launch issued behind it
    // (by a host thread that drives several GPUs, or several shards on one GPU) must never be the one that loads code.
    auto load = [&](const void *fn) {
        cudaFuncAttributes at;
        if (e == cudaSuccess) e = cudaFuncGetAttributes(&at, fn);
    };
    load((const void *)k_key_setup);
    load((const void *)k_values_leaves);
    load((const void *)k_values_level);
    load((const void *)k_uvalues_leaves);
    load((const void *)k_uvalues_level);
    load((const void *)k_linear_cta);
    load((const void *)k_linear_level);
    load((const void *)k_items_pre);
    load((const void *)k_tainted);
    load((const void *)k_chunk_cv);
    load((const void *)k_rep_hash);
    load((const void *)k_cv_tree_level);
    load((const void *)k_zrep_hash);
    load((const void *)k_verify_leaves);
    load((const void *)k_verify_items_online);
    load((const void *)k_verify_items_pre);
    load((const void *)k_challenge);
    load((const void *)k_xfinish);
    load((const void *)k_extract);
    load((const void *)k_seg_import);
    load((const void *)k_seg_export);
    load((const void *)k_seg_extract);
    if (e != cudaSuccess) return (int)e;
    done[device >> 6] |= 1ull << (device & 63);
    return 0;
}

}  // namespace rv
