// rv_maskgen.cuh -- device-only core shared by the GF(2) and Z64 mask generators: bitsliced AES-128-CTR for one
// (slice, counter block) per thread with the slice's round-key planes in shared memory.
#pragma once
#include <stdint.h>

#include "rv_aes_bs.cuh"

namespace rv {

constexpr int MG_SLICES = 8;    // slices per CTA (their round keys live in shared memory: 8 x 5632 B = 44 KB)
constexpr int MG_COUNTERS = 8;  // counter blocks per CTA (64 threads: fine-grained CTAs balance the 148 SMs)
constexpr int MG_THREADS = MG_SLICES * MG_COUNTERS;

struct SmemRoundKeys {
    const uint4 *base;  // [(round*32 + plane/4)][slice] uint4
    uint32_t sl;
    __device__ __forceinline__ uint4 quad(int round, int g) const { return base[(round * 32 + g) * MG_SLICES + sl]; }
};

// Same dataflow as bs_aes128_ctr_block (rv_aes_bs.cuh), with the round-key planes fetched four at a time (LDS.128).
__device__ __forceinline__ void aes_ctr_block_smem(uint64_t ctr, const SmemRoundKeys &rk, uint32_t *s) {
#pragma unroll
    for (int g = 0; g < 32; g++) {
        const uint4 k4 = rk.quad(0, g);
        const uint32_t kk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int k = 4 * g + i, B = k >> 3, b = k & 7;
            uint32_t in = 0;
            if (B >= 8) in = 0u - (uint32_t)((ctr >> (8 * (15 - B) + b)) & 1);
            s[k] = in ^ kk[i];
        }
    }
#pragma unroll 1
    for (int round = 1; round <= 10; round++) {
#pragma unroll
        for (int B = 0; B < 16; B++) bs_sbox<uint32_t>(s + 8 * B, 0xFFFFFFFFu);
        uint32_t t[128];
#pragma unroll
        for (int c = 0; c < 4; c++)
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int b = 0; b < 8; b++) t[8 * (4 * c + r) + b] = s[8 * (4 * ((c + r) & 3) + r) + b];
        if (round < 10) {
#pragma unroll
            for (int c = 0; c < 4; c++) bs_mix_column<uint32_t>(t + 32 * c, t + 32 * c + 8, t + 32 * c + 16, t + 32 * c + 24, s + 32 * c);
        } else {
#pragma unroll
            for (int k = 0; k < 128; k++) s[k] = t[k];
        }
#pragma unroll
        for (int g = 0; g < 32; g++) {
            const uint4 k4 = rk.quad(round, g);
            s[4 * g + 0] ^= k4.x;
            s[4 * g + 1] ^= k4.y;
            s[4 * g + 2] ^= k4.z;
            s[4 * g + 3] ^= k4.w;
        }
    }
}


// cooperative load of the round-key planes of slices [w0, w0 + MG_SLICES) into the CTA's shared memory image
__device__ __forceinline__ void load_round_keys(uint4 *sk, const uint32_t *__restrict__ ks, uint32_t w0, uint32_t nslices) {
    uint32_t *sk32 = reinterpret_cast<uint32_t *>(sk);
    for (uint32_t idx = threadIdx.x; idx < MG_SLICES * 1408; idx += MG_THREADS) {
        const uint32_t sl = idx / 1408, e = idx % 1408;
        const uint32_t v = (w0 + sl < nslices) ? ks[(size_t)(w0 + sl) * 1408 + e] : 0u;
        sk32[((e >> 2) * MG_SLICES + sl) * 4 + (e & 3)] = v;
    }
}

}  // namespace rv
