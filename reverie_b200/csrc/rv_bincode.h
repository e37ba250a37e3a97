// rv_bincode.h -- host-only helpers of Proof::verify: the bincode 1.3 layout of `Proof` (src/proof/mod.rs:40-66) and the
// small host-side hashes (combine_hashes / challenge_to_opening, src/proof/mod.rs:74-108) built from the same
// __host__ __device__ BLAKE3 code the kernels use.
#pragma once
#include <algorithm>
#include <cstring>
#include <initializer_list>
#include <vector>

#include "rv_planes.cuh"

namespace rv {
struct Slice {
    size_t off = 0, len = 0;
};
struct POnline {
    uint8_t omit;
    size_t keys;  // offset of the 8 x 16 key bytes
    Slice recons, corrs, inputs;
};
struct PPre {
    size_t seed, comm_online;  // offsets
};
struct PDomain {
    std::vector<POnline> online;
    std::vector<PPre> pre;
};

// bincode 1.3 default layout of ProofSingle (src/proof/mod.rs:40-60).  false = malformed (the reference's deserialize fails).
inline bool parse_domain(const uint8_t *p, size_t n, size_t &pos, PDomain &d) {
    auto u64 = [&](uint64_t &v) {
        if (n - pos < 8) return false;
        memcpy(&v, p + pos, 8);
        pos += 8;
        return true;
    };
    uint64_t cnt;
    if (!u64(cnt) || cnt > (n - pos) / 153) return false;
    d.online.resize((size_t)cnt);
    for (POnline &o : d.online) {
        if (n - pos < 129) return false;
        o.omit = p[pos];
        o.keys = pos + 1;
        pos += 129;
        for (Slice *f : {&o.recons, &o.corrs, &o.inputs}) {
            uint64_t l;
            if (!u64(l) || l > n - pos) return false;
            f->off = pos;
            f->len = (size_t)l;
            pos += (size_t)l;
        }
    }
    if (!u64(cnt) || cnt > (n - pos) / 48) return false;
    d.pre.resize((size_t)cnt);
    for (PPre &q : d.pre) {
        q.seed = pos;
        q.comm_online = pos + 16;
        pos += 48;
    }
    return true;
}

// combine_hashes + challenge_to_opening on the host (8 KiB of hashing; the same __host__ __device__ code as k_challenge)
inline void host_hash(const uint8_t *data, uint32_t len, uint32_t out[8]) {
    const uint32_t n_chunks = len == 0 ? 1 : (len + 1023) / 1024;
    std::vector<uint32_t> cvs((size_t)n_chunks * 8), buf(256);
    for (uint32_t c = 0; c < n_chunks; c++) {
        const uint32_t off = c * 1024, clen = std::min<uint32_t>(1024, len - off);
        memset(buf.data(), 0, 1024);
        if (clen) memcpy(buf.data(), data + off, clen);
        b3_chunk_cv(buf.data(), clen, c, n_chunks == 1, &cvs[(size_t)c * 8]);
    }
    uint32_t nn = n_chunks;
    while (nn > 1) {
        const uint32_t pairs = nn / 2, outn = (nn + 1) / 2;
        std::vector<uint32_t> nxt((size_t)outn * 8);
        for (uint32_t q = 0; q < outn; q++) {
            if (q < pairs) b3_parent_cv(&cvs[(size_t)2 * q * 8], &cvs[(size_t)(2 * q + 1) * 8], nn == 2, &nxt[(size_t)q * 8]);
            else memcpy(&nxt[(size_t)q * 8], &cvs[(size_t)2 * q * 8], 32);
        }
        cvs.swap(nxt);
        nn = outn;
    }
    memcpy(out, cvs.data(), 32);
}
inline void host_challenge(const uint8_t comm[32], uint8_t omit[RV_TOTAL_REPS]) {
    uint32_t cw[8], m[16];
    memcpy(cw, comm, 32);
    challenge_block(cw, m);
    memset(omit, RV_PLAYERS, RV_TOTAL_REPS);
    int distinct = 0;
    for (uint64_t t = 0; distinct < RV_ONLINE_REPS; t++) {
        uint32_t o[16];
        challenge_xof_block(m, t, o);
        challenge_consume(o, omit, &distinct);
    }
}

}  // namespace rv
