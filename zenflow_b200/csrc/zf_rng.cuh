// Counter-based RNG for the latent samplers (distributions.py:61-62,75-78,106-112,125-126).
// Philox4x32-10 (Salmon et al. 2011): every (event m, column j) owns an independent stream
// keyed by the seed, so a draw does not depend on tiling, kernel variant or GPU count.
// The reference draws with jax.random (threefry) streams that cannot be reproduced without JAX;
// these samplers are checked statistically, as the reference's own tests do.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "zf_math.cuh"

namespace zf {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0;
        k.y += W1;
    }
    return c;
}

struct LatentRng {
    uint4 ctr;     // (m lo, m hi, column, block index)
    uint2 key;
    uint4 buf;
    int have;
    __device__ __forceinline__ LatentRng(unsigned long long seed, long long m, int j)
        : ctr(make_uint4((uint32_t)m, (uint32_t)((unsigned long long)m >> 32), (uint32_t)j, 0u)),
          key(make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x5A17F10Bu)), buf(make_uint4(0, 0, 0, 0)), have(0) {}
    __device__ __forceinline__ uint32_t next() {
        if (have == 0) {
            buf = philox4x32_10(ctr, key);
            ctr.w++;
            have = 4;
        }
        const uint32_t r = have == 4 ? buf.x : have == 3 ? buf.y : have == 2 ? buf.z : buf.w;
        --have;
        return r;
    }
    __device__ __forceinline__ float uniform() { return (float)(next() >> 8) * (1.0f / 16777216.0f); }          // [0, 1)
    __device__ __forceinline__ float uniform_open() { return ((float)(next() >> 8) + 0.5f) * (1.0f / 16777216.0f); }  // (0, 1)
    __device__ __forceinline__ float normal() {  // Box-Muller (one of the pair)
        const float u1 = uniform_open(), u2 = uniform();
        return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
    }
};

// one draw of the latent for (event m, column j)
__device__ __forceinline__ float latent_draw(int kind, float peakness, unsigned long long seed, long long m, int j) {
    LatentRng g(seed, m, j);
    switch (kind) {
        case kLatentNormal: return 0.5f + 0.1f * g.normal();
        case kLatentTruncNormal: {
            float z = g.normal();
            for (int it = 0; it < 16 && fabsf(z) > 5.0f; ++it) z = g.normal();
            return 0.5f + 0.1f * fminf(fmaxf(z, -5.0f), 5.0f);
        }
        case kLatentBeta: {
            // Symmetric Beta(p, p), p >= 1 (distributions.py:96-97), without rejection: Ulrich (1984), Devroye IX.4:
            // 1/2 + 1/2 sqrt(1 - U^(2 / (2p - 1))) cos(2 pi V).  Two uniforms (one Philox block) instead of the ~2.1 rounds of
            // two Marsaglia-Tsang gammas (jax.random.beta's G1 / (G1 + G2) construction): same distribution, a third of the
            // instructions, no divergent rejection loop in the tile fill of Flow.sample.
            const float u = g.uniform_open(), v = g.uniform();
            const float t = exp2f(log2f(u) * (2.0f / (2.0f * peakness - 1.0f)));
            return 0.5f + 0.5f * sqrtf(fmaxf(1.0f - t, 0.f)) * cospif(2.0f * v);
        }
        default: return g.uniform();
    }
}

}  // namespace zf
