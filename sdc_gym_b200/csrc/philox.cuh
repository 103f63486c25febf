// philox.cuh - Philox4x32-10 counter-based generator (Salmon et al., SC'11), written out so that the
// device kernels, the host C++ and the Python restatement (sdc_gym_b200/rng.py) produce the same stream.
//
// The reference draws lambda with gym's `np_random.uniform` twice per reset, real part first
// (sdc_env.py:282-300), one independent generator per env seeded `seed + i` (utils/utils.py:284-289).
// gym's generator is unpinned third-party state; the replacement keys a counter-based stream by
//   key     = (seed_lo, seed_hi)
//   counter = (global_env_index_lo, global_env_index_hi, draw_index, 0)
// and turns the 128 output bits into two 53-bit uniforms (re, im), so lambda of env i at its k-th reset is
// independent of how the envs are sharded over GPUs.
#pragma once
#include <stdint.h>

#ifndef SDCGYM_HD
#ifdef __CUDACC__
#define SDCGYM_HD __host__ __device__ __forceinline__
#else
#define SDCGYM_HD inline
#endif
#endif

namespace sdcgym {

struct philox4 {
    uint32_t v[4];
};

SDCGYM_HD philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0;
        k1 += W1;
    }
    philox4 out;
    out.v[0] = c0; out.v[1] = c1; out.v[2] = c2; out.v[3] = c3;
    return out;
}

// 53-bit uniform in [0, 1) from two 32-bit words (same construction as numpy's random_sample)
SDCGYM_HD double u53(uint32_t a, uint32_t b) {
    return (double)((((uint64_t)(a >> 5)) << 26) | (uint64_t)(b >> 6)) * (1.0 / 9007199254740992.0);
}

}  // namespace sdcgym
