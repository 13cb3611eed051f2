// Philox4x32-10 counter-based RNG (Salmon et al., SC'11). Layout must match oracle/philox_np.py:
//   key = (seed_lo, seed_hi); counter = (idx_lo, idx_hi, block, stream).
#pragma once
#include "common.cuh"

#define PQ_STREAM_MH 1u
#define PQ_STREAM_UNIFORM_SWEEP 2u
#define PQ_STREAM_ESS 3u

struct Philox4 {
    uint32_t x, y, z, w;
};

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                 uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}

__device__ __forceinline__ Philox4 philox_block(uint64_t seed, uint32_t stream, uint64_t idx, uint32_t block) {
    return philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), block, stream, (uint32_t)seed, (uint32_t)(seed >> 32));
}

// two words -> double in [0,1) with 53 random bits: ((hi>>5)*2^26 + (lo>>6)) * 2^-53
__device__ __forceinline__ double philox_u53(uint32_t hi, uint32_t lo) {
    const unsigned long long m = ((unsigned long long)(hi >> 5) << 26) | (unsigned long long)(lo >> 6);
    return (double)m * 1.1102230246251565e-16;
}

// Box-Muller on one block: u1 = 1 - u_a in (0,1], angle = 2*pi*u_b
__device__ __forceinline__ void philox_normal_pair(Philox4 b, double& z0, double& z1) {
    const double ua = philox_u53(b.x, b.y), ub = philox_u53(b.z, b.w);
    const double r = sqrt(-2.0 * log(1.0 - ua));
    double s, c;
    sincos(6.283185307179586 * ub, &s, &c);
    z0 = r * c;
    z1 = r * s;
}
