// philox.cuh -- Philox4x32-10 (Salmon et al., SC'11; Random123 known answers in tests/test_oracle_mc.py) and the
// 53-bit uniform used by every randomised kernel (K10 rollouts, free-state sampling); bit-identical to oracle/mc.c.
#pragma once
#include "common.cuh"
#include "predicates.cuh"

namespace mpb {

struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ void operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t *out) const {
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            const uint32_t n0 = hi1 ^ c1 ^ a, n2 = hi0 ^ c3 ^ b;
            c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
};
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
    return dmul(dadd(dadd(dmul((double)(hi >> 5), 67108864.0), (double)(lo >> 6)), 0.5), 1.0 / 9007199254740992.0);
}

}  // namespace mpb
