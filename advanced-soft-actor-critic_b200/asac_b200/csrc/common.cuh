// Shared host/device helpers for libasac_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "asac_b200.h"

namespace asac {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define ASAC_REQUIRE(cond, ...)              \
    do {                                     \
        if (!(cond)) {                       \
            ::asac::set_error(__VA_ARGS__);  \
            return ASAC_EINVAL;              \
        }                                    \
    } while (0)

#define ASAC_UNSUPPORTED(cond, ...)          \
    do {                                     \
        if (cond) {                          \
            ::asac::set_error(__VA_ARGS__);  \
            return ASAC_EUNSUPPORTED;        \
        }                                    \
    } while (0)

// call after every kernel launch
#define ASAC_LAUNCHED(name)                                                       \
    do {                                                                          \
        ::asac::count_launch();                                                   \
        cudaError_t e__ = cudaPeekAtLastError();                                  \
        if (e__ != cudaSuccess) {                                                 \
            ::asac::set_error("%s: %s", name, cudaGetErrorString(e__));           \
            return (int)e__;                                                      \
        }                                                                         \
    } while (0)

#define ASAC_CUDA(call)                                                           \
    do {                                                                          \
        cudaError_t e__ = (call);                                                 \
        if (e__ != cudaSuccess) {                                                 \
            ::asac::set_error("%s: %s", #call, cudaGetErrorString(e__));          \
            return (int)e__;                                                      \
        }                                                                         \
    } while (0)

// ---------------------------------------------------------------- Philox4x32-10
struct Philox {
    uint32_t c[4];
    uint32_t k[2];
};

__host__ __device__ __forceinline__ void philox_round(uint32_t c[4], const uint32_t k[2]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(M0, c[0]), hi1 = __umulhi(M1, c[2]);
#else
    uint32_t hi0 = (uint32_t)(((uint64_t)M0 * c[0]) >> 32), hi1 = (uint32_t)(((uint64_t)M1 * c[2]) >> 32);
#endif
    uint32_t lo0 = M0 * c[0], lo1 = M1 * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

// 4 x 32 random bits for (key = seed, counter = (a, b))
__host__ __device__ __forceinline__ void philox4(uint64_t seed, uint64_t a, uint64_t b, uint32_t out[4]) {
    uint32_t c[4] = {(uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32)};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

// uniform double in [0, 1) with 53 random bits
__host__ __device__ __forceinline__ double u01_double(uint32_t hi, uint32_t lo) {
    uint64_t x = ((uint64_t)hi << 32) | lo;
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}

}  // namespace asac
