// Shared host/device helpers for libmzb200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mzb200.h"

namespace mzb {

void set_error(const char *fmt, ...);
void count_launch(uint64_t n = 1);

#define MZB_CHECK_ARG(cond, msg)                              \
    do {                                                      \
        if (!(cond)) {                                        \
            mzb::set_error("%s: %s", __func__, msg);          \
            return -1;                                        \
        }                                                     \
    } while (0)

#define MZB_CUDA(call)                                                                  \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess) {                                                        \
            mzb::set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e_)); \
            return -2;                                                                  \
        }                                                                               \
    } while (0)

#define MZB_LAUNCH_CHECK()                                                              \
    do {                                                                                \
        cudaError_t e_ = cudaGetLastError();                                            \
        if (e_ != cudaSuccess) {                                                        \
            mzb::set_error("%s: kernel launch failed: %s", __func__, cudaGetErrorString(e_)); \
            return -3;                                                                  \
        }                                                                               \
        mzb::count_launch();                                                            \
    } while (0)

// counter-based u32 stream shared with oracle/mcts_oracle.c:mto_rng_u32 (splitmix64 finaliser)
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint32_t mz_rng_u32(uint64_t seed, uint32_t tree, uint32_t ctr)
{
    return (uint32_t)(mix64(seed + 0x9E3779B97F4A7C15ULL * ((((uint64_t)tree) << 32) | (uint64_t)ctr)) >> 32);
}

constexpr int kNumSMs = 148;  // B200

// cudaFuncSetAttribute applies to the current device's copy of a kernel: per-device "already done" flags (a host process may
// drive several GPUs even though the usual deployment is one process per GPU)
inline bool first_use_on_device(bool (&seen)[64])
{
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
    if (seen[d]) return false;
    seen[d] = true;
    return true;
}

}  // namespace mzb
