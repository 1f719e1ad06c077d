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

// Programmatic dependent launch for the kernel CHAINS of a training step (a layer is 3-4 small dependent kernels; the hand-over between two
// graph nodes costs ~2.3 us, a sixth of a layer): a kernel launched with launch_chain() may become resident while its predecessor in the
// stream still runs; every kernel of a chain calls pdl_trigger() at its top (lets ITS successor be scheduled) and pdl_wait() before its
// first access to global memory (returns when the predecessor grid has completed and its writes are visible).  Without the launch attribute
// both instructions are no-ops.  Measured on the graphed training iteration (512 x K = 5): with EVERY chain kernel launched this way 43.6 ms
// against 35.8 ms plain -- the early-resident CTAs of the big kernels (2560-CTA BatchNorm apply, 227 KB convolutions) sit on SMs that the
// other stream's chain (prediction next to dynamics network) needs; with only the SMALL kernels (32-CTA BatchNorm finalize) launched early
// 37.6 ms.  A negative result either way, so the default is MZB_PDL=0 (plain launches); 1: the small kernels, 2: every chain kernel.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
int pdl_level();
inline bool pdl_enabled() { return pdl_level() >= 2; }
// small: a kernel of a few dozen CTAs (launched early from level 1 on)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain_lvl(int level, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_level() >= level ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args)
{
    return launch_chain_lvl(2, kernel, grid, block, smem, st, args...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain_small(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args)
{
    return launch_chain_lvl(1, kernel, grid, block, smem, st, args...);
}

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
