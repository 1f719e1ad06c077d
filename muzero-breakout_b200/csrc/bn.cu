// Training-mode BatchNorm2d of the ConvBlocks / ResidualBlocks (src/networks.py:12,16-17,26-35 under MuZeroAgent.train_mode(),
// train_torch.py:372) on channels-last [M = samples x pixels][C] tensors: batch statistics, normalise (+ residual) + activation,
// running-statistics update, and the backward pass (ReLU mask, d gamma, d beta, d input, d residual).  All HBM-bound:
//   forward   read z (4 B) twice + write y (2 B [+ 4 B])            backward   read dy, z (+ y) twice + write dz (4 B [+ 2 B])
// Reductions over the M rows go through per-CTA fp64 partial sums that one warp per channel adds in a fixed order: deterministic, and the
// sum / sum-of-squares variance does not lose digits to cancellation.
#include <stdlib.h>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

constexpr int BN_THREADS = 256;
const int BN_ROWS = [] { const char *e = getenv("MZB_BN_ROWS"); const int v = e ? atoi(e) : 128; return v >= 16 && v <= 1024 ? v : 128; }();   // rows per CTA of the reduction kernels (128: 80 CTAs at a 512-sample minibatch of 4x5 latents; MZB_BN_ROWS = 32 / 64 / 128: 31.0 / 30.2 / 30.0 ms per graphed training iteration, four rows' loads in flight per thread)

__device__ __forceinline__ float from16(uint16_t u, bool f16)
{
    return f16 ? __half2float(*reinterpret_cast<const __half *>(&u)) : __bfloat162float(*reinterpret_cast<const __nv_bfloat16 *>(&u));
}
__device__ __forceinline__ uint16_t to16(float v, bool f16)
{
    if (f16) { const __half h = __float2half_rn(v); return *reinterpret_cast<const uint16_t *>(&h); }
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    return *reinterpret_cast<const uint16_t *>(&h);
}
__device__ __forceinline__ void load4_16(const uint16_t *p, bool f16, float (&o)[4])
{
    const uint2 u = *reinterpret_cast<const uint2 *>(p);
    o[0] = from16((uint16_t)(u.x & 0xffff), f16); o[1] = from16((uint16_t)(u.x >> 16), f16);
    o[2] = from16((uint16_t)(u.y & 0xffff), f16); o[3] = from16((uint16_t)(u.y >> 16), f16);
}
__device__ __forceinline__ void store4_16(uint16_t *p, bool f16, const float (&v)[4])
{
    uint2 u;
    u.x = (uint32_t)to16(v[0], f16) | ((uint32_t)to16(v[1], f16) << 16);
    u.y = (uint32_t)to16(v[2], f16) | ((uint32_t)to16(v[3], f16) << 16);
    *reinterpret_cast<uint2 *>(p) = u;
}
__device__ __forceinline__ float act_fwd(float v, int act)
{
    switch (act) {
        case MZ_ACT_RELU: return fmaxf(v, 0.0f);
        case MZ_ACT_LEAKY_RELU: return v > 0.0f ? v : 0.01f * v;
        default: return v;
    }
}
__device__ __forceinline__ float act_grad(float pre, int act)      // d act / d pre-activation
{
    switch (act) {
        case MZ_ACT_RELU: return pre > 0.0f ? 1.0f : 0.0f;
        case MZ_ACT_LEAKY_RELU: return pre > 0.0f ? 1.0f : 0.01f;
        default: return 1.0f;
    }
}

// Two per-channel sums over a slab of rows.  mode 0 (forward): (z, z^2).  mode 1 (backward): (g, g * xhat) with
// g = dy * act'(pre), xhat = (z - mean) * invstd, pre = gamma * xhat + beta (+ res).
// Thread = 4 consecutive channels x every (256 / (C/4))-th row of the slab; partial[blockIdx][2][C] in fp64.
__global__ void __launch_bounds__(BN_THREADS)
bn_reduce_kernel(int M, int C, int rows_per, int mode, const float *__restrict__ z, const float *__restrict__ dy, const float *__restrict__ mean,
                 const float *__restrict__ invstd, const float *__restrict__ gamma, const float *__restrict__ beta, const uint16_t *__restrict__ res,
                 int f16, int act, double *__restrict__ partial)
{
    mzb::pdl_trigger();
    mzb::pdl_wait();
    extern __shared__ double s_red[];                       // [row lanes][2][C]
    const int tpr = C / 4, lanes = BN_THREADS / tpr;
    const int cq = (threadIdx.x % tpr) * 4, rl = threadIdx.x / tpr;
    const int r0 = blockIdx.x * rows_per, r1 = min(M, r0 + rows_per);
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
    float mu[4] = {0.f, 0.f, 0.f, 0.f}, is[4] = {1.f, 1.f, 1.f, 1.f}, ga[4] = {1.f, 1.f, 1.f, 1.f}, be[4] = {0.f, 0.f, 0.f, 0.f};
    if (mode == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { mu[i] = mean[cq + i]; is[i] = invstd[cq + i]; ga[i] = gamma[cq + i]; be[i] = beta[cq + i]; }
    }
    // four rows per trip: every load of the trip is issued before its first use (the order of the additions stays row by row)
    constexpr int RU = 4;
    for (int r = r0 + rl; r < r1; r += RU * lanes) {
        float4 z4[RU], d4[RU];
        uint2 r2[RU];
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            const int ru = r + u * lanes;
            if (ru < r1) {
                z4[u] = *reinterpret_cast<const float4 *>(z + (size_t)ru * C + cq);
                if (mode == 1) {
                    d4[u] = *reinterpret_cast<const float4 *>(dy + (size_t)ru * C + cq);
                    if (res) r2[u] = *reinterpret_cast<const uint2 *>(res + (size_t)ru * C + cq);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            if (r + u * lanes >= r1) break;
            const float zz[4] = {z4[u].x, z4[u].y, z4[u].z, z4[u].w};
            if (mode == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) { a[i] += zz[i]; b[i] = fmaf(zz[i], zz[i], b[i]); }
            } else {
                const float dd[4] = {d4[u].x, d4[u].y, d4[u].z, d4[u].w};
                float rr[4] = {0.f, 0.f, 0.f, 0.f};
                if (res) {
                    rr[0] = from16((uint16_t)(r2[u].x & 0xffff), f16 != 0); rr[1] = from16((uint16_t)(r2[u].x >> 16), f16 != 0);
                    rr[2] = from16((uint16_t)(r2[u].y & 0xffff), f16 != 0); rr[3] = from16((uint16_t)(r2[u].y >> 16), f16 != 0);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float xh = (zz[i] - mu[i]) * is[i];
                    const float g = dd[i] * act_grad(fmaf(ga[i], xh, be[i]) + rr[i], act);
                    a[i] += g;
                    b[i] = fmaf(g, xh, b[i]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        s_red[(rl * 2 + 0) * C + cq + i] = (double)a[i];
        s_red[(rl * 2 + 1) * C + cq + i] = (double)b[i];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 2 * C; j += BN_THREADS) {
        double v = 0.0;
        for (int l = 0; l < lanes; ++l) v += s_red[l * 2 * C + j];
        partial[(size_t)blockIdx.x * 2 * C + j] = v;
    }
}

// sum of the per-CTA partials of channel c (both quantities) by one warp: lanes stride over the CTAs, then a shuffle tree -- a fixed order,
// and the loads of a warp are all in flight at once (a serial loop over the partials took 20-35 us per call, profiles/r1_block_step_launches.csv)
__device__ __forceinline__ void bn_warp_sums(int C, int nblocks, const double *__restrict__ partial, int c, int lane, double &s, double &q)
{
    s = 0.0; q = 0.0;
    for (int b = lane; b < nblocks; b += 32) { s += partial[(size_t)b * 2 * C + c]; q += partial[(size_t)b * 2 * C + C + c]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); q += __shfl_down_sync(0xffffffffu, q, o); }
}

// forward finalize (one warp per channel): mean, biased variance -> invstd; running statistics (unbiased variance, torch's momentum rule)
__global__ void __launch_bounds__(BN_THREADS)
bn_fwd_finalize_kernel(int M, int C, int nblocks, const double *__restrict__ partial, double eps, double momentum, float *__restrict__ running_mean,
                       float *__restrict__ running_var, float *__restrict__ save_mean, float *__restrict__ save_invstd)
{
    mzb::pdl_trigger();
    mzb::pdl_wait();
    const int c = (blockIdx.x * BN_THREADS + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= C) return;                                   // warp-uniform
    double s, q;
    bn_warp_sums(C, nblocks, partial, c, lane, s, q);
    if (lane != 0) return;
    const double mean = s / M;
    double var = q / M - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    save_mean[c] = (float)mean;
    save_invstd[c] = (float)(1.0 / sqrt(var + eps));
    if (running_mean) running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * mean);
    if (running_var) running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * (M > 1 ? var * M / (M - 1) : var));
}

// backward finalize (one warp per channel): d beta = sum g, d gamma = sum g * xhat
__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_finalize_kernel(int C, int nblocks, const double *__restrict__ partial, float *__restrict__ dgamma, float *__restrict__ dbeta,
                       float *__restrict__ dgamma_acc, float *__restrict__ dbeta_acc)
{
    mzb::pdl_trigger();
    mzb::pdl_wait();
    const int c = (blockIdx.x * BN_THREADS + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= C) return;
    double s, q;
    bn_warp_sums(C, nblocks, partial, c, lane, s, q);
    if (lane != 0) return;
    dbeta[c] = (float)s;
    dgamma[c] = (float)q;
    // the parameters' own .grad (the K unroll steps share the layer: one rounding per call, in call order, like autograd's accumulation)
    if (dbeta_acc) dbeta_acc[c] += (float)s;
    if (dgamma_acc) dgamma_acc[c] += (float)q;
}

// column sums (the bias gradient of a convolution that no BatchNorm follows): out[c] (+)= sum of the per-CTA partial sums, one warp per channel
__global__ void __launch_bounds__(BN_THREADS)
colsum_finalize_kernel(int C, int nblocks, const double *__restrict__ partial, float *__restrict__ out, int accumulate)
{
    mzb::pdl_trigger();
    mzb::pdl_wait();
    const int c = (blockIdx.x * BN_THREADS + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= C) return;
    double s, q;
    bn_warp_sums(C, nblocks, partial, c, lane, s, q);
    if (lane != 0) return;
    out[c] = accumulate ? out[c] + (float)s : (float)s;
}

// y = act(gamma * (z - mean) * invstd + beta (+ res)).  A thread takes BN_U float4 groups a grid stride apart (the stride is a multiple of
// C / 4 groups: the same 4 channels, whose constants are loaded once); all its loads are issued before the first use.
constexpr int BN_U = 4;
template <int U>
__global__ void __launch_bounds__(BN_THREADS)
bn_fwd_apply_kernel(size_t total4, int C, const float *__restrict__ z, const float *__restrict__ mean, const float *__restrict__ invstd,
                    const float *__restrict__ gamma, const float *__restrict__ beta, const uint16_t *__restrict__ res, int f16, int act,
                    uint16_t *__restrict__ y, float *__restrict__ y_f32)
{
    mzb::pdl_trigger();
    mzb::pdl_wait();
    const size_t i0 = (size_t)blockIdx.x * BN_THREADS + threadIdx.x, stride = (size_t)gridDim.x * BN_THREADS;
    if (i0 >= total4) return;
    const int c = (int)((i0 * 4) % C);
    float ga[4], mu[4], is[4], be[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { ga[k] = gamma[c + k]; mu[k] = mean[c + k]; is[k] = invstd[c + k]; be[k] = beta[c + k]; }
    float4 z4[U];
    uint2 r2[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const size_t i = i0 + u * stride;
        if (i < total4) {
            z4[u] = __ldcs(reinterpret_cast<const float4 *>(z) + i);
            if (res) r2[u] = *reinterpret_cast<const uint2 *>(res + i * 4);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const size_t i = i0 + u * stride;
        if (i >= total4) break;
        const float zz[4] = {z4[u].x, z4[u].y, z4[u].z, z4[u].w};
        float rr[4] = {0.f, 0.f, 0.f, 0.f}, v[4];
        if (res) {
            rr[0] = from16((uint16_t)(r2[u].x & 0xffff), f16 != 0); rr[1] = from16((uint16_t)(r2[u].x >> 16), f16 != 0);
            rr[2] = from16((uint16_t)(r2[u].y & 0xffff), f16 != 0); rr[3] = from16((uint16_t)(r2[u].y >> 16), f16 != 0);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = act_fwd(fmaf(ga[k], (zz[k] - mu[k]) * is[k], be[k]) + rr[k], act);
        if (y) store4_16(y + i * 4, f16 != 0, v);
        if (y_f32) reinterpret_cast<float4 *>(y_f32)[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// dz = gamma * invstd * (g - dbeta / M - xhat * dgamma / M), g = dy * act'(pre); dres = g.  Same thread mapping as the forward kernel.
template <int U>
__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_apply_kernel(size_t total4, int M, int C, const float *__restrict__ z, const float *__restrict__ dy, const float *__restrict__ mean,
                    const float *__restrict__ invstd, const float *__restrict__ gamma, const float *__restrict__ beta, const uint16_t *__restrict__ res,
                    int f16, int dz_f16, int act, const float *__restrict__ dgamma, const float *__restrict__ dbeta, float *__restrict__ dz,
                    uint16_t *__restrict__ dz16, float *__restrict__ dres)
{
    mzb::pdl_trigger();
    mzb::pdl_wait();
    const size_t i0 = (size_t)blockIdx.x * BN_THREADS + threadIdx.x, stride = (size_t)gridDim.x * BN_THREADS;
    if (i0 >= total4) return;
    const int c = (int)((i0 * 4) % C);
    const float inv_m = 1.0f / (float)M;
    float ga[4], mu[4], is[4], be[4], dgm[4], dbm[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        ga[k] = gamma[c + k]; mu[k] = mean[c + k]; is[k] = invstd[c + k]; be[k] = beta[c + k];
        dgm[k] = dgamma[c + k] * inv_m; dbm[k] = dbeta[c + k] * inv_m;
    }
    float4 z4[U], d4[U];
    uint2 r2[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const size_t i = i0 + u * stride;
        if (i < total4) {
            z4[u] = __ldcs(reinterpret_cast<const float4 *>(z) + i);
            d4[u] = __ldcs(reinterpret_cast<const float4 *>(dy) + i);
            if (res) r2[u] = *reinterpret_cast<const uint2 *>(res + i * 4);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const size_t i = i0 + u * stride;
        if (i >= total4) break;
        const float zz[4] = {z4[u].x, z4[u].y, z4[u].z, z4[u].w}, dd[4] = {d4[u].x, d4[u].y, d4[u].z, d4[u].w};
        float rr[4] = {0.f, 0.f, 0.f, 0.f}, g[4], o[4];
        if (res) {
            rr[0] = from16((uint16_t)(r2[u].x & 0xffff), f16 != 0); rr[1] = from16((uint16_t)(r2[u].x >> 16), f16 != 0);
            rr[2] = from16((uint16_t)(r2[u].y & 0xffff), f16 != 0); rr[3] = from16((uint16_t)(r2[u].y >> 16), f16 != 0);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float xh = (zz[k] - mu[k]) * is[k];
            g[k] = dd[k] * act_grad(fmaf(ga[k], xh, be[k]) + rr[k], act);
            o[k] = ga[k] * is[k] * (g[k] - dbm[k] - xh * dgm[k]);
        }
        if (dz) reinterpret_cast<float4 *>(dz)[i] = make_float4(o[0], o[1], o[2], o[3]);
        if (dz16) store4_16(dz16 + i * 4, dz_f16 != 0, o);
        if (dres) reinterpret_cast<float4 *>(dres)[i] = make_float4(g[0], g[1], g[2], g[3]);
    }
}

// float4 groups per thread of the apply kernels as launched: BN_U, or 1 with MZB_BN_U=1 (profiling: one group per thread, the form before)
int bn_u()
{
    static const int v = [] { const char *e = getenv("MZB_BN_U"); const int u = e ? atoi(e) : BN_U; return (u == 1 || u == 2 || u == 8) ? u : BN_U; }();
    return v;
}
#define BN_DISPATCH_U(call_with_U)          \
    switch (bn_u()) {                       \
        case 1: { constexpr int U_ = 1; call_with_U; break; } \
        case 2: { constexpr int U_ = 2; call_with_U; break; } \
        case 8: { constexpr int U_ = 8; call_with_U; break; } \
        default: { constexpr int U_ = 4; call_with_U; break; } \
    }

bool bn_shape_ok(int M, int C) { return M > 0 && C >= 4 && C % 4 == 0 && C / 4 <= BN_THREADS && BN_THREADS % (C / 4) == 0; }
int bn_blocks(int M) { return (M + BN_ROWS - 1) / BN_ROWS; }

}  // namespace

extern "C" {

size_t mz_bn_scratch_bytes(int M, int C) { return bn_shape_ok(M, C) ? (size_t)bn_blocks(M) * 2 * C * sizeof(double) : 0; }

int mz_bn_train_fwd(int M, int C, const float *z, const float *gamma, const float *beta, const void *res, int dtype, int act, double eps,
                    double momentum, float *running_mean, float *running_var, float *save_mean, float *save_invstd, void *y, float *y_f32,
                    void *scratch, void *stream)
{
    MZB_CHECK_ARG(bn_shape_ok(M, C), "M must be positive and C one of 4 * {1, 2, 4, ..., 256}");
    MZB_CHECK_ARG(z && gamma && beta && save_mean && save_invstd && scratch && (y || y_f32), "null pointer");
    MZB_CHECK_ARG(dtype == MZ_BF16 || dtype == MZ_F16, "y / res are 16-bit: dtype must be MZ_BF16 or MZ_F16");
    MZB_CHECK_ARG(act == MZ_ACT_NONE || act == MZ_ACT_RELU || act == MZ_ACT_LEAKY_RELU, "activation not built for training (none / relu / leaky_relu)");
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = bn_blocks(M), lanes = BN_THREADS / (C / 4);
    const size_t smem = (size_t)lanes * 2 * C * sizeof(double);
    MZB_CUDA(mzb::launch_chain(bn_reduce_kernel, dim3(nb), dim3(BN_THREADS), smem, st, M, C, BN_ROWS, 0, z, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, (double *)scratch));
    MZB_LAUNCH_CHECK();
    MZB_CUDA(mzb::launch_chain_small(bn_fwd_finalize_kernel, dim3((C * 32 + BN_THREADS - 1) / BN_THREADS), dim3(BN_THREADS), 0, st, M, C, nb, (const double *)scratch, eps, momentum, running_mean, running_var, save_mean, save_invstd));
    MZB_LAUNCH_CHECK();
    const size_t total4 = (size_t)M * C / 4;
    BN_DISPATCH_U(MZB_CUDA(mzb::launch_chain(bn_fwd_apply_kernel<U_>, dim3((unsigned)((total4 + (size_t)BN_THREADS * bn_u() - 1) / ((size_t)BN_THREADS * bn_u()))), dim3(BN_THREADS), 0, st, total4, C, z, save_mean, save_invstd, gamma, beta, (const uint16_t *)res,
                                                                                                 dtype == MZ_F16, act, (uint16_t *)y, y_f32)));
    MZB_LAUNCH_CHECK();
    return 0;
}

int mz_bn_train_fwd_pre(int M, int C, int nblocks, const double *partial, const float *z, const float *gamma, const float *beta, const void *res,
                        int dtype, int act, double eps, double momentum, float *running_mean, float *running_var, float *save_mean,
                        float *save_invstd, void *y, float *y_f32, void *stream)
{
    MZB_CHECK_ARG(bn_shape_ok(M, C) && nblocks > 0, "M must be positive and C one of 4 * {1, 2, 4, ..., 256}");
    MZB_CHECK_ARG(partial && z && gamma && beta && save_mean && save_invstd && (y || y_f32), "null pointer");
    MZB_CHECK_ARG(dtype == MZ_BF16 || dtype == MZ_F16, "y / res are 16-bit: dtype must be MZ_BF16 or MZ_F16");
    MZB_CHECK_ARG(act == MZ_ACT_NONE || act == MZ_ACT_RELU || act == MZ_ACT_LEAKY_RELU, "activation not built for training (none / relu / leaky_relu)");
    cudaStream_t st = (cudaStream_t)stream;
    MZB_CUDA(mzb::launch_chain_small(bn_fwd_finalize_kernel, dim3((C * 32 + BN_THREADS - 1) / BN_THREADS), dim3(BN_THREADS), 0, st, M, C, nblocks, partial, eps, momentum,
                                     running_mean, running_var, save_mean, save_invstd));
    MZB_LAUNCH_CHECK();
    const size_t total4 = (size_t)M * C / 4;
    BN_DISPATCH_U(MZB_CUDA(mzb::launch_chain(bn_fwd_apply_kernel<U_>, dim3((unsigned)((total4 + (size_t)BN_THREADS * bn_u() - 1) / ((size_t)BN_THREADS * bn_u()))), dim3(BN_THREADS), 0, st, total4, C, z, save_mean, save_invstd,
                               gamma, beta, (const uint16_t *)res, dtype == MZ_F16, act, (uint16_t *)y, y_f32)));
    MZB_LAUNCH_CHECK();
    return 0;
}

int mz_colsum(int M, int C, const float *x, float *out, int accumulate, void *scratch, void *stream)
{
    MZB_CHECK_ARG(bn_shape_ok(M, C), "M must be positive and C one of 4 * {1, 2, 4, ..., 256}");
    MZB_CHECK_ARG(x && out && scratch, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = bn_blocks(M), lanes = BN_THREADS / (C / 4);
    const size_t smem = (size_t)lanes * 2 * C * sizeof(double);
    MZB_CUDA(mzb::launch_chain(bn_reduce_kernel, dim3(nb), dim3(BN_THREADS), smem, st, M, C, BN_ROWS, 0, x, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, (double *)scratch));
    MZB_LAUNCH_CHECK();
    MZB_CUDA(mzb::launch_chain_small(colsum_finalize_kernel, dim3((C * 32 + BN_THREADS - 1) / BN_THREADS), dim3(BN_THREADS), 0, st, C, nb, (const double *)scratch, out, accumulate));
    MZB_LAUNCH_CHECK();
    return 0;
}

int mz_bn_train_bwd(int M, int C, const float *z, const float *dy, const float *gamma, const float *beta, const void *res, int dtype, int act,
                    const float *save_mean, const float *save_invstd, float *dgamma, float *dbeta, float *dz, void *dz16, float *dres, void *scratch,
                    void *stream)
{
    return mz_bn_train_bwd_mixed(M, C, z, dy, gamma, beta, res, dtype, dtype, act, save_mean, save_invstd, dgamma, dbeta, dz, dz16, dres, scratch, stream);
}

int mz_bn_train_bwd_mixed(int M, int C, const float *z, const float *dy, const float *gamma, const float *beta, const void *res, int dtype, int dz_dtype,
                          int act, const float *save_mean, const float *save_invstd, float *dgamma, float *dbeta, float *dz, void *dz16, float *dres,
                          void *scratch, void *stream)
{
    return mz_bn_train_bwd_acc(M, C, z, dy, gamma, beta, res, dtype, dz_dtype, act, save_mean, save_invstd, dgamma, dbeta, nullptr, nullptr, dz, dz16, dres,
                               scratch, stream);
}

int mz_bn_train_bwd_acc(int M, int C, const float *z, const float *dy, const float *gamma, const float *beta, const void *res, int dtype, int dz_dtype,
                        int act, const float *save_mean, const float *save_invstd, float *dgamma, float *dbeta, float *dgamma_acc, float *dbeta_acc,
                        float *dz, void *dz16, float *dres, void *scratch, void *stream)
{
    MZB_CHECK_ARG(dz_dtype == MZ_BF16 || dz_dtype == MZ_F16, "dz16 is 16-bit: dz_dtype must be MZ_BF16 or MZ_F16");
    MZB_CHECK_ARG(bn_shape_ok(M, C), "M must be positive and C one of 4 * {1, 2, 4, ..., 256}");
    MZB_CHECK_ARG(z && dy && gamma && beta && save_mean && save_invstd && dgamma && dbeta && scratch && (dz || dz16), "null pointer");
    MZB_CHECK_ARG(dtype == MZ_BF16 || dtype == MZ_F16, "dz16 / res are 16-bit: dtype must be MZ_BF16 or MZ_F16");
    MZB_CHECK_ARG(act == MZ_ACT_NONE || act == MZ_ACT_RELU || act == MZ_ACT_LEAKY_RELU, "activation not built for training (none / relu / leaky_relu)");
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = bn_blocks(M), lanes = BN_THREADS / (C / 4);
    const size_t smem = (size_t)lanes * 2 * C * sizeof(double);
    MZB_CUDA(mzb::launch_chain(bn_reduce_kernel, dim3(nb), dim3(BN_THREADS), smem, st, M, C, BN_ROWS, 1, z, dy, save_mean, save_invstd, gamma, beta, (const uint16_t *)res, dtype == MZ_F16, act,
                                                   (double *)scratch));
    MZB_LAUNCH_CHECK();
    MZB_CUDA(mzb::launch_chain_small(bn_bwd_finalize_kernel, dim3((C * 32 + BN_THREADS - 1) / BN_THREADS), dim3(BN_THREADS), 0, st, C, nb, (const double *)scratch, dgamma, dbeta, dgamma_acc, dbeta_acc));
    MZB_LAUNCH_CHECK();
    const size_t total4 = (size_t)M * C / 4;
    BN_DISPATCH_U(MZB_CUDA(mzb::launch_chain(bn_bwd_apply_kernel<U_>, dim3((unsigned)((total4 + (size_t)BN_THREADS * bn_u() - 1) / ((size_t)BN_THREADS * bn_u()))), dim3(BN_THREADS), 0, st, total4, M, C, z, dy, save_mean, save_invstd, gamma, beta,
                                                                                                 (const uint16_t *)res, dtype == MZ_F16, dz_dtype == MZ_F16, act, dgamma, dbeta, dz,
                                                                                                 (uint16_t *)dz16, dres)));
    MZB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
