// MuZero network evaluation on B200: the op-program runner (mz_run) and the CUDA-core kernels:
// exact fp32 implicit-GEMM convolution (the 1e-5 parity path; also runs bf16 activations as the
// on-device cross-check of the tcgen05 kernel in conv_tc.cu), fused heads (Linear + softmax +
// support expectation + inverse transform), _scale_state, average pool and layout conversion.
//
// Replaces (behaviour, not code) src/networks.py ConvBlock :7-17, ResidualBlock :19-35, the three
// network forwards :94-99,151-167,225-241, MuZeroAgent._scale_state :314-328 and
// utils.py ScalarTransforms.inverted_softmax_expectation :74-81 of the reference.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cooperative_groups.h>
#include <cuda_fp8.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace mzb {
int conv_tc_launch(const mz_op &op, int nsamples, cudaStream_t st);   // conv_tc.cu
}

namespace {

namespace cg = cooperative_groups;

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// 16-bit residual stream + e4m3 correction (see tc_common.cuh: split2 / lo2): LO_SCALE = 2048 (fp16) / 256 (bf16)
template <typename T> __device__ __forceinline__ float lo_scale_of() { return 1.0f; }
template <> __device__ __forceinline__ float lo_scale_of<__half>() { return 2048.0f; }
template <> __device__ __forceinline__ float lo_scale_of<__nv_bfloat16>() { return 256.0f; }
__device__ __forceinline__ float lo_decode(uint8_t b)
{
    const __half_raw r = __nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)b, __NV_E4M3);
    return __half2float(*reinterpret_cast<const __half *>(&r));
}
__device__ __forceinline__ uint8_t lo_encode(float d) { return (uint8_t)__nv_cvt_float_to_fp8(d, __NV_SATFINITE, __NV_E4M3); }

__device__ __forceinline__ float activate(float v, int act)
{
    switch (act) {
        case MZ_ACT_RELU: return fmaxf(v, 0.0f);
        case MZ_ACT_LEAKY_RELU: return v > 0.0f ? v : 0.01f * v;
        case MZ_ACT_SILU: return v / (1.0f + expf(-v));
        case MZ_ACT_GELU: return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
        default: return v;
    }
}

// ------------------------------------------------------------------------------------------------
// Implicit-GEMM convolution on CUDA cores.  M = n*H*W output pixels, N = cout, K = k*k*cin.
// 64x64 output tile per CTA, K in chunks of 16, 256 threads x (4x4) outputs, fp32 FFMA accumulation.
constexpr int TM = 64, TN = 64, TK = 16;

template <typename T>
__global__ void __launch_bounds__(256)
conv_simt_kernel(int M, int H, int W, int cin, int cout, int ksize, int act, const T *__restrict__ src, T *__restrict__ dst,
                 const T *__restrict__ res, float *__restrict__ dst_f32, const T *__restrict__ w, const float *__restrict__ scale,
                 const float *__restrict__ shift, const float *__restrict__ act_bias, const int *__restrict__ act_idx,
                 const uint8_t *__restrict__ res_lo, uint8_t *__restrict__ dst_lo, const float *__restrict__ res_f32)
{
    __shared__ float As[TK][TM + 4];
    __shared__ float Bs[TK][TN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
    const int HW = H * W, K = ksize * ksize * cin, pad = ksize / 2;
    // loader roles: thread -> (row, 4-channel group) of the 64 x 16 A chunk / (col, 4-k group) of B
    const int lr = tid >> 2, lc = (tid & 3) * 4;
    const int am = m0 + lr;
    const int a_s = am / HW, a_p = am - a_s * HW, a_y = a_p / W, a_x = a_p - a_y * W;
    const int ty = tid >> 4, tx = tid & 15;   // compute roles: 4 rows x 4 cols each
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    for (int tap = 0; tap < ksize * ksize; ++tap) {
        const int dy = tap / ksize - pad, dx = tap % ksize - pad;
        const int yy = a_y + dy, xx = a_x + dx;
        const bool inb = am < M && yy >= 0 && yy < H && xx >= 0 && xx < W;
        const T *ap = src + ((size_t)(a_s * HW + yy * W + xx)) * cin + lc;
        const T *bp = w + (size_t)(n0 + lr) * K + tap * cin + lc;
        for (int c0 = 0; c0 < cin; c0 += TK) {
            float a4[4] = {0.f, 0.f, 0.f, 0.f}, b4[4];
            if (inb) {
#pragma unroll
                for (int q = 0; q < 4; ++q) a4[q] = to_f(ap[c0 + q]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) b4[q] = to_f(bp[c0 + q]);
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 4; ++q) { As[lc + q][lr] = a4[q]; Bs[lc + q][lr] = b4[q]; }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < TK; ++k) {
                const float4 a = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
                const float4 b = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
                const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
        }
    }
    // epilogue: (+ action bias) * scale + shift (+ residual) -> activation
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
        const int s = m / HW, p = m - s * HW;
        const float *ab = act_bias ? act_bias + ((size_t)act_idx[s] * HW + p) * cout : nullptr;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            float v = acc[i][j];
            if (ab) v += ab[n];
            v = v * (scale ? scale[n] : 1.0f) + shift[n];
            if (res) v += to_f(res[(size_t)m * cout + n]);
            if (res_lo) v += lo_decode(res_lo[(size_t)m * cout + n]) * (1.0f / lo_scale_of<T>());
            if (res_f32) v += res_f32[(size_t)m * cout + n];
            v = activate(v, act);
            const T hi = from_f<T>(v);
            dst[(size_t)m * cout + n] = hi;
            if (dst_lo) dst_lo[(size_t)m * cout + n] = lo_encode((v - to_f(hi)) * lo_scale_of<T>());
            if (dst_f32) dst_f32[(size_t)m * cout + n] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void pool2_kernel(size_t total, int H, int W, int C, const T *__restrict__ src, T *__restrict__ dst, float *__restrict__ dst_f32)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int Ho = H / 2, Wo = W / 2;
    const int c = (int)(i % C);
    size_t r = i / C;
    const int xo = (int)(r % Wo); r /= Wo;
    const int yo = (int)(r % Ho);
    const size_t s = r / Ho;
    const T *b = src + ((s * H + 2 * yo) * W + 2 * xo) * C + c;
    const float v = ((to_f(b[0]) + to_f(b[C])) + (to_f(b[(size_t)W * C]) + to_f(b[(size_t)W * C + C]))) * 0.25f;
    dst[i] = from_f<T>(v);
    if (dst_f32) dst_f32[i] = v;
}

// _scale_state: one CTA per sample.  Fast path (elems == 5120: the 256x4x5 latent): each of the 256 threads
// keeps its 20 values in registers (five float4 loads), block min/max, normalise, vectorised stores -- one read
// of the fp32 source and one write per destination.
template <typename T> __device__ __forceinline__ void store4(T *p, float a, float b, float c, float d);
template <> __device__ __forceinline__ void store4<float>(float *p, float a, float b, float c, float d)
{
    *reinterpret_cast<float4 *>(p) = make_float4(a, b, c, d);
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16 *p, float a, float b, float c, float d)
{
    __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t *>(&lo);
    u.y = *reinterpret_cast<uint32_t *>(&hi);
    *reinterpret_cast<uint2 *>(p) = u;
}

template <> __device__ __forceinline__ void store4<__half>(__half *p, float a, float b, float c, float d)
{
    __half2 lo = __floats2half2_rn(a, b), hi = __floats2half2_rn(c, d);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t *>(&lo);
    u.y = *reinterpret_cast<uint32_t *>(&hi);
    *reinterpret_cast<uint2 *>(p) = u;
}

template <typename T>
__global__ void __launch_bounds__(256)
scale_state_kernel(int elems, const float *__restrict__ src, T *__restrict__ dst, T *__restrict__ dst2,
                   const int *__restrict__ dst2_slot, long long dst2_stride)
{
    __shared__ float s_lo[8], s_hi[8];
    const int i = blockIdx.x, tid = threadIdx.x;
    const float *x = src + (size_t)i * elems;
    T *d1 = dst ? dst + (size_t)i * elems : nullptr;
    T *d2 = dst2 ? dst2 + ((size_t)i * dst2_stride + (dst2_slot ? dst2_slot[i] : 0)) * elems : nullptr;
    const bool fast = elems == 256 * 20;
    float4 v[5];
    float lo = INFINITY, hi = -INFINITY;
    if (fast) {
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            v[q] = __ldcs(reinterpret_cast<const float4 *>(x) + q * 256 + tid);
            lo = fminf(fminf(lo, fminf(v[q].x, v[q].y)), fminf(v[q].z, v[q].w));
            hi = fmaxf(fmaxf(hi, fmaxf(v[q].x, v[q].y)), fmaxf(v[q].z, v[q].w));
        }
    } else {
        for (int e = tid; e < elems; e += 256) { const float t = x[e]; lo = fminf(lo, t); hi = fmaxf(hi, t); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if ((tid & 31) == 0) { s_lo[tid >> 5] = lo; s_hi[tid >> 5] = hi; }
    __syncthreads();
    lo = s_lo[0]; hi = s_hi[0];
#pragma unroll
    for (int q = 1; q < 8; ++q) { lo = fminf(lo, s_lo[q]); hi = fmaxf(hi, s_hi[q]); }
    const float den = __fadd_rn(__fsub_rn(hi, lo), 1e-8f);                 // s_max - s_min + 1e-8  (:327)
    if (fast) {
        // fp32 outputs (the 1e-5 parity path): the reference's correctly rounded division.  16-bit outputs: one reciprocal per
        // sample and a multiply -- the <= 1.5 ulp(fp32) difference disappears in the rounding to bf16 / fp16, and the IEEE
        // division sequence was what bound this kernel (ncu: SM 72 % busy at 30 % of the DRAM rate)
        const float inv = __frcp_rn(den);
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            float a, b, c, d;
            if (sizeof(T) == 4) {
                a = __fdiv_rn(__fsub_rn(v[q].x, lo), den); b = __fdiv_rn(__fsub_rn(v[q].y, lo), den);
                c = __fdiv_rn(__fsub_rn(v[q].z, lo), den); d = __fdiv_rn(__fsub_rn(v[q].w, lo), den);
            } else {
                a = __fmul_rn(__fsub_rn(v[q].x, lo), inv); b = __fmul_rn(__fsub_rn(v[q].y, lo), inv);
                c = __fmul_rn(__fsub_rn(v[q].z, lo), inv); d = __fmul_rn(__fsub_rn(v[q].w, lo), inv);
            }
            const int e = (q * 256 + tid) * 4;
            if (d1) store4<T>(d1 + e, a, b, c, d);
            if (d2) store4<T>(d2 + e, a, b, c, d);
        }
    } else {
        for (int e = tid; e < elems; e += 256) {
            const T t = from_f<T>(__fdiv_rn(__fsub_rn(x[e], lo), den));
            if (d1) d1[e] = t;
            if (d2) d2[e] = t;
        }
    }
}

// Flatten + Linear + (softmax -> expectation -> inverse transform | softmax).  One CTA per HEAD_SAMPLES
// samples: every thread owns a strided slice of the feature axis and keeps HEAD_SAMPLES x nout partial sums,
// so each weight element is read once per CTA (not once per sample) and the reads are coalesced.
constexpr int HEAD_MAX_OUT = 16;
constexpr int HEAD_SAMPLES = 4;      // 4096 samples = 1024 CTAs, 4 per SM: 8 samples per CTA gave 512 CTAs on 444 slots, a half-empty second wave
constexpr int HEAD_THREADS = 128;

// eight consecutive features of one sample as floats (one 16-byte load for bf16, two for fp32)
__device__ __forceinline__ void load8(const float *p, float (&f)[8])
{
    const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16 *p, float (&f)[8])
{
    const uint4 u = __ldg(reinterpret_cast<const uint4 *>(p));
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
    for (int q = 0; q < 4; ++q) { const float2 t = __bfloat1622float2(h[q]); f[2 * q] = t.x; f[2 * q + 1] = t.y; }
}

__device__ __forceinline__ void load8(const __half *p, float (&f)[8])
{
    const uint4 u = __ldg(reinterpret_cast<const uint4 *>(p));
    const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
    for (int q = 0; q < 4; ++q) { const float2 t = __half22float2(h[q]); f[2 * q] = t.x; f[2 * q + 1] = t.y; }
}

template <typename T, int NOUT>
__global__ void __launch_bounds__(HEAD_THREADS, 4)
head_kernel(int n, int feat, int cin, int cstride, int mode, const T *__restrict__ src, const float *__restrict__ w, const float *__restrict__ bias,
            float *__restrict__ out, float *__restrict__ out_logits)
{
    // src rows are pixels of cstride channels of which this head reads cin (cstride == cin: a dense [n][feat] tensor; the policy and value
    // heads read the two halves of one 256-channel buffer that a single trunk layer wrote)
    const int pixels = feat / cin;
    __shared__ float s_part[HEAD_THREADS / 32][HEAD_SAMPLES][NOUT];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int s0 = blockIdx.x * HEAD_SAMPLES;
    const int ns = min(HEAD_SAMPLES, n - s0);
    float acc[HEAD_SAMPLES][NOUT];
#pragma unroll
    for (int s = 0; s < HEAD_SAMPLES; ++s)
#pragma unroll
        for (int o = 0; o < NOUT; ++o) acc[s][o] = 0.0f;
    for (int e = tid * 8; e < feat; e += HEAD_THREADS * 8) {      // feat is a multiple of 8 (checked by the launcher)
        float xv[HEAD_SAMPLES][8];
#pragma unroll
        for (int s = 0; s < HEAD_SAMPLES; ++s) {
            if (s < ns) load8(src + ((size_t)(s0 + s) * pixels + e / cin) * cstride + e % cin, xv[s]);
            else {
#pragma unroll
                for (int q = 0; q < 8; ++q) xv[s][q] = 0.0f;
            }
        }
#pragma unroll
        for (int o = 0; o < NOUT; ++o) {
            float wv[8];
            load8(w + (size_t)o * feat + e, wv);
#pragma unroll
            for (int s = 0; s < HEAD_SAMPLES; ++s)
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[s][o] = fmaf(xv[s][q], wv[q], acc[s][o]);
        }
    }
#pragma unroll
    for (int s = 0; s < HEAD_SAMPLES; ++s)
#pragma unroll
        for (int o = 0; o < NOUT; ++o) {
            float v = acc[s][o];
#pragma unroll
            for (int sh = 16; sh > 0; sh >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sh);
            if (lane == 0) s_part[wid][s][o] = v;
        }
    __syncthreads();
    if (tid >= ns) return;                                       // thread s finishes sample s
    const int i = s0 + tid;
    float logit[NOUT];
    float mx = -INFINITY;
#pragma unroll
    for (int o = 0; o < NOUT; ++o) {
        float v = 0.0f;
#pragma unroll
        for (int q = 0; q < HEAD_THREADS / 32; ++q) v += s_part[q][tid][o];
        logit[o] = v + bias[o];
        mx = fmaxf(mx, logit[o]);
        if (out_logits) out_logits[(size_t)i * NOUT + o] = logit[o];
    }
    if (mode == 0) return;
    float den = 0.0f, e[NOUT];
#pragma unroll
    for (int o = 0; o < NOUT; ++o) { e[o] = expf(logit[o] - mx); den += e[o]; }
    if (mode == 2) {                                                   // softmax probabilities (mcts.py:100,199)
#pragma unroll
        for (int o = 0; o < NOUT; ++o) out[(size_t)i * NOUT + o] = e[o] / den;
        return;
    }
    // utils.py:66-81: supports = linspace(-5, 5, 11) (integers), x = sum p*s, y = sign(x)((|x| + 0.999)^2 - 1)
    const float half = 0.5f * (float)(NOUT - 1);
    float ex = 0.0f;
#pragma unroll
    for (int o = 0; o < NOUT; ++o) ex += (e[o] / den) * ((float)o - half);
    const float sg = ex > 0.0f ? 1.0f : (ex < 0.0f ? -1.0f : 0.0f);
    const float t = fabsf(ex) + 0.999f;                                  // float32(1 - epsilon), utils.py:14,28
    out[i] = sg * (t * t - 1.0f);
}

template <typename T>
int launch_head(const mz_op &o, int n, cudaStream_t st)
{
    const int feat = o.H * o.W * o.cin;
    if (feat % 8) { mzb::set_error("head: feature count %d is not a multiple of 8", feat); return -1; }
    const int grid = (n + HEAD_SAMPLES - 1) / HEAD_SAMPLES;
    const int cstride = o.cout > 0 ? o.cout : o.cin;
    if (o.cin % 8 || cstride % 8 || cstride < o.cin) { mzb::set_error("head: channels %d / row stride %d", o.cin, cstride); return -1; }
#define MZB_HEAD(NO) head_kernel<T, NO><<<grid, HEAD_THREADS, 0, st>>>(n, feat, o.cin, cstride, o.head_mode, (const T *)o.src, (const float *)o.w, o.shift, o.out, o.out_logits)
    switch (o.nout) {
        case 3: MZB_HEAD(3); break;
        case 11: MZB_HEAD(11); break;
        default:
            mzb::set_error("head: nout=%d is not built (3 actions / 11 supports, config.yaml:6,30)", o.nout);
            return -1;
    }
#undef MZB_HEAD
    return 0;
}


// ------------------------------------------------------------------------------------------------
// 16-bit heads on the warp MMA (round 2).  The CUDA-core head_kernel above re-reads the fp32 weight matrix once per 4 samples
// (225 KB x 1024 CTAs through L1 at 4096 samples: 37-49 us for 42 MB of activations).  Here a cluster of HC_KS CTAs takes 128 samples:
// CTA r of the cluster owns the r-th slice of the feature axis, stages that slice of the weights ONCE in shared memory -- split into
// hi + lo / LO 16-bit halves (two MMAs: the products carry the full fp32 weight, the 16-bit activations are exact) and already in
// `mma.sync.m16n8k16` B-fragment order -- and each of its 8 warps runs one 16-sample m-tile over the slice with the activations going
// from global memory straight into A fragments by 16-byte loads: the reduction index of a dot product may be permuted as long as both
// operands agree, so the thread that owns k columns {2t, 2t+1, 2t+8, 2t+9} of two consecutive MMAs simply loads the 8 consecutive
// features [8t, 8t+8) of its row.  The slices' partial sums meet through distributed shared memory in a fixed order (deterministic):
// CTA r adds the HC_KS partials of samples [16r, 16r+16) and finishes them (bias, softmax / support expectation / inverse transform).
// Up to two heads that share the batch (policy + value) run as the two z-slices of one launch.
struct HeadDesc {
    const void *src;
    const float *w, *bias;
    float *out, *out_logits;
    int nout, mode, feat, cin, cstride;
};
struct HeadPair {
    HeadDesc h[2];
};

template <typename T> __device__ __forceinline__ uint32_t pack_hi_lo(float a, float b, uint32_t &lo);
template <> __device__ __forceinline__ uint32_t pack_hi_lo<__half>(float a, float b, uint32_t &lo)
{
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn((a - hf.x) * 2048.0f, (b - hf.y) * 2048.0f);
    lo = *reinterpret_cast<const uint32_t *>(&l);
    return *reinterpret_cast<const uint32_t *>(&h);
}
template <> __device__ __forceinline__ uint32_t pack_hi_lo<__nv_bfloat16>(float a, float b, uint32_t &lo)
{
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn((a - hf.x) * 256.0f, (b - hf.y) * 256.0f);
    lo = *reinterpret_cast<const uint32_t *>(&l);
    return *reinterpret_cast<const uint32_t *>(&h);
}
template <typename T> __device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1);
template <> __device__ __forceinline__ void mma16816<__half>(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <> __device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int HC_KS = 8;            // CTAs per cluster = slices of the feature axis
constexpr int HC_WARPS = 8;         // one 16-sample m-tile per warp
constexpr int HC_SAMPLES = 16 * HC_WARPS;
constexpr int HC_UNROLL = 5;        // 32-feature blocks whose activation loads are in flight together (10 x 16 B per thread)
constexpr int HC_MAX_KB = 20;       // 32-feature blocks per slice: 5120 features / 32 / HC_KS
constexpr size_t HC_SMEM = (size_t)HC_MAX_KB * 2 * 2 * 32 * 16 + HC_SAMPLES * 16 * 4 + 16 * 16 * 4;

template <typename T>
__global__ void __cluster_dims__(1, HC_KS, 1) __launch_bounds__(HC_WARPS * 32)
head_mma_kernel(int n, const __grid_constant__ HeadPair hp)
{
    extern __shared__ __align__(16) uint8_t hc_smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const HeadDesc &hd = hp.h[blockIdx.z];
    const int rank = blockIdx.y;                                  // cluster dims (1, HC_KS, 1): the cluster rank is blockIdx.y
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, t = lane & 3;
    const int nkb = hd.feat / (32 * HC_KS);                        // launcher: a multiple of HC_UNROLL, <= HC_MAX_KB
    const int nt = hd.nout <= 8 ? 1 : 2;                           // n-tiles of 8 outputs
    uint4 *s_b = reinterpret_cast<uint4 *>(hc_smem);              // [nkb][2 n-tiles][hi, lo][32 lanes]: B fragments of both MMAs of a block
    float(*s_part)[16] = reinterpret_cast<float(*)[16]>(hc_smem + (size_t)HC_MAX_KB * 2 * 2 * 32 * 16);
    float(*s_fin)[16] = s_part + HC_SAMPLES;

    // weight slice -> shared memory, split and in fragment order; every global load of the staging and of the first activation blocks is
    // issued before the first use (one DRAM / L2 latency for the lot)
    constexpr int ST = HC_MAX_KB * 2 * 32 / (HC_WARPS * 32);       // staging units per thread (5)
    float4 w0[ST], w1[ST];
#pragma unroll
    for (int it = 0; it < ST; ++it) {
        const int idx = tid + it * HC_WARPS * 32;
        const int l = idx & 31, j = (idx >> 5) % nt, kb = (idx >> 5) / nt;
        const int o = j * 8 + (l >> 2), e = (rank * nkb + kb) * 32 + (l & 3) * 8;
        w0[it] = w1[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < nkb * nt * 32 && o < hd.nout) {
            w0[it] = __ldg(reinterpret_cast<const float4 *>(hd.w + (size_t)o * hd.feat + e));
            w1[it] = __ldg(reinterpret_cast<const float4 *>(hd.w + (size_t)o * hd.feat + e) + 1);
        }
    }
    const int s0 = blockIdx.x * HC_SAMPLES;
    const T *src = (const T *)hd.src;
    const int pixels = hd.feat / hd.cin;
    const T *row0 = src + (size_t)min(s0 + wid * 16 + g, n - 1) * pixels * hd.cstride;
    const T *row1 = src + (size_t)min(s0 + wid * 16 + 8 + g, n - 1) * pixels * hd.cstride;
    const int e_base = rank * nkb * 32 + t * 8;
    uint4 a[HC_UNROLL][2];
#pragma unroll
    for (int q = 0; q < HC_UNROLL; ++q) {
        const int e = e_base + q * 32;
        const int off = (e / hd.cin) * hd.cstride + e % hd.cin;
        a[q][0] = __ldcs(reinterpret_cast<const uint4 *>(row0 + off));
        a[q][1] = __ldcs(reinterpret_cast<const uint4 *>(row1 + off));
    }
#pragma unroll
    for (int it = 0; it < ST; ++it) {
        const int idx = tid + it * HC_WARPS * 32;
        if (idx < nkb * nt * 32) {
            const int l = idx & 31, j = (idx >> 5) % nt, kb = (idx >> 5) / nt;
            uint4 hi, lo;
            hi.x = pack_hi_lo<T>(w0[it].x, w0[it].y, lo.x); hi.y = pack_hi_lo<T>(w0[it].z, w0[it].w, lo.y);
            hi.z = pack_hi_lo<T>(w1[it].x, w1[it].y, lo.z); hi.w = pack_hi_lo<T>(w1[it].z, w1[it].w, lo.w);
            s_b[((kb * 2 + j) * 2 + 0) * 32 + l] = hi;
            s_b[((kb * 2 + j) * 2 + 1) * 32 + l] = lo;
        }
    }
    __syncthreads();

    float acc[2][2][4];                                            // [n-tile][hi, lo][fragment]
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[q >> 3][(q >> 2) & 1][q & 3] = 0.0f;
    for (int kb0 = 0; kb0 < nkb; kb0 += HC_UNROLL) {
        if (kb0 > 0) {
#pragma unroll
            for (int q = 0; q < HC_UNROLL; ++q) {
                const int e = e_base + (kb0 + q) * 32;
                const int off = (e / hd.cin) * hd.cstride + e % hd.cin;
                a[q][0] = __ldcs(reinterpret_cast<const uint4 *>(row0 + off));
                a[q][1] = __ldcs(reinterpret_cast<const uint4 *>(row1 + off));
            }
        }
#pragma unroll
        for (int q = 0; q < HC_UNROLL; ++q) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (j < nt) {
                    const uint4 bh = s_b[(((kb0 + q) * 2 + j) * 2 + 0) * 32 + lane], bl = s_b[(((kb0 + q) * 2 + j) * 2 + 1) * 32 + lane];
                    mma16816<T>(acc[j][0], a[q][0].x, a[q][1].x, a[q][0].y, a[q][1].y, bh.x, bh.y);
                    mma16816<T>(acc[j][0], a[q][0].z, a[q][1].z, a[q][0].w, a[q][1].w, bh.z, bh.w);
                    mma16816<T>(acc[j][1], a[q][0].x, a[q][1].x, a[q][0].y, a[q][1].y, bl.x, bl.y);
                    mma16816<T>(acc[j][1], a[q][0].z, a[q][1].z, a[q][0].w, a[q][1].w, bl.z, bl.w);
                }
            }
        }
    }
    const float inv_lo = 1.0f / lo_scale_of<T>();
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q)                               // c0,c1: row g, cols 2t,2t+1; c2,c3: row g+8
            s_part[wid * 16 + (q >> 1) * 8 + g][j * 8 + 2 * t + (q & 1)] = acc[j][0][q] + acc[j][1][q] * inv_lo;
    cluster.sync();
    {   // CTA `rank` adds the slices' partials of samples [16 rank, 16 rank + 16) in slice order
        const int sl = tid >> 4, o = tid & 15;
        float v = 0.0f;
#pragma unroll
        for (int r = 0; r < HC_KS; ++r) v += cluster.map_shared_rank(&s_part[0][0], r)[(rank * 16 + sl) * 16 + o];
        s_fin[sl][o] = v;
    }
    cluster.sync();                                                // nobody leaves while its partials are still being read
    const int i = s0 + rank * 16 + tid;
    if (tid >= 16 || i >= n) return;                              // thread s finishes sample s
    const int nout = hd.nout;
    float logit[16];
    float mx = -INFINITY;
#pragma unroll
    for (int o = 0; o < 16; ++o) {
        logit[o] = o < nout ? s_fin[tid][o] + hd.bias[o] : -INFINITY;
        mx = fmaxf(mx, logit[o]);
        if (hd.out_logits && o < nout) hd.out_logits[(size_t)i * nout + o] = logit[o];
    }
    if (hd.mode == 0) return;
    float den = 0.0f, e[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) { e[o] = o < nout ? expf(logit[o] - mx) : 0.0f; den += e[o]; }
    if (hd.mode == 2) {                                                // softmax probabilities (mcts.py:100,199)
#pragma unroll
        for (int o = 0; o < 16; ++o)
            if (o < nout) hd.out[(size_t)i * nout + o] = e[o] / den;
        return;
    }
    // utils.py:66-81, as in head_kernel
    const float half = 0.5f * (float)(nout - 1);
    float ex = 0.0f;
#pragma unroll
    for (int o = 0; o < 16; ++o)
        if (o < nout) ex += (e[o] / den) * ((float)o - half);
    const float sg = ex > 0.0f ? 1.0f : (ex < 0.0f ? -1.0f : 0.0f);
    const float tt = fabsf(ex) + 0.999f;
    hd.out[i] = sg * (tt * tt - 1.0f);
}

inline bool head_mma_ok(const mz_op &o)
{
    const int cstride = o.cout > 0 ? o.cout : o.cin;
    const int feat = o.H * o.W * o.cin;
    return o.op == MZ_OP_HEAD && (o.dtype == MZ_BF16 || o.dtype == MZ_F16) && o.cin % 32 == 0 && cstride % 8 == 0 && cstride >= o.cin && o.nout >= 1 &&
           o.nout <= 16 && feat % (32 * HC_KS * HC_UNROLL) == 0 && feat / (32 * HC_KS) <= HC_MAX_KB && o.src && o.w && o.shift &&
           (o.head_mode == 0 ? o.out_logits != nullptr : o.out != nullptr) && ((uintptr_t)o.src % 16 == 0) && ((uintptr_t)o.w % 16 == 0);
}

// one launch for ops[0 .. cnt) (cnt 1 or 2, all head_mma_ok, same dtype)
int launch_head_mma(const mz_op *ops, int cnt, int n, cudaStream_t st)
{
    HeadPair hp;
    for (int i = 0; i < 2; ++i) {
        const mz_op &o = ops[i < cnt ? i : 0];
        hp.h[i] = HeadDesc{o.src, (const float *)o.w, o.shift, o.out, o.out_logits, o.nout, o.head_mode, o.H * o.W * o.cin, o.cin, o.cout > 0 ? o.cout : o.cin};
    }
    static bool attr_done[64];
    if (mzb::first_use_on_device(attr_done)) {
        MZB_CUDA(cudaFuncSetAttribute(head_mma_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HC_SMEM));
        MZB_CUDA(cudaFuncSetAttribute(head_mma_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HC_SMEM));
    }
    const dim3 grid((n + HC_SAMPLES - 1) / HC_SAMPLES, HC_KS, cnt);
    if (ops[0].dtype == MZ_F16) head_mma_kernel<__half><<<grid, HC_WARPS * 32, HC_SMEM, st>>>(n, hp);
    else head_mma_kernel<__nv_bfloat16><<<grid, HC_WARPS * 32, HC_SMEM, st>>>(n, hp);
    MZB_LAUNCH_CHECK();
    return 0;
}

template <typename T>
__global__ void nchw_in_kernel(size_t total, int C, int HW, const float *__restrict__ src, T *__restrict__ dst, T *__restrict__ dst2,
                               const int *__restrict__ dst2_slot, long long dst2_stride)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;     // index into channels-last dst
    if (i >= total) return;
    const int c = (int)(i % C);
    const size_t r = i / C;
    const int p = (int)(r % HW);
    const size_t s = r / HW;
    const T v = from_f<T>(src[(s * C + c) * HW + p]);
    if (dst) dst[i] = v;
    if (dst2) dst2[((s * dst2_stride + (dst2_slot ? dst2_slot[s] : 0)) * HW + p) * C + c] = v;
}

template <typename T>
__global__ void nhwc_out_kernel(size_t total, int C, int HW, const T *__restrict__ src, float *__restrict__ dst)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;     // index into NCHW dst
    if (i >= total) return;
    const int p = (int)(i % HW);
    const size_t r = i / HW;
    const int c = (int)(r % C);
    const size_t s = r / C;
    dst[i] = to_f(src[(s * HW + p) * C + c]);
}

template <typename T>
int run_op(const mz_op &o, int n, cudaStream_t st)
{
    switch (o.op) {
        case MZ_OP_CONV: {
            MZB_CHECK_ARG(o.src && o.dst && o.w && o.shift, "conv: null pointer");
            MZB_CHECK_ARG((!o.res_lo || o.res) && (sizeof(T) == 2 || (!o.res_lo && !o.dst_lo)), "conv: correction planes belong to a 16-bit residual stream");
            MZB_CHECK_ARG((o.ksize == 1 || o.ksize == 3) && o.cin % TK == 0 && o.cout % TN == 0, "conv: unsupported shape");
            MZB_CHECK_ARG(!o.act_bias || o.act_idx, "conv: act_bias needs act_idx");
            const int M = n * o.H * o.W;
            dim3 grid((M + TM - 1) / TM, o.cout / TN);
            conv_simt_kernel<T><<<grid, 256, 0, st>>>(M, o.H, o.W, o.cin, o.cout, o.ksize, o.act, (const T *)o.src, (T *)o.dst,
                                                     (const T *)o.res, o.dst_f32, (const T *)o.w, o.scale, o.shift, o.act_bias, o.act_idx,
                                                     (const uint8_t *)o.res_lo, (uint8_t *)o.dst_lo, o.res_f32);
            break;
        }
        case MZ_OP_POOL2: {
            MZB_CHECK_ARG(o.src && o.dst && o.H % 2 == 0 && o.W % 2 == 0, "pool: bad argument");
            const size_t total = (size_t)n * (o.H / 2) * (o.W / 2) * o.cin;
            pool2_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(total, o.H, o.W, o.cin, (const T *)o.src, (T *)o.dst, o.dst_f32);
            break;
        }
        case MZ_OP_SCALE: {
            MZB_CHECK_ARG(o.src && (o.dst || o.dst2), "scale: null pointer");
            scale_state_kernel<T><<<n, 256, 0, st>>>(o.H * o.W * o.cin, (const float *)o.src, (T *)o.dst, (T *)o.dst2, o.dst2_slot, o.dst2_stride);
            break;
        }
        case MZ_OP_HEAD: {
            MZB_CHECK_ARG(o.src && o.w && o.shift && o.nout > 0 && o.nout <= HEAD_MAX_OUT, "head: bad argument");
            MZB_CHECK_ARG(o.head_mode == 0 ? o.out_logits != nullptr : o.out != nullptr, "head: missing output");
            if (int rc = launch_head<T>(o, n, st)) return rc;
            break;
        }
        case MZ_OP_NCHW_IN: {
            MZB_CHECK_ARG(o.src && (o.dst || o.dst2), "nchw_in: null pointer");
            const size_t total = (size_t)n * o.H * o.W * o.cin;
            nchw_in_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(total, o.cin, o.H * o.W, (const float *)o.src, (T *)o.dst, (T *)o.dst2,
                                                                               o.dst2_slot, o.dst2_stride);
            break;
        }
        case MZ_OP_NHWC_OUT: {
            MZB_CHECK_ARG(o.src && o.dst, "nhwc_out: null pointer");
            const size_t total = (size_t)n * o.H * o.W * o.cin;
            nhwc_out_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(total, o.cin, o.H * o.W, (const T *)o.src, (float *)o.dst);
            break;
        }
        default:
            mzb::set_error("mz_run: unknown op %d", o.op);
            return -1;
    }
    MZB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" int mz_run(const mz_op *ops, int n_ops, int nsamples, void *stream)
{
    static const bool g_head_simt = getenv("MZB_HEAD_SIMT") && atoi(getenv("MZB_HEAD_SIMT")) != 0;   // A/B switch: the CUDA-core heads
    MZB_CHECK_ARG(ops && n_ops > 0 && nsamples > 0, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    for (int i = 0; i < n_ops; ++i) {
        const mz_op &o = ops[i];
        int rc;
        if (head_mma_ok(o) && !g_head_simt) {                 // 16-bit heads: warp-MMA kernel, two heads of one batch per launch
            const int cnt = (i + 1 < n_ops && head_mma_ok(ops[i + 1]) && ops[i + 1].dtype == o.dtype) ? 2 : 1;
            if (int rc2 = launch_head_mma(ops + i, cnt, nsamples, st)) return rc2;
            i += cnt - 1;
            continue;
        }
        if (o.dtype == MZ_F32) rc = run_op<float>(o, nsamples, st);
        else if (o.dtype == MZ_BF16 || o.dtype == MZ_F16) {
            if (o.op == MZ_OP_CONV && o.use_tc) rc = mzb::conv_tc_launch(o, nsamples, st);
            else if (o.dtype == MZ_BF16) rc = run_op<__nv_bfloat16>(o, nsamples, st);
            else rc = run_op<__half>(o, nsamples, st);
        } else {
            mzb::set_error("mz_run: op %d has unknown dtype %d", i, o.dtype);
            return -1;
        }
        if (rc) return rc;
    }
    return 0;
}
