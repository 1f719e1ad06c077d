// The small layers of a training step that sit around the tensor-core convolutions, forward and backward, on
// channels-last float32 rows -- all HBM / launch bound, CUDA cores:
//   mz_cvt16              float32 -> 16-bit operand of the next convolution
//   mz_pool2_train_*      nn.AvgPool2d(2, 2) of the representation network (src/networks.py:43,82,92)
//   mz_linear_*           nn.Flatten + nn.Linear of the three heads (:147-149,207-209,221-223): raw logits, d input, d weight, d bias
//   mz_scale_train_*      MuZeroAgent._scale_state (:314-328) with the gradient torch's autograd gives it (through min / max as well)
//   mz_planes_conv_*      the action-plane input channels of the dynamics ConvBlock (:117-122, torch.cat at :295): their share of the 3x3
//                         convolution and of its weight gradient (3 of 259 input channels: not worth a tensor-core tile)
// Reductions go through per-CTA partial sums that are added in index order: results are deterministic.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ uint16_t tl_to16(float v, bool f16)
{
    if (f16) { const __half h = __float2half_rn(v); return *reinterpret_cast<const uint16_t *>(&h); }
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    return *reinterpret_cast<const uint16_t *>(&h);
}
__device__ __forceinline__ void tl_store4_16(uint16_t *p, bool f16, float a, float b, float c, float d)
{
    uint2 u;
    u.x = (uint32_t)tl_to16(a, f16) | ((uint32_t)tl_to16(b, f16) << 16);
    u.y = (uint32_t)tl_to16(c, f16) | ((uint32_t)tl_to16(d, f16) << 16);
    *reinterpret_cast<uint2 *>(p) = u;
}

// ---------------------------------------------------------------- float32 -> 16-bit
__global__ void __launch_bounds__(256) cvt16_kernel(size_t n4, const float4 *__restrict__ src, uint16_t *__restrict__ dst, int f16)
{
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n4) return;
    const float4 v = __ldg(src + i);
    tl_store4_16(dst + i * 4, f16 != 0, v.x, v.y, v.z, v.w);
}

// ---------------------------------------------------------------- 2x2 average pool, stride 2
// thread = 4 channels of one output pixel
__global__ void __launch_bounds__(256) pool2_fwd_kernel(size_t total4, int H, int W, int C, const float *__restrict__ x, float *__restrict__ y,
                                                        uint16_t *__restrict__ y16, int f16)
{
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= total4) return;
    const int c4 = C / 4, Wo = W / 2, Ho = H / 2;
    const int cq = (int)(i % c4);
    size_t r = i / c4;
    const int xo = (int)(r % Wo); r /= Wo;
    const int yo = (int)(r % Ho);
    const size_t n = r / Ho;
    const float4 *base = reinterpret_cast<const float4 *>(x) + ((n * H + 2 * yo) * W + 2 * xo) * c4 + cq;
    const float4 a = __ldg(base), b = __ldg(base + c4), c = __ldg(base + (size_t)W * c4), d = __ldg(base + (size_t)W * c4 + c4);
    // torch's avg_pool2d sums the window in row-major order and divides by the window size
    const float4 o = make_float4((((a.x + b.x) + c.x) + d.x) / 4.0f, (((a.y + b.y) + c.y) + d.y) / 4.0f, (((a.z + b.z) + c.z) + d.z) / 4.0f,
                                 (((a.w + b.w) + c.w) + d.w) / 4.0f);
    if (y) reinterpret_cast<float4 *>(y)[i] = o;
    if (y16) tl_store4_16(y16 + i * 4, f16 != 0, o.x, o.y, o.z, o.w);
}

// thread = 4 channels of one INPUT pixel: dx = dy[y/2][x/2] / 4
__global__ void __launch_bounds__(256) pool2_bwd_kernel(size_t total4, int H, int W, int C, const float *__restrict__ dy, float *__restrict__ dx)
{
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= total4) return;
    const int c4 = C / 4;
    const int cq = (int)(i % c4);
    size_t r = i / c4;
    const int xi = (int)(r % W); r /= W;
    const int yi = (int)(r % H);
    const size_t n = r / H;
    const float4 g = __ldg(reinterpret_cast<const float4 *>(dy) + ((n * (H / 2) + yi / 2) * (W / 2) + xi / 2) * c4 + cq);
    reinterpret_cast<float4 *>(dx)[i] = make_float4(g.x / 4.0f, g.y / 4.0f, g.z / 4.0f, g.w / 4.0f);
}

// ---------------------------------------------------------------- Flatten + Linear
// x float32 [n][F] in CHANNELS-LAST flatten order j = pixel * C + channel; w float32 [O][F] in nn.Flatten's (channel, pixel) order, so the
// weight of x[j] is w[o][perm(j)], perm(j) = (j % C) * HW + j / C.  O <= 16.
constexpr int LIN_MAX_O = 16;
constexpr int LIN_S = 4;             // samples per CTA of the forward kernel

__device__ __forceinline__ int lin_perm(int j, int C, int HW) { return (j % C) * HW + j / C; }

__global__ void __launch_bounds__(256) linear_fwd_kernel(int n, int HW, int C, int O, const float *__restrict__ x, const float *__restrict__ w,
                                                         const float *__restrict__ bias, float *__restrict__ out)
{
    __shared__ float s_red[8][LIN_S * LIN_MAX_O];
    const int F = HW * C, s0 = blockIdx.x * LIN_S;
    float acc[LIN_S][LIN_MAX_O];
#pragma unroll
    for (int s = 0; s < LIN_S; ++s)
#pragma unroll
        for (int o = 0; o < LIN_MAX_O; ++o) acc[s][o] = 0.0f;
    for (int j = threadIdx.x; j < F; j += 256) {
        const int pj = lin_perm(j, C, HW);
        float xv[LIN_S];
#pragma unroll
        for (int s = 0; s < LIN_S; ++s) xv[s] = s0 + s < n ? __ldg(x + (size_t)(s0 + s) * F + j) : 0.0f;
#pragma unroll
        for (int o = 0; o < LIN_MAX_O; ++o) {
            if (o < O) {
                const float wv = __ldg(w + (size_t)o * F + pj);
#pragma unroll
                for (int s = 0; s < LIN_S; ++s) acc[s][o] = fmaf(xv[s], wv, acc[s][o]);
            }
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int s = 0; s < LIN_S; ++s)
#pragma unroll
        for (int o = 0; o < LIN_MAX_O; ++o) {
            if (o < O) {
                float v = acc[s][o];
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
                if (lane == 0) s_red[warp][s * LIN_MAX_O + o] = v;
            }
        }
    __syncthreads();
    if (threadIdx.x < LIN_S * LIN_MAX_O) {
        const int s = threadIdx.x / LIN_MAX_O, o = threadIdx.x % LIN_MAX_O;
        if (o < O && s0 + s < n) {
            float v = 0.0f;
            for (int k = 0; k < 8; ++k) v += s_red[k][threadIdx.x];
            out[(size_t)(s0 + s) * O + o] = v + bias[o];
        }
    }
}

// dx[s][j] = sum_o g[s][o] * w[o][perm(j)]; grid (F / 256, ceil(n / 16)), thread = j
constexpr int LIN_BS = 16;
__global__ void __launch_bounds__(256) linear_bwd_data_kernel(int n, int HW, int C, int O, const float *__restrict__ g, const float *__restrict__ w,
                                                              float *__restrict__ dx)
{
    __shared__ float s_g[LIN_BS][LIN_MAX_O];
    const int F = HW * C, s0 = blockIdx.y * LIN_BS;
    for (int t = threadIdx.x; t < LIN_BS * LIN_MAX_O; t += 256) {
        const int s = t / LIN_MAX_O, o = t % LIN_MAX_O;
        s_g[s][o] = (o < O && s0 + s < n) ? g[(size_t)(s0 + s) * O + o] : 0.0f;
    }
    __syncthreads();
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= F) return;
    const int pj = lin_perm(j, C, HW);
    float wv[LIN_MAX_O];
#pragma unroll
    for (int o = 0; o < LIN_MAX_O; ++o) wv[o] = o < O ? __ldg(w + (size_t)o * F + pj) : 0.0f;
    for (int s = 0; s < LIN_BS && s0 + s < n; ++s) {
        float v = 0.0f;
#pragma unroll
        for (int o = 0; o < LIN_MAX_O; ++o) v = fmaf(s_g[s][o], wv[o], v);
        dx[(size_t)(s0 + s) * F + j] = v;
    }
}

// partial[chunk][o][j] = sum over the chunk's samples of g[s][o] * x[s][j]; grid (F / 256, chunks), thread = j
constexpr int LIN_CHUNK = 64;
__global__ void __launch_bounds__(256) linear_bwd_weight_kernel(int n, int F, int O, const float *__restrict__ g, const float *__restrict__ x,
                                                                float *__restrict__ partial)
{
    __shared__ float s_g[LIN_CHUNK][LIN_MAX_O];
    const int s0 = blockIdx.y * LIN_CHUNK;
    for (int t = threadIdx.x; t < LIN_CHUNK * LIN_MAX_O; t += 256) {
        const int s = t / LIN_MAX_O, o = t % LIN_MAX_O;
        s_g[s][o] = (o < O && s0 + s < n) ? g[(size_t)(s0 + s) * O + o] : 0.0f;
    }
    __syncthreads();
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= F) return;
    float acc[LIN_MAX_O];
#pragma unroll
    for (int o = 0; o < LIN_MAX_O; ++o) acc[o] = 0.0f;
    const int s1 = min(LIN_CHUNK, n - s0);
    for (int s = 0; s < s1; ++s) {
        const float xv = __ldg(x + (size_t)(s0 + s) * F + j);
#pragma unroll
        for (int o = 0; o < LIN_MAX_O; ++o) acc[o] = fmaf(s_g[s][o], xv, acc[o]);
    }
#pragma unroll
    for (int o = 0; o < LIN_MAX_O; ++o)
        if (o < O) partial[((size_t)blockIdx.y * O + o) * F + j] = acc[o];
}

// dw[o][perm(j)] (+)= sum over chunks (index order); the extra last CTA: db[o] (+)= sum_s g[s][o] (strided per-thread sums, then index order)
__global__ void __launch_bounds__(256) linear_bwd_reduce_kernel(int n, int HW, int C, int O, int chunks, const float *__restrict__ partial,
                                                                const float *__restrict__ g, float *__restrict__ dw, float *__restrict__ db, int accumulate)
{
    const int F = HW * C;
    if (blockIdx.x == gridDim.x - 1) {
        __shared__ float s_b[256][LIN_MAX_O + 1];
        float acc[LIN_MAX_O];
#pragma unroll
        for (int o = 0; o < LIN_MAX_O; ++o) acc[o] = 0.0f;
        for (int s = threadIdx.x; s < n; s += 256)
#pragma unroll
            for (int o = 0; o < LIN_MAX_O; ++o)
                if (o < O) acc[o] += g[(size_t)s * O + o];
#pragma unroll
        for (int o = 0; o < LIN_MAX_O; ++o) s_b[threadIdx.x][o] = acc[o];
        __syncthreads();
        if (threadIdx.x < O && db) {
            float v = 0.0f;
            for (int t = 0; t < 256; ++t) v += s_b[t][threadIdx.x];
            db[threadIdx.x] = accumulate ? db[threadIdx.x] + v : v;
        }
        return;
    }
    const int i = blockIdx.x * 256 + threadIdx.x;            // o * F + j
    if (i >= O * F) return;
    const int o = i / F, j = i - o * F;
    float v = 0.0f;
    for (int c = 0; c < chunks; ++c) v += __ldcs(partial + ((size_t)c * O + o) * F + j);
    float *dst = dw + (size_t)o * F + lin_perm(j, C, HW);
    *dst = accumulate ? *dst + v : v;
}

// ---------------------------------------------------------------- _scale_state
// (value, index) pairs ordered by value, ties to the smaller index (what torch's CPU min / max return; on the trunks' ReLU outputs the
// many tied zeros have a zero ReLU mask, so the choice never reaches a parameter)
struct MinMax {
    float mn, mx;
    int imn, imx;
};
__device__ __forceinline__ void mm_merge(MinMax &a, float mn, int imn, float mx, int imx)
{
    if (mn < a.mn || (mn == a.mn && imn < a.imn)) { a.mn = mn; a.imn = imn; }
    if (mx > a.mx || (mx == a.mx && imx < a.imx)) { a.mx = mx; a.imx = imx; }
}

// one CTA per sample: y = (x - min) / (max - min + 1e-8); stats[n] = {min, max, argmin, argmax}
__global__ void __launch_bounds__(256) scale_fwd_kernel(int E, const float *__restrict__ x, float *__restrict__ y, uint16_t *__restrict__ y16, int f16,
                                                        float4 *__restrict__ stats)
{
    __shared__ MinMax s_mm[8];
    const float *xs = x + (size_t)blockIdx.x * E;
    MinMax m{INFINITY, -INFINITY, 0x7fffffff, 0x7fffffff};
    for (int i = threadIdx.x * 4; i < E; i += 1024) {
        const float4 v = *reinterpret_cast<const float4 *>(xs + i);
        mm_merge(m, v.x, i, v.x, i); mm_merge(m, v.y, i + 1, v.y, i + 1); mm_merge(m, v.z, i + 2, v.z, i + 2); mm_merge(m, v.w, i + 3, v.w, i + 3);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const float mn = __shfl_down_sync(0xffffffffu, m.mn, d), mx = __shfl_down_sync(0xffffffffu, m.mx, d);
        const int imn = __shfl_down_sync(0xffffffffu, m.imn, d), imx = __shfl_down_sync(0xffffffffu, m.imx, d);
        mm_merge(m, mn, imn, mx, imx);
    }
    if ((threadIdx.x & 31) == 0) s_mm[threadIdx.x >> 5] = m;
    __syncthreads();
    m = s_mm[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) mm_merge(m, s_mm[k].mn, s_mm[k].imn, s_mm[k].mx, s_mm[k].imx);
    const float r = m.mx - m.mn + 1e-8f;
    if (threadIdx.x == 0 && stats) stats[blockIdx.x] = make_float4(m.mn, m.mx, __int_as_float(m.imn), __int_as_float(m.imx));
    for (int i = threadIdx.x * 4; i < E; i += 1024) {
        const float4 v = *reinterpret_cast<const float4 *>(xs + i);
        const float4 o = make_float4((v.x - m.mn) / r, (v.y - m.mn) / r, (v.z - m.mn) / r, (v.w - m.mn) / r);
        if (y) *reinterpret_cast<float4 *>(y + (size_t)blockIdx.x * E + i) = o;
        if (y16) tl_store4_16(y16 + (size_t)blockIdx.x * E + i, f16 != 0, o.x, o.y, o.z, o.w);
    }
}

// dx_i = g_i / r  (+ at argmin: -S1 / r + S2 / r^2;  + at argmax: -S2 / r^2),  S1 = sum g, S2 = sum g * (x - min), r = max - min + 1e-8
__global__ void __launch_bounds__(256) scale_bwd_kernel(int E, const float *__restrict__ x, const float *__restrict__ g, const float4 *__restrict__ stats,
                                                        float *__restrict__ dx)
{
    __shared__ double s_sum[8][2];
    const float *xs = x + (size_t)blockIdx.x * E, *gs = g + (size_t)blockIdx.x * E;
    float *ds = dx + (size_t)blockIdx.x * E;
    const float4 st = stats[blockIdx.x];
    const float mn = st.x, r = st.y - st.x + 1e-8f;
    const int imn = __float_as_int(st.z), imx = __float_as_int(st.w);
    double s1 = 0.0, s2 = 0.0;
    for (int i = threadIdx.x * 4; i < E; i += 1024) {
        const float4 v = *reinterpret_cast<const float4 *>(xs + i), gg = *reinterpret_cast<const float4 *>(gs + i);
        s1 += (double)gg.x + (double)gg.y + (double)gg.z + (double)gg.w;
        s2 += (double)gg.x * (double)(v.x - mn) + (double)gg.y * (double)(v.y - mn) + (double)gg.z * (double)(v.z - mn) + (double)gg.w * (double)(v.w - mn);
        *reinterpret_cast<float4 *>(ds + i) = make_float4(gg.x / r, gg.y / r, gg.z / r, gg.w / r);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { s1 += __shfl_down_sync(0xffffffffu, s1, d); s2 += __shfl_down_sync(0xffffffffu, s2, d); }
    if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5][0] = s1; s_sum[threadIdx.x >> 5][1] = s2; }
    __syncthreads();                                     // also orders every thread's dx stores before thread 0's two updates
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int k = 0; k < 8; ++k) { a += s_sum[k][0]; b += s_sum[k][1]; }
        const double rr = (double)r;
        ds[imn] += (float)(-a / rr + b / (rr * rr));
        ds[imx] += (float)(-b / (rr * rr));
    }
}

// ---------------------------------------------------------------- action-plane channels of the dynamics ConvBlock
// planes float32 [n][A][H][W] through element strides (the reference hands an expanded (n, A, 1, 1) view: strides 0 over H and W);
// w float32 [cout][cin_total][3][3], the planes are input channels c0 .. c0 + A - 1.  A <= 4, H * W <= 64.
constexpr int PL_MAX_A = 4;
constexpr int PL_MAX_HW = 64;
constexpr int PL_S = 2;              // samples per CTA (forward): 1280 CTAs at the 2560 samples of a training step's unroll
constexpr int PL_CHUNK = 4;          // samples per CTA (weight gradient): 640 CTAs; 16 samples = 160 CTAs took 175 us per call

struct PlaneArgs {
    int n, H, W, A, cout, cin_total, c0;
    long long sn, sa, sy, sx;        // element strides of planes
};

__device__ __forceinline__ void pl_stage(const PlaneArgs &a, const float *__restrict__ planes, int s, float (*dst)[PL_MAX_HW])
{
    const int HW = a.H * a.W;
    for (int t = threadIdx.x; t < a.A * HW; t += blockDim.x) {
        const int ch = t / HW, p = t - ch * HW;
        dst[ch][p] = planes[s * a.sn + ch * a.sa + (p / a.W) * a.sy + (p % a.W) * a.sx];
    }
}

// z[s][p][co] += sum_{a, tap in bounds} planes[s][a][p + tap] * w[co][c0 + a][tap]; thread = co
__global__ void planes_conv_fwd_kernel(const PlaneArgs a, const float *__restrict__ planes, const float *__restrict__ w, float *__restrict__ z)
{
    __shared__ float s_pl[PL_S][PL_MAX_A][PL_MAX_HW];
    const int co = threadIdx.x, HW = a.H * a.W, s0 = blockIdx.x * PL_S;
    for (int s = 0; s < PL_S && s0 + s < a.n; ++s) pl_stage(a, planes, s0 + s, s_pl[s]);
    float wr[PL_MAX_A][9];
#pragma unroll
    for (int ch = 0; ch < PL_MAX_A; ++ch)
#pragma unroll
        for (int t = 0; t < 9; ++t) wr[ch][t] = ch < a.A ? __ldg(w + ((size_t)co * a.cin_total + a.c0 + ch) * 9 + t) : 0.0f;
    __syncthreads();
    for (int s = 0; s < PL_S && s0 + s < a.n; ++s)
        for (int p = 0; p < HW; ++p) {
            const int y = p / a.W, x = p - y * a.W;
            float acc = 0.0f;
#pragma unroll
            for (int ch = 0; ch < PL_MAX_A; ++ch)
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
                    if (ch < a.A && yy >= 0 && yy < a.H && xx >= 0 && xx < a.W) acc = fmaf(s_pl[s][ch][yy * a.W + xx], wr[ch][t], acc);
                }
            float *dst = z + ((size_t)(s0 + s) * HW + p) * a.cout + co;
            *dst += acc;
        }
}

// partial[chunk][co][a * 9 + tap] = sum over the chunk's samples and pixels of dz[s][p][co] * planes[s][a][p + tap]; thread = co
__global__ void planes_dw_kernel(const PlaneArgs a, const float *__restrict__ planes, const float *__restrict__ dz, float *__restrict__ partial)
{
    __shared__ float s_pl[PL_MAX_A][PL_MAX_HW];
    const int co = threadIdx.x, HW = a.H * a.W, s0 = blockIdx.x * PL_CHUNK;
    float acc[PL_MAX_A][9];
#pragma unroll
    for (int ch = 0; ch < PL_MAX_A; ++ch)
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[ch][t] = 0.0f;
    for (int s = s0; s < s0 + PL_CHUNK && s < a.n; ++s) {
        __syncthreads();
        pl_stage(a, planes, s, s_pl);
        __syncthreads();
        for (int p = 0; p < HW; ++p) {
            const int y = p / a.W, x = p - y * a.W;
            const float d = __ldg(dz + ((size_t)s * HW + p) * a.cout + co);
#pragma unroll
            for (int ch = 0; ch < PL_MAX_A; ++ch)
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
                    if (ch < a.A && yy >= 0 && yy < a.H && xx >= 0 && xx < a.W) acc[ch][t] = fmaf(d, s_pl[ch][yy * a.W + xx], acc[ch][t]);
                }
        }
    }
#pragma unroll
    for (int ch = 0; ch < PL_MAX_A; ++ch)
#pragma unroll
        for (int t = 0; t < 9; ++t)
            if (ch < a.A) partial[((size_t)blockIdx.x * a.cout + co) * (PL_MAX_A * 9) + ch * 9 + t] = acc[ch][t];
}

// dw[co][c0 + a][tap] (+)= sum over chunks in index order; thread = (co, a, tap)
__global__ void __launch_bounds__(256) planes_dw_reduce_kernel(const PlaneArgs a, int chunks, const float *__restrict__ partial, float *__restrict__ dw,
                                                                       int accumulate)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= a.cout * a.A * 9) return;
    const int co = i / (a.A * 9), k = i - co * (a.A * 9);
    float v = 0.0f;
    const float *src = partial + (size_t)co * (PL_MAX_A * 9) + k;
    const size_t stride = (size_t)a.cout * (PL_MAX_A * 9);
    int c = 0;
    for (; c + 8 <= chunks; c += 8) {                      // eight loads in flight, added in index order
        float t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = __ldcs(src + (size_t)(c + u) * stride);
#pragma unroll
        for (int u = 0; u < 8; ++u) v += t[u];
    }
    for (; c < chunks; ++c) v += __ldcs(src + (size_t)c * stride);
    float *dst = dw + ((size_t)co * a.cin_total + a.c0) * 9 + k;
    *dst = accumulate ? *dst + v : v;
}

// ---------------------------------------------------------------- weight packs of a convolution for the tensor-core kernels
// w float32 [cout][cin_total][taps] (PyTorch's (cout, cin, k, k)), the first cin input channels are packed:
//   fwd   [tap][cin/64][cout][64]   element (t, cb, co, c) = w[co][cb*64 + c][t]                 (forward operand, 16-bit type of the activations)
//   dgrad [tap][cout/64][cin][64]   element (t, ob, ci, o) = w[ob*64 + o][ci][taps - 1 - t]      (transposed, tap-flipped filter: bf16)
// one thread per packed element, both packs in one launch (a torch permute / flip / contiguous / convert chain is 5-7 tiny kernels per
// convolution, ~550 per training iteration, in front of each layer's first use)
__global__ void __launch_bounds__(256) pack_conv_kernel(int cout, int cin_total, int cin, int taps, const float *__restrict__ w, uint16_t *__restrict__ fwd,
                                                        int fwd_f16, uint16_t *__restrict__ dgrad)
{
    const int total = taps * cin * cout;
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i < total) {
        if (!fwd) return;
        const int c = i & 63;
        int r = i >> 6;
        const int co = r % cout; r /= cout;
        const int cb = r % (cin >> 6), t = r / (cin >> 6);
        fwd[i] = tl_to16(__ldg(w + ((size_t)co * cin_total + cb * 64 + c) * taps + t), fwd_f16 != 0);
        return;
    }
    i -= total;
    if (i >= total || !dgrad) return;
    const int o = i & 63;
    int r = i >> 6;
    const int ci = r % cin; r /= cin;
    const int ob = r % (cout >> 6), t = r / (cout >> 6);
    dgrad[i] = tl_to16(__ldg(w + ((size_t)(ob * 64 + o) * cin_total + ci) * taps + (taps - 1 - t)), false);
}

bool pl_args_ok(const PlaneArgs &a)
{
    return a.n > 0 && a.H > 0 && a.W > 0 && a.H * a.W <= PL_MAX_HW && a.A >= 1 && a.A <= PL_MAX_A && a.cout >= 32 && a.cout <= 1024 && a.cout % 32 == 0 &&
           a.c0 >= 0 && a.c0 + a.A <= a.cin_total;
}

}  // namespace

extern "C" {

int mz_cvt16(long long n, const float *src, void *dst, int dtype, void *stream)
{
    MZB_CHECK_ARG(n > 0 && n % 4 == 0 && src && dst, "n must be a positive multiple of 4");
    MZB_CHECK_ARG(dtype == MZ_BF16 || dtype == MZ_F16, "dtype must be MZ_BF16 or MZ_F16");
    MZB_CHECK_ARG(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 7) == 0, "src must be 16-byte, dst 8-byte aligned");
    const size_t n4 = (size_t)n / 4;
    cvt16_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n4, (const float4 *)src, (uint16_t *)dst, dtype == MZ_F16);
    MZB_LAUNCH_CHECK();
    return 0;
}

int mz_pack_conv(int cout, int cin_total, int cin, int ksize, const float *w, void *fwd, int fwd_dtype, void *dgrad, void *stream)
{
    MZB_CHECK_ARG(cout > 0 && cout % 64 == 0 && cin > 0 && cin % 64 == 0 && cin <= cin_total && (ksize == 1 || ksize == 3), "cout and cin must be multiples of 64");
    MZB_CHECK_ARG(w && (fwd || dgrad), "null pointer");
    MZB_CHECK_ARG(!fwd || fwd_dtype == MZ_BF16 || fwd_dtype == MZ_F16, "fwd_dtype must be MZ_BF16 or MZ_F16");
    const int total = ksize * ksize * cin * cout;
    pack_conv_kernel<<<(2 * total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(cout, cin_total, cin, ksize * ksize, w, (uint16_t *)fwd, fwd_dtype == MZ_F16,
                                                                              (uint16_t *)dgrad);
    MZB_LAUNCH_CHECK();
    return 0;
}

int mz_pool2_train_fwd(int n, int H, int W, int C, const float *x, float *y, void *y16, int dtype, void *stream)
{
    MZB_CHECK_ARG(n > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 4 == 0, "H and W must be even, C a multiple of 4");
    MZB_CHECK_ARG(x && (y || y16), "null pointer");
    MZB_CHECK_ARG(!y16 || dtype == MZ_BF16 || dtype == MZ_F16, "dtype must be MZ_BF16 or MZ_F16");
    const size_t total4 = (size_t)n * (H / 2) * (W / 2) * (C / 4);
    pool2_fwd_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(total4, H, W, C, x, y, (uint16_t *)y16, dtype == MZ_F16);
    MZB_LAUNCH_CHECK();
    return 0;
}

int mz_pool2_train_bwd(int n, int H, int W, int C, const float *dy, float *dx, void *stream)
{
    MZB_CHECK_ARG(n > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 4 == 0 && dy && dx, "H and W must be even, C a multiple of 4");
    const size_t total4 = (size_t)n * H * W * (C / 4);
    pool2_bwd_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(total4, H, W, C, dy, dx);
    MZB_LAUNCH_CHECK();
    return 0;
}

int mz_linear_fwd(int n, int HW, int C, int O, const float *x, const float *w, const float *bias, float *out, void *stream)
{
    MZB_CHECK_ARG(n > 0 && HW > 0 && C > 0 && O >= 1 && O <= LIN_MAX_O && x && w && bias && out, "bad argument (O <= 16)");
    linear_fwd_kernel<<<(n + LIN_S - 1) / LIN_S, 256, 0, (cudaStream_t)stream>>>(n, HW, C, O, x, w, bias, out);
    MZB_LAUNCH_CHECK();
    return 0;
}

size_t mz_linear_scratch_bytes(int n, int HW, int C, int O)
{
    if (n <= 0 || HW <= 0 || C <= 0 || O < 1 || O > LIN_MAX_O) return 0;
    return (size_t)((n + LIN_CHUNK - 1) / LIN_CHUNK) * O * HW * C * sizeof(float);
}

int mz_linear_bwd(int n, int HW, int C, int O, const float *x, const float *w, const float *g, float *dx, float *dw, float *db, int accumulate,
                  void *scratch, void *stream)
{
    MZB_CHECK_ARG(n > 0 && HW > 0 && C > 0 && O >= 1 && O <= LIN_MAX_O && x && w && g, "bad argument (O <= 16)");
    MZB_CHECK_ARG(!dw || scratch, "the weight gradient needs mz_linear_scratch_bytes() of scratch");
    cudaStream_t st = (cudaStream_t)stream;
    const int F = HW * C;
    if (dx) {
        linear_bwd_data_kernel<<<dim3((F + 255) / 256, (n + LIN_BS - 1) / LIN_BS), 256, 0, st>>>(n, HW, C, O, g, w, dx);
        MZB_LAUNCH_CHECK();
    }
    if (dw) {
        const int chunks = (n + LIN_CHUNK - 1) / LIN_CHUNK;
        linear_bwd_weight_kernel<<<dim3((F + 255) / 256, chunks), 256, 0, st>>>(n, F, O, g, x, (float *)scratch);
        MZB_LAUNCH_CHECK();
        linear_bwd_reduce_kernel<<<(O * F + 255) / 256 + 1, 256, 0, st>>>(n, HW, C, O, chunks, (const float *)scratch, g, dw, db, accumulate);
        MZB_LAUNCH_CHECK();
    }
    return 0;
}

int mz_scale_train_fwd(int n, int E, const float *x, float *y, void *y16, int dtype, float *stats, void *stream)
{
    MZB_CHECK_ARG(n > 0 && E > 0 && E % 4 == 0 && x && (y || y16), "E must be a positive multiple of 4");
    MZB_CHECK_ARG(!y16 || dtype == MZ_BF16 || dtype == MZ_F16, "dtype must be MZ_BF16 or MZ_F16");
    MZB_CHECK_ARG((((uintptr_t)x | (uintptr_t)y | (uintptr_t)stats) & 15) == 0, "buffers must be 16-byte aligned");
    scale_fwd_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(E, x, y, (uint16_t *)y16, dtype == MZ_F16, (float4 *)stats);
    MZB_LAUNCH_CHECK();
    return 0;
}

int mz_scale_train_bwd(int n, int E, const float *x, const float *g, const float *stats, float *dx, void *stream)
{
    MZB_CHECK_ARG(n > 0 && E > 0 && E % 4 == 0 && x && g && stats && dx, "E must be a positive multiple of 4");
    MZB_CHECK_ARG((((uintptr_t)x | (uintptr_t)g | (uintptr_t)stats | (uintptr_t)dx) & 15) == 0, "buffers must be 16-byte aligned");
    scale_bwd_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(E, x, g, (const float4 *)stats, dx);
    MZB_LAUNCH_CHECK();
    return 0;
}

int mz_planes_conv_fwd(int n, int H, int W, int A, int cout, int cin_total, int c0, const float *planes, long long sn, long long sa, long long sy,
                       long long sx, const float *w, float *z, void *stream)
{
    const PlaneArgs a{n, H, W, A, cout, cin_total, c0, sn, sa, sy, sx};
    MZB_CHECK_ARG(pl_args_ok(a) && planes && w && z, "bad argument (A <= 4, H * W <= 64, cout a multiple of 32 up to 1024)");
    planes_conv_fwd_kernel<<<(n + PL_S - 1) / PL_S, cout, 0, (cudaStream_t)stream>>>(a, planes, w, z);
    MZB_LAUNCH_CHECK();
    return 0;
}

size_t mz_planes_wgrad_scratch_bytes(int n, int cout) { return n <= 0 || cout <= 0 ? 0 : (size_t)((n + PL_CHUNK - 1) / PL_CHUNK) * cout * PL_MAX_A * 9 * sizeof(float); }

int mz_planes_conv_wgrad(int n, int H, int W, int A, int cout, int cin_total, int c0, const float *planes, long long sn, long long sa, long long sy,
                         long long sx, const float *dz, float *dw, int accumulate, void *scratch, void *stream)
{
    const PlaneArgs a{n, H, W, A, cout, cin_total, c0, sn, sa, sy, sx};
    MZB_CHECK_ARG(pl_args_ok(a) && planes && dz && dw && scratch, "bad argument (A <= 4, H * W <= 64, cout a multiple of 32 up to 1024)");
    const int chunks = (n + PL_CHUNK - 1) / PL_CHUNK;
    planes_dw_kernel<<<chunks, cout, 0, (cudaStream_t)stream>>>(a, planes, dz, (float *)scratch);
    MZB_LAUNCH_CHECK();
    planes_dw_reduce_kernel<<<(cout * A * 9 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a, chunks, (const float *)scratch, dw, accumulate);
    MZB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
