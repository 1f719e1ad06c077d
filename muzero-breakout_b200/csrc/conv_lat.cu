// Latency-mode residual trunk for SMALL leaf batches (config.yaml's default acting stage: 24 roots): the same run of
// 3x3 256->256 convolutions on the 4x5 latent as conv_stack.cu, one persistent launch, but tiled for latency instead of
// throughput.
//
// Why a second kernel.  The tcgen05 trunk works on pixel tiles of 128 samples x 256 output channels: at 24 samples a layer
// is 20 tiles (20 of the 148 SMs), each of which streams the whole 1.18 MB weight tensor through one SM's TMA port and
// issues 256x256x16 MMAs with 24 useful rows -- ~20 us per layer, 57 dependent layers per simulation step.  Here a layer
// is cut into (3-sample row tile) x (16-output-channel slice) work items, 16 * ceil(n/3) of them (128 CTAs at n = 24): every
// SM streams only its 72 KB slice of the weights (one TMA box of 36 [16 channels][64 inputs] units, SWIZZLE_128B = the layout
// ldmatrix wants, requested one item ahead by one thread, double-buffered, mbarrier-tracked), the 60 activation rows of its
// samples stay in shared memory for all 9 taps, and the math is warp-level mma.sync m16n8k16 (bf16/fp16 in, fp32
// accumulate) with the K dimension split over the 8 warps (warp w owns input channels [32 w, 32 w + 32) of every tap).
// tcgen05 needs M = 128 rows per instruction; at 60 rows x 16 channels per CTA the legacy warp MMA is the unit that fits.
// Measured on this B200 (profiles/micro/hmma_rate.cu): 1 HMMA.16816 per 2 cycles per SM, 1 LDSM.x4 per 4.06 cycles per SM;
// an item is 1152 HMMA + 720 LDSM.x4, i.e. shared-memory-read-bound at 2900 cycles = 1.5 us (the math phase takes 2.2 us).
// Tried and backed out (profiles/experiments/conv_lat_cluster16.cu.txt): one 16-CTA cluster per row tile with the layer
// hand-off through distributed shared memory -- only 7 such clusters fit on the chip at once (24 roots need 8), and the
// 16 x 16 slice broadcast + barrier.cluster cost 2.7 us per layer against 2.1 us for the first global-memory hand-off below.
// The network's other pieces ride in the same launch: 1x1 256->256 layers (the reward head's ConvBlock), a split last layer
// (policy 3x3 + value 1x1 256->128 ConvBlocks) and per-sample tail ops (heads, _scale_state) -- 3 launches per simulation step.
//
// Layer ordering without kernel boundaries: a 3x3 convolution never mixes samples, so item (layer L, row tile r) needs
// exactly the 16 channel slices of (L-1, r).  The hand-off is "flag in data" (the idea of NCCL's LL protocol): every pair of
// output channels is stored as one 8-byte word {2 x 16-bit, flag = (launch epoch, layer)} into a per-row-tile buffer (two buffers,
// by layer parity); a 64-bit store is indivisible, so a consumer that polls the words and finds this layer's flag has this
// layer's data -- no fence, no counter, no second round trip.  A buffer is overwritten two layers later, by items that have
// consumed the layer in between, which exists only when all consumers of the older data were done with it.  The launch epoch
// lives in device memory and is advanced by the last CTA to finish, so a CUDA-graph replay needs no reset of the buffers.
// (First version: a counter per (layer, row tile) -- stores, bar.sync, fence + red.release / ld.acquire spin, then cp.async of
// the rows: 2.05 of 5.4 us per layer.)  Every CTA walks its items in (layer, tile) order and the grid never exceeds the SM
// count (1 CTA per SM by shared memory), so a dependency always points at an item that is running or finished.
#include <string.h>

#include "tc_common.cuh"

namespace {

constexpr int HW = 20, LAT_W = 5, LAT_H = 4, CH = 256;
constexpr int RS = 3;                    // samples per row tile
constexpr int ROWS = RS * HW;            // 60 real rows, padded to 4 m16 tiles
constexpr int MT = 4;
constexpr int NS = 16;                   // output channels per item (two n8 tiles)
constexpr int NSLICES = CH / NS;         // 16
constexpr int KSTEPS = 9 * CH / 16;      // 144 k16 steps per item
constexpr int WARPS = 8, THREADS = WARPS * 32;
constexpr int STEPS_PER_WARP = KSTEPS / WARPS;   // 18
constexpr int KEEP_WAVES = 4;            // items per CTA and layer whose residual stream stays in registers (covers mz_lat_max_samples() = 3 waves; MZB_LAT_MAX_SAMPLES may raise it to 4)
constexpr int A_PITCH = CH * 2 + 16;     // 528 B: consecutive rows start 16 bytes apart modulo 128 -> conflict-free ldmatrix without an XOR
constexpr int A_BYTES = 64 * A_PITCH;    // 33 KB: [64 rows][528 B]
constexpr int W_UNITS = 9 * (CH / 64);   // 36 (tap, 64-channel chunk) units of [16 rows][128 B]
constexpr int W_BYTES = W_UNITS * NS * 128;      // 72 KB per item, double-buffered
constexpr int RED_BYTES = WARPS * 32 * 32 * 4;   // 32 KB: per-warp partial accumulators in fragment order
constexpr int MAX_LAYERS = 32;           // descriptors are staged in shared memory (a trunk has 28-29 layers)
static_assert(KSTEPS % WARPS == 0 && STEPS_PER_WARP == 18 && WARPS * 32 == CH, "K split: two k16 steps of each of the 9 taps per warp");

struct LatOperands {                     // 64 bytes; staged in shared memory
    const float *shift;                  // [256] (the BatchNorm scale is folded into the weights: mz_op.scale == NULL)
    uint8_t *lo;                         // e4m3 correction plane of the residual stream, [row][channel] bytes, or NULL (flags: bits 1-2)
    const float *act_bias;               // [3][20][256] or NULL
    float *dst_f32;                      // optional fp32 copy of the output or NULL
    const void *src;                     // activations [n][20][256], 16-bit
    void *dst;
    const void *res;                     // or NULL
    int act;
    short k1;                            // bit 0: 1x1 convolution (centre tap only, 4 weight units): the reward / value heads' ConvBlocks (networks.py:138-146, 212-218);
                                         // bit 1: the residual's correction is read from lo; bit 2: the output's correction is written to lo;
                                         // bit 3: res points at a float32 residual
    short cout;                          // 256, or 128 for the two halves of a split last layer (policy + value head convolutions)
};
struct alignas(64) LatLayer {            // device-resident descriptor of one convolution of the trunk
    CUtensorMap map_w;                   // tile-contiguous weights [9][4][256][64] as 3-D (64, 256, 36) with box (64, 16, 9), SWIZZLE_128B
    LatOperands o;
};
static_assert(sizeof(LatLayer) == 192 && sizeof(LatOperands) == 64, "layout");
// A tail op: per-sample work that follows the last convolution and would otherwise be two more launches per trunk -- the heads'
// Flatten + Linear + softmax / support expectation (networks.py:147-149,207-209,221-223; utils.py:74-81) and _scale_state
// (networks.py:314-328).  Stored in the blob slot that follows the convolution records (one LatLayer-sized slot per op).
struct LatTail {
    int kind;                            // MZ_OP_HEAD or MZ_OP_SCALE
    int mode, nout, feat;                // head: head_mode, outputs (3 / 11), features per sample; scale: feat = elements per sample (5120)
    const void *src;                     // head: 16-bit [n][feat]; scale: float32 [n][feat]
    const float *w, *bias;               // head: [nout][feat], [nout]
    float *out, *out_logits;             // head
    void *dst, *dst2;                    // scale: 16-bit outputs (dst2: the tree's latent store)
    const int *dst2_slot;
    long long dst2_stride;
};
static_assert(sizeof(LatTail) <= sizeof(LatLayer), "a tail op fits a blob slot");
constexpr int MAX_TAILS = 2;
constexpr int LL_PER_THREAD = ROWS * (CH / 4) / THREADS;       // 15 16-byte hand-off loads per thread and item
static_assert(LL_PER_THREAD * THREADS == ROWS * CH / 4, "hand-off loads");
constexpr int W_BOX = 36;                // units per TMA box: the whole 72 KB slice in one instruction (each cp.async.bulk.tensor issue holds its thread ~0.1-0.15 us; 36 boxes of 2 KB took 6.5 us to land)
constexpr int OFF_W = 0, OFF_A = 2 * W_BYTES, OFF_RED = OFF_A + A_BYTES, OFF_OPS = OFF_RED + RED_BYTES, OFF_BAR = OFF_OPS + MAX_LAYERS * 64;
constexpr int LAT_SMEM = OFF_BAR + 16;   // 210 KB
static_assert(LAT_SMEM <= 232448 && OFF_A % 1024 == 0 && OFF_OPS % 16 == 0 && OFF_BAR % 8 == 0, "shared-memory map (weight units are SWIZZLE_128B atoms: 1024-byte aligned)");

struct LatParams {
    const LatLayer *layers;
    int nlayers, n, rtiles, f16;
    uint2 *ll;                           // hand-off buffers: [2 (layer parity)][rtiles][60 rows][128 channel pairs] of {2 x 16-bit data, flag}
    int *sync;                           // [0] launch epoch, [1] finished-CTA count, [2 + (0|1) * rtiles + rt] item counters of the last two layers
    const int *act_idx;
    int ntails;                          // tail ops in the blob slots nlayers + split_last ...
    int split_last;                      // the last layer is two 128-channel convolutions of the same input (records nlayers-1 and nlayers): slices 0-7 / 8-15
    int trace;                           // profiling (env MZB_LAT_TRACE=1): CTA 0's thread 0 records phase timestamps per layer
    int w_early;                         // 1 (default): the next item's weights are requested before this item's math, 0: right after it
    int reg_stream;                      // 1 (default): launches with one item per CTA and layer keep the residual stream in registers (env MZB_LAT_REGSTREAM)
};

__device__ unsigned long long g_lat_trace[8 * 64];
#define LTRACE(slot) do { if (p.trace && blockIdx.x == 0 && tid == 0 && seq < 64) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_lat_trace[(slot) * 64 + seq] = t_; } } while (0)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4])
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
template <bool F16>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    if (F16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// W_BOX weight units [16 output channels][64 input channels] -> shared memory, bytes signalled on `bar`
__device__ __forceinline__ void tma_load_units(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// ---------------------------------------------------------------- tail ops (one sample, the whole CTA)
template <bool F16>
__device__ __forceinline__ void load8_16(const void *p, float (&f)[8])
{
    const uint4 u = __ldcg(reinterpret_cast<const uint4 *>(p));          // written by other CTAs of this launch: L2, never L1
    const uint32_t h[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) { const float2 t = unpack2(h[q], F16); f[2 * q] = t.x; f[2 * q + 1] = t.y; }
}

// Flatten + Linear, then head_mode 0: raw logits; 1: inverted_softmax_expectation (utils.py:74-81); 2: softmax probabilities
// (same arithmetic as nets.cu:head_kernel; the summation order over the features differs)
template <bool F16, int NOUT>
__device__ void tail_head(const LatTail &t, int s, float *scratch)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    float acc[NOUT];
#pragma unroll
    for (int o = 0; o < NOUT; ++o) acc[o] = 0.0f;
    const uint16_t *x = reinterpret_cast<const uint16_t *>(t.src) + (size_t)s * t.feat;
    for (int e = tid * 8; e < t.feat; e += THREADS * 8) {
        float xv[8];
        load8_16<F16>(x + e, xv);
#pragma unroll
        for (int o = 0; o < NOUT; ++o) {
            const float4 w0 = __ldg(reinterpret_cast<const float4 *>(t.w + (size_t)o * t.feat + e)), w1 = __ldg(reinterpret_cast<const float4 *>(t.w + (size_t)o * t.feat + e) + 1);
            acc[o] = fmaf(xv[0], w0.x, acc[o]); acc[o] = fmaf(xv[1], w0.y, acc[o]); acc[o] = fmaf(xv[2], w0.z, acc[o]); acc[o] = fmaf(xv[3], w0.w, acc[o]);
            acc[o] = fmaf(xv[4], w1.x, acc[o]); acc[o] = fmaf(xv[5], w1.y, acc[o]); acc[o] = fmaf(xv[6], w1.z, acc[o]); acc[o] = fmaf(xv[7], w1.w, acc[o]);
        }
    }
#pragma unroll
    for (int o = 0; o < NOUT; ++o) {
        float v = acc[o];
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sh);
        if (lane == 0) scratch[wid * NOUT + o] = v;
    }
    __syncthreads();
    if (tid == 0) {
        float logit[NOUT], mx = -INFINITY;
#pragma unroll
        for (int o = 0; o < NOUT; ++o) {
            float v = 0.0f;
#pragma unroll
            for (int q = 0; q < WARPS; ++q) v += scratch[q * NOUT + o];
            logit[o] = v + t.bias[o];
            mx = fmaxf(mx, logit[o]);
            if (t.out_logits) t.out_logits[(size_t)s * NOUT + o] = logit[o];
        }
        if (t.mode != 0) {
            float den = 0.0f, e[NOUT];
#pragma unroll
            for (int o = 0; o < NOUT; ++o) { e[o] = expf(logit[o] - mx); den += e[o]; }
            if (t.mode == 2) {                                                 // softmax probabilities (mcts.py:100,199)
#pragma unroll
                for (int o = 0; o < NOUT; ++o) t.out[(size_t)s * NOUT + o] = e[o] / den;
            } else {                                                           // utils.py:66-81
                const float half = 0.5f * (float)(NOUT - 1);
                float ex = 0.0f;
#pragma unroll
                for (int o = 0; o < NOUT; ++o) ex += (e[o] / den) * ((float)o - half);
                const float sg = ex > 0.0f ? 1.0f : (ex < 0.0f ? -1.0f : 0.0f);
                const float a = fabsf(ex) + 0.999f;                            // float32(1 - epsilon), utils.py:14,28
                t.out[s] = sg * (a * a - 1.0f);
            }
        }
    }
    __syncthreads();
}

// MuZeroAgent._scale_state (networks.py:314-328) of one sample's 5120 fp32 values -> 16-bit dst / dst2 (as nets.cu:scale_state_kernel)
template <bool F16>
__device__ void tail_scale(const LatTail &t, int s, float *scratch)
{
    const int tid = threadIdx.x;
    const float *x = t.src ? reinterpret_cast<const float *>(t.src) + (size_t)s * t.feat : nullptr;
    float4 v[5];
    float lo = INFINITY, hi = -INFINITY;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        v[q] = __ldcg(reinterpret_cast<const float4 *>(x) + q * THREADS + tid);
        lo = fminf(fminf(lo, fminf(v[q].x, v[q].y)), fminf(v[q].z, v[q].w));
        hi = fmaxf(fmaxf(hi, fmaxf(v[q].x, v[q].y)), fmaxf(v[q].z, v[q].w));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if ((tid & 31) == 0) { scratch[tid >> 5] = lo; scratch[8 + (tid >> 5)] = hi; }
    __syncthreads();
    lo = scratch[0]; hi = scratch[8];
#pragma unroll
    for (int q = 1; q < WARPS; ++q) { lo = fminf(lo, scratch[q]); hi = fmaxf(hi, scratch[8 + q]); }
    const float inv = __frcp_rn(__fadd_rn(__fsub_rn(hi, lo), 1e-8f));          // 1 / (s_max - s_min + 1e-8)  (:327)
    uint16_t *d1 = t.dst ? reinterpret_cast<uint16_t *>(t.dst) + (size_t)s * t.feat : nullptr;
    uint16_t *d2 = t.dst2 ? reinterpret_cast<uint16_t *>(t.dst2) + ((size_t)s * t.dst2_stride + (t.dst2_slot ? t.dst2_slot[s] : 0)) * t.feat : nullptr;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const uint2 o = make_uint2(pack2(__fmul_rn(__fsub_rn(v[q].x, lo), inv), __fmul_rn(__fsub_rn(v[q].y, lo), inv), F16),
                                   pack2(__fmul_rn(__fsub_rn(v[q].z, lo), inv), __fmul_rn(__fsub_rn(v[q].w, lo), inv), F16));
        const int e = (q * THREADS + tid) * 4;
        if (d1) *reinterpret_cast<uint2 *>(d1 + e) = o;
        if (d2) *reinterpret_cast<uint2 *>(d2 + e) = o;
    }
    __syncthreads();
}

template <bool F16>
__global__ void __launch_bounds__(THREADS, 1) conv_lat_kernel(const LatParams p)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sA = smem + OFF_A;
    float *sRed = reinterpret_cast<float *>(smem + OFF_RED);
    const LatOperands *sOps = reinterpret_cast<const LatOperands *>(smem + OFF_OPS);   // no pointer chase through L2 per layer
    const uint32_t sA_u = smem_u32(sA), sW_u = smem_u32(smem + OFF_W), bar_w = smem_u32(smem + OFF_BAR);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = p.rtiles * NSLICES;
    const int tpc = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // items of this CTA per layer (>= 1)
    const int total = p.nlayers * tpc;
    // one item per layer: the residual stream stays in this thread's registers between the layers that produce and consume it
    const bool reg_stream = tpc <= KEEP_WAVES && p.reg_stream;
    float2 kept_all[KEEP_WAVES][2];
#pragma unroll
    for (int w = 0; w < KEEP_WAVES; ++w) kept_all[w][0] = kept_all[w][1] = make_float2(0.f, 0.f);
    if (sW_u & 1023u) __trap();

    for (int i = tid; i < (p.nlayers + p.split_last) * 4; i += THREADS)
        reinterpret_cast<uint4 *>(smem + OFF_OPS)[i] = __ldg(reinterpret_cast<const uint4 *>(reinterpret_cast<const uint8_t *>(p.layers + (i >> 2)) + 128) + (i & 3));
    // the zero rows 60..63 are never overwritten
    for (int i = tid; i < 4 * A_PITCH / 16; i += THREADS) reinterpret_cast<uint4 *>(sA + ROWS * A_PITCH)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_w + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    // an item's [36 units][16 rows][128 B] weight slice -> buffer seq & 1: 4 TMA boxes issued by one thread; the buffer's
    // previous readers (ldmatrix of item seq - 2) are behind a __syncthreads
    auto weights_async = [&](int seq) {
        if (tid == 0) {
            const int layer = seq / tpc, tile = (int)blockIdx.x + (seq - layer * tpc) * (int)gridDim.x;
            int ns = tile % NSLICES, rec = layer;
            if (p.split_last && layer == p.nlayers - 1 && ns >= NSLICES / 2) { rec = layer + 1; ns -= NSLICES / 2; }
            const uint32_t bar = bar_w + 8 * (seq & 1), dst0 = sW_u + (uint32_t)(seq & 1) * W_BYTES;
            if (sOps[rec].k1 & 1) {                                // one box of 4 units (the map's box is (64, 16, 4))
                mbar_expect_tx(bar, 4 * NS * 128);
                tma_load_units(dst0, &p.layers[rec].map_w, bar, 0, ns * NS, 0);
            } else {
                mbar_expect_tx(bar, W_BYTES);
#pragma unroll
                for (int b = 0; b < W_UNITS / W_BOX; ++b) tma_load_units(dst0 + b * W_BOX * NS * 128, &p.layers[rec].map_w, bar, 0, ns * NS, b * W_BOX);
            }
        }
    };

    // ldmatrix source rows of this lane: A matrices (m16 x k16, row-major) are [rows 0-7 | rows 8-15] x [k 0-7 | k 8-15]
    const int mj = lane >> 3;
    int a_row[MT];                    // row inside the tile, or -1
    uint32_t a_taps[MT];              // in-bounds taps of that row's pixel
    for (int mt = 0; mt < MT; ++mt) {
        const int r = mt * 16 + (mj & 1) * 8 + (lane & 7);
        a_row[mt] = r;
        uint32_t m = 0u;
        if (r < ROWS) {
            const int pix = r % HW, y = pix / LAT_W, x = pix - y * LAT_W;
            for (int tap = 0; tap < 9; ++tap) {
                const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                if (y + dy >= 0 && y + dy < LAT_H && x + dx >= 0 && x + dx < LAT_W) m |= 1u << tap;
            }
        }
        a_taps[mt] = m;
    }
    const int a_khalf = mj >> 1;
    // B matrices (k16 x n8, "col"): [n 0-7 | n 8-15] x [k 0-7 | k 8-15] -> {b0, b1} of n-tile 0, {b0, b1} of n-tile 1
    const int b_n = (mj >> 1) * 8 + (lane & 7), b_khalf = mj & 1;
    // epilogue ownership: warp -> (m-tile, n-tile), thread -> rows g and g+8, two adjacent channels
    const int e_mt = warp >> 1, e_nt = warp & 1;
    const int e_r0 = e_mt * 16 + (lane >> 2), e_c = e_nt * 8 + (lane & 3) * 2;

    __syncthreads();
    const uint32_t epoch = *reinterpret_cast<const volatile uint32_t *>(p.sync);     // advanced by the previous launch's last CTA
    const uint32_t flag_base = epoch << 6;                        // + layer + 1 (<= 33): never 0, never a flag of another launch or layer
    if (total > 0) weights_async(0);

    for (int seq = 0; seq < total; ++seq) {
        const int layer = seq / tpc, wave = seq - layer * tpc, tile = (int)blockIdx.x + wave * (int)gridDim.x;
        const int rt = tile / NSLICES;
        int ns = tile - rt * NSLICES, rec = layer;                // ns: 16-channel slice of this item's convolution
        if (p.split_last && layer == p.nlayers - 1 && ns >= NSLICES / 2) { rec = layer + 1; ns -= NSLICES / 2; }
        const LatOperands *L = sOps + rec;
        const int cout = L->cout;
        const int s0 = rt * RS;                                   // first sample of the row tile
        const int nrows = min(RS, p.n - s0) * HW;                 // rows that exist

        LTRACE(0);
        if (layer == 0) {
            // the trunk's input comes from global memory (written by an earlier kernel): all 256 channels of the tile's rows
            const uint8_t *src = reinterpret_cast<const uint8_t *>(L->src) + (size_t)s0 * HW * CH * 2;
            for (int i = tid; i < nrows * 32; i += THREADS) {
                const int row = i >> 5, c = i & 31;
                cp_async16(sA_u + row * A_PITCH + c * 16, src + (size_t)row * (CH * 2) + c * 16);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        LTRACE(1);
        // epilogue operands that do not depend on the math: fetched now, used after the reduction
        const int co = ns * NS + e_c;
        const float2 sf = __ldg(reinterpret_cast<const float2 *>(L->shift + co));
        float2 ab[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, rs[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = e_r0 + h * 8;
            if (r < nrows) {
                const int s = s0 + r / HW, pix = r % HW;
                if (L->act_bias) ab[h] = __ldg(reinterpret_cast<const float2 *>(L->act_bias + ((size_t)p.act_idx[s] * HW + pix) * cout + co));
                if (reg_stream && (L->k1 & 32)) {                // this thread's own fp32 output of two layers ago (same item = same wave)
#pragma unroll
                    for (int w = 0; w < KEEP_WAVES; ++w) if (w == wave) rs[h] = kept_all[w][h];
                }
                else if (L->k1 & 8) rs[h] = __ldcg(reinterpret_cast<const float2 *>(reinterpret_cast<const float *>(L->res) + ((size_t)s0 * HW + r) * cout + co));
                else if (L->res) rs[h] = unpack2(__ldcg(reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint16_t *>(L->res) + ((size_t)s0 * HW + r) * cout + co)), F16);
                if ((L->k1 & 2) && !(reg_stream && (L->k1 & 32))) {   // 16-bit residual stream + e4m3 correction (tc_common.cuh: split2 / lo2)
                    const float2 l = lo2(__ldcg(reinterpret_cast<const uint16_t *>(L->lo + ((size_t)s0 * HW + r) * cout + co)), F16);
                    rs[h].x += l.x; rs[h].y += l.y;
                }
            }
        }
        if (layer == 0) {
            asm volatile("cp.async.wait_all;" ::: "memory");
        } else {
            // Layer hand-off, flag in data (the NCCL "LL" idea): the previous layer's 16 items wrote every pair of output channels as one
            // 8-byte word {2 x 16-bit, flag}, flag = (launch epoch, layer).  A 64-bit store is indivisible, so a word that carries this
            // layer's flag carries this layer's data: no fence on the producer side, no counter, no second round trip -- the consumer
            // polls the data itself (16-byte loads = 2 words) and moves what has arrived into the row buffer.
            const uint32_t want = flag_base + (uint32_t)layer;                     // = flag of layer - 1's outputs
            const uint4 *llp = reinterpret_cast<const uint4 *>(p.ll + ((size_t)((layer - 1) & 1) * p.rtiles + rt) * (ROWS * CH / 2));
            const int limit = nrows * 64;
            uint32_t pending = 0;
#pragma unroll
            for (int k = 0; k < LL_PER_THREAD; ++k) pending |= (tid + k * THREADS < limit) ? 1u << k : 0u;
            uint32_t spins = 0;
            while (pending) {
                uint4 v[LL_PER_THREAD];
#pragma unroll
                for (int k = 0; k < LL_PER_THREAD; ++k)
                    if ((pending >> k) & 1u) {
                        const uint4 *q = llp + tid + k * THREADS;
                        // two 64-bit elements: a naturally aligned 64-bit access is single-copy atomic in the PTX memory model, so data and
                        // flag of a word are always from the same store
                        unsigned long long w0, w1;
                        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(q) : "memory");
                        v[k] = make_uint4((uint32_t)w0, (uint32_t)(w0 >> 32), (uint32_t)w1, (uint32_t)(w1 >> 32));
                    }
#pragma unroll
                for (int k = 0; k < LL_PER_THREAD; ++k)
                    if (((pending >> k) & 1u) && v[k].y == want && v[k].w == want) {
                        const int i = tid + k * THREADS;
                        *reinterpret_cast<uint2 *>(sA + (i >> 6) * A_PITCH + (i & 63) * 8) = make_uint2(v[k].x, v[k].z);
                        pending &= ~(1u << k);
                    }
                if (++spins > (1u << 22)) __trap();               // a protocol bug traps instead of hanging the GPU
            }
        }
        mbar_wait(bar_w + 8 * (seq & 1), (seq >> 1) & 1);         // this item's weights (requested one item ago)
        __syncthreads();
        LTRACE(2);
        if (p.w_early && seq + 1 < total) weights_async(seq + 1);      // default: lands while this item's math runs
        LTRACE(6);

        float acc[MT][2][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[mt][nt][q] = 0.0f;

        const uint32_t wbuf = sW_u + (uint32_t)(seq & 1) * W_BYTES;
        uint32_t a_mask[MT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) a_mask[mt] = a_row[mt] < nrows ? a_taps[mt] : 0u;
        // K split: warp w owns input channels [32 w, 32 w + 32) = two k16 steps of every tap, so everything that depends on the
        // warp (channel offsets, swizzled weight addresses) is loop-invariant and everything that depends on the tap is a
        // compile-time constant of the unrolled loop.  Software-pipelined: the fragments of step i+1 are in flight while step
        // i's 8 MMAs issue (two warps per scheduler cannot hide the ldmatrix -> mma latency by themselves).
        uint32_t af[2][MT][4], bf[2][4];
        uint32_t rowa[MT];
        const uint32_t a_base = sA_u + (uint32_t)(warp * 4 + a_khalf) * 16;                         // + row * A_PITCH (+ 32 for the second k16)
        const uint32_t zero_base = sA_u + ROWS * A_PITCH;                                           // rows 60..63: 2112 bytes of zeros; [0, 160) are used
        const uint32_t b_base0 = wbuf + (warp >> 1) * (NS * 128) + b_n * 128 + (((((warp * 2) & 3) * 2 + b_khalf) ^ (b_n & 7)) << 4);
        const uint32_t b_base1 = wbuf + (warp >> 1) * (NS * 128) + b_n * 128 + (((((warp * 2 + 1) & 3) * 2 + b_khalf) ^ (b_n & 7)) << 4);
        // i = tap * 2 + j (compile-time after unrolling); wtap = the tap's index in the weight tensor
        auto load_frags = [&](int i, int wtap, uint32_t (&a)[MT][4], uint32_t (&b)[4]) {
            const int tap = i >> 1, j = i & 1;
            if (j == 0) {
                const int doff = (tap / 3 - 1) * LAT_W + (tap % 3 - 1);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    // zero padding, padding rows, samples past n: a 16-byte chunk of the zero rows IN THE BANK GROUP of the row it
                    // replaces -- a fixed zero chunk collides with whichever valid row of the same 8x16-byte matrix shares its
                    // bank group (border pixels are 14 of 20: ncu showed 24 % extra shared-memory wavefronts from bank conflicts)
                    const uint32_t addr = a_base + (uint32_t)((a_row[mt] + doff) * A_PITCH);
                    rowa[mt] = ((a_mask[mt] >> tap) & 1u) ? addr : zero_base + ((addr - zero_base) & 127u);
                }
            }
            ldmatrix_x4((j ? b_base1 : b_base0) + wtap * 4 * (NS * 128), b);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) ldmatrix_x4(rowa[mt] + j * 32, a[mt]);
        };
        auto mma_step = [&](const uint32_t (&a)[MT][4], const uint32_t (&b)[4]) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                mma16816<F16>(acc[mt][0], a[mt], b[0], b[1]);
                mma16816<F16>(acc[mt][1], a[mt], b[2], b[3]);
            }
        };
        if (L->k1 & 1) {                                            // centre tap only: this warp's two k16 steps
            load_frags(8, 0, af[0], bf[0]);
            load_frags(9, 0, af[1], bf[1]);
            mma_step(af[0], bf[0]);
            mma_step(af[1], bf[1]);
        } else {
            load_frags(0, 0, af[0], bf[0]);
#pragma unroll
            for (int i = 0; i < STEPS_PER_WARP; ++i) {
                if (i + 1 < STEPS_PER_WARP) load_frags(i + 1, (i + 1) >> 1, af[(i + 1) & 1], bf[(i + 1) & 1]);
                mma_step(af[i & 1], bf[i & 1]);
            }
        }
        LTRACE(7);
        // K reduction over the warps: partials in fragment order (conflict-free), summed in warp order (deterministic)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int q = 0; q < 4; ++q) sRed[(warp * 32 + (mt * 2 + nt) * 4 + q) * 32 + lane] = acc[mt][nt][q];
        LTRACE(3);
        __syncthreads();
        if (!p.w_early && seq + 1 < total) weights_async(seq + 1);    // MZB_LAT_W_EARLY=0: requested after the math instead (measured: no gain)
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float s = 0.0f;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) s += sRed[(w * 32 + (e_mt * 2 + e_nt) * 4 + q) * 32 + lane];
            v[q] = s;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = e_r0 + h * 8;
            if (r < nrows) {
                float x0 = (v[h * 2] + ab[h].x) + sf.x + rs[h].x;
                float x1 = (v[h * 2 + 1] + ab[h].y) + sf.y + rs[h].y;
                if (L->act == MZ_ACT_RELU) { x0 = fmaxf(x0, 0.0f); x1 = fmaxf(x1, 0.0f); }
                else if (L->act != MZ_ACT_NONE) { x0 = activate(x0, L->act); x1 = activate(x1, L->act); }
                const size_t o = ((size_t)s0 * HW + r) * cout + co;
                uint32_t packed;
                if (reg_stream && (L->k1 & 16)) {
#pragma unroll
                    for (int w = 0; w < KEEP_WAVES; ++w) if (w == wave) kept_all[w][h] = make_float2(x0, x1);
                }
                if ((L->k1 & 4) && !(reg_stream && (L->k1 & 16))) {
                    uint16_t l;
                    packed = split2(x0, x1, F16, l);
                    *reinterpret_cast<uint16_t *>(L->lo + o) = l;
                } else packed = pack2(x0, x1, F16);
                if (layer + 1 < p.nlayers) {                      // the next layer's items poll this word
                    uint2 *w = p.ll + ((size_t)(layer & 1) * p.rtiles + rt) * (ROWS * CH / 2) + (size_t)r * (CH / 2) + (co >> 1);
                    const unsigned long long word = (unsigned long long)packed | ((unsigned long long)(flag_base + (uint32_t)layer + 1u) << 32);
                    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(w), "l"(word) : "memory");          // one indivisible 64-bit store
                }
                *reinterpret_cast<uint32_t *>(reinterpret_cast<uint16_t *>(L->dst) + o) = packed;
                if (L->dst_f32) *reinterpret_cast<float2 *>(L->dst_f32 + o) = make_float2(x0, x1);
            }
        }
        // The global copies of the last two layers' outputs feed the tail ops (other CTAs read them through L2): those two layers
        // also publish with a counter -- stores -> bar.sync -> one thread's fence + red.release (cumulative), counted per launch epoch.
        LTRACE(4);
        __syncthreads();
        if (tid == 0 && p.ntails && layer >= p.nlayers - 2) {
            __threadfence();
            asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(p.sync + 2 + (layer - (p.nlayers - 2)) * p.rtiles + rt) : "memory");
        }
        LTRACE(5);
    }

    // Tail ops, after this CTA's last convolution item (they wait on convolution items only, which never wait on a tail): the CTAs
    // that computed channel slices 0..2 of the last layer of a row tile take one of its three samples each.
    if (p.ntails) {
        const LatLayer *tail_slots = p.layers + p.nlayers + p.split_last;      // one LatLayer-sized blob slot per tail op
        for (int k = 0; k < tpc; ++k) {
            const int tile = (int)blockIdx.x + k * (int)gridDim.x, rt = tile / NSLICES, j = tile - rt * NSLICES, s = rt * RS + j;
            if (j >= RS || s >= p.n) continue;
            if (tid == 0) {
                const int target = (int)(epoch + 1u) * NSLICES;      // the counters are never reset: 16 arrivals per launch
                for (int q = p.nlayers >= 2 ? 0 : 1; q < 2; ++q) {
                    const int *flag = p.sync + 2 + q * p.rtiles + rt;
                    uint32_t spins = 0;
                    for (;;) {
                        int v;
                        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
                        if (v - target >= 0) break;
                        if (++spins > (1u << 26)) __trap();
                    }
                }
            }
            __syncthreads();
            for (int i = 0; i < p.ntails; ++i) {
                const LatTail t = *reinterpret_cast<const LatTail *>(tail_slots + i);
                if (t.kind == MZ_OP_SCALE) tail_scale<F16>(t, s, sRed);
                else if (t.nout == 3) tail_head<F16, 3>(t, s, sRed);
                else tail_head<F16, 11>(t, s, sRed);
            }
        }
    }
    // the last CTA to finish advances the launch epoch (every CTA has read it by then)
    __syncthreads();
    if (tid == 0) {
        if (atomicAdd(p.sync + 1, 1) == (int)gridDim.x - 1) {
            p.sync[1] = 0;
            __threadfence();
            atomicAdd(p.sync, 1);
        }
    }
}

}  // namespace

extern "C" {

size_t mz_lat_layer_bytes(void) { return sizeof(LatLayer); }

int mz_lat_max_layers(void) { return MAX_LAYERS; }

// Measured crossover (profiles/README.md, round 2, ms per simulation step, latency mode vs the tcgen05 trunk with output-channel-split
// items): 27 samples (one wave of 144 items) 0.29 vs 0.79, 54 (two waves) 0.55 vs 0.79, 81 (three) 0.81 vs 0.80, 108 (four) 1.07 vs 0.83
// -> up to three waves
int mz_lat_max_samples(void) { return 3 * RS * (mzb::kNumSMs / NSLICES); }    // 81

int mz_lat_trace(unsigned long long *host_out)   // profiling aid: copies the 8 x 64 trace words
{
    return cudaMemcpyFromSymbol(host_out, g_lat_trace, sizeof(unsigned long long) * 8 * 64) == cudaSuccess ? 0 : -2;
}

int mz_lat_build(const mz_op *ops, int n_ops, void *blob_host, size_t blob_bytes)
{
    // (n_ops is reduced to the number of convolution records below)
    MZB_CHECK_ARG(ops && n_ops > 0 && n_ops <= MAX_LAYERS + MAX_TAILS && blob_host, "bad argument (at most 32 convolution records + 2 tail ops per launch)");
    MZB_CHECK_ARG(blob_bytes >= (size_t)n_ops * sizeof(LatLayer), "blob too small");
    MZB_CHECK_ARG((reinterpret_cast<uintptr_t>(blob_host) & 63) == 0, "blob must be 64-byte aligned");
    EncodeTiledFn enc = encode_fn();
    if (!enc) { mzb::set_error("mz_lat_build: cuTensorMapEncodeTiled not available from the driver"); return -2; }
    LatLayer *L = reinterpret_cast<LatLayer *>(blob_host);
    // trailing tail ops (heads, _scale_state): one blob slot each after the convolution records
    int ntails = 0;
    while (ntails < MAX_TAILS && n_ops - ntails > 1 && (ops[n_ops - 1 - ntails].op == MZ_OP_HEAD || ops[n_ops - 1 - ntails].op == MZ_OP_SCALE)) ++ntails;
    const int n_all = n_ops;
    n_ops -= ntails;
    for (int i = 0; i < ntails; ++i) {
        const mz_op &o = ops[n_ops + i];
        MZB_CHECK_ARG(o.dtype == ops[0].dtype && o.src, "tail op: element type / source");
        LatTail t{};
        t.kind = o.op;
        t.feat = o.H * o.W * o.cin;
        t.src = o.src;
        if (o.op == MZ_OP_HEAD) {
            MZB_CHECK_ARG(o.w && o.shift && (o.nout == 3 || o.nout == 11) && t.feat % 8 == 0, "tail head: 3 or 11 outputs, features a multiple of 8");
            MZB_CHECK_ARG(o.head_mode == 0 ? o.out_logits != nullptr : o.out != nullptr, "tail head: missing output");
            t.mode = o.head_mode; t.nout = o.nout; t.w = reinterpret_cast<const float *>(o.w); t.bias = o.shift; t.out = o.out; t.out_logits = o.out_logits;
        } else {
            MZB_CHECK_ARG(t.feat == HW * CH && (o.dst || o.dst2), "tail scale: 5120 elements per sample, a destination");
            t.dst = o.dst; t.dst2 = o.dst2; t.dst2_slot = o.dst2_slot; t.dst2_stride = o.dst2_stride;
        }
        memcpy(reinterpret_cast<uint8_t *>(L + n_ops + i), &t, sizeof(t));
    }
    (void)n_all;
    MZB_CHECK_ARG(n_ops <= MAX_LAYERS, "at most 32 convolution records per launch");
    const bool split = n_ops >= 3 && ops[n_ops - 1].cout == CH / 2 && ops[n_ops - 2].cout == CH / 2;
    for (int i = 0; i < n_ops; ++i) {
        const mz_op &o = ops[i];
        const bool half = split && i >= n_ops - 2;
        MZB_CHECK_ARG(o.op == MZ_OP_CONV && (o.dtype == MZ_BF16 || o.dtype == MZ_F16) && o.dtype == ops[0].dtype && o.w_layout == 1 && (o.ksize == 3 || o.ksize == 1) && o.cin == CH &&
                          o.cout == (half ? CH / 2 : CH) && o.H == LAT_H && o.W == LAT_W,
                      "op is not a 3x3 / 1x1 256->256 16-bit convolution on the 4x5 latent with tile-contiguous weights (or one of two final 256->128 ones)");
        MZB_CHECK_ARG(o.src && o.dst && o.w && o.shift && o.src != o.dst, "missing operand, or a convolution in place on its own input");
        MZB_CHECK_ARG(!o.scale, "the latency-mode trunk takes weights with the BatchNorm scale folded in (scale == NULL)");
        MZB_CHECK_ARG((!o.res_lo || o.res) && (!o.res_lo || !o.dst_lo || o.res_lo == o.dst_lo), "correction planes: res_lo needs res; a layer that reads and writes one uses the same plane (in-place stream)");
        MZB_CHECK_ARG(!o.act_bias || o.act_idx, "act_bias without act_idx");
        // every layer reads the previous layer's output (the layer counters order exactly that); the two halves of a split last layer
        // both read the output of the layer before them and write different buffers
        const int prev = half ? n_ops - 3 : i - 1;
        MZB_CHECK_ARG(i == 0 || o.src == ops[prev].dst, "the layers must form a chain: each one reads the previous one's output");
        MZB_CHECK_ARG(o.dst != ops[0].src, "no layer may overwrite the trunk's input (its rows are loaded by every CTA at its own pace)");
        MZB_CHECK_ARG(!half || (!o.res && !o.act_bias && ops[n_ops - 1].dst != ops[n_ops - 2].dst), "split last layer: no residual / action bias, two destinations");
        uint8_t *lo = reinterpret_cast<uint8_t *>(o.dst_lo ? o.dst_lo : const_cast<void *>(o.res_lo));
        MZB_CHECK_ARG(!o.res_f32 || !o.res, "res_f32 replaces res / res_lo");
        L[i].o = LatOperands{o.shift, lo, o.act_bias, o.dst_f32, o.src, o.dst, o.res_f32 ? (const void *)o.res_f32 : o.res, o.act,
                             (short)((o.ksize == 1 ? 1 : 0) | (o.res_lo ? 2 : 0) | (o.dst_lo ? 4 : 0) | (o.res_f32 ? 8 : 0)), (short)o.cout};
        cuuint64_t dims[3] = {64, (cuuint64_t)o.cout, (cuuint64_t)(o.ksize == 1 ? CH / 64 : W_UNITS)};
        cuuint64_t strides[2] = {128, (cuuint64_t)o.cout * 128};
        cuuint32_t box[3] = {64, NS, (cuuint32_t)(o.ksize == 1 ? CH / 64 : W_BOX)};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&L[i].map_w, o.dtype == MZ_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(o.w), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { mzb::set_error("mz_lat_build: cuTensorMapEncodeTiled(weights) failed: %d", (int)r); return -2; }
    }
    // Residual stream in registers: a layer whose residual is the output of the layer two before it (conv2 of a ResidualBlock adding the
    // previous block's output) finds that output -- in fp32, before any 16-bit rounding -- in the registers of the very thread that
    // computed it, because the item -> CTA -> thread mapping is the same in every 256 -> 256 layer.  Bit 4: keep the outputs; bit 5: the
    // residual is the kept value (no 16-bit residual load, no correction plane traffic).  Used by launches with one item per CTA and layer.
    for (int i = 2; i < n_ops; ++i) {
        const bool half = split && i >= n_ops - 2;
        if (!half && ops[i].res && !ops[i].res_f32 && ops[i].res == ops[i - 2].dst && ops[i].cout == CH && ops[i - 2].cout == CH) {
            L[i].o.k1 = (short)(L[i].o.k1 | 32);
            L[i - 2].o.k1 = (short)(L[i - 2].o.k1 | 16);
        }
    }
    return (split ? 1 : 0) | (ntails << 1);       // flags for mz_lat_run: bit 0 = the last two convolution records are the halves of a split layer, bits 1-2 = tail ops
}

static size_t lat_ll_bytes(int nsamples) { return (size_t)2 * ((nsamples + RS - 1) / RS) * ROWS * (CH / 2) * sizeof(uint2); }

size_t mz_lat_scratch_bytes(int nsamples)
{
    return lat_ll_bytes(nsamples) + sizeof(int) * (2 + 2 * (size_t)((nsamples + RS - 1) / RS));
}

int mz_lat_run(const void *blob_dev, int n_ops, int flags, int nsamples, const int32_t *act_idx, void *scratch, int dtype, void *stream)
{
    const int split_last = flags & 1, ntails = (flags >> 1) & 3;
    n_ops -= ntails;                         // convolution records
    MZB_CHECK_ARG(n_ops <= MAX_LAYERS && (!split_last || n_ops >= 3), "too many layers for one launch");
    const int n_layers = n_ops - (split_last ? 1 : 0);
    MZB_CHECK_ARG(blob_dev && n_layers > 0 && nsamples > 0 && scratch && (reinterpret_cast<uintptr_t>(scratch) & 15) == 0 && (dtype == MZ_BF16 || dtype == MZ_F16), "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    LatParams p{};
    p.layers = reinterpret_cast<const LatLayer *>(blob_dev);
    p.nlayers = n_layers;
    p.split_last = split_last ? 1 : 0;
    p.ntails = ntails;
    p.n = nsamples;
    p.rtiles = (nsamples + RS - 1) / RS;
    p.f16 = dtype == MZ_F16;
    p.ll = reinterpret_cast<uint2 *>(scratch);
    p.sync = reinterpret_cast<int *>(reinterpret_cast<uint8_t *>(scratch) + lat_ll_bytes(nsamples));
    p.act_idx = act_idx;
    { static int tr = -1; if (tr < 0) { const char *e = getenv("MZB_LAT_TRACE"); tr = e ? atoi(e) : 0; } p.trace = tr; }
    { static int we = -1; if (we < 0) { const char *e = getenv("MZB_LAT_W_EARLY"); we = e ? atoi(e) : 1; } p.w_early = we; }
    { static int rg = -1; if (rg < 0) { const char *e = getenv("MZB_LAT_REGSTREAM"); rg = e ? atoi(e) : 1; } p.reg_stream = rg; }
    static bool attr_set[64] = {};
    if (mzb::first_use_on_device(attr_set)) {
        MZB_CUDA(cudaFuncSetAttribute(conv_lat_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LAT_SMEM));
        MZB_CUDA(cudaFuncSetAttribute(conv_lat_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LAT_SMEM));
    }
    const int ntiles = p.rtiles * NSLICES;
    // all CTAs must be co-resident (items wait on other CTAs' outputs): the grid never exceeds what this device / context can
    // hold at once (one CTA per SM by shared memory; fewer SMs than a full B200 under MIG / MPS limits), and the launch is
    // cooperative so that it starts only when the whole grid fits beside whatever else is running
    static int resident[64] = {};
    int d = 0;
    MZB_CUDA(cudaGetDevice(&d));
    if (d < 0 || d >= 64 || resident[d] == 0) {
        int per_sm = 0, sms = 0;
        MZB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, conv_lat_kernel<true>, THREADS, LAT_SMEM));
        MZB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d));
        if (per_sm * sms < 1) { mzb::set_error("mz_lat_run: the latency-mode trunk does not fit on this device"); return -2; }
        if (d >= 0 && d < 64) resident[d] = per_sm * sms;
    }
    const int cap = resident[d] < mzb::kNumSMs ? resident[d] : mzb::kNumSMs;
    const int grid = ntiles < cap ? ntiles : cap;
    static int coop = -1;
    if (coop < 0) { const char *e = getenv("MZB_LAT_COOP"); coop = e ? atoi(e) : 1; }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = LAT_SMEM; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = coop ? 1 : 0;
    if (p.f16) MZB_CUDA(cudaLaunchKernelEx(&cfg, conv_lat_kernel<true>, p));
    else MZB_CUDA(cudaLaunchKernelEx(&cfg, conv_lat_kernel<false>, p));
    MZB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
