// A whole residual trunk (a run of 3x3 256->256 convolutions on the 4x5 latent: the dynamics network's
// ConvBlock + 14 ResidualBlocks, or the prediction network's 14 ResidualBlocks) in ONE persistent launch.
//
// Same tile kernel as conv_tc.cu (per-pixel tiles that skip the zero-padding taps, cta_group::2 pair MMAs,
// TMA-fed 5-stage ring, two TMEM accumulators, coalesced staging epilogue), but the CTA pairs walk the layers
// back to back instead of returning to the host between them.  A layer only needs the previous layer's output for
// the SAME 128-sample group (a 3x3 conv mixes pixels, never samples) at the tile's 3x3 pixel neighbourhood, so
// there is no grid-wide barrier: done[layer][group][pixel] counts the epilogue warps that have stored (and fenced)
// their part of that pixel tile, and the TMA producer of tile (pixel p, group g) of layer L+1 waits for
// done[L][g][q] == 8 for the in-bounds neighbours q of p before it issues the first load.  The CTA pairs therefore
// flow from one layer into the next without draining their pipelines.  This removes, per convolution, the launch gap, the prologue (barrier init, TMEM allocation,
// cluster sync), the exposed last epilogue and the tail of the wave -- about 10 of 55 us at 4096 samples, and
// most of the time at small batches (60 launches of ~20 us at 24 samples).
//
// Residual blocks run in place on two activation buffers (conv1: X -> Y, conv2: Y + X -> X).  The neighbour wait
// also covers the write-after-read hazards: the tiles of the previous layer that READ buffer[g][p] are exactly the
// neighbours q of p (p is in N(q) iff q is in N(p)), the same tiles whose outputs tile p needs.
// Cross-proxy ordering: epilogue stores are generic-proxy writes that a later TMA (async proxy) reads, so the writer
// does st.global -> __threadfence -> fence.proxy.async -> red.release and the reader ld.acquire -> fence.proxy.async.
#include "tc_common.cuh"

namespace {

constexpr int MAX_BUFS = 3;
constexpr int HW = 20, LAT_W = 5, LAT_H = 4, CH = 256;

struct alignas(64) StackLayer {          // device-resident descriptor of one convolution of the trunk
    CUtensorMap map_b;                   // tile-contiguous weights [9][4][256][64], box = 128 rows (half of N = 256)
    CUtensorMap map_b_half;              // same tensor, box = 64 rows (half of an N = 128 slice: small-batch mode)
    const float *scale, *shift;          // [256]
    const float *act_bias;               // [3][20][256] or NULL
    float *dst_f32;                      // optional fp32 copy of the output or NULL
    int src, dst, res;                   // activation buffer ids; res = -1: no residual
    int act;
};

struct StackParams {
    CUtensorMap map_act[MAX_BUFS];       // activation buffers as (channel, x, y, sample) TMA tensors
    __nv_bfloat16 *act[MAX_BUFS];
    const StackLayer *layers;
    int nlayers;
    int *done;                           // [nlayers][groups][20 pixels], zeroed before the launch
    const int *act_idx;
    int f16;                             // 16-bit element type: 0 = bf16, 1 = fp16
    int trace;                           // profiling (env MZB_STACK_TRACE=1): cluster 0's leader records per-layer timestamps
    int fine;                            // 1: wait for the 3x3 neighbour pixel tiles only; 0: for all 20 pixel tiles of the group
    int rot;                             // tile -> CTA-pair assignment is rotated by rot pairs per layer (evens out the 4/6/9-tap tile costs)
    int n, groups, pairs, ntiles;        // samples, 128-sample groups, group pairs, pair-tiles per layer (= 20 * pairs * N/NT)
    long long f32_off;                   // element offset of this launch's first sample in the layers' dst_f32 tensors
};

__device__ unsigned long long g_stack_trace[6 * 64];
__device__ __forceinline__ unsigned long long gtime_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define STRACE(slot, layer) do { if (p.trace == 1 && blockIdx.x == 0 && (layer) < 64) g_stack_trace[(slot) * 64 + (layer)] = gtime_ns(); } while (0)

__device__ __forceinline__ uint32_t tap_mask(int y, int x)
{
    uint32_t m = 0u;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        if (y + dy >= 0 && y + dy < LAT_H && x + dx >= 0 && x + dx < LAT_W) m |= 1u << tap;
    }
    return m;
}

// Shared-memory geometry of one kernel variant: A ring (AROWS x 128 B per k-step; the UMMA always reads 128 rows, the rows
// past AROWS alias the following slots / the tail pad and only feed output rows that are discarded), B ring (this CTA's
// NT/2 weight rows per k-step), the epilogue staging tiles, scale/shift, barriers.  Smaller slots = more k-steps in flight:
// a k-step's TMA round trip is ~1.4 us regardless of its size, so a small batch is bound by stages / latency.
template <int NT, int AROWS>
struct Geo {
    static constexpr int A_SLOT = AROWS * BLOCK_K * 2;
    static constexpr int B_SLOT = (NT / 2) * BLOCK_K * 2;
    static constexpr int A_PAD = BLOCK_M * BLOCK_K * 2 - A_SLOT;
    static constexpr int EPI_WARP = 32 * (NT / 2) * 2;
    static constexpr int FIXED = A_PAD + NUM_EPI_WARPS * EPI_WARP + 2 * 256 * (int)sizeof(float) + 384;
    static constexpr int RAW = (232448 - FIXED) / (A_SLOT + B_SLOT);
    static constexpr int STAGES = RAW > 16 ? 16 : RAW;
    static constexpr int A_OFF = 0;
    static constexpr int B_OFF = STAGES * A_SLOT + A_PAD;
    static constexpr int EPI_OFF = B_OFF + STAGES * B_SLOT;
    static constexpr int SS_OFF = EPI_OFF + NUM_EPI_WARPS * EPI_WARP;
    static constexpr int BAR_OFF = SS_OFF + 2 * 256 * (int)sizeof(float);
    static constexpr size_t SMEM = BAR_OFF + 384;
    static_assert(STAGES >= 4 && SMEM <= 232448, "shared-memory budget");
    static_assert(B_OFF % 1024 == 0 && A_SLOT % 1024 == 0 && B_SLOT % 1024 == 0, "SWIZZLE_128B tiles need 1024-byte alignment");
};

// NT = output channels per tile: 256 (one tile per pixel and group pair) or 128 (two tiles: small batches have too few
// pixel tiles to fill the chip, so the N dimension is split to halve the per-layer latency and double the busy SMs)
constexpr int STACK_THREADS = NUM_THREADS + 32;       // + warp 10: the dependency scout

template <int NT, int AROWS>
__global__ void __launch_bounds__(STACK_THREADS, 1) conv_stack_kernel(const __grid_constant__ StackParams p)
{
    constexpr int N = CH;
    constexpr int nsplit = CH / NT;
    extern __shared__ __align__(1024) uint8_t smem[];
    using G = Geo<NT, AROWS>;
    constexpr int STAGES = G::STAGES;                 // shadows the one-layer kernel's constant
    uint8_t *epi_stage = smem + G::EPI_OFF;
    float *s_scale = reinterpret_cast<float *>(smem + G::SS_OFF);   // [scale 256 | shift 256] of the current layer
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + G::BAR_OFF);
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + STAGES);
    const uint32_t bar_tfull = smem_u32(bars + 2 * STAGES), bar_tempty = smem_u32(bars + 2 * STAGES + 2);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 4);
    int *s_ready = reinterpret_cast<int *>(bars + 2 * STAGES + 5);          // tiles of this CTA's list whose inputs are known to be complete

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
    constexpr int kchunks = CH / BLOCK_K;
    const uint32_t smem_base = smem_u32(smem);
    if (smem_base & 1023u) __trap();
    // first tile of this CTA pair in `layer`; the pair then strides by nclusters.  All warp roles walk the same list.
    auto first_tile = [&](int layer) { return (cluster_id + layer * p.rot) % nclusters; };

    if (threadIdx.x == 0) *s_ready = 0;
    if (warp == 0 && lane == 0) {
        for (int b = 0; b < MAX_BUFS; ++b) asm volatile("prefetch.tensormap [%0];" ::"l"(&p.map_act[b]) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 2); mbar_init(bar_empty + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, 2 * NUM_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            constexpr uint32_t a_bytes = G::A_SLOT, b_bytes = G::B_SLOT;
            const uint32_t lead_full = map_to_cta(bar_full, 0);
            int stage = 0, nks = 0, seq = 0;
            uint32_t phase = 0;
            for (int layer = 0; layer < p.nlayers; ++layer) {
                const StackLayer *L = p.layers + layer;
                const int src = L->src;
                const CUtensorMap *map_b = NT == CH ? &L->map_b : &L->map_b_half;
                const int first = first_tile(layer);
                for (int tile = first; tile < p.ntiles; tile += nclusters) {
                    const int ns = tile % nsplit, t2 = tile / nsplit;
                    const int pix = t2 / p.pairs, g = 2 * (t2 - pix * p.pairs) + rank;
                    const int y0 = pix / LAT_W, x0 = pix - y0 * LAT_W;
                    const uint32_t taps = tap_mask(y0, x0);
                    if (tile == first) STRACE(0, layer);
                    ++seq;
                    if (layer > 0) {
                        // the scout warp has seen the previous layer's output of this sample group complete at the in-bounds
                        // neighbour pixels (and acquired it); it polls ahead of this loop, so this wait is a shared-memory read
                        uint32_t spins = 0;
                        for (;;) {
                            int v;
                            asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(s_ready)) : "memory");
                            if (v >= seq) break;
                            if (++spins > (1u << 28)) __trap();
                        }
                        asm volatile("fence.proxy.async;" ::: "memory");
                    }
                    if (tile == first) STRACE(1, layer);
                    for (int tap = 0; tap < 9; ++tap) {
                        if (!((taps >> tap) & 1u)) continue;
                        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
#pragma unroll
                        for (int kc = 0; kc < kchunks; ++kc) {
                            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                            const uint32_t sa = smem_base + G::A_OFF + stage * G::A_SLOT, sb = smem_base + G::B_OFF + stage * G::B_SLOT;
                            if (rank == 0) mbar_expect_tx(bar_full + 8 * stage, 2 * (a_bytes + b_bytes));
                            else mbar_arrive_cluster(lead_full + 8 * stage);
                            if (p.trace == 2 && layer == 0 && tile == first && blockIdx.x == 0 && nks < 64) g_stack_trace[nks++] = gtime_ns();
                            tma_load_4d(sa, &p.map_act[src], lead_full + 8 * stage, kc * BLOCK_K, x0 + dx, y0 + dy, g * BLOCK_M);
                            tma_load_2d(sb, map_b, lead_full + 8 * stage, 0, (tap * kchunks + kc) * N + ns * NT + rank * (NT / 2));
                            if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (pair leader only) =====================
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = instr_desc(NT, p.f16 != 0);
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (int layer = 0; layer < p.nlayers; ++layer) {
                const int first = first_tile(layer);
                for (int tile = first; tile < p.ntiles; tile += nclusters, ++it) {
                    const int buf = it & 1;
                    const int pix = (tile / nsplit) / p.pairs, y0 = pix / LAT_W, x0 = pix - y0 * LAT_W;
                    if (p.trace == 4 && blockIdx.x == 0 && it < 96) g_stack_trace[it * 4] = gtime_ns();
                    mbar_wait(bar_tempty + 8 * buf, ((it >> 1) & 1) ^ 1);
                    tc_fence_after();
                    if (p.trace == 4 && blockIdx.x == 0 && it < 96) g_stack_trace[it * 4 + 1] = gtime_ns();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(buf * NT);
                    const int ksteps = __popc(tap_mask(y0, x0)) * kchunks;
                    for (int ks = 0; ks < ksteps; ++ks) {
                        mbar_wait(bar_full + 8 * stage, phase);
                        tc_fence_after();
                        if (ks == 0 && tile == first) STRACE(2, layer);
                        if (p.trace == 4 && ks == 0 && blockIdx.x == 0 && it < 96) g_stack_trace[it * 4 + 2] = gtime_ns();
                        if (p.trace == 3 && layer == 0 && tile == first && blockIdx.x == 0 && ks < 64) g_stack_trace[ks] = gtime_ns();
                        const uint32_t sa = smem_base + G::A_OFF + stage * G::A_SLOT, sb = smem_base + G::B_OFF + stage * G::B_SLOT;
                        const uint64_t adesc = smem_desc(sa), bdesc = smem_desc(sb);
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                            umma_bf16_pair(d_tmem, adesc + (uint64_t)(k * UMMA_K * 2 / 16), bdesc + (uint64_t)(k * UMMA_K * 2 / 16), idesc,
                                           (ks | k) ? 1u : 0u);
                        umma_commit_pair(bar_empty + 8 * stage);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma_commit_pair(bar_tfull + 8 * buf);
                    if (tile == first) STRACE(3, layer);
                    if (p.trace == 4 && blockIdx.x == 0 && it < 96) g_stack_trace[it * 4 + 3] = (gtime_ns() - g_stack_trace[it * 4 + 2]) | ((unsigned long long)ksteps << 40);   // issue time | k-steps
                }
            }
        }
    } else if (warp < 2 + NUM_EPI_WARPS) {
        // ===================== epilogue (warps 2..9) =====================
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        const int r = quarter * 32 + lane;
        constexpr int ncols = NT / 2, nchunks = ncols / 32, units = ncols / 8, row_bytes = ncols * 2, rows_per_it = 32 / units;
        const int my_u = lane % units, my_rsub = lane / units;
        const int etid = threadIdx.x - 64;                                   // 0..255 among the epilogue threads
        uint8_t *stg = epi_stage + (warp - 2) * G::EPI_WARP;
        const uint32_t lead_tempty = map_to_cta(bar_tempty, 0);
        int it = 0;
        for (int layer = 0; layer < p.nlayers; ++layer) {
            const StackLayer *L = p.layers + layer;
            // per-layer BN scale/shift: the first barrier proves every epilogue warp has finished the previous layer
            // (nobody reads the old values any more), the second that the new ones are in place
            asm volatile("bar.sync 1, 256;" ::: "memory");
            float *sc = s_scale, *sf = sc + 256;
            sc[etid] = L->scale[etid];
            sf[etid] = L->shift[etid];
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const __nv_bfloat16 *res_base = L->res >= 0 ? p.act[L->res] : nullptr;
            __nv_bfloat16 *dst_base = p.act[L->dst];
            const float *act_bias = L->act_bias;
            float *dst_f32 = L->dst_f32 ? L->dst_f32 + p.f32_off : nullptr;
            const int act = L->act;
            const int first = first_tile(layer);
            for (int tile = first; tile < p.ntiles; tile += nclusters, ++it) {
                const int buf = it & 1;
                const int ns = tile % nsplit, t2 = tile / nsplit;
                const int pix = t2 / p.pairs, g = 2 * (t2 - pix * p.pairs) + rank;
                const int col0 = ns * NT + half * ncols;
                const int s = g * BLOCK_M + r;
                const bool valid = s < p.n;
                const long long m = valid ? (long long)s * HW + pix : -1;        // global output row
                const float *ab = (valid && act_bias) ? act_bias + ((size_t)p.act_idx[s] * HW + pix) * N : nullptr;
                if (res_base) {
#pragma unroll
                    for (int k = 0; k < units; ++k) {
                        const int rr = k * rows_per_it + my_rsub;
                        const long long mr = __shfl_sync(0xffffffffu, m, rr);
                        if (mr >= 0) {
                            const uint32_t sdst = smem_u32(stg + rr * row_bytes + 16 * (my_u ^ (rr & (units - 1))));
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst), "l"(res_base + mr * N + col0 + my_u * 8) : "memory");
                        }
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                }
                uint32_t acc[2][32];
                mbar_wait(bar_tfull + 8 * buf, (it >> 1) & 1);
                tc_fence_after();
                if (warp == 2 && lane == 0 && tile == first) STRACE(4, layer);
                if (res_base) {
                    asm volatile("cp.async.wait_all;" ::: "memory");
                    __syncwarp();
                }
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * NT + half * ncols);
                tmem_ld32_async(taddr, acc[0]);
#pragma unroll
                for (int c = 0; c < nchunks; ++c) {
                    tmem_wait(acc[c & 1]);
                    if (c + 1 < nchunks) tmem_ld32_async(taddr + (uint32_t)((c + 1) * 32), acc[(c + 1) & 1]);
                    if (valid) {
                        const int c0 = col0 + c * 32;
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[c & 1][j]);
                        if (ab) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float4 t4 = __ldg(reinterpret_cast<const float4 *>(ab + c0) + q);
                                v[q * 4] += t4.x; v[q * 4 + 1] += t4.y; v[q * 4 + 2] += t4.z; v[q * 4 + 3] += t4.w;
                            }
                        }
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 a4 = *reinterpret_cast<const float4 *>(sc + c0 + q * 4), b4 = *reinterpret_cast<const float4 *>(sf + c0 + q * 4);
                            v[q * 4] = v[q * 4] * a4.x + b4.x; v[q * 4 + 1] = v[q * 4 + 1] * a4.y + b4.y;
                            v[q * 4 + 2] = v[q * 4 + 2] * a4.z + b4.z; v[q * 4 + 3] = v[q * 4 + 3] * a4.w + b4.w;
                        }
                        uint8_t *srow = stg + lane * row_bytes;
                        if (res_base) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint4 u4 = *reinterpret_cast<const uint4 *>(srow + 16 * ((c * 4 + q) ^ (lane & (units - 1))));
                                const uint32_t *h = reinterpret_cast<const uint32_t *>(&u4);
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 f = unpack2(h[e], p.f16);
                                    v[q * 8 + e * 2] += f.x;
                                    v[q * 8 + e * 2 + 1] += f.y;
                                }
                            }
                        }
                        if (act == MZ_ACT_RELU) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = activate(v[j], act);
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint4 u4;
                            uint32_t *h = reinterpret_cast<uint32_t *>(&u4);
#pragma unroll
                            for (int e = 0; e < 4; ++e) h[e] = pack2(v[q * 8 + e * 2], v[q * 8 + e * 2 + 1], p.f16);
                            *reinterpret_cast<uint4 *>(srow + 16 * ((c * 4 + q) ^ (lane & (units - 1)))) = u4;
                        }
                        if (dst_f32) {
                            float4 *fp = reinterpret_cast<float4 *>(dst_f32 + m * N + c0);
#pragma unroll
                            for (int q = 0; q < 8; ++q) fp[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(lead_tempty + 8 * buf);
#pragma unroll
                for (int k = 0; k < units; ++k) {
                    const int rr = k * rows_per_it + my_rsub;
                    const long long mr = __shfl_sync(0xffffffffu, m, rr);
                    if (mr >= 0) {
                        const uint4 v4 = *reinterpret_cast<const uint4 *>(stg + rr * row_bytes + 16 * (my_u ^ (rr & (units - 1))));
                        *(reinterpret_cast<uint4 *>(dst_base + mr * N + col0) + my_u) = v4;
                    }
                }
                // publish: this warp's part of (layer, group g, pixel pix) is in global memory
                __threadfence();
                asm volatile("fence.proxy.async;" ::: "memory");
                __syncwarp();
                if (lane == 0 && g < p.groups) asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(p.done + ((size_t)layer * p.groups + g) * HW + pix) : "memory");
                if (warp == 2 && lane == 0 && tile == first) STRACE(5, layer);
            }
        }
    } else if (lane == 0) {
        // ===================== dependency scout (warp 10) =====================
        // Walks this CTA's tile list AHEAD of the TMA producer: for tile (pixel p, group g) of layer L+1 it polls
        // done[L][g][q] of the in-bounds neighbour pixels q (one L2 round trip for all of them), acquires, and publishes the
        // running count of cleared tiles in shared memory.  The producer's own wait is then a shared-memory read, so the
        // flag round trip + fence (~2 us) no longer sits between the last load of one tile and the first of the next.
        int seq = 0;
        for (int layer = 0; layer < p.nlayers; ++layer) {
            const int first = first_tile(layer);
            for (int tile = first; tile < p.ntiles; tile += nclusters) {
                const int t2 = tile / nsplit;
                const int pix = t2 / p.pairs, g = 2 * (t2 - pix * p.pairs) + rank;
                const int y0 = pix / LAT_W, x0 = pix - y0 * LAT_W;
                ++seq;
                if (layer == 0) continue;
                if (g < p.groups) {
                    const uint32_t taps = tap_mask(y0, x0);
                    const int *flags = p.done + ((size_t)(layer - 1) * p.groups + g) * HW;
                    uint32_t spins = 0;
                    for (;;) {
                        int ready = 1;
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            if (p.fine && ((taps >> tap) & 1u)) {
                                const int v = *reinterpret_cast<const volatile int *>(flags + (y0 + tap / 3 - 1) * LAT_W + (x0 + tap % 3 - 1));
                                ready &= v >= NUM_EPI_WARPS * nsplit;
                            }
                        }
                        if (!p.fine) {
#pragma unroll
                            for (int q = 0; q < HW; ++q) ready &= *reinterpret_cast<const volatile int *>(flags + q) >= NUM_EPI_WARPS * nsplit;
                        }
                        if (ready) break;
                        if (++spins > (1u << 26)) __trap();
                        __nanosleep(32);
                    }
                    asm volatile("fence.acq_rel.gpu;" ::: "memory");       // acquire: pairs with the epilogues' red.release
                }
                asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(s_ready)), "r"(seq) : "memory");
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}


}  // namespace

extern "C" {

size_t mz_stack_layer_bytes(void) { return sizeof(StackLayer); }

int mz_stack_trace(unsigned long long *host_out)   // profiling aid: copies the 6 x 64 trace words
{
    return cudaMemcpyFromSymbol(host_out, g_stack_trace, sizeof(unsigned long long) * 6 * 64) == cudaSuccess ? 0 : -2;
}

int mz_stack_build(const mz_op *ops, int n_ops, void *blob_host, size_t blob_bytes, const void *const *bufs, int n_bufs)
{
    MZB_CHECK_ARG(ops && n_ops > 0 && blob_host && bufs && n_bufs > 0 && n_bufs <= MAX_BUFS, "bad argument");
    MZB_CHECK_ARG(blob_bytes >= (size_t)n_ops * sizeof(StackLayer), "blob too small");
    MZB_CHECK_ARG((reinterpret_cast<uintptr_t>(blob_host) & 63) == 0, "blob must be 64-byte aligned");
    EncodeTiledFn enc = encode_fn();
    if (!enc) { mzb::set_error("mz_stack_build: cuTensorMapEncodeTiled not available from the driver"); return -2; }
    StackLayer *L = reinterpret_cast<StackLayer *>(blob_host);
    auto buf_id = [&](const void *ptr) { for (int i = 0; i < n_bufs; ++i) if (bufs[i] == ptr) return i; return -1; };
    for (int i = 0; i < n_ops; ++i) {
        const mz_op &o = ops[i];
        MZB_CHECK_ARG(o.op == MZ_OP_CONV && (o.dtype == MZ_BF16 || o.dtype == MZ_F16) && o.use_tc && o.w_layout == 1 && o.ksize == 3 && o.cin == CH && o.cout == CH &&
                          o.H == LAT_H && o.W == LAT_W, "op is not a stackable 3x3 256->256 convolution on the 4x5 latent");
        StackLayer &l = L[i];
        l.src = buf_id(o.src); l.dst = buf_id(o.dst); l.res = o.res ? buf_id(o.res) : -1;
        MZB_CHECK_ARG(l.src >= 0 && l.dst >= 0 && (!o.res || l.res >= 0), "op buffer is not one of the stack's activation buffers");
        MZB_CHECK_ARG(l.src != l.dst, "a convolution cannot run in place on its own input");
        l.scale = o.scale; l.shift = o.shift; l.act_bias = o.act_bias; l.dst_f32 = o.dst_f32; l.act = o.act;
        cuuint64_t dims[2] = {BLOCK_K, (cuuint64_t)9 * (CH / BLOCK_K) * CH};
        cuuint64_t strides[1] = {BLOCK_K * 2};
        cuuint32_t box[2] = {BLOCK_K, CH / 2};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&l.map_b, o.dtype == MZ_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(o.w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { mzb::set_error("mz_stack_build: cuTensorMapEncodeTiled(weights) failed: %d", (int)r); return -2; }
        cuuint32_t box_half[2] = {BLOCK_K, CH / 4};
        r = enc(&l.map_b_half, o.dtype == MZ_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(o.w), dims, strides,
                box_half, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { mzb::set_error("mz_stack_build: cuTensorMapEncodeTiled(weights, half box) failed: %d", (int)r); return -2; }
    }
    return 0;
}

int mz_stack_run(const void *blob_dev, int n_layers, int sample0, int nsamples, void *const *bufs, int n_bufs, const int32_t *act_idx, int32_t *done,
                 int dtype, void *stream)
{
    MZB_CHECK_ARG(blob_dev && n_layers > 0 && sample0 >= 0 && nsamples > 0 && bufs && n_bufs > 0 && n_bufs <= MAX_BUFS && done && (dtype == MZ_BF16 || dtype == MZ_F16), "bad argument");
    EncodeTiledFn enc = encode_fn();
    if (!enc) { mzb::set_error("mz_stack_run: cuTensorMapEncodeTiled not available from the driver"); return -2; }
    cudaStream_t st = (cudaStream_t)stream;
    StackParams p{};
    static int want_split0 = -1;
    if (want_split0 < 0) { const char *e = getenv("MZB_STACK_SPLIT"); want_split0 = e ? atoi(e) : 0; }
    const int arows = (want_split0 && nsamples <= 32) ? 32 : BLOCK_M;   // rows of the activation box (split mode: tiny batches load only what exists)
    for (int b = 0; b < MAX_BUFS; ++b) {
        void *ptr = (__nv_bfloat16 *)bufs[b < n_bufs ? b : 0] + (size_t)sample0 * HW * CH;    // this launch's slice of the samples
        p.act[b] = (__nv_bfloat16 *)ptr;
        cuuint64_t dims[4] = {CH, LAT_W, LAT_H, (cuuint64_t)nsamples};
        cuuint64_t strides[3] = {CH * 2, LAT_W * CH * 2, HW * CH * 2};
        cuuint32_t box[4] = {BLOCK_K, 1, 1, (cuuint32_t)arows};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&p.map_act[b], dtype == MZ_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { mzb::set_error("mz_stack_run: cuTensorMapEncodeTiled(activations) failed: %d", (int)r); return -2; }
    }
    p.f16 = dtype == MZ_F16;
    p.layers = reinterpret_cast<const StackLayer *>(blob_dev);
    p.nlayers = n_layers;
    p.done = done;
    p.act_idx = act_idx ? act_idx + sample0 : nullptr;
    p.f32_off = (long long)sample0 * HW * CH;
    { static int tr = -1; if (tr < 0) { const char *e = getenv("MZB_STACK_TRACE"); tr = e ? atoi(e) : 0; } p.trace = tr; }
    { static int fine = -1; if (fine < 0) { const char *e = getenv("MZB_STACK_FINE"); fine = e ? atoi(e) : 1; } p.fine = fine; }
    { static int rot = -1; if (rot < 0) { const char *e = getenv("MZB_STACK_ROT"); rot = e ? atoi(e) : 13; } p.rot = rot; }
    p.n = nsamples;
    p.groups = (nsamples + BLOCK_M - 1) / BLOCK_M;
    p.pairs = (p.groups + 1) / 2;
    // Few pixel tiles (<= half of the 74 CTA pairs): N can be split in two so that twice as many pairs work on half-size
    // tiles with a deeper pipeline.  Measured at 24 samples: no gain (51 ms per 50-simulation search either way) -- a
    // k-step costs the single MMA-issuing thread ~300 ns (barrier wait + fence + 4 tcgen05.mma + commit, traced with
    // MZB_STACK_TRACE=3) whatever N is, so small batches are issue-bound, not tensor-bound.  Off unless MZB_STACK_SPLIT=1.
    static int want_split = -1;
    if (want_split < 0) { const char *e = getenv("MZB_STACK_SPLIT"); want_split = e ? atoi(e) : 0; }
    const bool split = want_split && HW * p.pairs * 2 <= mzb::kNumSMs / 2;
    p.ntiles = HW * p.pairs * (split ? 2 : 1);
    MZB_CUDA(cudaMemsetAsync(done, 0, sizeof(int) * (size_t)n_layers * p.groups * HW, st));
    static bool attr_set[64] = {};
    if (mzb::first_use_on_device(attr_set)) {
        MZB_CUDA(cudaFuncSetAttribute(conv_stack_kernel<256, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo<256, 128>::SMEM));
        MZB_CUDA(cudaFuncSetAttribute(conv_stack_kernel<128, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo<128, 128>::SMEM));
        MZB_CUDA(cudaFuncSetAttribute(conv_stack_kernel<128, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo<128, 32>::SMEM));
    }
    const int clusters = p.ntiles < mzb::kNumSMs / 2 ? p.ntiles : mzb::kNumSMs / 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(STACK_THREADS);
    cfg.dynamicSmemBytes = !split ? Geo<256, 128>::SMEM : (arows == 32 ? Geo<128, 32>::SMEM : Geo<128, 128>::SMEM);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (!split) MZB_CUDA(cudaLaunchKernelEx(&cfg, conv_stack_kernel<256, 128>, p));
    else if (arows == 32) MZB_CUDA(cudaLaunchKernelEx(&cfg, conv_stack_kernel<128, 32>, p));
    else MZB_CUDA(cudaLaunchKernelEx(&cfg, conv_stack_kernel<128, 128>, p));
    MZB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
