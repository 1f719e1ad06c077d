// 3x3 / 1x1 convolution on the B200 5th-generation tensor cores (sm_100a): implicit GEMM with
// tcgen05.mma (bf16 x bf16 -> fp32 accumulators in TMEM), operands staged in shared memory by TMA.
//
//   M = output pixels.  Two tilings of the 128-lane UMMA tile, chosen per launch by predicted cost:
//         "pixel"   one tile = ONE pixel position (y,x) of 128 consecutive samples.  Every row of the tile
//                   shares the same set of in-bounds taps, so the zero-padding taps of the 3x3 conv are
//                   skipped outright (a 4x5 latent executes 130 of the 180 (pixel,tap) pairs) and all 128
//                   rows are used;
//         "spatial" one tile = S samples x hb rows x W columns (<= 128 rows), for small batches of large
//                   images (the representation network), taps fully outside the image rows skipped
//   N = cout (128 or 256: the whole output-channel range, one UMMA N)
//   K = taps x cin, walked as (tap, 64-channel chunk); one pipeline stage = A[128 x 64] + B[N x 64] bf16
//
// The 3x3 taps need no im2col and no padded copy: the activation tensor is described to TMA as a 4-D
// tensor (channel, x, y, sample) and the tap (dy,dx) is just the box start coordinate (dx, y0+dy);
// out-of-range x / y (the conv's zero padding) and samples past the end are zero-filled by the TMA
// unit.  Both operands are K-major with the 128-byte swizzle, so a stage is one swizzle atom wide and
// the four K=16 MMAs of a stage advance the descriptor start address by 32 bytes.
//
// CTA pairs (cluster of 2, tcgen05 cta_group::2): the two CTAs of a pair work on two 128-sample groups of the
// SAME pixel / image rows, so they need the same weight tile at every k-step.  One MMA instruction issued by
// the leader CTA computes the pair's 256 x N tile: rows 0-127 from the leader's A tile, rows 128-255 from the
// peer's, and each CTA holds only HALF of the weight tile (N/2 rows) in its shared memory.  Per CTA and k-step
// that is 16 KB (A) + 16 KB (B half) of TMA writes and the same of operand reads instead of 48 + 48 KB, which
// keeps the 128 B/clk shared-memory port from capping the tensor pipe at ~2/3 (measured: 61-69 % with the
// 1-CTA form), and the 192 KB of stages hold 6 k-steps in flight instead of 4.
//
// Warp roles (320 threads, one persistent CTA per SM): warp 0 = TMA producer, warp 1 = TMEM allocator +
// MMA issuer (one elected lane), warps 2-9 = epilogue (TMEM lane quarter = warp_idx % 4, two warps per
// quarter split the N columns; TMEM loads and residual loads are software-pipelined): tcgen05.ld ->
// (+ per-action bias) * scale + shift (+ residual) -> activation -> bf16 (and optional fp32) stores.
// Two TMEM accumulator buffers (2 x N columns) overlap the epilogue of tile i with the MMAs of tile i+1.
//
// Replaces the cuDNN convolutions + separate BN / ReLU / add kernels the reference launches for
// src/networks.py ConvBlock :7-17 and ResidualBlock :19-35 (K4-K6 of SURVEY.md section 2d).
#include "tc_common.cuh"

namespace {

struct ConvParams {
    int n, H, W, cin, cout, taps, pad, act;
    int mode;          // 0 spatial tiles, 1 pixel tiles
    int debug;         // profiling experiments only (env MZB_TC_DEBUG): 1 = epilogue without memory traffic, 2 = TMA only for the
                       // first k-step of a tile (MMAs run on stale shared memory), 4 = no MMAs
    int f16;           // 16-bit element type: 0 = bf16, 1 = fp16
    int w_tiled;       // weights stored tile-contiguous [tap][cin/64][cout][64] instead of [cout][taps*cin]
    int S, hb, tile_rows, ytiles, groups, ntiles;   // groups = sample-group PAIRS per pixel (pixel mode); ntiles = pair-tiles
    __nv_bfloat16 *dst;
    const __nv_bfloat16 *res;
    float *dst_f32;
    const float *scale, *shift, *act_bias;
    const int *act_idx;
    const uint8_t *res_lo;   // e4m3 correction planes of the residual / the output (tc_common.cuh: split2 / lo2), tile-private layout
    uint8_t *dst_lo;
    const float *res_f32;    // fp32 residual instead of res / res_lo
    double *bn_partial;      // training form: per 32-row group of a tile, the column sums and sums of squares of the float32 output
                             // [(tile * 2 + CTA rank) * 4 + lane quarter][2][cout] -- the partial sums of the BatchNorm that follows
};

// profiling trace (debug & 8): per-tile timestamps of cluster 0's leader CTA, read back with mz_conv_trace()
__device__ unsigned long long g_trace[8 * 64];
__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define TRACE(slot, it) do { if ((p.debug & 8) && blockIdx.x == 0 && (it) < 64) g_trace[(slot) * 64 + (it)] = gtime(); } while (0)
// finer: stages of the first tile's 32-column chunks of epilogue warp 2 (slots 32 + 4 * chunk + stage of row 0)
#define TRACE_CHUNK(c, j) do { if ((p.debug & 8) && blockIdx.x == 0 && warp == 2 && lane == 0 && it == 0) g_trace[32 + 4 * (c) + (j)] = gtime(); } while (0)

struct Tile {
    int s0, y0, x0;
    uint32_t taps;   // bit t set = tap t has in-bounds rows for this tile (others contribute exact zeros)
};
__device__ __forceinline__ Tile decode_tile(const ConvParams &p, int tile, int rank)
{
    Tile t;
    int ny, nx;      // tile extent in y / x
    if (p.mode == 1) {
        const int pix = tile / p.groups, g = tile - pix * p.groups;   // group-fastest: neighbours share the pixel's cost
        t.y0 = pix / p.W; t.x0 = pix - t.y0 * p.W; t.s0 = (2 * g + rank) * p.S;
        ny = 1; nx = 1;
    } else {
        const int sg = tile / p.ytiles, yt = tile - sg * p.ytiles;
        t.s0 = (2 * sg + rank) * p.S; t.y0 = yt * p.hb; t.x0 = 0;
        ny = p.hb; nx = p.W;
    }
    if (p.taps == 1) { t.taps = 1u; return t; }
    t.taps = 0u;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        if (t.y0 + dy + ny > 0 && t.y0 + dy < p.H && t.x0 + dx + nx > 0 && t.x0 + dx < p.W) t.taps |= 1u << tap;
    }
    return t;
}

// N = cout (one UMMA N); kRelu: the activation is ReLU (else the runtime switch); kTrain: the form every convolution of a training step has --
// float32 output only, no activation, no 16-bit residual / action bias / correction planes (an optional float32 addend stays): its epilogue
// is compiled without the other paths.  Measured on the general instantiation (15 000 SASS instructions, four unrolled copies of a chunk
// body that carries every variant): 0.7 us per 32-column chunk between the accumulator read and the staged tile, 1.3 us per chunk, 5 us to
// drain a tile -- instruction fetch, not arithmetic.
template <int N, bool kRelu, bool kTrain = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const ConvParams p)
{
    extern __shared__ __align__(1024) uint8_t smem[];          // SWIZZLE_128B operand tiles need 1024-byte alignment
    uint8_t *epi_stage = smem + STAGES * STAGE_BYTES;
    float *s_scale = reinterpret_cast<float *>(epi_stage + NUM_EPI_WARPS * EPI_STAGE_BYTES);
    float *s_shift = s_scale + 256;
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_shift + 256);
    // bars: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then the TMEM base address word
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + STAGES);
    const uint32_t bar_tfull = smem_u32(bars + 2 * STAGES), bar_tempty = smem_u32(bars + 2 * STAGES + 2);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();          // 0 / 1 inside the CTA pair
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
    const int kchunks = p.cin / BLOCK_K;
    const uint32_t smem_base = smem_u32(smem);
    if (smem_base & 1023u) __trap();                  // the driver honours __align__(1024) on the dynamic segment; fail loudly if not

    mzb::pdl_trigger();                               // (training-step chains, common.cuh) the successor may become resident
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 2); mbar_init(bar_empty + 8 * s, 1); }   // full: both producers arrive (used in the leader)
        for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, 2 * NUM_EPI_WARPS); }   // both CTAs' epilogue warps (leader's copy)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                               // the peer's barriers are initialised before any remote arrive / multicast
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    // everything above touched only shared / tensor memory: it may overlap the tail of the previous kernel in the stream.  From here on
    // global memory is read (per-channel constants, operands) and written
    mzb::pdl_wait();
    for (int i = threadIdx.x; i < N; i += NUM_THREADS) { s_scale[i] = p.scale ? p.scale[i] : 1.0f; s_shift[i] = p.shift[i]; }
    __syncthreads();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            const uint32_t a_bytes = (uint32_t)p.tile_rows * BLOCK_K * 2, b_bytes = (uint32_t)(N / 2) * BLOCK_K * 2;
            const uint32_t lead_full = map_to_cta(bar_full, 0);          // the pair leader's full barriers
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < p.ntiles; tile += nclusters) {
                const Tile t = decode_tile(p, tile, rank);
                bool first = true;
                for (int tap = 0; tap < p.taps; ++tap) {
                    if (!((t.taps >> tap) & 1u)) continue;
                    const int dy = p.taps == 1 ? 0 : tap / 3 - 1, dx = p.taps == 1 ? 0 : tap % 3 - 1;
                    for (int kc = 0; kc < kchunks; ++kc) {
                        mbar_wait(bar_empty + 8 * stage, phase ^ 1);     // the pair's MMAs have retired this slot (in both CTAs)
                        const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_STAGE_BYTES;
                        const bool skip_tma = (p.debug & 2) && !first;
                        first = false;
                        if (rank == 0) mbar_expect_tx(bar_full + 8 * stage, skip_tma ? 0u : 2 * (a_bytes + b_bytes));   // bytes of both CTAs
                        else mbar_arrive_cluster(lead_full + 8 * stage);
                        if (skip_tma) { if (++stage == STAGES) { stage = 0; phase ^= 1; } continue; }
                        tma_load_4d(sa, &map_a, lead_full + 8 * stage, kc * BLOCK_K, t.x0 + dx, t.y0 + dy, t.s0);
                        if (p.w_tiled) tma_load_2d(sb, &map_b, lead_full + 8 * stage, 0, (tap * kchunks + kc) * N + rank * (N / 2));
                        else tma_load_2d(sb, &map_b, lead_full + 8 * stage, tap * p.cin + kc * BLOCK_K, rank * (N / 2));   // my half of the weight tile
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (pair leader only) =====================
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = instr_desc(N, p.f16 != 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = cluster_id; tile < p.ntiles; tile += nclusters, ++it) {
                const int buf = it & 1;
                TRACE(0, it);
                mbar_wait(bar_tempty + 8 * buf, ((it >> 1) & 1) ^ 1);       // epilogue has drained this accumulator
                tc_fence_after();
                TRACE(1, it);
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * N);
                const int ksteps = __popc(decode_tile(p, tile, rank).taps) * kchunks;
                for (int ks = 0; ks < ksteps; ++ks) {
                    mbar_wait(bar_full + 8 * stage, phase);                 // TMA bytes have landed
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_STAGE_BYTES;
                    const uint64_t adesc = smem_desc(sa), bdesc = smem_desc(sb);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                        if (!(p.debug & 4)) umma_bf16_pair(d_tmem, adesc + (uint64_t)(k * UMMA_K * 2 / 16), bdesc + (uint64_t)(k * UMMA_K * 2 / 16), idesc,
                                  (ks | k) ? 1u : 0u);
                    umma_commit_pair(bar_empty + 8 * stage);                // frees slot `stage` in both CTAs when the MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_pair(bar_tfull + 8 * buf);                      // accumulator complete -> both CTAs' epilogues
                TRACE(2, it);
            }
        }
    } else {
        // ===================== epilogue (warps 2..9) =====================
        // Thread = one row of the tile (TMEM lane).  Global traffic goes through a per-warp shared-memory
        // staging tile so that every global instruction is coalesced (16 lanes cover one row's contiguous
        // N/2-column segment) instead of 32 lanes touching 32 rows 10 KB apart: the residual is prefetched into
        // the staging tile while the MMAs of this tile are still running, the result overwrites it in place
        // (16-byte units XOR-swizzled by row: conflict-free both for row-per-lane and 16-lanes-per-row access).
        const int quarter = warp & 3;                  // TMEM lanes [32*quarter, 32*quarter+32)
        const int half = (warp - 2) >> 2;              // which half of the N columns this warp drains
        const int r = quarter * 32 + lane;             // row of the tile
        const int rows_per_sample = p.hb * p.W;
        constexpr int ncols = N / 2, nchunks = ncols / 32;
        constexpr int units = ncols / 8;               // 16-byte units per staged row (16 or 8)
        constexpr int row_bytes = ncols * 2;
        constexpr int rows_per_it = 32 / units;        // staged rows covered by one cooperative instruction (2 or 4)
        const int col0 = half * ncols;
        const int my_u = lane % units, my_rsub = lane / units;
        uint8_t *stg = epi_stage + (warp - 2) * EPI_STAGE_BYTES;
        const uint32_t lead_tempty = map_to_cta(bar_tempty, 0);
        int it = 0;
        for (int tile = cluster_id; tile < p.ntiles; tile += nclusters, ++it) {
            const int buf = it & 1;
            const Tile t = decode_tile(p, tile, rank);
            int s, y, x;
            if (p.mode == 1) { s = t.s0 + r; y = t.y0; x = t.x0; }
            else {
                const int sl = r / rows_per_sample, rr = r - sl * rows_per_sample;
                s = t.s0 + sl; y = t.y0 + rr / p.W; x = rr % p.W;
            }
            const bool valid = r < p.tile_rows && s < p.n && !(p.debug & 1);
            const long long m = valid ? ((long long)s * p.H + y) * p.W + x : -1;   // global output row, -1 = nothing to write
            const float *ab = (!kTrain && valid && p.act_bias) ? p.act_bias + ((size_t)p.act_idx[s] * p.H * p.W + (y * p.W + x)) * N : nullptr;
            const bool has_res = !kTrain && p.res != nullptr;
            // fp32 output without a 16-bit residual (every convolution of a training step: BatchNorm reads the fp32 sums): each 32-column
            // chunk goes through the staging tile and leaves as whole 128-byte row segments.  (Row-per-lane float4 stores touch 32 rows
            // 20 KB apart per instruction: 44 us per 512-sample layer against 10 us of MMAs.)  The 16-bit copy is optional then.
            const bool f32_staged = kTrain || (p.dst_f32 != nullptr && !has_res);
            // correction planes: [CTA tile][epilogue warp][2 * nchunks][32 lanes] x 16 bytes (same as conv_stack.cu in pixel mode)
            const size_t lo_tile = p.mode == 1 ? (size_t)(t.s0 / BLOCK_M) * (p.H * p.W) + (size_t)(t.y0 * p.W + t.x0) : (size_t)2 * tile + rank;
            const size_t lo_off = (lo_tile * NUM_EPI_WARPS + (warp - 2)) * (2 * nchunks * 512) + (size_t)lane * 16;
            uint4 lo_in[2 * nchunks];
            if (!kTrain && p.res_lo && valid) {
#pragma unroll
                for (int i = 0; i < 2 * nchunks; ++i) lo_in[i] = __ldcg(reinterpret_cast<const uint4 *>(p.res_lo + lo_off + i * 512));
            }
            if (warp == 2 && lane == 0) TRACE(3, it);
            if (has_res) {                                               // coalesced residual prefetch into the staging tile:
                // cp.async (global -> shared, 16 bytes each, no register staging) so that all 16 requests of a lane are
                // in flight at once and their latency hides behind the wait for this tile's MMAs
#pragma unroll
                for (int k = 0; k < units; ++k) {                        // units iterations x rows_per_it rows = 32 rows
                    const int rr = k * rows_per_it + my_rsub;
                    const long long mr = __shfl_sync(0xffffffffu, m, rr);
                    if (mr >= 0) {
                        const uint32_t sdst = smem_u32(stg + rr * row_bytes + 16 * (my_u ^ (rr & (units - 1))));
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst), "l"(p.res + mr * N + col0 + my_u * 8) : "memory");
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            uint32_t acc[2][32];
            if (warp == 2 && lane == 0) TRACE(4, it);
            mbar_wait(bar_tfull + 8 * buf, (it >> 1) & 1);
            tc_fence_after();
            if (warp == 2 && lane == 0) TRACE(5, it);
            if (has_res) {
                asm volatile("cp.async.wait_all;" ::: "memory");
                __syncwarp();                                            // every lane's residual pieces are visible to the whole warp
            }
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * N + col0);
            tmem_ld32_async(taddr, acc[0]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (c < nchunks) {
                    TRACE_CHUNK(c, 0);
                    tmem_wait(acc[c & 1]);
                    TRACE_CHUNK(c, 1);
                    if (c + 1 < nchunks) tmem_ld32_async(taddr + (uint32_t)((c + 1) * 32), acc[(c + 1) & 1]);
                    if (valid) {
                        const int c0 = col0 + c * 32;
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[c & 1][j]);
                        if (ab) {                                        // per-action bias row (dynamics first layer), 128-bit loads
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float4 t4 = __ldg(reinterpret_cast<const float4 *>(ab + c0) + q);
                                v[q * 4] += t4.x; v[q * 4 + 1] += t4.y; v[q * 4 + 2] += t4.z; v[q * 4 + 3] += t4.w;
                            }
                        }
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 sc = *reinterpret_cast<const float4 *>(s_scale + c0 + q * 4), sf = *reinterpret_cast<const float4 *>(s_shift + c0 + q * 4);
                            v[q * 4] = v[q * 4] * sc.x + sf.x; v[q * 4 + 1] = v[q * 4 + 1] * sc.y + sf.y;
                            v[q * 4 + 2] = v[q * 4 + 2] * sc.z + sf.z; v[q * 4 + 3] = v[q * 4 + 3] * sc.w + sf.w;
                        }
                        uint8_t *srow = stg + lane * row_bytes;
                        if (has_res) {
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
                            if (p.res_lo) {
#pragma unroll
                                for (int i = 0; i < 2; ++i) {
                                    const uint4 l4 = lo_in[(2 * c + i) < 2 * nchunks ? 2 * c + i : 0];
                                    const uint32_t w[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
                                    for (int e = 0; e < 8; ++e) {
                                        const float2 f = lo2((uint16_t)(w[e >> 1] >> ((e & 1) * 16)), p.f16);
                                        v[i * 16 + e * 2] += f.x;
                                        v[i * 16 + e * 2 + 1] += f.y;
                                    }
                                }
                            }
                        }
                        if (p.res_f32) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float4 t4 = __ldcg(reinterpret_cast<const float4 *>(p.res_f32 + m * N + c0) + q);
                                v[q * 4] += t4.x; v[q * 4 + 1] += t4.y; v[q * 4 + 2] += t4.z; v[q * 4 + 3] += t4.w;
                            }
                        }
                        if (kTrain) {
                        } else if (kRelu) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
                        } else if (p.act != MZ_ACT_NONE) {               // one uniform branch per chunk: the per-element switch of activate() costs
#pragma unroll                                                          // ~8 dependent branches per element (6 us per 32-column chunk: every
                            for (int j = 0; j < 32; ++j) v[j] = activate(v[j], p.act);   // convolution of a training step has act = none)
                        }
                        uint32_t hi[16];
                        if (kTrain) {
                        } else if (p.dst_lo) {
                            uint32_t lw[8];
#pragma unroll
                            for (int e = 0; e < 16; e += 2) {
                                uint16_t l0, l1;
                                hi[e] = split2(v[e * 2], v[e * 2 + 1], p.f16, l0);
                                hi[e + 1] = split2(v[e * 2 + 2], v[e * 2 + 3], p.f16, l1);
                                lw[e >> 1] = (uint32_t)l0 | ((uint32_t)l1 << 16);
                            }
                            __stcg(reinterpret_cast<uint4 *>(p.dst_lo + lo_off + (2 * c) * 512), make_uint4(lw[0], lw[1], lw[2], lw[3]));
                            __stcg(reinterpret_cast<uint4 *>(p.dst_lo + lo_off + (2 * c + 1) * 512), make_uint4(lw[4], lw[5], lw[6], lw[7]));
                        } else {
#pragma unroll
                            for (int e = 0; e < 16; ++e) hi[e] = pack2(v[e * 2], v[e * 2 + 1], p.f16);
                        }
                        if (f32_staged) {
#pragma unroll
                            for (int q = 0; q < 8; ++q)
                                *reinterpret_cast<float4 *>(stg + lane * 128 + 16 * (q ^ (lane & 7))) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                            if (!kTrain && p.dst) {
#pragma unroll
                                for (int q = 0; q < 4; ++q)
                                    *reinterpret_cast<uint4 *>(stg + 4096 + lane * 64 + 16 * (q ^ (lane & 3))) = make_uint4(hi[q * 4], hi[q * 4 + 1], hi[q * 4 + 2], hi[q * 4 + 3]);
                            }
                        } else {
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                *reinterpret_cast<uint4 *>(srow + 16 * ((c * 4 + q) ^ (lane & (units - 1)))) = make_uint4(hi[q * 4], hi[q * 4 + 1], hi[q * 4 + 2], hi[q * 4 + 3]);
                            if (p.dst_f32) {
                                float4 *fp = reinterpret_cast<float4 *>(p.dst_f32 + m * N + c0);
#pragma unroll
                                for (int q = 0; q < 8; ++q) fp[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                            }
                        }
                    }
                    if (f32_staged) {
                        const int c0 = col0 + c * 32;
                        TRACE_CHUNK(c, 2);
                        __syncwarp();
                        if (kTrain && p.bn_partial) {
                            // lane = column c0 + lane: sum and sum of squares over the warp's 32 staged rows (rows outside the tensor skipped;
                            // a row's 32 floats are one conflict-free shared-memory request), two float32 runs of 16 rows added in fp64
                            const uint32_t vm = __ballot_sync(0xffffffffu, valid);
                            double ds = 0.0, dq = 0.0;
#pragma unroll
                            for (int hrow = 0; hrow < 2; ++hrow) {
                                float sa = 0.0f, sq = 0.0f;
#pragma unroll
                                for (int r2 = 0; r2 < 16; ++r2) {
                                    const int rr = hrow * 16 + r2;
                                    const float xv = *reinterpret_cast<const float *>(stg + rr * 128 + 16 * ((lane >> 2) ^ (rr & 7)) + (lane & 3) * 4);
                                    if ((vm >> rr) & 1u) { sa += xv; sq = fmaf(xv, xv, sq); }
                                }
                                ds += (double)sa; dq += (double)sq;
                            }
                            double *bp = p.bn_partial + ((((size_t)tile * 2 + rank) * 4 + quarter) * 2) * N + c0 + lane;
                            bp[0] = ds;
                            bp[N] = dq;
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k) {                    // 8 lanes cover one row's 128-byte fp32 segment, 4 rows per instruction
                            const int rr = k * 4 + (lane >> 3), u = lane & 7;
                            const long long mr = __shfl_sync(0xffffffffu, m, rr);
                            if (mr >= 0)
                                *(reinterpret_cast<float4 *>(p.dst_f32 + mr * N + c0) + u) = *reinterpret_cast<const float4 *>(stg + rr * 128 + 16 * (u ^ (rr & 7)));
                        }
                        if (!kTrain && p.dst) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {                // 4 lanes cover one row's 64-byte 16-bit segment, 8 rows per instruction
                                const int rr = k * 8 + (lane >> 2), u = lane & 3;
                                const long long mr = __shfl_sync(0xffffffffu, m, rr);
                                if (mr >= 0)
                                    *(reinterpret_cast<uint4 *>(p.dst + mr * N + c0) + u) = *reinterpret_cast<const uint4 *>(stg + 4096 + rr * 64 + 16 * (u ^ (rr & 3)));
                            }
                        }
                        __syncwarp();
                        TRACE_CHUNK(c, 3);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lead_tempty + 8 * buf);   // accumulator drained: the next tile's MMAs may reuse it
            if (warp == 2 && lane == 0) TRACE(6, it);
            // coalesced store of the staged result rows
#pragma unroll
            for (int k = 0; k < units; ++k) {
                const int rr = k * rows_per_it + my_rsub;
                const long long mr = f32_staged ? -1 : __shfl_sync(0xffffffffu, m, rr);
                if (mr >= 0) {
                    const uint4 v4 = *reinterpret_cast<const uint4 *>(stg + rr * row_bytes + 16 * (my_u ^ (rr & (units - 1))));
                    *(reinterpret_cast<uint4 *>(p.dst + mr * N + col0) + my_u) = v4;
                }
            }
            __syncwarp();                                                // staging tile is reused by the next residual prefetch
            if (warp == 2 && lane == 0) TRACE(7, it);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                               // the peer may still signal my barriers / multicast into my smem until it is done
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

}  // namespace

extern "C" int mz_conv_trace(unsigned long long *host_out)   // profiling aid: copies the 8 x 64 trace words
{
    return cudaMemcpyFromSymbol(host_out, g_trace, sizeof(unsigned long long) * 8 * 64) == cudaSuccess ? 0 : -2;
}

namespace {

// the two tilings of conv_tc_launch and which one a shape gets (predicted cost = waves x taps per tile)
struct Tiling { int mode, S, hb, tile_rows, ytiles, groups, ntiles; };
Tiling choose_tiling(int n, int H, int W, int ksize)
{
    // spatial tiling: S samples x hb rows x W columns with S*hb*W <= 128, hb | H, as many rows as possible
    int best = 0, s_hb = 1, s_S = 1;
    for (int hb = 1; hb <= H; ++hb) {
        if (H % hb || hb * W > BLOCK_M) continue;
        int S = BLOCK_M / (hb * W);
        if (S > 256) S = 256;
        const int rows = S * hb * W;
        if (rows > best || (rows == best && hb > s_hb)) { best = rows; s_hb = hb; s_S = S; }
    }
    const long long sp_pairs = best > 0 ? (((n + s_S - 1) / s_S + 1) / 2) : 0;                       // sample-group pairs
    const long long sp_tiles = sp_pairs * (best > 0 ? H / s_hb : 0);
    // pixel tiling: 128 samples x one pixel; predicted cost = waves x average in-bounds taps
    const long long px_groups = ((n + BLOCK_M - 1) / BLOCK_M + 1) / 2, px_tiles = px_groups * H * W;   // pairs of 128-sample groups
    double px_taps = 1.0;
    if (ksize == 3) px_taps = (double)(3 * H - 2) * (3 * W - 2) / (H * W);
    auto waves = [](long long tiles) { return (double)((tiles + mzb::kNumSMs / 2 - 1) / (mzb::kNumSMs / 2)); };
    const double cost_px = waves(px_tiles) * px_taps, cost_sp = best > 0 ? waves(sp_tiles) * (ksize * ksize) : 1e30;
    if (cost_px <= cost_sp) return Tiling{1, BLOCK_M, 1, BLOCK_M, 1, (int)px_groups, (int)px_tiles};
    return Tiling{0, s_S, s_hb, best, H / s_hb, 1, (int)sp_tiles};
}

}  // namespace

// a correction plane covers whole CTA tiles (128 rows x cout bytes each, two per pair-tile); the CUDA-core and latency-mode
// kernels use the first n*H*W*cout bytes of it as a plain [row][channel] array
extern "C" size_t mz_conv_lo_bytes(int nsamples, int H, int W, int cout, int ksize)
{
    if (nsamples <= 0 || H <= 0 || W <= 0 || cout <= 0 || W > 256) return 0;
    const Tiling t3 = choose_tiling(nsamples, H, W, ksize == 1 ? 1 : 3);
    return (size_t)t3.ntiles * 2 * BLOCK_M * cout;
}

extern "C" int mz_conv_stats_blocks(int nsamples, int H, int W, int ksize)
{
    if (nsamples <= 0 || H <= 0 || W <= 0 || W > 256 || (ksize != 1 && ksize != 3)) return 0;
    return choose_tiling(nsamples, H, W, ksize).ntiles * 8;
}

namespace mzb {

int conv_tc_launch(const mz_op &o, int n, cudaStream_t st)
{
    MZB_CHECK_ARG(o.src && o.w && o.shift && (o.dst || (o.dst_f32 && !o.res && !o.dst_lo)), "conv_tc: null pointer (dst may be NULL only for an fp32-only output without residual)");
    MZB_CHECK_ARG((o.ksize == 1 || o.ksize == 3) && o.cin % BLOCK_K == 0 && (o.cout == 128 || o.cout == 256), "conv_tc: unsupported shape");
    MZB_CHECK_ARG(!o.act_bias || o.act_idx, "conv_tc: act_bias needs act_idx");
    EncodeTiledFn enc = encode_fn();
    if (!enc) { set_error("conv_tc: cuTensorMapEncodeTiled not available from the driver"); return -2; }

    ConvParams p{};
    p.n = n; p.H = o.H; p.W = o.W; p.cin = o.cin; p.cout = o.cout; p.taps = o.ksize * o.ksize; p.pad = o.ksize / 2; p.act = o.act;
    MZB_CHECK_ARG(o.W <= 256, "conv_tc: image too wide");
    const Tiling tl = choose_tiling(n, o.H, o.W, o.ksize);
    p.mode = tl.mode; p.S = tl.S; p.hb = tl.hb; p.tile_rows = tl.tile_rows; p.ytiles = tl.ytiles; p.groups = tl.groups; p.ntiles = tl.ntiles;
    p.dst = (__nv_bfloat16 *)o.dst; p.res = (const __nv_bfloat16 *)o.res; p.dst_f32 = o.dst_f32;
    p.scale = o.scale; p.shift = o.shift; p.act_bias = o.act_bias; p.act_idx = o.act_idx;
    p.res_lo = reinterpret_cast<const uint8_t *>(o.res_lo); p.dst_lo = reinterpret_cast<uint8_t *>(o.dst_lo);
    MZB_CHECK_ARG(!o.res_lo || o.res, "conv_tc: res_lo without res");
    MZB_CHECK_ARG(!o.res_f32 || !o.res, "conv_tc: res_f32 replaces res / res_lo");
    p.res_f32 = o.res_f32;
    p.bn_partial = o.bn_partial;
    p.w_tiled = o.w_layout == 1;
    p.f16 = o.dtype == MZ_F16;
    const CUtensorMapDataType tm_type = p.f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    { static int dbg = -1; if (dbg < 0) { const char *e = getenv("MZB_TC_DEBUG"); dbg = e ? atoi(e) : 0; } p.debug = dbg; }

    CUtensorMap map_a, map_b;
    {
        cuuint64_t dims[4] = {(cuuint64_t)o.cin, (cuuint64_t)o.W, (cuuint64_t)o.H, (cuuint64_t)n};
        cuuint64_t strides[3] = {(cuuint64_t)o.cin * 2, (cuuint64_t)o.W * o.cin * 2, (cuuint64_t)o.H * o.W * o.cin * 2};
        cuuint32_t box[4] = {BLOCK_K, (cuuint32_t)(p.mode == 1 ? 1 : o.W), (cuuint32_t)p.hb, (cuuint32_t)p.S};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&map_a, tm_type, 4, const_cast<void *>(o.src), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled(A) failed: %d", (int)r); return -2; }
    }
    {
        const cuuint64_t K = (cuuint64_t)p.taps * o.cin;
        cuuint64_t dims[2] = {K, (cuuint64_t)o.cout};
        cuuint64_t strides[1] = {K * 2};
        if (p.w_tiled) {                                   // [tap][cin/64][cout][64]: every (tap, chunk) tile is 128-byte rows back to back
            dims[0] = BLOCK_K; dims[1] = (cuuint64_t)p.taps * (o.cin / BLOCK_K) * o.cout;
            strides[0] = BLOCK_K * 2;
        }
        cuuint32_t box[2] = {BLOCK_K, (cuuint32_t)(o.cout / 2)};     // each CTA of a pair loads half of the rows
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&map_b, tm_type, 2, const_cast<void *>(o.w), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled(B) failed: %d", (int)r); return -2; }
    }
    static bool attr_set[64] = {};
    if (mzb::first_use_on_device(attr_set)) {
        MZB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        MZB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        MZB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        MZB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        MZB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<256, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        MZB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<128, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    }
    const int clusters = p.ntiles < kNumSMs / 2 ? p.ntiles : kNumSMs / 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchAttribute attr2[2];
    if (!o.dst && mzb::pdl_enabled()) {               // fp32-only output = a convolution of a training step: part of a kernel chain (common.cuh)
        attr2[0] = attr[0];
        attr2[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr2[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr2;
        cfg.numAttrs = 2;
    }
    const bool relu = o.act == MZ_ACT_RELU;
    // the training form (see the template comment): float32 output only, no activation, no 16-bit residual / action bias / correction planes
    const bool train_form = !o.dst && o.dst_f32 && !o.res && !o.res_lo && !o.dst_lo && !o.act_bias && o.act == MZ_ACT_NONE &&
                            !(getenv("MZB_TC_GENERAL_EPILOGUE") && atoi(getenv("MZB_TC_GENERAL_EPILOGUE")));
    MZB_CHECK_ARG(!o.bn_partial || train_form, "conv_tc: bn_partial needs the training form (float32 output only, no activation / residual)");
    if (train_form) {
        if (o.cout == 256) MZB_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<256, false, true>, map_a, map_b, p));
        else MZB_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<128, false, true>, map_a, map_b, p));
    } else if (o.cout == 256) {
        if (relu) MZB_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<256, true>, map_a, map_b, p));
        else MZB_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<256, false>, map_a, map_b, p));
    } else {
        if (relu) MZB_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<128, true>, map_a, map_b, p));
        else MZB_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<128, false>, map_a, map_b, p));
    }
    MZB_LAUNCH_CHECK();
    return 0;
}

}  // namespace mzb
