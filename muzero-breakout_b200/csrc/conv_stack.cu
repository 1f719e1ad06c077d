// A whole residual trunk (a run of 3x3 256->256 convolutions on the 4x5 latent: the dynamics network's
// ConvBlock + 14 ResidualBlocks, or the prediction network's 14 ResidualBlocks, each optionally followed by the
// head ConvBlocks that read the trunk output) in ONE persistent launch.
//
// Same tile scheme as conv_tc.cu (per-pixel tiles that skip the zero-padding taps, cta_group::2 pair MMAs,
// TMA-fed 5-stage ring, two TMEM accumulators), but the CTA pairs walk the layers back to back instead of
// returning to the host between them.  A layer only needs the previous layer's output for the SAME 128-sample
// group (a 3x3 conv mixes pixels, never samples) at the tile's 3x3 pixel neighbourhood, so there is no grid-wide
// barrier: done[layer][group][pixel] counts the epilogue warps that have stored their part of that pixel tile, and
// the TMA producer of tile (pixel p, group g) of layer L+1 waits for the in-bounds neighbours q of p before it
// issues the first load.  The CTA pairs therefore flow from one layer into the next without draining their pipelines.
//
// Residual blocks run in place on two activation buffers (conv1: X -> Y, conv2: Y + X -> X).  The neighbour wait
// also covers the write-after-read hazards: the tiles of the previous layer that READ buffer[g][p] are exactly the
// neighbours q of p (p is in N(q) iff q is in N(p)), the same tiles whose outputs tile p needs.
//
// Epilogue (8 warps, one TMEM lane quarter x half of the 256 columns each).  All activation traffic is TMA:
//   residual  : two [32 samples][64 channels] boxes of the stream buffer -> the warp's SWIZZLE_128B staging tile, issued
//               before the wait for the tile's MMAs, completion on a per-warp mbarrier;
//   result    : tcgen05.ld -> + shift (+ residual) -> activation -> 16-bit -> st.shared into the same staging tile
//               -> fence.proxy.async.shared::cta -> ONE lane issues two TMA tile stores (cp.async.bulk.tensor
//               shared -> global), waits for the bulk group and publishes the tile with a single red.release.
// No thread does a generic global store of activations, so there is no per-warp __threadfence (MEMBAR + L1
// invalidate) and no second pass over the staging tile by the LSU.  The BatchNorm scale is folded into the 16-bit
// weights by the packer (mz_op.scale == NULL), so the per-layer epilogue constants are the 256 shifts only.
// The residual STREAM of the blocks is carried as 16-bit value + e4m3 correction (tc_common.cuh: split2 / lo2);
// the correction plane lives in a tile-private layout -- [group][pixel][epilogue warp][8][32 lanes] 16-byte units --
// so its loads and stores are fully coalesced register traffic that never touches shared memory.
//
// The dependency counters are never reset: every launch adds 8 to each of them, the launch epoch lives next to them
// in device memory and is advanced by the last CTA to finish, so a CUDA-graph replay needs no memset node.
#include "tc_common.cuh"

namespace {

constexpr int MAX_BUFS = MZ_STACK_MAX_BUFS;
#define MZ_STACK_MAX_PAIRS 74
#define MZ_STACK_MAX_LPT_ITEMS 448
constexpr int HW = 20, LAT_W = 5, LAT_H = 4, CH = 256;

struct alignas(64) StackLayer {          // device-resident descriptor of one convolution of the trunk
    CUtensorMap map_b;                   // tile-contiguous weights [taps][4][256][64], box = 128 rows (half of N = 256)
    CUtensorMap map_b64;                 // same tensor, box = 64 rows (half of N = 128: the output-channel-split items of small batches)
    const float *shift;                  // [256]
    const float *act_bias;               // [3][20][256] or NULL
    float *dst_f32;                      // optional fp32 copy of the output or NULL
    const uint8_t *res_lo;               // correction plane of the residual or NULL
    uint8_t *dst_lo;                     // correction plane of the output or NULL
    const float *res_f32;                // fp32 residual instead of a 16-bit one (res = -1) or NULL
    int src, dst, res;                   // activation buffer ids; res = -1: no 16-bit residual
    int act;
    uint32_t taps;                       // taps of the kernel: 0x1FF (3x3) or 0x010 (1x1: the centre)
    int k1;                              // 1x1: the weight tensor holds the one tap at index 0
};
static_assert(sizeof(StackLayer) == 384, "layout");

struct StackParams {
    CUtensorMap map_act[MAX_BUFS];       // activation buffers as (channel, x, y, sample) TMA tensors, box = 64 ch x 128 samples (A operand)
    CUtensorMap map_epi[MAX_BUFS];       // same tensors, box = 64 ch x 32 samples (epilogue: residual in, result out)
    const StackLayer *layers;
    int nlayers;
    int *done;                           // [nlayers][groups][20 pixels], += 8 per launch
    int *sync;                           // [0] launch epoch, [1] finished-CTA count
    const int *act_idx;
    int f16;                             // 16-bit element type: 0 = bf16, 1 = fp16
    int trace;                           // profiling (env MZB_STACK_TRACE): cluster 0's leader records timestamps
    int debug;                           // timing experiments only (env MZB_STACK_DEBUG; results are garbage): 1 = weight tiles loaded for a tile's
                                         // first k-step only, 2 = activation tiles likewise, 4 = the epilogue skips the TMA store wait
    int nap;                             // ns the scout sleeps between two polls of the counters
    int fine;                            // 1: wait for the 3x3 neighbour pixel tiles only; 0: for all 20 pixel tiles of the group
    int rot;                             // tile -> CTA-pair assignment is rotated by rot pairs per layer (evens out the 4/6/9-tap tile costs)
    int n, groups;                       // samples, 128-sample groups
    int nsplit;                          // items per (pixel, group pair) tile: 1, or 2 output-channel halves (NT = 128)
    int nslices, slice_groups;           // the samples are walked as nslices slices of slice_groups groups (all layers of a slice before the
                                         // next slice, inside the one launch: the live activations of a slice stay in the L2)
    int sticky_groups;                   // sample groups g < sticky_groups move their activations with an L2 evict_last policy (they stay resident
                                         // in the L2 from layer to layer), the others with other_policy (0 evict_normal, 1 evict_first)
    int other_policy;
    float sticky_frac;                   // fraction of a sticky group's cache lines that get evict_last (createpolicy.fractional)
    long long elem_off;                  // element offset of this launch's first sample in [n][20][256] side tensors (dst_f32, correction planes)
    // static balanced schedule (single-slice launches): the items of a layer cost 4, 6 or 9 tap-units; lpt_items[lpt_off[c] .. lpt_off[c + 1])
    // are the items CTA pair c runs in every layer (longest-processing-time-first assignment, ascending item order within a pair)
    int use_lpt;
    uint16_t lpt_off[MZ_STACK_MAX_PAIRS + 2];
    uint16_t lpt_items[MZ_STACK_MAX_LPT_ITEMS];
};

__device__ unsigned long long g_stack_trace[6 * 64];
__device__ __forceinline__ unsigned long long gtime_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define STRACE(slot, layer) do { if (p.trace == 1 && blockIdx.x == 0 && (layer) < 64) g_stack_trace[(slot) * 64 + (layer)] = gtime_ns(); } while (0)

// generic proxy <-> async proxy ordering of GLOBAL memory (acquired counters -> TMA loads of what they guard; TMA stores -> the release
// that publishes them).  mode (experiments, StackParams::debug bits 8 / 16): 0 = .global form, 8 = the all-spaces form, 16 = none (timing only)
__device__ __forceinline__ void proxy_fence_global(int mode)
{
    if (mode & 16) return;
    if (mode & 8) asm volatile("fence.proxy.async;" ::: "memory");
    else asm volatile("fence.proxy.async.global;" ::: "memory");
}

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

// shared-memory map: A ring | B ring | 8 epilogue staging tiles (2 x [32][128 B] SWIZZLE_128B each) | shift[256] | barriers
constexpr int A_SLOT = BLOCK_M * BLOCK_K * 2, B_SLOT = (CH / 2) * BLOCK_K * 2;     // 16 KB each
constexpr int NSTAGE = 5;
constexpr int EPI_WARP = 2 * 32 * 128;                                               // 8 KB
constexpr int A_OFF = 0, B_OFF = NSTAGE * A_SLOT, EPI_OFF = B_OFF + NSTAGE * B_SLOT, SS_OFF = EPI_OFF + NUM_EPI_WARPS * EPI_WARP;
constexpr int BAR_OFF = SS_OFF + 256 * (int)sizeof(float);
constexpr int NBARS = 2 * NSTAGE + 4 + NUM_EPI_WARPS;                                // full, empty, tfull[2], tempty[2], residual[8]
constexpr size_t STACK_SMEM = BAR_OFF + 8 * NBARS + 16;
static_assert(STACK_SMEM <= 232448 && EPI_OFF % 1024 == 0 && EPI_WARP % 1024 == 0, "shared-memory budget / SWIZZLE_128B alignment");

constexpr int STACK_THREADS = NUM_THREADS + 32;       // + warp 10: the dependency scout

// virtual layer vl = slice * nlayers + layer: every warp role walks (vl, this pair's tiles of vl) in the same order
struct VLayer { int layer, g0, sgroups, spairs, ntiles; };
__device__ __forceinline__ VLayer vlayer(const StackParams &p, int vl)
{
    VLayer v;
    const int sl = vl / p.nlayers;
    v.layer = vl - sl * p.nlayers;
    v.g0 = sl * p.slice_groups;
    v.sgroups = min(p.slice_groups, p.groups - v.g0);
    v.spairs = (v.sgroups + 1) >> 1;
    v.ntiles = HW * v.spairs * p.nsplit;          // items: (pixel, group pair) tiles x output-channel splits
    return v;
}

// kF16: the 16-bit element type is a compile-time constant (fp16 / bf16) -- as a run-time flag the epilogue's conversions were
// compiled as both variants + a select per element
// NT: output channels per work item.  256 = a whole pixel tile (large batches).  128 = the tile is split into two output-channel
// halves that run on different CTA pairs (batches of ~100 ... ~1500 samples, where a layer has fewer pixel tiles than the chip has CTA
// pairs and the 9-tap tiles are the critical path of every layer: the halves take half the tensor time each and need no reduction;
// the price is that both halves load the tile's activations, which the L2 absorbs at these sizes).
template <bool kF16, int NT>
__global__ void __launch_bounds__(STACK_THREADS, 1) conv_stack_kernel(const __grid_constant__ StackParams p)
{
    constexpr int N = CH;
    constexpr int NSPLIT = CH / NT;                 // items per pixel tile
    constexpr int WCOLS = NT / 2;                   // columns per epilogue warp
    constexpr int NCHUNK = WCOLS / 32;              // 32-column TMEM chunks per epilogue warp
    constexpr int NSUB = WCOLS / 64;                // 64-channel staging sub-tiles (TMA boxes) per epilogue warp
    constexpr int B_BYTES = (NT / 2) * BLOCK_K * 2; // this CTA's half of an item's weight tile per k-chunk
    extern __shared__ __align__(1024) uint8_t smem[];
    float *s_shift = reinterpret_cast<float *>(smem + SS_OFF);       // the current layer's 256 shifts
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + BAR_OFF);
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + NSTAGE);
    const uint32_t bar_tfull = smem_u32(bars + 2 * NSTAGE), bar_tempty = smem_u32(bars + 2 * NSTAGE + 2);
    const uint32_t bar_res0 = smem_u32(bars + 2 * NSTAGE + 4);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + NBARS);
    int *s_ready = reinterpret_cast<int *>(bars + NBARS) + 1;        // tiles of this CTA's list whose inputs are known to be complete

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
    constexpr int kchunks = CH / BLOCK_K;
    const uint32_t smem_base = smem_u32(smem);
    if (smem_base & 1023u) __trap();
    // first tile of this CTA pair in `layer`; the pair then strides by nclusters.  All warp roles walk the same list.
    auto first_tile = [&](int vl) { return (cluster_id + vl * p.rot) % nclusters; };
    // this pair's items of virtual layer vl: count and k-th item (rotated round-robin, or the static balanced schedule)
    const int lpt0 = p.use_lpt ? p.lpt_off[cluster_id] : 0, lptn = p.use_lpt ? p.lpt_off[cluster_id + 1] - lpt0 : 0;
    auto item_count = [&](int vl, int ntiles) {
        if (p.use_lpt) return lptn;
        const int f = first_tile(vl);
        return f < ntiles ? (ntiles - f + nclusters - 1) / nclusters : 0;
    };
    auto item_at = [&](int vl, int k) { return p.use_lpt ? (int)p.lpt_items[lpt0 + k] : first_tile(vl) + k * nclusters; };
    const int nvl = p.nslices * p.nlayers;

    if (threadIdx.x == 0) *s_ready = 0;
    if (warp == 0 && lane == 0) {
        for (int b = 0; b < MAX_BUFS; ++b) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&p.map_act[b]) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&p.map_epi[b]) : "memory");
        }
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar_full + 8 * s, 2); mbar_init(bar_empty + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, 2 * NUM_EPI_WARPS); }
        for (int w = 0; w < NUM_EPI_WARPS; ++w) mbar_init(bar_res0 + 8 * w, 1);
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
    const int epoch = *reinterpret_cast<const volatile int *>(p.sync);   // advanced by the previous launch's last CTA
    const uint64_t pol_sticky = l2_policy_evict_last(p.sticky_frac), pol_other = p.other_policy ? l2_policy_evict_first() : l2_policy_evict_normal();

    if (warp == 0) {
        // ===================== TMA producer =====================
        // the whole warp walks the loops (converged); one elected lane issues (see elect_one)
        {
            const uint32_t lead_full = map_to_cta(bar_full, 0);
            int stage = 0, seq = 0;
            uint32_t phase = 0;
            for (int vl = 0; vl < nvl; ++vl) {
                const VLayer V = vlayer(p, vl);
                const int layer = V.layer;
                const StackLayer *L = p.layers + layer;
                const int src = L->src, k1 = L->k1;
                const uint32_t ltaps = L->taps;
                const int nit = item_count(vl, V.ntiles);
                for (int ik = 0; ik < nit; ++ik) {
                    const int item = item_at(vl, ik);
                    const bool item0 = ik == 0;
                    const int tile = item / NSPLIT, nh = item - tile * NSPLIT;
                    const int pix = tile / V.spairs, g = V.g0 + 2 * (tile - pix * V.spairs) + rank;
                    const int y0 = pix / LAT_W, x0 = pix - y0 * LAT_W;
                    const uint32_t taps = tap_mask(y0, x0) & ltaps;
                    if (item0 && lane == 0) STRACE(0, layer);
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
                        proxy_fence_global(p.debug);
                    }
                    if (item0 && lane == 0) STRACE(1, layer);
                    bool first_ks = true;
                    for (int tap = 0; tap < 9; ++tap) {
                        if (!((taps >> tap) & 1u)) continue;
                        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                        const int wrow = (k1 ? 0 : tap) * kchunks;
#pragma unroll
                        for (int kc = 0; kc < kchunks; ++kc) {
                            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                            const uint32_t sa = smem_base + A_OFF + stage * A_SLOT, sb = smem_base + B_OFF + stage * B_SLOT;
                            const bool do_a = first_ks || !(p.debug & 2), do_b = first_ks || !(p.debug & 1);
                            first_ks = false;
                            if (elect_one()) {
                                if (rank == 0) mbar_expect_tx(bar_full + 8 * stage, 2 * ((do_a ? A_SLOT : 0) + (do_b ? B_BYTES : 0)));
                                else mbar_arrive_cluster(lead_full + 8 * stage);
                                if (do_a) tma_load_4d_hint(sa, &p.map_act[src], lead_full + 8 * stage, kc * BLOCK_K, x0 + dx, y0 + dy, g * BLOCK_M,
                                                           g - V.g0 < p.sticky_groups ? pol_sticky : pol_other);
                                if (do_b) {
                                    if (NT == CH) tma_load_2d(sb, &L->map_b, lead_full + 8 * stage, 0, (wrow + kc) * N + rank * (N / 2));
                                    else tma_load_2d(sb, &L->map_b64, lead_full + 8 * stage, 0, (wrow + kc) * N + nh * NT + rank * (NT / 2));
                                }
                            }
                            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (pair leader only) =====================
        // the whole warp walks the loops (converged: the barrier waits are warp-wide); one elected lane issues the MMAs and commits
        if (rank == 0) {
            const uint32_t idesc = instr_desc(NT, kF16);
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (int vl = 0; vl < nvl; ++vl) {
                const VLayer V = vlayer(p, vl);
                const int layer = V.layer;
                const uint32_t ltaps = p.layers[layer].taps;
                const int nit = item_count(vl, V.ntiles);
                for (int ik = 0; ik < nit; ++ik, ++it) {
                    const int item = item_at(vl, ik);
                    const bool item0 = ik == 0;
                    const int buf = it & 1;
                    const int tile = item / NSPLIT;
                    const int pix = tile / V.spairs, y0 = pix / LAT_W, x0 = pix - y0 * LAT_W;
                    const bool tr4 = p.trace == 4 && blockIdx.x == 0 && it < 96 && lane == 0;
                    if (tr4) g_stack_trace[it * 4] = gtime_ns();
                    mbar_wait(bar_tempty + 8 * buf, ((it >> 1) & 1) ^ 1);
                    tc_fence_after();
                    if (tr4) g_stack_trace[it * 4 + 1] = gtime_ns();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(buf * N);
                    const int ksteps = __popc(tap_mask(y0, x0) & ltaps) * kchunks;
                    for (int ks = 0; ks < ksteps; ++ks) {
                        mbar_wait(bar_full + 8 * stage, phase);
                        tc_fence_after();
                        if (ks == 0 && item0 && lane == 0) STRACE(2, layer);
                        if (tr4 && ks == 0) g_stack_trace[it * 4 + 2] = gtime_ns();
                        const uint32_t sa = smem_base + A_OFF + stage * A_SLOT, sb = smem_base + B_OFF + stage * B_SLOT;
                        const uint64_t adesc = smem_desc(sa), bdesc = smem_desc(sb);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                                umma_bf16_pair(d_tmem, adesc + (uint64_t)(k * UMMA_K * 2 / 16), bdesc + (uint64_t)(k * UMMA_K * 2 / 16), idesc,
                                               (ks | k) ? 1u : 0u);
                            umma_commit_pair(bar_empty + 8 * stage);
                        }
                        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                    }
                    if (elect_one()) umma_commit_pair(bar_tfull + 8 * buf);
                    if (item0 && lane == 0) STRACE(3, layer);
                    if (tr4) g_stack_trace[it * 4 + 3] = (gtime_ns() - g_stack_trace[it * 4 + 2]) | ((unsigned long long)ksteps << 40);   // issue time | k-steps
                }
            }
        }
    } else if (warp < 2 + NUM_EPI_WARPS) {
        // ===================== epilogue (warps 2..9) =====================
        const int ew = warp - 2, quarter = warp & 3, half = ew >> 2;
        const int etid = threadIdx.x - 64;                                   // 0..255 among the epilogue threads
        const uint32_t stg = smem_base + EPI_OFF + ew * EPI_WARP;            // two SWIZZLE_128B tiles [32 rows][64 channels]
        const uint32_t srow = stg + lane * 128;
        const uint32_t bar_res = bar_res0 + 8 * ew;
        const uint32_t lead_tempty = map_to_cta(bar_tempty, 0);
        constexpr bool f16 = kF16;
        uint32_t res_phase = 0;
        int it = 0;
        for (int vl = 0; vl < nvl; ++vl) {
            const VLayer V = vlayer(p, vl);
            const int layer = V.layer;
            const StackLayer *L = p.layers + layer;
            // per-layer shifts: the first barrier proves every epilogue warp has finished the previous layer
            // (nobody reads the old values any more), the second that the new ones are in place
            asm volatile("bar.sync 1, 256;" ::: "memory");
            s_shift[etid] = L->shift[etid];
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const int res = L->res, dst = L->dst;
            const float *act_bias = L->act_bias;
            float *dst_f32 = L->dst_f32 ? L->dst_f32 + p.elem_off : nullptr;
            const uint8_t *res_lo = L->res_lo ? L->res_lo + p.elem_off : nullptr;
            uint8_t *dst_lo = L->dst_lo ? L->dst_lo + p.elem_off : nullptr;
            const float *res_f32 = L->res_f32 ? L->res_f32 + p.elem_off : nullptr;
            const int act = L->act;
            const int nit = item_count(vl, V.ntiles);
            for (int ik = 0; ik < nit; ++ik, ++it) {
                const int item = item_at(vl, ik);
                const bool item0 = ik == 0;
                const int buf = it & 1;
                const int tile = item / NSPLIT, nh = item - tile * NSPLIT;
                const int col0 = nh * NT + half * WCOLS;                     // this warp's WCOLS output channels
                const int pix = tile / V.spairs, gl = 2 * (tile - pix * V.spairs) + rank, g = V.g0 + gl;
                const int y0 = pix / LAT_W, x0 = pix - y0 * LAT_W;
                const int s0w = g * BLOCK_M + quarter * 32;                  // first sample of this warp's 32 rows
                const int s = s0w + lane;
                const bool live = gl < V.sgroups;                            // an odd group count leaves the last pair's second tile empty
                const bool valid = live && s < p.n;
                const long long m = (long long)s * HW + pix;                 // global output row
                const float *ab = (valid && act_bias) ? act_bias + ((size_t)p.act_idx[s] * HW + pix) * N : nullptr;
                const uint64_t pol = gl < p.sticky_groups ? pol_sticky : pol_other;
                // correction planes, item-private layout: [group][pixel][split][epilogue warp][2 * NCHUNK][32 lanes] x 16 bytes
                const size_t lo_off = ((((size_t)g * HW + pix) * NSPLIT + nh) * NUM_EPI_WARPS + ew) * (2 * NCHUNK * 32 * 16) + (size_t)lane * 16;
                if (layer > 0) {
                    // This tile's inputs -- the residual and its correction, written by an earlier layer's epilogue of the same
                    // (group, pixel), possibly on another SM -- are ordered before the previous layer's counters the scout has
                    // acquired; take the same acquire here (the scout runs ahead: normally one shared-memory read)
                    uint32_t spins = 0;
                    for (;;) {
                        int v;
                        asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(s_ready)) : "memory");
                        if (v >= it + 1) break;
                        if (++spins > (1u << 28)) __trap();
                    }
                }
                if (res >= 0 && live && elect_one()) {
                    proxy_fence_global(p.debug);
                    mbar_expect_tx(bar_res, NSUB * 4096);
#pragma unroll
                    for (int sub = 0; sub < NSUB; ++sub) tma_load_4d_cta_hint(stg + sub * 4096, &p.map_epi[res], bar_res, col0 + 64 * sub, x0, y0, s0w, pol);
                }
                uint4 lo_in[2 * NCHUNK];
                if (res_lo && valid) {
#pragma unroll
                    for (int i = 0; i < 2 * NCHUNK; ++i) lo_in[i] = ldcg_hint(res_lo + lo_off + i * 512, pol);
                }
                uint32_t acc[2][32];
                mbar_wait(bar_tfull + 8 * buf, (it >> 1) & 1);
                tc_fence_after();
                if (warp == 2 && lane == 0 && item0) STRACE(4, layer);
                if (res >= 0 && live) {
                    mbar_wait(bar_res, res_phase);
                    res_phase ^= 1;
                }
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * N + half * WCOLS);
                tmem_ld32_async(taddr, acc[0]);
#pragma unroll
                for (int c = 0; c < NCHUNK; ++c) {
                    tmem_wait(acc[c & 1]);
                    if (c + 1 < NCHUNK) tmem_ld32_async(taddr + (uint32_t)((c + 1) * 32), acc[(c + 1) & 1]);
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
                            const float4 b4 = *reinterpret_cast<const float4 *>(s_shift + c0 + q * 4);
                            v[q * 4] += b4.x; v[q * 4 + 1] += b4.y; v[q * 4 + 2] += b4.z; v[q * 4 + 3] += b4.w;
                        }
                        // this lane's row of the staging tile: 16-byte unit j of sub-tile c/2 sits at (j ^ (row & 7))
                        const uint32_t sbase = srow + (c >> 1) * 4096;
                        if (res >= 0) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                uint4 u4;
                                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(u4.x), "=r"(u4.y), "=r"(u4.z), "=r"(u4.w)
                                             : "r"(sbase + 16 * ((((c & 1) << 2) + q) ^ (lane & 7))) : "memory");
                                const uint32_t h[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 f = unpack2(h[e], f16);
                                    v[q * 8 + e * 2] += f.x;
                                    v[q * 8 + e * 2 + 1] += f.y;
                                }
                            }
                            if (res_lo) {
#pragma unroll
                                for (int i = 0; i < 2; ++i) {
                                    const uint32_t w[4] = {lo_in[2 * c + i].x, lo_in[2 * c + i].y, lo_in[2 * c + i].z, lo_in[2 * c + i].w};
#pragma unroll
                                    for (int e = 0; e < 8; ++e) {
                                        const float2 f = lo2((uint16_t)(w[e >> 1] >> ((e & 1) * 16)), f16);
                                        v[i * 16 + e * 2] += f.x;
                                        v[i * 16 + e * 2 + 1] += f.y;
                                    }
                                }
                            }
                        }
                        if (res_f32) {                                       // rare: a stream that enters as fp32 (one layer of an API-level call)
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float4 t4 = __ldcg(reinterpret_cast<const float4 *>(res_f32 + m * N + c0) + q);
                                v[q * 4] += t4.x; v[q * 4 + 1] += t4.y; v[q * 4 + 2] += t4.z; v[q * 4 + 3] += t4.w;
                            }
                        }
                        if (act == MZ_ACT_RELU) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
                        } else if (act != MZ_ACT_NONE) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = activate(v[j], act);
                        }
                        uint32_t hi[16];
                        if (dst_lo) {
                            uint32_t lw[8];
#pragma unroll
                            for (int e = 0; e < 16; e += 2) {
                                uint16_t l0, l1;
                                hi[e] = split2(v[e * 2], v[e * 2 + 1], f16, l0);
                                hi[e + 1] = split2(v[e * 2 + 2], v[e * 2 + 3], f16, l1);
                                lw[e >> 1] = (uint32_t)l0 | ((uint32_t)l1 << 16);
                            }
                            stcg_hint(dst_lo + lo_off + (2 * c) * 512, make_uint4(lw[0], lw[1], lw[2], lw[3]), pol);
                            stcg_hint(dst_lo + lo_off + (2 * c + 1) * 512, make_uint4(lw[4], lw[5], lw[6], lw[7]), pol);
                        } else {
#pragma unroll
                            for (int e = 0; e < 16; ++e) hi[e] = pack2(v[e * 2], v[e * 2 + 1], f16);
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sbase + 16 * ((((c & 1) << 2) + q) ^ (lane & 7))),
                                         "r"(hi[q * 4]), "r"(hi[q * 4 + 1]), "r"(hi[q * 4 + 2]), "r"(hi[q * 4 + 3]) : "memory");
                        if (dst_f32) {
                            float4 *fp = reinterpret_cast<float4 *>(dst_f32 + m * N + c0);
#pragma unroll
                            for (int q = 0; q < 8; ++q) fp[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                        }
                    }
                }
                tc_fence_before();
                fence_async_smem();                                          // my st.shared -> visible to the TMA store
                __syncwarp();
                if (elect_one()) {
                    mbar_arrive_cluster(lead_tempty + 8 * buf);              // accumulator drained: the pair's next-but-one tile may reuse it
                    if (live) {
#pragma unroll
                        for (int sub = 0; sub < NSUB; ++sub) tma_store_4d_hint(&p.map_epi[dst], stg + sub * 4096, col0 + 64 * sub, x0, y0, s0w, pol);
                        tma_store_commit();
                        tma_store_wait();                                    // global writes performed (and the staging tile is free again)
                        proxy_fence_global(p.debug);
                        // publish: this warp's part of (layer, group g, pixel pix) is in global memory -- the TMA stores above and, through
                        // the __syncwarp, every lane's correction-plane / fp32 stores (release is cumulative)
                        asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(p.done + ((size_t)layer * p.groups + g) * HW + pix) : "memory");
                    }
                }
                __syncwarp();                                                // nobody overwrites the staging tile before the store has read it
                if (warp == 2 && lane == 0 && item0) STRACE(5, layer);
            }
        }
    } else if (lane == 0) {
        // ===================== dependency scout (warp 10) =====================
        // Walks this CTA's tile list AHEAD of the TMA producer: for tile (pixel p, group g) of layer L+1 it polls
        // done[L][g][q] of the in-bounds neighbour pixels q (one L2 round trip for all of them), acquires, and publishes the
        // running count of cleared tiles in shared memory.  The producer's own wait is then a shared-memory read, so the
        // flag round trip + fence (~2 us) no longer sits between the last load of one tile and the first of the next.
        const int target = (epoch + 1) * NUM_EPI_WARPS * NSPLIT;      // every item's 8 epilogue warps arrive once per launch
        int seq = 0;
        for (int vl = 0; vl < nvl; ++vl) {
            const VLayer V = vlayer(p, vl);
            const int layer = V.layer;
            const uint32_t ltaps = p.layers[layer].taps;
            const int nit = item_count(vl, V.ntiles);
            for (int ik = 0; ik < nit; ++ik) {
                const int item = item_at(vl, ik);
                const int tile = item / NSPLIT;
                const int pix = tile / V.spairs, gl = 2 * (tile - pix * V.spairs) + rank, g = V.g0 + gl;
                const int y0 = pix / LAT_W, x0 = pix - y0 * LAT_W;
                ++seq;
                if (layer == 0) continue;
                if (gl < V.sgroups) {
                    const uint32_t taps = tap_mask(y0, x0) & ltaps;
                    const int *flags = p.done + ((size_t)(layer - 1) * p.groups + g) * HW;
                    uint32_t spins = 0;
                    for (;;) {
                        int ready = 1;
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            if (p.fine && ((taps >> tap) & 1u)) {
                                const int *fp = flags + (y0 + tap / 3 - 1) * LAT_W + (x0 + tap % 3 - 1);
                                int v;
                                // acquire loads (pair with the epilogues' red.release): no separate fence.acq_rel.gpu after the poll
                                // (256 roots +3.7 %, 1024 +1.4 %, 4096 +0.5 %; debug bit 32 = the old volatile poll + fence)
                                if (!(p.debug & 32)) asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(fp) : "memory");
                                else v = *reinterpret_cast<const volatile int *>(fp);
                                ready &= (v - target) >= 0;
                            }
                        }
                        if (!p.fine) {
#pragma unroll
                            for (int q = 0; q < HW; ++q) ready &= (*reinterpret_cast<const volatile int *>(flags + q) - target) >= 0;
                        }
                        if (ready) break;
                        if (++spins > (1u << 26)) __trap();
                        __nanosleep(p.nap);
                    }
                    if ((p.debug & 32) || !p.fine) asm volatile("fence.acq_rel.gpu;" ::: "memory");       // acquire for the volatile polls
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
    // the last CTA to finish advances the launch epoch (every CTA has read it by then)
    if (threadIdx.x == 0) {
        if (atomicAdd(p.sync + 1, 1) == (int)gridDim.x - 1) {
            p.sync[1] = 0;
            __threadfence();
            atomicAdd(p.sync, 1);
        }
    }
}

size_t done_ints(int n_layers, int nsamples) { return (size_t)n_layers * ((nsamples + BLOCK_M - 1) / BLOCK_M) * HW; }

}  // namespace

extern "C" {

size_t mz_stack_layer_bytes(void) { return sizeof(StackLayer); }

size_t mz_stack_scratch_bytes(int n_layers, int nsamples) { return (done_ints(n_layers, nsamples) + 4) * sizeof(int); }

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
        MZB_CHECK_ARG(o.op == MZ_OP_CONV && (o.dtype == MZ_BF16 || o.dtype == MZ_F16) && o.dtype == ops[0].dtype && o.use_tc && o.w_layout == 1 &&
                          (o.ksize == 3 || o.ksize == 1) && o.cin == CH && o.cout == CH && o.H == LAT_H && o.W == LAT_W,
                      "op is not a stackable 3x3 / 1x1 256->256 convolution on the 4x5 latent");
        MZB_CHECK_ARG(!o.scale, "the fused trunk takes weights with the BatchNorm scale folded in (scale == NULL)");
        MZB_CHECK_ARG(o.src && o.dst && o.w && o.shift && (!o.act_bias || o.act_idx), "missing operand");
        StackLayer &l = L[i];
        l.src = buf_id(o.src); l.dst = buf_id(o.dst); l.res = o.res ? buf_id(o.res) : -1;
        MZB_CHECK_ARG(l.src >= 0 && l.dst >= 0 && (!o.res || l.res >= 0), "op buffer is not one of the stack's activation buffers");
        MZB_CHECK_ARG(l.src != l.dst, "a convolution cannot run in place on its own input");
        MZB_CHECK_ARG(!o.res_lo || o.res, "res_lo without res");
        MZB_CHECK_ARG(!o.res_f32 || !o.res, "res_f32 replaces res / res_lo");
        // a layer's readers of older data are ordered by the 3x3 neighbour wait; a 1x1 layer waits for its own pixel only, so nothing
        // after it may overwrite what an earlier layer still reads: 1x1 layers come last and write buffers nobody else uses
        if (o.ksize == 1)
            for (int j = 0; j < n_ops; ++j)
                MZB_CHECK_ARG(j == i || (ops[j].src != o.dst && ops[j].dst != o.dst && ops[j].res != o.dst && (j < i || ops[j].ksize == 1)),
                              "a 1x1 convolution must be a trailing record with a destination of its own");
        l.shift = o.shift; l.act_bias = o.act_bias; l.dst_f32 = o.dst_f32; l.act = o.act;
        l.res_lo = reinterpret_cast<const uint8_t *>(o.res_lo); l.dst_lo = reinterpret_cast<uint8_t *>(o.dst_lo);
        l.res_f32 = o.res_f32;
        l.taps = o.ksize == 3 ? 0x1FFu : 0x010u;
        l.k1 = o.ksize == 1;
        const int wtaps = o.ksize * o.ksize;
        cuuint64_t dims[2] = {BLOCK_K, (cuuint64_t)wtaps * (CH / BLOCK_K) * CH};
        cuuint64_t strides[1] = {BLOCK_K * 2};
        cuuint32_t box[2] = {BLOCK_K, CH / 2};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&l.map_b, o.dtype == MZ_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(o.w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cuuint32_t box64[2] = {BLOCK_K, CH / 4};
        if (r == CUDA_SUCCESS)
            r = enc(&l.map_b64, o.dtype == MZ_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(o.w), dims, strides, box64, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { mzb::set_error("mz_stack_build: cuTensorMapEncodeTiled(weights) failed: %d", (int)r); return -2; }
    }
    // every layer but the first must read what an earlier layer of the run (or the caller) wrote; the counters only order
    // consecutive layers, so a layer's source has to be the previous 3x3 layer's destination or the trunk output the 1x1 heads share
    for (int i = 1; i < n_ops; ++i) {
        int prev = i - 1;
        while (prev > 0 && ops[prev].ksize == 1) --prev;          // trailing 1x1 records all read the last 3x3 layer's output
        MZB_CHECK_ARG(ops[i].src == ops[prev].dst, "the layers must form a chain: each one reads the previous 3x3 layer's output");
    }
    return 0;
}

int mz_stack_run(const void *blob_dev, int n_layers, int sample0, int nsamples, int slice_samples, void *const *bufs, int n_bufs, const int32_t *act_idx,
                 void *scratch, int dtype, void *stream)
{
    MZB_CHECK_ARG(blob_dev && n_layers > 0 && sample0 >= 0 && sample0 % (2 * BLOCK_M) == 0 && nsamples > 0 && bufs && n_bufs > 0 && n_bufs <= MAX_BUFS && scratch &&
                      (reinterpret_cast<uintptr_t>(scratch) & 15) == 0 && (dtype == MZ_BF16 || dtype == MZ_F16), "bad argument");
    MZB_CHECK_ARG(slice_samples >= 0 && slice_samples % (2 * BLOCK_M) == 0, "slice_samples must be a multiple of 256 (whole CTA-pair tiles), or 0 for one slice");
    EncodeTiledFn enc = encode_fn();
    if (!enc) { mzb::set_error("mz_stack_run: cuTensorMapEncodeTiled not available from the driver"); return -2; }
    cudaStream_t st = (cudaStream_t)stream;
    StackParams p{};
    static int promo_a = -1;
    if (promo_a < 0) { const char *e = getenv("MZB_STACK_L2PROMO"); promo_a = e ? atoi(e) : 3; }
    const CUtensorMapL2promotion l2p = promo_a == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : promo_a == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                       : promo_a == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    for (int b = 0; b < MAX_BUFS; ++b) {
        void *ptr = (__nv_bfloat16 *)bufs[b < n_bufs ? b : 0] + (size_t)sample0 * HW * CH;    // this launch's slice of the samples
        cuuint64_t dims[4] = {CH, LAT_W, LAT_H, (cuuint64_t)nsamples};
        cuuint64_t strides[3] = {CH * 2, LAT_W * CH * 2, HW * CH * 2};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        const CUtensorMapDataType tm = dtype == MZ_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
        cuuint32_t box[4] = {BLOCK_K, 1, 1, BLOCK_M};
        CUresult r = enc(&p.map_act[b], tm, 4, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, l2p, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cuuint32_t box_e[4] = {BLOCK_K, 1, 1, 32};
        if (r == CUDA_SUCCESS)
            r = enc(&p.map_epi[b], tm, 4, ptr, dims, strides, box_e, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, l2p, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { mzb::set_error("mz_stack_run: cuTensorMapEncodeTiled(activations) failed: %d", (int)r); return -2; }
    }
    p.f16 = dtype == MZ_F16;
    p.layers = reinterpret_cast<const StackLayer *>(blob_dev);
    p.nlayers = n_layers;
    p.done = reinterpret_cast<int *>(scratch) + 4;
    p.sync = reinterpret_cast<int *>(scratch);
    p.act_idx = act_idx ? act_idx + sample0 : nullptr;
    p.elem_off = (long long)sample0 * HW * CH;
    { static int tr = -1; if (tr < 0) { const char *e = getenv("MZB_STACK_TRACE"); tr = e ? atoi(e) : 0; } p.trace = tr; }
    { static int nap = -1; if (nap < 0) { const char *e = getenv("MZB_STACK_NAP"); nap = e ? atoi(e) : 32; } p.nap = nap; }
    { static int dbg = -1; if (dbg < 0) { const char *e = getenv("MZB_STACK_DEBUG"); dbg = e ? atoi(e) : 0; } p.debug = dbg; }
    { static int fine = -1; if (fine < 0) { const char *e = getenv("MZB_STACK_FINE"); fine = e ? atoi(e) : 1; } p.fine = fine; }
    { static int rot = -1; if (rot < 0) { const char *e = getenv("MZB_STACK_ROT"); rot = e ? atoi(e) : 13; } p.rot = rot; }
    p.n = nsamples;
    p.groups = (nsamples + BLOCK_M - 1) / BLOCK_M;
    p.slice_groups = slice_samples > 0 && slice_samples < nsamples ? slice_samples / BLOCK_M : p.groups;
    p.nslices = (p.groups + p.slice_groups - 1) / p.slice_groups;
    {   // L2 residency of the activations (see StackParams::sticky_groups).  A slice's live activations (two 16-bit buffers + the correction
        // plane = 128 x 20 x 256 x 5 bytes per group) stay in the L2 from layer to layer up to ~2048 samples (52 MB); beyond that the L2
        // thrashes (ncu: 2.7 GB of DRAM traffic per 28-layer launch at 4096 samples = every layer's input read from and output written to
        // DRAM).  From 3072 samples per slice on, the first MZB_STACK_STICKY_MB (default 45) megabytes' worth of sample groups move with
        // evict_last and the rest with evict_first (MZB_STACK_OTHER_POLICY, 0 = evict_normal): 2.71 -> 2.14 GB, +0.6 % simulations/s
        // (profiles/r2_l2_policy.txt; evict_last alone: 2.56 GB -- the hardware treats the hint as a priority, not as a reservation).
        static int mb = -1, other = 1; static float frac = 1.0f;
        if (mb < 0) {
            const char *e = getenv("MZB_STACK_STICKY_MB"); mb = e ? atoi(e) : 45;
            e = getenv("MZB_STACK_OTHER_POLICY"); other = e ? atoi(e) : 1;
            e = getenv("MZB_STACK_STICKY_FRAC"); frac = e ? (float)atof(e) : 1.0f;
        }
        const double group_mb = BLOCK_M * HW * CH * 5.0 / 1e6 * frac;
        const bool thrash = p.slice_groups * BLOCK_M >= 3072;
        p.sticky_groups = thrash && mb > 0 ? (int)(mb / group_mb) : 0;
        p.other_policy = thrash && mb > 0 ? other : 0;
        p.sticky_frac = frac;
    }
    // output-channel-split items (two CTA pairs per pixel tile) while a layer has fewer pixel tiles than ~2 per CTA pair: a pure function
    // of the batch size (the scratch counters of a (trunk, nsamples) always see the same split); MZB_STACK_NSPLIT_MAX overrides the limit
    static int nsplit_max = -1;
    if (nsplit_max < 0) { const char *e = getenv("MZB_STACK_NSPLIT_MAX"); nsplit_max = e ? atoi(e) : 768; }
    p.nsplit = nsamples <= nsplit_max ? 2 : 1;
    const int max_tiles = HW * ((p.slice_groups + 1) / 2) * p.nsplit;          // items of a (full) slice per layer
    static bool attr_set[64] = {};
    if (mzb::first_use_on_device(attr_set)) {
        MZB_CUDA(cudaFuncSetAttribute(conv_stack_kernel<true, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STACK_SMEM));
        MZB_CUDA(cudaFuncSetAttribute(conv_stack_kernel<false, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STACK_SMEM));
        MZB_CUDA(cudaFuncSetAttribute(conv_stack_kernel<true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STACK_SMEM));
        MZB_CUDA(cudaFuncSetAttribute(conv_stack_kernel<false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STACK_SMEM));
        // every CTA pair must be resident at once (pairs wait on each other's counters): one CTA per SM by shared memory, so the
        // grid may not exceed what this device / context can co-schedule
    }
    int max_clusters = 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(mzb::kNumSMs);
    cfg.blockDim = dim3(STACK_THREADS);
    cfg.dynamicSmemBytes = STACK_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    // cooperative: the launch starts only when the whole grid is resident, so two persistent trunks enqueued on different
    // streams can never hold part of the chip each and wait for CTAs that cannot start (MZB_STACK_COOP=0 switches it off)
    static int coop = -1;
    if (coop < 0) { const char *e = getenv("MZB_STACK_COOP"); coop = e ? atoi(e) : 1; }
    attr[1].id = cudaLaunchAttributeCooperative;
    attr[1].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    {
        static int cached[64] = {};
        int d = 0;
        cudaGetDevice(&d);
        if (d < 0 || d >= 64 || cached[d] == 0) {
            MZB_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, conv_stack_kernel<true, 256>, &cfg));
            if (d >= 0 && d < 64) cached[d] = max_clusters;
        } else max_clusters = cached[d];
    }
    if (max_clusters < 1) { mzb::set_error("mz_stack_run: no CTA pair of the fused trunk fits on this device"); return -2; }
    int clusters = max_tiles < mzb::kNumSMs / 2 ? max_tiles : mzb::kNumSMs / 2;
    if (clusters > max_clusters) clusters = max_clusters;       // fewer SMs than a full B200 (MIG / MPS limits): still all co-resident
    { static int cap = -1; if (cap < 0) { const char *e = getenv("MZB_STACK_MAX_PAIRS"); cap = e ? atoi(e) : 0; } if (cap > 0 && clusters > cap) clusters = cap; }   // experiments
    // Static balanced schedule for single-slice launches whose layers have only a few items per CTA pair (the rotation of a round-robin
    // map evens the 4 / 6 / 9-tap item costs out over SEVERAL layers, which needs dependency slack that small batches do not have):
    // longest-processing-time-first assignment of a layer's items to the pairs, the same in every layer.
    static int lpt_max = -1;
    if (lpt_max < 0) { const char *e = getenv("MZB_STACK_LPT_MAX"); lpt_max = e ? atoi(e) : 1792; }
    if (p.nslices == 1 && nsamples <= lpt_max && max_tiles <= MZ_STACK_MAX_LPT_ITEMS && clusters <= MZ_STACK_MAX_PAIRS) {
        const int spairs = (p.groups + 1) / 2, nitems = HW * spairs * p.nsplit;
        int load[MZ_STACK_MAX_PAIRS] = {}, owner[MZ_STACK_MAX_LPT_ITEMS], count[MZ_STACK_MAX_PAIRS] = {};
        for (int cost = 9; cost >= 1; --cost)                     // items in descending cost (stable in item order)
            for (int item = 0; item < nitems; ++item) {
                const int pix = (item / p.nsplit) / spairs, y0 = pix / LAT_W, x0 = pix - y0 * LAT_W;
                int taps = 0;
                for (int t = 0; t < 9; ++t) taps += (y0 + t / 3 - 1 >= 0 && y0 + t / 3 - 1 < LAT_H && x0 + t % 3 - 1 >= 0 && x0 + t % 3 - 1 < LAT_W);
                if (taps != cost) continue;
                int best = 0;
                for (int c = 1; c < clusters; ++c) if (load[c] < load[best] || (load[c] == load[best] && count[c] < count[best])) best = c;
                owner[item] = best; load[best] += cost; count[best] += 1;
            }
        int off = 0;
        for (int c = 0; c < clusters; ++c) {
            p.lpt_off[c] = (uint16_t)off;
            for (int item = 0; item < nitems; ++item) if (owner[item] == c) p.lpt_items[off++] = (uint16_t)item;     // ascending item (= pixel) order
        }
        p.lpt_off[clusters] = (uint16_t)off;
        p.use_lpt = 1;
    }
    cfg.gridDim = dim3(2 * clusters);
    cfg.numAttrs = coop ? 2 : 1;
    if (p.nsplit == 2) {
        if (p.f16) MZB_CUDA(cudaLaunchKernelEx(&cfg, conv_stack_kernel<true, 128>, p));
        else MZB_CUDA(cudaLaunchKernelEx(&cfg, conv_stack_kernel<false, 128>, p));
    } else {
        if (p.f16) MZB_CUDA(cudaLaunchKernelEx(&cfg, conv_stack_kernel<true, 256>, p));
        else MZB_CUDA(cudaLaunchKernelEx(&cfg, conv_stack_kernel<false, 256>, p));
    }
    MZB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
