// Weight gradient of the 3x3 / 1x1 convolutions (256 -> 256 trunks; 64 / 128 / 256 -> 128 / 256 stems, head ConvBlocks and the
// representation network's 128-channel blocks: autograd of nn.Conv2d, src/networks.py:11,24-25,47,65,117,139,201,213, inside
// loss.backward() train_torch.py:515) on the B200 tensor cores:  dW[co][ci][ky][kx] = sum over (sample n, pixel (y,x)) of
// dY[n][y][x][co] * X[n][y+ky-1][x+kx-1][ci].
//
// As a GEMM per tap: M = co (256 = one CTA pair, tcgen05 cta_group::2), N = ci (256, one UMMA N), K = samples x pixels.  In the
// channels-last activation layout K is the OUTER index of both operands, so they are first transposed to channel-major,
// pixel, sample ([C][P][ns], ns = samples padded to 64 with zeros; wgrad_transpose_kernel, 2 x 5 MB per layer at 512 samples):
// then both operands are K-major exactly like the forward convolution's (same TMA boxes, SWIZZLE_128B, same UMMA descriptors),
// the tap is a constant column offset (p' - p) * ns between the two operands, and the zero-padding pixels of a tap are simply
// not in its K range (no masking, no zero work -- the forward kernel's per-pixel tap skipping, seen from the other side).
//
// Tiles: (tap, K-split) -- 9 taps x up to 8 splits of the sample chunks = 72 CTA pairs on the 148 SMs; every pair accumulates
// its 256 x 256 fp32 partial in TMEM and writes it to `partial[tap][split]`; wgrad_reduce_kernel sums the splits in a fixed order
// (deterministic) into PyTorch's weight layout (cout, cin, k, k).
// Warp roles as in conv_tc.cu: warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer (pair leader), warps 2-9 epilogue.
//
// Other channel counts (mz_conv_wgrad_any): the GEMM is P[tap][r][c] = sum_pix A[r][pix] * B[c][pix + tap offset] with Ca <= 256 rows of A
// (the pair's M = 256; with Ca = 128 the second CTA's rows are never loaded nor read) and N = Cb in {64, 128, 256} columns.  cout = 256:
// A = dY, B = X.  cout = 128, cin = 256: operands SWAPPED (A = X, B = dY, so that M stays 256): P[tap] is then the transposed gradient of
// the mirrored tap, which the reduction kernel undoes.  cout = 128, cin <= 128: A = dY with half of the pair's rows idle.
#include "tc_common.cuh"

namespace {

constexpr int WG_C = 256;                       // cin = cout
constexpr int WG_MAX_SPLITS = 8;
constexpr size_t WG_SMEM = (size_t)STAGES * STAGE_BYTES + 256;

struct WgradParams {
    int ns, H, W, taps, chunks, splits, ntiles, a_f16, b_f16;     // element types of the A and B operands
    int Ca, Cb;                                 // rows of A that exist (128 or 256), columns N of the product (64, 128, 256)
    float *partial;                             // [taps][splits][256][Cb]
};

// k-steps of tile (tap, split): pixels whose shifted partner is inside the image x the split's 64-sample chunks
struct WgTile {
    int dy, dx, c0, c1;
};
__device__ __forceinline__ WgTile wg_decode(const WgradParams &p, int tile)
{
    WgTile t;
    const int tap = tile / p.splits, sp = tile - tap * p.splits;
    t.dy = p.taps == 1 ? 0 : tap / 3 - 1;
    t.dx = p.taps == 1 ? 0 : tap % 3 - 1;
    t.c0 = (int)((long long)sp * p.chunks / p.splits);
    t.c1 = (int)((long long)(sp + 1) * p.chunks / p.splits);
    return t;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, const WgradParams p)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + STAGES);
    const uint32_t bar_tfull = smem_u32(bars + 2 * STAGES), bar_tempty = smem_u32(bars + 2 * STAGES + 2);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
    const uint32_t smem_base = smem_u32(smem);
    if (smem_base & 1023u) __trap();

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dy) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
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
    const int P = p.H * p.W;

    if (warp == 0) {
        // ===================== TMA producer (one lane per CTA) =====================
        if (lane == 0) {
            const uint32_t lead_full = map_to_cta(bar_full, 0);
            const uint32_t tx_bytes = (uint32_t)((p.Ca > 128 ? 2 : 1) * A_STAGE_BYTES + p.Cb * BLOCK_K * 2);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < p.ntiles; tile += nclusters) {
                const WgTile t = wg_decode(p, tile);
                for (int pix = 0; pix < P; ++pix) {
                    const int y = pix / p.W, x = pix - y * p.W;
                    if (y + t.dy < 0 || y + t.dy >= p.H || x + t.dx < 0 || x + t.dx >= p.W) continue;
                    const int pix2 = pix + t.dy * p.W + t.dx;
                    for (int c = t.c0; c < t.c1; ++c) {
                        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                        const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_STAGE_BYTES;
                        if (rank == 0) mbar_expect_tx(bar_full + 8 * stage, tx_bytes);              // bytes of both CTAs
                        else mbar_arrive_cluster(lead_full + 8 * stage);
                        if (rank * 128 < p.Ca) tma_load_2d(sa, &map_dy, lead_full + 8 * stage, pix * p.ns + c * BLOCK_K, rank * 128);    // my 128 rows of A
                        tma_load_2d(sb, &map_x, lead_full + 8 * stage, pix2 * p.ns + c * BLOCK_K, rank * (p.Cb >> 1));                   // my half of the B rows
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (pair leader only) =====================
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = instr_desc_ab(p.Cb, p.a_f16 != 0, p.b_f16 != 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = cluster_id; tile < p.ntiles; tile += nclusters, ++it) {
                const int buf = it & 1;
                mbar_wait(bar_tempty + 8 * buf, ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * WG_C);
                const WgTile t = wg_decode(p, tile);
                int npix = 0;
                for (int pix = 0; pix < P; ++pix) {
                    const int y = pix / p.W, x = pix - y * p.W;
                    npix += (y + t.dy >= 0 && y + t.dy < p.H && x + t.dx >= 0 && x + t.dx < p.W) ? 1 : 0;
                }
                const int ksteps = npix * (t.c1 - t.c0);
                for (int ks = 0; ks < ksteps; ++ks) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_STAGE_BYTES;
                    const uint64_t adesc = smem_desc(sa), bdesc = smem_desc(sb);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                        umma_bf16_pair(d_tmem, adesc + (uint64_t)(k * UMMA_K * 2 / 16), bdesc + (uint64_t)(k * UMMA_K * 2 / 16), idesc, (ks | k) ? 1u : 0u);
                    umma_commit_pair(bar_empty + 8 * stage);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_pair(bar_tfull + 8 * buf);
            }
        }
    } else {
        // ===================== epilogue (warps 2..9): TMEM -> fp32 partial =====================
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        const int co = rank * 128 + quarter * 32 + lane;                 // TMEM lane = output row
        const uint32_t lead_tempty = map_to_cta(bar_tempty, 0);
        int it = 0;
        for (int tile = cluster_id; tile < p.ntiles; tile += nclusters, ++it) {
            const int buf = it & 1;
            const WgTile t = wg_decode(p, tile);
            int npix = 0;
            for (int pix = 0; pix < P; ++pix) {
                const int y = pix / p.W, x = pix - y * p.W;
                npix += (y + t.dy >= 0 && y + t.dy < p.H && x + t.dx >= 0 && x + t.dx < p.W) ? 1 : 0;
            }
            const bool empty = npix * (t.c1 - t.c0) == 0;                // no k-step: nothing was accumulated, the partial is zero
            const int hcols = p.Cb >> 1, nch = p.Cb >> 6;                // columns / 32-column chunks of this warp's half
            float *dst = p.partial + ((size_t)tile * WG_C + co) * p.Cb + half * hcols;
            uint32_t acc[2][32];
            mbar_wait(bar_tfull + 8 * buf, (it >> 1) & 1);
            tc_fence_after();
            if (rank * 128 < p.Ca) {                                     // CTA-uniform: with Ca = 128 the second CTA's accumulator rows are idle
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * WG_C + half * hcols);
                tmem_ld32_async(taddr, acc[0]);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c < nch) {
                        tmem_wait(acc[c & 1]);
                        if (c + 1 < nch) tmem_ld32_async(taddr + (uint32_t)((c + 1) * 32), acc[(c + 1) & 1]);
                        float4 *fp = reinterpret_cast<float4 *>(dst + c * 32);
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            fp[q] = empty ? make_float4(0.f, 0.f, 0.f, 0.f)
                                          : make_float4(__uint_as_float(acc[c & 1][q * 4]), __uint_as_float(acc[c & 1][q * 4 + 1]),
                                                        __uint_as_float(acc[c & 1][q * 4 + 2]), __uint_as_float(acc[c & 1][q * 4 + 3]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lead_tempty + 8 * buf);
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

// [n][P][C] (channels-last, 16-bit) -> [C][P][ns], samples n..ns-1 zero.  One CTA = 64 samples x 64 channels of one pixel: 32-byte
// loads along the channels, 32-byte stores along the samples, through a padded shared-memory tile.  grid (ns/64, C/64, P), 256 threads
// cvt: 1 = the source is fp16 and the destination bf16 (the activations of an fp16 forward pass as the weight gradient's bf16 operand: one MMA
// cannot mix A / B element types in kind::f16 -- the hardware answers "illegal instruction")
// ns = row length of dst (all samples of the K-concatenated operand), s_off = first sample column of this call
// A launch may carry a second tensor (src2 / dst2 with C2 channels: the other operand of the same weight gradient): blockIdx.y >= C / 64
// belongs to it -- one launch per (dy, x) pair instead of two.
__global__ void __launch_bounds__(256) wgrad_transpose_kernel(int n, int ns, int s_off, int P, int C, const uint16_t *__restrict__ src, uint16_t *__restrict__ dst, int cvt,
                                                              int C2, const uint16_t *__restrict__ src2, uint16_t *__restrict__ dst2, int cvt2)
{
    mzb::pdl_trigger();
    mzb::pdl_wait();
    __shared__ uint16_t tile[64][64 + 2];                    // row pitch 132 bytes = 33 words: column reads hit 32 different banks
    int by = blockIdx.y;
    if (by >= C / 64) { by -= C / 64; C = C2; src = src2; dst = dst2; cvt = cvt2; }
    const int n0 = blockIdx.x * 64, c0 = by * 64, pix = blockIdx.z;
    const int r = threadIdx.x >> 2, q = (threadIdx.x & 3) * 16;
    {
        const int s = n0 + r;
        uint4 v[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
        if (s < n) {
            const uint4 *g = reinterpret_cast<const uint4 *>(src + ((size_t)s * P + pix) * C + c0 + q);
            v[0] = __ldg(g); v[1] = __ldg(g + 1);
        }
        const uint32_t *w = reinterpret_cast<const uint32_t *>(v);
        uint32_t *t32 = reinterpret_cast<uint32_t *>(&tile[r][q]);   // (r * 66 + q) * 2 bytes: 4-byte aligned (q even)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t u = w[i];
            if (cvt) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&u));
                const __nv_bfloat162 b = __floats2bfloat162_rn(f.x, f.y);
                u = *reinterpret_cast<const uint32_t *>(&b);
            }
            t32[i] = u;
        }
    }
    __syncthreads();
    {
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = (uint32_t)tile[q + 2 * i][r] | ((uint32_t)tile[q + 2 * i + 1][r] << 16);   // channel c0 + r, samples n0 + q ..
        uint4 *g = reinterpret_cast<uint4 *>(dst + ((size_t)(c0 + r) * P + pix) * ns + s_off + n0 + q);
        g[0] = make_uint4(w[0], w[1], w[2], w[3]);
        g[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
}

// dw[co][ci][tap] = sum over splits of partial[tap'][split][row][col], splits added in index order.  Not swapped: row = co, col = ci,
// tap' = tap.  Swapped operands: row = ci, col = co, tap' = taps - 1 - tap (the mirrored offset).  One thread per (tap, row, col):
// its <= 8 loads are independent and coalesced over col; grid (rows * Cb / 256, taps)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(int taps, int splits, int rows, int Cb, int swap, int dw_cin, const float *__restrict__ partial,
                                                           float *__restrict__ dw, int accumulate)
{
    const int i = blockIdx.x * 256 + threadIdx.x;            // row * Cb + col
    if (i >= rows * Cb) return;
    const int row = i / Cb, col = i - row * Cb;
    const int tp = blockIdx.y;                               // tap' of the partial
    const size_t tile_stride = (size_t)WG_C * Cb;
    const float *src = partial + (size_t)tp * splits * tile_stride + (size_t)row * Cb + col;
    float v[WG_MAX_SPLITS];
#pragma unroll
    for (int s = 0; s < WG_MAX_SPLITS; ++s) v[s] = s < splits ? __ldcs(src + (size_t)s * tile_stride) : 0.0f;
    float acc = 0.0f;
#pragma unroll
    for (int s = 0; s < WG_MAX_SPLITS; ++s) acc += v[s];     // index order; the padding terms are exact zeros
    // PyTorch's (cout, dw_cin, k, k); the gradient covers its first cin input channels (not swapped cin = Cb, swapped cin = rows)
    float *out = swap ? dw + ((size_t)col * dw_cin + row) * taps + (taps - 1 - tp) : dw + ((size_t)row * dw_cin + col) * taps + tp;
    *out = accumulate ? *out + acc : acc;                    // accumulate: the K unroll steps of a training step add into the parameter's .grad
}

int wg_splits(int ns) { const int chunks = ns / BLOCK_K; return chunks < WG_MAX_SPLITS ? chunks : WG_MAX_SPLITS; }
bool wg_shape_ok(int cout, int cin) { return (cout == 128 || cout == 256) && (cin == 64 || cin == 128 || cin == 256); }

}  // namespace

extern "C" {

int mz_wgrad_padded_samples(int n) { return n <= 0 ? 0 : (n + BLOCK_K - 1) / BLOCK_K * BLOCK_K; }

size_t mz_wgrad_partial_bytes(int ksize, int n) { return mz_wgrad_partial_bytes_any(ksize, n, WG_C, WG_C); }

size_t mz_wgrad_partial_bytes_any(int ksize, int n, int cout, int cin)
{
    if ((ksize != 1 && ksize != 3) || n <= 0 || !wg_shape_ok(cout, cin)) return 0;
    const int Cb = (cout == 128 && cin == 256) ? cout : cin;
    return (size_t)ksize * ksize * wg_splits(mz_wgrad_padded_samples(n)) * WG_C * Cb * sizeof(float);
}

int mz_wgrad_transpose(int n, int P, int C, const void *src, void *dst, void *stream)
{
    return mz_wgrad_transpose_cvt(n, P, C, src, dst, 0, stream);
}

int mz_wgrad_transpose_cvt(int n, int P, int C, const void *src, void *dst, int f16_to_bf16, void *stream)
{
    return mz_wgrad_transpose_into(n, P, C, src, dst, mz_wgrad_padded_samples(n), 0, f16_to_bf16, stream);
}

int mz_wgrad_transpose_into(int n, int P, int C, const void *src, void *dst, int ns_total, int s_offset, int f16_to_bf16, void *stream)
{
    MZB_CHECK_ARG(n > 0 && P > 0 && C > 0 && C % 64 == 0 && src && dst, "bad argument");
    MZB_CHECK_ARG((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "buffers must be 16-byte aligned");
    const int ns = mz_wgrad_padded_samples(n);
    MZB_CHECK_ARG(s_offset >= 0 && s_offset % BLOCK_K == 0 && ns_total % BLOCK_K == 0 && s_offset + ns <= ns_total, "sample window outside the destination rows");
    MZB_CHECK_ARG(P <= 65535 && C / 64 <= 65535, "image or channel count too large");
    MZB_CUDA(mzb::launch_chain(wgrad_transpose_kernel, dim3(ns / 64, C / 64, P), dim3(256), 0, (cudaStream_t)stream, n, ns_total, s_offset, P, C,
                               (const uint16_t *)src, (uint16_t *)dst, f16_to_bf16 != 0, 0, (const uint16_t *)nullptr, (uint16_t *)nullptr, 0));
    MZB_LAUNCH_CHECK();
    return 0;
}

int mz_wgrad_transpose_pair(int n, int P, int Ca, const void *src_a, void *dst_a, int cvt_a, int Cb, const void *src_b, void *dst_b, int cvt_b, int ns_total,
                            int s_offset, void *stream)
{
    MZB_CHECK_ARG(n > 0 && P > 0 && Ca > 0 && Ca % 64 == 0 && Cb > 0 && Cb % 64 == 0 && src_a && dst_a && src_b && dst_b, "bad argument");
    MZB_CHECK_ARG((((uintptr_t)src_a | (uintptr_t)dst_a | (uintptr_t)src_b | (uintptr_t)dst_b) & 15) == 0, "buffers must be 16-byte aligned");
    const int ns = mz_wgrad_padded_samples(n);
    MZB_CHECK_ARG(s_offset >= 0 && s_offset % BLOCK_K == 0 && ns_total % BLOCK_K == 0 && s_offset + ns <= ns_total, "sample window outside the destination rows");
    MZB_CHECK_ARG(P <= 65535 && (Ca + Cb) / 64 <= 65535, "image or channel count too large");
    MZB_CUDA(mzb::launch_chain(wgrad_transpose_kernel, dim3(ns / 64, (Ca + Cb) / 64, P), dim3(256), 0, (cudaStream_t)stream, n, ns_total, s_offset, P, Ca,
                               (const uint16_t *)src_a, (uint16_t *)dst_a, cvt_a != 0, Cb, (const uint16_t *)src_b, (uint16_t *)dst_b, cvt_b != 0));
    MZB_LAUNCH_CHECK();
    return 0;
}

int mz_conv_wgrad(int n, int H, int W, int ksize, int dtype, const void *dy_t, const void *x_t, float *partial, float *dw, void *stream)
{
    return mz_conv_wgrad_accum(n, H, W, ksize, dtype, dy_t, x_t, partial, dw, 0, stream);
}

int mz_conv_wgrad_accum(int n, int H, int W, int ksize, int dtype, const void *dy_t, const void *x_t, float *partial, float *dw, int accumulate, void *stream)
{
    return mz_conv_wgrad_any(n, H, W, ksize, dtype, WG_C, WG_C, WG_C, dy_t, x_t, partial, dw, accumulate, stream);
}

int mz_conv_wgrad_any(int n, int H, int W, int ksize, int dtype, int cout, int cin, int dw_cin, const void *dy_t, const void *x_t, float *partial,
                      float *dw, int accumulate, void *stream)
{
    MZB_CHECK_ARG(dw_cin >= cin, "dw_cin (input channels of the dw tensor) must be >= cin");
    MZB_CHECK_ARG(n > 0 && H > 0 && W > 0 && (ksize == 1 || ksize == 3) && (dtype == MZ_BF16 || dtype == MZ_F16), "bad argument");
    MZB_CHECK_ARG(wg_shape_ok(cout, cin), "cout must be 128 or 256, cin 64, 128 or 256");
    MZB_CHECK_ARG(dy_t && x_t && partial && dw, "null pointer");
    MZB_CHECK_ARG((((uintptr_t)dy_t | (uintptr_t)x_t | (uintptr_t)partial) & 15) == 0, "buffers must be 16-byte aligned");
    EncodeTiledFn enc = encode_fn();
    if (!enc) { mzb::set_error("mz_conv_wgrad: cuTensorMapEncodeTiled not available from the driver"); return -2; }
    const bool swap = cout == 128 && cin == 256;              // keep M = 256: A = x (rows ci), B = dy (columns co)
    WgradParams p{};
    p.ns = mz_wgrad_padded_samples(n);
    p.H = H; p.W = W; p.taps = ksize * ksize;
    p.chunks = p.ns / BLOCK_K;
    p.splits = wg_splits(p.ns);
    p.ntiles = p.taps * p.splits;
    p.a_f16 = dtype == MZ_F16;
    p.b_f16 = dtype == MZ_F16;
    p.Ca = swap ? cin : cout;
    p.Cb = swap ? cout : cin;
    p.partial = partial;
    const long long K = (long long)H * W * p.ns;
    MZB_CHECK_ARG(K < (1ll << 31), "samples x pixels too large for one launch");
    CUtensorMap maps[2];
    const void *ptrs[2] = {swap ? x_t : dy_t, swap ? dy_t : x_t};
    const int chans[2] = {p.Ca, p.Cb};
    for (int i = 0; i < 2; ++i) {
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)chans[i]};
        cuuint64_t strides[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {BLOCK_K, (cuuint32_t)(i == 0 ? 128 : p.Cb / 2)};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&maps[i], (i == 0 ? p.a_f16 : p.b_f16) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptrs[i]), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { mzb::set_error("mz_conv_wgrad: cuTensorMapEncodeTiled failed: %d", (int)r); return -2; }
    }
    static bool attr_set[64] = {};
    if (mzb::first_use_on_device(attr_set))
        MZB_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_SMEM));
    const int clusters = p.ntiles < mzb::kNumSMs / 2 ? p.ntiles : mzb::kNumSMs / 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = WG_SMEM;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MZB_CUDA(cudaLaunchKernelEx(&cfg, wgrad_kernel, maps[0], maps[1], p));
    MZB_LAUNCH_CHECK();
    wgrad_reduce_kernel<<<dim3((p.Ca * p.Cb + 255) / 256, p.taps), 256, 0, (cudaStream_t)stream>>>(p.taps, p.splits, p.Ca, p.Cb, swap ? 1 : 0, dw_cin, partial, dw, accumulate);
    MZB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
