// Shared device-side wrappers (PTX for mbarrier / TMA / tcgen05 / cluster), tile constants and the tensor-map
// encoder of the tensor-core convolution kernels (conv_tc.cu: one layer per launch; conv_stack.cu: a whole
// residual trunk per launch).  Everything has internal linkage: include inside the .cu file's anonymous use.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;          // bf16 per smem row = 128 bytes = one SWIZZLE_128B atom
constexpr int UMMA_K = 16;
constexpr int STAGES = 5;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;    // 16 KB
constexpr int B_STAGE_BYTES = 128 * BLOCK_K * 2;        // 16 KB: this CTA's half (N/2 <= 128 rows) of the weight tile
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int NUM_EPI_WARPS = 8;                        // two per TMEM lane quarter, each takes half of the N columns
constexpr int NUM_THREADS = 64 + NUM_EPI_WARPS * 32;
constexpr int TMEM_COLS = 512;
constexpr int EPI_STAGE_BYTES = 32 * 256;               // per epilogue warp: 32 rows x (N/2 <= 128 cols) bf16, residual in / result out
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + (size_t)NUM_EPI_WARPS * EPI_STAGE_BYTES + 2 * 256 * sizeof(float) + 128;   // 226.1 KB of the 227 KB

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// One lane of a CONVERGED warp (all 32 lanes must call it).  Unlike `lane == 0`, the elect.sync predicate tells ptxas that exactly one
// lane runs the guarded code, so instructions of the uniform datapath (UTCHMMA, UTMALDG / UTMASTG, UTCBAR) are issued straight-line
// with their operands in uniform registers; behind a divergent `lane == 0` branch every one of them is wrapped in an
// ELECT / BRA.U.ANY "waterfall" loop with R2UR moves (seen in the SASS of the round-1 kernels' MMA issue loop).
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred = 0, laneid = 0;
    asm volatile(
        "{\n.reg .b32 rx;\n.reg .pred px;\n"
        "elect.sync rx|px, %2;\n"
        "@px mov.s32 %1, 1;\n"
        "mov.s32 %0, rx;\n}"
        : "+r"(laneid), "+r"(pred)
        : "r"(0xFFFFFFFFu));
    return pred != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    // bounded spin: a protocol bug traps (launch error reported to the host) instead of hanging the GPU
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 24); ++it) {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.b32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}
// cta_group::2 TMA loads: data lands in THIS CTA's shared memory, the transaction bytes are signalled on the
// mbarrier `bar`, a shared::cluster address that may belong to the peer (the pair leader's full barrier)
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// Plain (cta_group-less) TMA tile load into THIS CTA's shared memory, bytes signalled on this CTA's mbarrier `bar`
// (a shared::cta address is a valid shared::cluster address of the executing CTA): the epilogue's residual tiles
__device__ __forceinline__ void tma_load_4d_cta(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// TMA tile store shared -> global (bulk async-group completion); rows / pixels outside the tensor are clipped
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, uint32_t src, int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
        ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// ---- the same copies with an L2 cache-policy operand (createpolicy descriptors: the fused trunk keeps part of its activations
// resident in the L2 across layers, conv_stack.cu)
__device__ __forceinline__ uint64_t l2_policy_evict_last(float fraction)
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, %1;" : "=l"(pol) : "f"(fraction));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_4d_hint(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2, int c3, uint64_t pol)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_cta_hint(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2, int c3, uint64_t pol)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void tma_store_4d_hint(const CUtensorMap *map, uint32_t src, int c0, int c1, int c2, int c3, uint64_t pol)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4, %5}], [%1], %6;"
        ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(pol)
        : "memory");
}
__device__ __forceinline__ uint4 ldcg_hint(const void *ptr, uint64_t pol)
{
    uint4 v;
    asm volatile("ld.global.cg.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ void stcg_hint(void *ptr, uint4 v, uint64_t pol)
{
    asm volatile("st.global.cg.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all of this thread's bulk stores have READ their shared-memory source (the staging tile may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... have completed: their global writes are performed
__device__ __forceinline__ void tma_store_wait() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy shared-memory writes (st.shared) -> visible to the async proxy (the TMA store that reads them)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// commit of the pair's MMAs: arrives on the mbarrier at this offset in BOTH CTAs when they have retired
__device__ __forceinline__ void umma_commit_pair(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)0x3) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// tcgen05.ld of 32 consecutive fp32 columns of this thread's TMEM lane; asynchronous until tmem_wait().
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
// wait for the outstanding tcgen05.ld; the registers are in/out operands so that no use of them can be
// scheduled above the wait
__device__ __forceinline__ void tmem_wait(uint32_t (&r)[32])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                   "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): rows of 128 bytes,
// 8-row groups 1024 bytes apart (SBO), LBO unused (=1), version 1 (Blackwell), layout type 2.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16 (format 1) or fp16 (format 0), both
// K-major, M=256 (the pair), N
__device__ __forceinline__ uint32_t instr_desc(int N, bool f16 = false)
{
    const uint32_t fmt = f16 ? 0u : 1u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((2 * BLOCK_M) >> 4) << 24);
}
// the same with separate operand formats (A = dy bf16, B = x fp16 in the weight gradient of an fp16-forward training step)
__device__ __forceinline__ uint32_t instr_desc_ab(int N, bool a_f16, bool b_f16)
{
    return (1u << 4) | ((a_f16 ? 0u : 1u) << 7) | ((b_f16 ? 0u : 1u) << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((2 * BLOCK_M) >> 4) << 24);
}
// two packed 16-bit activations <-> floats; `f16` is uniform per launch
__device__ __forceinline__ float2 unpack2(uint32_t u, bool f16)
{
    if (f16) return __half22float2(*reinterpret_cast<const __half2 *>(&u));
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&u));
}
__device__ __forceinline__ uint32_t pack2(float a, float b, bool f16)
{
    if (f16) { const __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<const uint32_t *>(&h); }
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}

// ---------------------------------------------------------------- residual stream: 16-bit value + 8-bit correction
// The residual stream x_k of a ResidualBlock chain (networks.py:31-35: x_{k+1} = act(bn2(conv2(..)) + x_k)) is the one
// place where 16-bit STORAGE compounds: every block would add a fresh rounding of the whole stream (14 blocks per trunk).
// The stream is therefore kept as hi = 16-bit round(x) (the tensor-core operand of the next convolution) plus
// lo = e4m3((x - hi) * LO_SCALE) in a second, 1-byte plane that only the convolution epilogues read and write:
// x ~ hi + lo / LO_SCALE carries 15 (fp16) / 12 (bf16) significant bits.  |x - hi| <= ulp(hi)/2, so the scaled difference is at most
// |hi| and e4m3 (max 448, saturating) holds it for |x| < 448; x - hi and the scaling are exact in fp32.
// Measured against the fp32 reference (profiles/emulate_precision.py): fp16 trunks 1.3-1.7e-3 -> 4.4-6.1e-4 of range.
__device__ __forceinline__ float lo_scale(bool f16) { return f16 ? 2048.0f : 256.0f; }
// two fp32 results -> packed 16-bit pair (returned) and their two e4m3 corrections (low byte = first value)
__device__ __forceinline__ uint32_t split2(float a, float b, bool f16, uint16_t &lo)
{
    const uint32_t h = pack2(a, b, f16);
    const float2 back = unpack2(h, f16);
    const float sc = lo_scale(f16);
    lo = __nv_cvt_float2_to_fp8x2(make_float2((a - back.x) * sc, (b - back.y) * sc), __NV_SATFINITE, __NV_E4M3);
    return h;
}
// the two corrections of a packed e4m3 pair as floats, already divided by LO_SCALE
__device__ __forceinline__ float2 lo2(uint16_t lo, bool f16)
{
    const __half2_raw r = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)lo, __NV_E4M3);
    const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&r));
    const float inv = f16 ? (1.0f / 2048.0f) : (1.0f / 256.0f);
    return make_float2(f.x * inv, f.y * inv);
}

__device__ __forceinline__ float activate(float v, int act)
{
    switch (act) {
        case MZ_ACT_RELU: return fmaxf(v, 0.0f);
        case MZ_ACT_LEAKY_RELU: return v > 0.0f ? v : 0.01f * v;
        case MZ_ACT_SILU: return v / (1.0f + __expf(-v));
        case MZ_ACT_GELU: return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
        default: return v;
    }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

}  // namespace

