// Microbenchmark: issue rate of the legacy warp MMA (mma.sync.m16n8k16 bf16 -> fp32) and of ldmatrix.x4 on one SM of a B200,
// as a function of resident warps and independent accumulators per warp.  nvcc -arch=sm_100a -O3 -o hmma_rate hmma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int NACC, bool LDSM>
__global__ void k(float *out, long long *cyc, int iters)
{
    extern __shared__ __align__(128) uint8_t sm[];
    float acc[NACC][4];
    for (int i = 0; i < NACC; ++i) for (int q = 0; q < 4; ++q) acc[i][q] = 0.f;
    uint32_t a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b0 = 0x3f803f80u, b1 = 0x3f803f80u;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 528 + (threadIdx.x >> 5) * 16;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (LDSM) {
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(base + (it & 7) * 32));
        }
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                         : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < NACC; ++i) for (int q = 0; q < 4; ++q) s += acc[i][q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int NACC, bool LDSM>
void run(int warps, float *out, long long *cyc)
{
    const int iters = 2000;
    cudaFuncSetAttribute(k<NACC, LDSM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    k<NACC, LDSM><<<1, warps * 32, 64 * 1024>>>(out, cyc, iters);
    k<NACC, LDSM><<<1, warps * 32, 64 * 1024>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double mma = (double)iters * NACC * warps;
    printf("warps %2d  acc/warp %d  ldsm %d: %8lld cycles, %.2f cycles per MMA per SM (%.0f MAC/clk/SM), %.1f cycles per warp-iteration\n", warps, NACC, (int)LDSM, c,
           c / mma, mma * 2048 / c, (double)c / iters);
}

// the latency-mode trunk's step: 5 ldmatrix.x4 (prefetched one step ahead) + 8 MMAs on 8 accumulators
template <int DEPTH>
__global__ void k58(float *out, long long *cyc, int iters)
{
    extern __shared__ __align__(128) uint8_t sm[];
    float acc[8][4];
    for (int i = 0; i < 8; ++i) for (int q = 0; q < 4; ++q) acc[i][q] = 0.f;
    uint32_t f[DEPTH + 1][5][4];
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 7) * 528 + ((threadIdx.x >> 3) & 3) * 16 + (threadIdx.x >> 5) * 64;
    for (int d = 0; d <= DEPTH; ++d) for (int j = 0; j < 5; ++j) for (int q = 0; q < 4; ++q) f[d][j][q] = 0x3f803f80u;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it += DEPTH + 1) {
#pragma unroll
        for (int u = 0; u <= DEPTH; ++u) {
            // load the fragments DEPTH steps ahead into slot (u + DEPTH) % (DEPTH + 1), consume slot u
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                uint32_t (&r)[4] = f[(u + DEPTH) % (DEPTH + 1)][j];
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(base + j * 4224 + ((it + u) & 3) * 32));
            }
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int n = 0; n < 2; ++n)
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                                 : "+f"(acc[m * 2 + n][0]), "+f"(acc[m * 2 + n][1]), "+f"(acc[m * 2 + n][2]), "+f"(acc[m * 2 + n][3])
                                 : "r"(f[u][m][0]), "r"(f[u][m][1]), "r"(f[u][m][2]), "r"(f[u][m][3]), "r"(f[u][4][n * 2]), "r"(f[u][4][n * 2 + 1]));
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 8; ++i) for (int q = 0; q < 4; ++q) s += acc[i][q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int DEPTH>
void run58(int warps, float *out, long long *cyc)
{
    const int iters = 1800;
    cudaFuncSetAttribute(k58<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    k58<DEPTH><<<1, warps * 32, 64 * 1024>>>(out, cyc, iters);
    k58<DEPTH><<<1, warps * 32, 64 * 1024>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("5 ldsm.x4 + 8 mma per step, prefetch depth %d, warps %2d: %.1f cycles per step per warp, %.2f cycles per MMA per SM, %.2f per LDSM.x4 per SM\n", DEPTH, warps,
           (double)c / iters, (double)c / iters / (8.0 * warps), (double)c / iters / (5.0 * warps));
}

int main()
{
    float *out; long long *cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
    for (int w : {8}) { run<1, false>(w, out, cyc); run<2, false>(w, out, cyc); run<4, false>(w, out, cyc); run<8, false>(w, out, cyc); }
    for (int w : {8, 16}) { run<2, true>(w, out, cyc); run<8, true>(w, out, cyc); }
    for (int w : {4, 8, 12, 16}) { run58<1>(w, out, cyc); run58<2>(w, out, cyc); }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
