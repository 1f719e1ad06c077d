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

int main()
{
    float *out; long long *cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
    for (int w : {1, 4, 8, 16, 32}) { run<1, false>(w, out, cyc); run<2, false>(w, out, cyc); run<4, false>(w, out, cyc); run<8, false>(w, out, cyc); }
    for (int w : {4, 8, 16, 32}) { run<2, true>(w, out, cyc); run<8, true>(w, out, cyc); }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
