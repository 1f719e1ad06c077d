// How many clusters of 2 / 4 / 8 / 16 CTAs with the fused trunk's footprint (one CTA per SM by shared memory) can be
// resident on this GPU at once?  nvcc -arch=sm_100a cluster_occupancy.cu -o cluster_occupancy && ./cluster_occupancy
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(352, 1) k(int *p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main()
{
    const int smem = 231000;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    printf("%s: %d SMs\n", pr.name, pr.multiProcessorCount);
    for (int cs : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(352); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim = {(unsigned)cs, 1, 1};
        cfg.attrs = a; cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
        printf("cluster size %2d: max active clusters %3d = %3d CTAs (%s)\n", cs, n, n * cs, cudaGetErrorString(e));
    }
    return 0;
}
