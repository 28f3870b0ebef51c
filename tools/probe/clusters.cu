// How many thread-block clusters of a given size can be resident at once (B200)?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int *x) { extern __shared__ int s[]; if (x) x[0] = s[0]; }
int main() {
    for (int smem : {30 * 1024, 120 * 1024}) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int cs : {1, 2, 4, 8, 16}) {
            if (cs == 16) cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            cudaLaunchConfig_t lc = {};
            lc.gridDim = dim3(cs, 64); lc.blockDim = dim3(288); lc.dynamicSmemBytes = smem;
            cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension;
            a[0].val.clusterDim.x = cs; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
            lc.attrs = a; lc.numAttrs = 1;
            int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &lc);
            printf("smem %d KB cluster %d: max active clusters %d (%s)\n", smem / 1024, cs, n, cudaGetErrorString(e));
        }
    }
    return 0;
}
