#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { extern __shared__ float s[]; if (p) p[0] = (int)s[0]; }
int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  printf("SMs %d smem/block optin %zu\n", pr.multiProcessorCount, pr.sharedMemPerBlockOptin);
  for (int cl : {2, 4, 8, 16}) for (size_t smem : {size_t(100*1024), size_t(223712)}) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cl > 8) cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(cl * 32); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster %2d smem %6zu: max active clusters %d (%s) -> %d SMs\n", cl, smem, n, cudaGetErrorString(e), n * cl);
  }
  return 0;
}
