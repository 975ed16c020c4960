// How many clusters of 2 / 4 / 8 CTAs (one CTA per SM: 227 KB of shared memory) can be co-resident on this GPU?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scratch/ubench/clusters scratch/ubench/clusters.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
  const int smem = 227 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  printf("%s: %d SMs\n", prop.name, prop.multiProcessorCount);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(prop.multiProcessorCount / cs * cs);
    cfg.blockDim = dim3(320);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster size %2d: max active clusters %d -> %d SMs busy (%s)\n", cs, n, n * cs, cudaGetErrorString(e));
  }
  return 0;
}
