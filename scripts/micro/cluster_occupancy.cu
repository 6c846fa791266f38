// Development aid: how many thread-block clusters of a given size can be co-resident with one 227 KB CTA per SM?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_occupancy cluster_occupancy.cu && ./cluster_occupancy
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(192, 1) dummy(int* p) { extern __shared__ unsigned char s[]; if (p) p[0] = s[0]; }
int main() {
  cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  printf("SMs %d\n", prop.multiProcessorCount);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(prop.multiProcessorCount / cs * cs); cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = 232448;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
    printf("cluster %2d: max active clusters %d (= %d SMs) %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
