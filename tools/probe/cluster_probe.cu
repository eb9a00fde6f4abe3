// How many clusters of 2 / 4 / 8 CTAs of the sampler's shape (640 threads, 232 KB dynamic shared memory, 1 CTA per SM)
// can be co-resident on this GPU?  nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_probe cluster_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(640, 1) k_dummy(int* p) {
  extern __shared__ unsigned char smem[];
  if (p != nullptr && threadIdx.x == 0) p[blockIdx.x] = smem[0];
}
int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int smem = 232000;   // of the 232,448 bytes a CTA may have
  cudaFuncSetAttribute(k_dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  printf("SMs %d\n", sms);
  for (int c = 1; c <= 16; c *= 2) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sms / c * c);
    cfg.blockDim = dim3(640);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = c;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k_dummy, &cfg);
    printf("cluster %2d: max active clusters %d (= %d CTAs)  %s\n", c, n, n * c, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
