// How many clusters of 8 / 16 CTAs with the decoder's footprint (256 threads, ~220 KB dynamic shared memory, 1 CTA per SM) can be
// co-resident on this GPU?  nvcc -arch=sm_100a -o tools/bin/cluster_occupancy tools/cluster_occupancy.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256, 1) probe(int* out) {
  extern __shared__ unsigned char smem[];
  if (threadIdx.x == 0 && out) out[blockIdx.x] = smem[0];
}

int main() {
  const int smem = 224256;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("%s: %d SMs\n", p.name, p.multiProcessorCount);
  for (int cs : {2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 32);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe, &cfg);
    printf("cluster size %2d: max active clusters %d (%d CTAs)%s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
