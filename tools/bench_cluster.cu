// Micro-benchmark for the cluster-resident decoder design: how many clusters of 8 / 16 CTAs (one CTA per SM, ~215 KB of
// shared memory) are co-resident on a B200, and what one "push my slice to every peer + cluster barrier + read" round costs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bench_cluster tools/bench_cluster.cu && /tmp/bench_cluster
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_size() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}

// Each iteration: every CTA writes `f4_per_peer` float4 into its slot of EVERY peer's buffer, cluster barrier, then sums what
// it received (ping-pong buffers, as the decoder would).
__global__ void __launch_bounds__(256, 1) k(int iters, int f4_per_peer, float* out) {
  extern __shared__ __align__(16) char smem[];
  float4* buf = reinterpret_cast<float4*>(smem);            // [2][16][f4_per_peer]
  const uint32_t rank = cluster_rank(), n = cluster_size();
  float acc = 0.f;
  cluster_sync();
  for (int it = 0; it < iters; ++it) {
    float4* mine = buf + ((it & 1) * 16 + rank) * f4_per_peer;
    for (int i = threadIdx.x; i < f4_per_peer * (int)n; i += blockDim.x) {
      const int peer = i / f4_per_peer, j = i % f4_per_peer;
      st_cluster_v4(mapa(smem_u32(mine + j), peer), make_float4(acc, it, rank, j));
    }
    cluster_sync();
    const float4* rd = buf + (it & 1) * 16 * f4_per_peer;
    for (int i = threadIdx.x; i < f4_per_peer * (int)n; i += blockDim.x) acc += rd[i].x + rd[i].w;
  }
  out[blockIdx.x * 256 + threadIdx.x] = acc;
}

static void run(int cluster, int nclusters, int f4_per_peer, int smem_bytes) {
  float* out;
  cudaMalloc(&out, 256 * 256 * 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cluster * nclusters);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int maxc = -1;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&maxc, k, &cfg);
  const int iters = 2000;
  float best = 1e9;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaLaunchKernelEx(&cfg, k, iters, f4_per_peer, out);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("cluster %d x %d: %s\n", cluster, nclusters, cudaGetErrorString(err)); cudaGetLastError(); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  printf("cluster %2d x %2d clusters, smem %3d KB, push %5d B/peer: maxActiveClusters=%d (%s)  %.3f us / round\n", cluster, nclusters,
         smem_bytes / 1024, f4_per_peer * 16, maxc, cudaGetErrorString(e), best * 1e3 / iters);
  cudaFree(out);
}

int main() {
  for (int smem : {64 * 1024, 215 * 1024}) {
    run(16, 1, 16, smem);
    run(16, 8, 16, smem);
    run(16, 9, 16, smem);
    run(8, 16, 16, smem);
    run(8, 18, 16, smem);
  }
  run(16, 8, 1, 215 * 1024);
  run(16, 8, 64, 215 * 1024);     // 1 KB per peer (8 rows x 32 fp32)
  run(16, 8, 256, 215 * 1024);    // 4 KB per peer (8 rows x 128 fp32)
  run(8, 16, 64, 215 * 1024);
  run(8, 16, 256, 215 * 1024);
  return 0;
}
