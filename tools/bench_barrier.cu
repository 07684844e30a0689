// Micro-benchmark of grid-wide barrier variants for the persistent decoder (one CTA per SM, 256 threads).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bench_barrier tools/bench_barrier.cu && /tmp/bench_barrier
#include <cuda_runtime.h>
#include <stdio.h>

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_release(unsigned* p) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }

// V0: as in ot_decoder.cu (first version): proxy fences by all threads, threadfence + atomicAdd + acquire spin + threadfence
template <int V>
__device__ __forceinline__ void barrier(unsigned* ctr, unsigned* flags, unsigned& target, unsigned& epoch) {
  if (V == 0) {
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
    target += gridDim.x;
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(ctr, 1u);
      while ((int)(ld_acquire(ctr) - target) < 0) {}
      __threadfence();
      asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncthreads();
  } else if (V == 1) {   // no proxy fences
    __syncthreads();
    target += gridDim.x;
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(ctr, 1u);
      while ((int)(ld_acquire(ctr) - target) < 0) {}
      __threadfence();
    }
    __syncthreads();
  } else if (V == 2) {   // red.release + ld.acquire, no explicit fences
    __syncthreads();
    target += gridDim.x;
    if (threadIdx.x == 0) {
      red_release(ctr);
      while ((int)(ld_acquire(ctr) - target) < 0) {}
    }
    __syncthreads();
  } else if (V == 3) {   // flag all-gather: one release store, thread i polls CTA i's flag
    __syncthreads();
    ++epoch;
    if (threadIdx.x == 0) st_release(flags + blockIdx.x, epoch);
    if (threadIdx.x < gridDim.x) {
      while ((int)(ld_acquire(flags + threadIdx.x) - epoch) < 0) {}
    }
    __syncthreads();
  } else if (V == 4) {   // flag all-gather, relaxed polls + one fence
    __syncthreads();
    ++epoch;
    if (threadIdx.x == 0) st_release(flags + blockIdx.x, epoch);
    if (threadIdx.x < gridDim.x) {
      while ((int)(ld_relaxed(flags + threadIdx.x) - epoch) < 0) {}
      __threadfence();
    }
    __syncthreads();
  } else if (V == 5) {   // V2 + proxy fence only by thread 0
    __syncthreads();
    target += gridDim.x;
    if (threadIdx.x == 0) {
      asm volatile("fence.proxy.async;" ::: "memory");
      red_release(ctr);
      while ((int)(ld_acquire(ctr) - target) < 0) {}
      asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncthreads();
  } else if (V == 6) {   // V3 with flags padded to 128 B each
    __syncthreads();
    ++epoch;
    if (threadIdx.x == 0) st_release(flags + blockIdx.x * 32, epoch);
    if (threadIdx.x < gridDim.x) {
      while ((int)(ld_acquire(flags + threadIdx.x * 32) - epoch) < 0) {}
    }
    __syncthreads();
  } else if (V == 7) {   // V2 + all-thread proxy fence before
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
    target += gridDim.x;
    if (threadIdx.x == 0) {
      red_release(ctr);
      while ((int)(ld_acquire(ctr) - target) < 0) {}
    }
    __syncthreads();
  }
}

template <int V>
__global__ void __launch_bounds__(256, 1) k(unsigned* ctr, unsigned* flags, int iters, float* sink) {
  extern __shared__ char smem[];
  unsigned target = 0, epoch = 0;
  float acc = 0.f;
  for (int i = 0; i < iters; ++i) {
    sink[blockIdx.x * 256 + threadIdx.x] = acc + i;     // some global write per phase
    barrier<V>(ctr, flags, target, epoch);
    acc += __ldcg(sink + ((blockIdx.x + 1) % gridDim.x) * 256 + threadIdx.x);
  }
  sink[blockIdx.x * 256 + threadIdx.x] = acc;
}

template <int V>
void run(const char* name, int grid) {
  unsigned *ctr, *flags;
  float* sink;
  cudaMalloc(&ctr, 4);
  cudaMalloc(&flags, 148 * 128);
  cudaMalloc(&sink, 148 * 256 * 4);
  cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
  const int iters = 2000;
  float best = 1e9;
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemset(ctr, 0, 4);
    cudaMemset(flags, 0, 148 * 128);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    void* args[] = {&ctr, &flags, (void*)&iters, &sink};
    cudaLaunchCooperativeKernel((void*)k<V>, dim3(grid), dim3(256), args, 120 * 1024, 0);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(err)); return; }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  printf("%-58s grid %3d: %.3f us / barrier\n", name, grid, best * 1e3 / iters);
  cudaFree(ctr); cudaFree(flags); cudaFree(sink);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run<0>("V0 proxy fences(all) + threadfence + atomicAdd + acquire", sms);
  run<1>("V1 threadfence + atomicAdd + acquire spin + threadfence", sms);
  run<2>("V2 red.release + ld.acquire spin", sms);
  run<5>("V5 V2 + proxy fences by thread 0", sms);
  run<7>("V7 V2 + proxy fence by all threads before", sms);
  run<3>("V3 flag all-gather (st.release, per-thread ld.acquire)", sms);
  run<4>("V4 flag all-gather (relaxed polls + fence)", sms);
  run<6>("V6 flag all-gather, 128 B per flag", sms);
  run<2>("V2 at grid 64", 64);
  run<3>("V3 at grid 64", 64);
  return 0;
}
