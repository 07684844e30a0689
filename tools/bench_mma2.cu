// Micro-benchmark: what bounds a chain of small tcgen05.mma.kind::i8 instructions -- the issuing thread, the accumulator dependency or
// the tensor pipe?  16 fully unrolled MMAs per group (descriptors precomputed, compile-time shape), spread round-robin over NACC
// accumulators (column offset = acc * N); cycles = (first issue .. commit arrival) / MMAs.  tools/bench_mma.cu has the one-accumulator table.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I onnx-transformer_b200/csrc -o tools/bin/bench_mma2 tools/bench_mma2.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#include "ot_ptx.cuh"
using namespace ot;

template <int M, int N, int NACC, bool UNI>
__global__ void __launch_bounds__(128, 1) k(int groups, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
    __syncwarp();
    tmem_alloc(smem_u32(&slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const int warp_u = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x) / 32, 0);
  // UNI: the whole warp takes the branch and one lane is elected inside (ptxas emits back-to-back UTCIMMA);
  // !UNI: `threadIdx.x == 0` -- a divergent branch, ptxas wraps every UTCIMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop
  if (UNI ? (warp_u == 0 && elect_one()) : (threadIdx.x == 0)) {
    constexpr uint32_t idesc = make_idesc_i8(M, N);
    const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem));
    const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem + 32768));
    uint32_t parity = 0;
    for (int round = 0; round < 3; ++round) {
      const long long t0 = clock64();
#pragma unroll 1
      for (int g = 0; g < groups; ++g) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          mma_i8_ss(tmem + (i % NACC) * N, a_desc + static_cast<uint64_t>((i & 3) * 2 + (i >> 2) * 64), b_desc + static_cast<uint64_t>((i & 3) * 2 + (i >> 2) * 512), idesc,
                    (g > 0 || i >= NACC) ? 1u : 0u);
      }
      mma_commit(smem_u32(&bar));
      const long long t1 = clock64();
      mbar_wait(smem_u32(&bar), parity);
      parity ^= 1u;
      const long long t2 = clock64();
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int M, int N, int NACC, bool UNI = true>
void run(long long* out) {
  cudaFuncSetAttribute(k<M, N, NACC, UNI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int groups : {1, 4}) {
    out[0] = out[1] = 0;
    k<M, N, NACC, UNI><<<1, 128, 96 * 1024>>>(groups, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("M=%d N=%d nacc=%d: %s\n", M, N, NACC, cudaGetErrorString(e)); exit(1); }
    printf("%6d %6d %6d %5s %8d %12lld %12lld %14.1f\n", M, N, NACC, UNI ? "elect" : "tid0", 16 * groups, out[0], out[1], (double)out[1] / (16 * groups));
  }
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 16);
  printf("tcgen05.mma kind::i8 K=32, smem operands, MMAs spread over nacc accumulators; cyc_per_mma = (issue..commit arrival)/MMAs\n");
  printf("%6s %6s %6s %5s %8s %12s %12s %14s\n", "M", "N", "nacc", "issue", "mmas", "issue_cyc", "total_cyc", "cyc_per_mma");
  run<64, 64, 1, false>(out); run<64, 192, 1, false>(out); run<64, 256, 1, false>(out); run<128, 256, 1, false>(out);
  run<64, 64, 1>(out); run<64, 64, 2>(out); run<64, 64, 4>(out); run<64, 64, 8>(out);
  run<64, 96, 1>(out); run<64, 96, 2>(out); run<64, 96, 4>(out);
  run<64, 128, 1>(out); run<64, 128, 2>(out); run<64, 128, 4>(out);
  run<64, 192, 1>(out); run<64, 192, 2>(out);
  run<64, 256, 1>(out); run<64, 256, 2>(out);
  run<128, 16, 1>(out); run<128, 16, 4>(out);
  run<128, 64, 1>(out); run<128, 64, 4>(out);
  run<128, 128, 1>(out); run<128, 128, 2>(out); run<128, 128, 4>(out);
  run<128, 256, 1>(out); run<128, 256, 2>(out);
  return 0;
}
