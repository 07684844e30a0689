// Micro-benchmark behind a design choice of csrc/ot_cdecoder.cu: what does ONE tcgen05.mma.cta_group::1.kind::i8 (K = 32) cost as a
// function of its shape when the operands are already in shared memory?  One CTA, one issuing thread, `reps` back-to-back MMAs into
// the same accumulator, tcgen05.commit, mbarrier wait; cycles from clock64.  Operands are whatever the shared memory holds.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I onnx-transformer_b200/csrc -o tools/bin/bench_mma tools/bench_mma.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#include "ot_ptx.cuh"
using namespace ot;

__global__ void __launch_bounds__(128, 1) k(int m, int n, int reps, int a_in_tmem, long long* out) {
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
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_i8(m, n);
    const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem));
    const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem + 32768));
    uint32_t parity = 0;
    for (int round = 0; round < 3; ++round) {      // last round is the one reported (warm instruction cache)
      const long long t0 = clock64();
      for (int i = 0; i < reps; ++i) {
        if (a_in_tmem) {
          // A operand from tensor memory (columns 256..): [d], [a], b-desc, idesc
          asm volatile(
              "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
              " tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem),
              "r"(tmem + 256 + 8 * (i & 3)), "l"(b_desc + static_cast<uint64_t>((i & 3) * 2)), "r"(idesc), "r"(i ? 1u : 0u)
              : "memory");
        } else {
          mma_i8_ss(tmem, a_desc + static_cast<uint64_t>((i & 3) * 2), b_desc + static_cast<uint64_t>((i & 3) * 2), idesc, i ? 1u : 0u);
        }
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

int main() {
  long long* out;
  cudaMallocManaged(&out, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  printf("tcgen05.mma kind::i8, K=32 per instruction, operands resident; cycles per MMA = (issue..mbarrier) / reps\n");
  printf("%6s %6s %8s %12s %12s %14s\n", "M", "N", "A", "reps", "total_cyc", "cyc_per_mma");
  for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
    for (int m : {64, 128})
      for (int n : {16, 64, 128, 192, 256})
        for (int reps : {16, 64}) {
          out[0] = out[1] = 0;
          k<<<1, 128, 96 * 1024>>>(m, n, reps, a_tmem, out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("M=%d N=%d a_tmem=%d: %s\n", m, n, a_tmem, cudaGetErrorString(e)); return 1; }
          printf("%6d %6d %8s %12d %12lld %14.1f\n", m, n, a_tmem ? "tmem" : "smem", reps, out[1], (double)out[1] / reps);
        }
  return 0;
}
