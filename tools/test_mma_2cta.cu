// Known-answer test of tcgen05.mma.cta_group::2.kind::i8 (a CTA pair computing one 256 x 256 x 128 int8 product: each CTA holds 128 rows
// of A, 128 of the 256 rows of B and 128 rows x 256 columns of the int32 accumulator in its own TMEM; the leader CTA issues the MMAs and
// a multicast commit signals a barrier in both CTAs).  Development aid for a 2-CTA form of the K = 2048 encoder GEMM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I onnx-transformer_b200/csrc -I include -o tools/bin/test_mma_2cta tools/test_mma_2cta.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "ot_ptx.cuh"

using namespace ot;

__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_i8_ss_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit_2cta(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k2cta(const int8_t* A, const int8_t* B, int32_t* D) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = smem;                    // [128 rows][128 B], 128-byte swizzle
  uint8_t* sB = smem + 16384;            // [128 rows of W][128 B]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 32768 + 64);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  if (warp == 0) {
    if (elect_one()) {
      mbar_init(smem_u32(bar), 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc2(smem_u32(slot), 256);
    tmem_relinquish2();
  }
  // operands: thread = row, 8 chunks of 16 bytes, chunk index XOR (row & 7)
  for (int c = 0; c < 8; ++c) {
    const uint4 a = *reinterpret_cast<const uint4*>(A + (rank * 128 + tid) * 128 + c * 16);
    const uint4 b = *reinterpret_cast<const uint4*>(B + (rank * 128 + tid) * 128 + c * 16);
    *reinterpret_cast<uint4*>(sA + tid * 128 + ((c ^ (tid & 7)) << 4)) = a;
    *reinterpret_cast<uint4*>(sB + tid * 128 + ((c ^ (tid & 7)) << 4)) = b;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  cluster_sync_all();                    // both CTAs: barrier initialised, TMEM allocated, operands in shared memory
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (rank == 0 && warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_i8(256, 256);
      const uint64_t a_desc = make_smem_desc_sw128(smem_u32(sA)), b_desc = make_smem_desc_sw128(smem_u32(sB));
      for (int k = 0; k < 4; ++k) mma_i8_ss_2cta(tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, k != 0);
      mma_commit_2cta(smem_u32(bar), 3);
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(bar), 0);
  tc_fence_after();
  const uint32_t tl = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  int32_t* drow = D + (rank * 128 + warp * 32 + lane) * 256;
  for (int c = 0; c < 16; ++c) {
    uint32_t r[16];
    tmem_ld_32x16(tl + 16 * c, r);
    tmem_wait_ld();
    for (int j = 0; j < 16; ++j) drow[16 * c + j] = static_cast<int32_t>(r[j]);
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2(tmem, 256);
}

int main() {
  const int M = 256, N = 256, K = 128;
  int8_t *hA = (int8_t*)malloc(M * K), *hB = (int8_t*)malloc(N * K);
  int32_t* hD = (int32_t*)malloc(M * N * 4);
  srand(1);
  for (int i = 0; i < M * K; ++i) hA[i] = (int8_t)(rand() % 255 - 127);
  for (int i = 0; i < N * K; ++i) hB[i] = (int8_t)(rand() % 255 - 127);
  int8_t *dA, *dB;
  int32_t* dD;
  cudaMalloc(&dA, M * K); cudaMalloc(&dB, N * K); cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, hA, M * K, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, N * K, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, M * N * 4);
  cudaFuncSetAttribute(k2cta, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  k2cta<<<2, 128, 40 * 1024>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost);
  long bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      int32_t ref = 0;
      for (int k = 0; k < K; ++k) ref += (int)hA[m * K + k] * (int)hB[n * K + k];
      if (ref != hD[m * N + n]) {
        if (bad < 5) printf("mismatch m=%d n=%d got %d want %d\n", m, n, hD[m * N + n], ref);
        ++bad;
      }
    }
  printf("2-CTA MMA 256x256x128: %ld mismatches of %d\n", bad, M * N);
  return bad != 0;
}
