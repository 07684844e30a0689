// Instruction-throughput microbenchmark for the epilogue / softmax building blocks on sm_100a: warp-instructions per clock per SM for
// each op with 16 resident warps per SM and 8 independent chains per thread (so the figure is the pipe's issue rate, not a latency).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/bench_ops tools/bench_ops.cu && tools/bin/bench_ops
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

constexpr int kIters = 512, kChains = 8;

template <int OP>
__global__ void __launch_bounds__(512) k(float* out, int seed, long long* cycles, unsigned long long ka, unsigned long long kb) {
  float f[kChains];
  unsigned long long d[kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c) d[c] = ka + c + seed;
  int i[kChains];
  unsigned p = 0;
#pragma unroll
  for (int c = 0; c < kChains; ++c) { f[c] = 1.0f + 0.001f * (threadIdx.x + c + seed); i[c] = threadIdx.x * 7 + c + seed; }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int c = 0; c < kChains; ++c) {
      if (OP == 0) f[c] = __fmaf_rn(f[c], 1.0001f, 0.5f);                        // FFMA
      if (OP == 1) f[c] = __fadd_rn(f[c], 1.5f);                                 // FADD
      if (OP == 2) { f[c] = __int2float_rn(i[c]); i[c] += __float_as_int(f[c]) & 3; }   // I2F (+ 2 int ops)
      if (OP == 3) { asm volatile("{.reg .pred q; setp.lt.f32 q, %1, %2; selp.u32 %0, 1, %0, q;}" : "+r"(p) : "f"(f[c]), "f"(2.0f)); f[c] += 1.0f; }   // FSETP (+ SEL + FADD)
      if (OP == 4) f[c] = fmaxf(f[c] * 1.0001f, 0.5f);                           // FMNMX (+ FMUL)
      if (OP == 5) i[c] = __byte_perm(i[c], i[(c + 1) % kChains], 0x5140);       // PRMT
      if (OP == 6) i[c] = (i[c] ^ 0x5a5a5a5a) + 12345;                           // LOP3 + IADD
      if (OP == 7) f[c] = __int_as_float(0x4B400000 + (i[c] & 0xFFFFF)) - 12582912.0f + f[c];   // magic int->float (LOP, IADD, FADD, FADD)
      if (OP == 8) f[c] = rintf(f[c] * 1.37f);                                   // FRND (+ FMUL)
      if (OP == 9) { i[c] = __float2int_rn(f[c]); f[c] = f[c] * 1.0001f + (i[c] & 1); }   // F2I (+ FFMA, LOP, I2F?)
      if (OP == 10) f[c] = __expf(f[c] * 0.001f);                                // MUFU.EX2 (+ 2 FMUL)
      if (OP == 11) f[c] = __frcp_rn(f[c] + 1.0f);                               // MUFU.RCP + fixup
      if (OP == 12) f[c] = expf(f[c] * 0.001f);                                  // accurate expf
      if (OP == 13) f[c] = __fdiv_rn(f[c], 1.0001f + f[(c + 1) % kChains] * 1e-9f);   // IEEE division
      if (OP == 14) { __half2 h = __floats2half2_rn(f[c], f[(c + 1) % kChains]); i[c] ^= *reinterpret_cast<int*>(&h); f[c] += 1.0f; }   // F2F.F16x2 pack
      if (OP == 15) f[c] = __fmul_rn(__fmul_rn(f[c], 1.0001f), 0.9999f);         // 2 x FMUL
      if (OP == 16) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[c]) : "l"(ka), "l"(kb));   // FFMA2: two fp32 FMAs per lane
      if (OP == 17) f[c] = fmaxf(fmaxf(f[c], fabsf(f[(c + 1) % kChains])), fabsf(f[(c + 3) % kChains]));   // FMNMX3
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
  int iacc = p;
#pragma unroll
  for (int c = 0; c < kChains; ++c) { acc += f[c]; iacc += i[c] + static_cast<int>(d[c] >> 3); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + iacc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, float inst_per_chain_step) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&cyc, 148 * 8);
  k<OP><<<148, 512>>>(out, 1, cyc, 0x3F8000013F800001ull, 0x3F0000003F000000ull);
  k<OP><<<148, 512>>>(out, 2, cyc, 0x3F8000013F800001ull, 0x3F0000003F000000ull);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0;
  for (int b = 0; b < 148; ++b) c += h[b];
  c /= 148;
  const double steps = 16.0 * kIters * kChains;   // warp-level chain steps per SM
  printf("%-40s %8.0f cycles  %6.3f chain-steps/clk/SM  (%5.2f clk per warp-step; ~%.1f instr per step)\n", name, c, steps / c, c / steps, inst_per_chain_step);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("FFMA", 1);
  run<16>("FFMA2 (fma.rn.f32x2)", 1);
  run<17>("FMNMX3", 1);
  run<1>("FADD", 1);
  run<15>("2 x FMUL", 2);
  run<2>("I2F.S32 + LOP + IADD", 3);
  run<7>("magic int->float (LOP,IADD,FADD,FADD)", 4);
  run<3>("FSETP + SEL + FADD", 3);
  run<4>("FMNMX + FMUL", 2);
  run<5>("PRMT", 1);
  run<6>("LOP3 + IADD", 2);
  run<8>("FRND + FMUL", 2);
  run<9>("F2I + FFMA + LOP + I2F", 4);
  run<10>("__expf (FMUL, FMUL, MUFU.EX2)", 3);
  run<11>("__frcp_rn", 3);
  run<12>("expf (accurate)", 8);
  run<13>("__fdiv_rn", 10);
  run<14>("F2F.F16x2 pack + LOP + FADD", 3);
  return 0;
}
