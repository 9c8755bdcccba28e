// Developer microbenchmark: per-SM-sub-partition issue rate of the instructions the attention softmax is made of
// (MUFU.EX2, FFMA2, FADD2, F2FP bf16x2 pack, FMNMX3, FFMA) with 1, 2 and 4 warps per sub-partition, 8 independent
// chains per thread.  Prints cycles per warp-instruction per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/microbench/alu_rate.cu -o tools/microbench/alu_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kIters = 2048;

template <int OP>
__global__ void rate(long long* out, float seed) {
  float a[8];
  unsigned long long p[8];
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; p[i] = (unsigned long long)__float_as_uint(a[i]) << 32 | __float_as_uint(a[i] + 1.f); u[i] = 0; }
  const unsigned long long c2 = ((unsigned long long)__float_as_uint(0.999f) << 32) | __float_as_uint(1.001f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(c2));
      if (OP == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(c2));
      if (OP == 3) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(__uint_as_float(u[i])));
      if (OP == 4) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]), "f"(seed));
      if (OP == 5) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(seed));
      if (OP == 6) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(c2)); asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(__uint_as_float(u[i]))); }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((uint32_t)p[i]) + __uint_as_float(u[i]);
  if (s == 12345.f) out[1] = 1;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int OP> void run(const char* name) {
  long long* d; cudaMalloc(&d, 16);
  for (int warps_per_smsp : {1, 2, 4}) {
    rate<OP><<<148, 128 * warps_per_smsp>>>(d, 0.5f);
    rate<OP><<<148, 128 * warps_per_smsp>>>(d, 0.5f);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    const double n = (double)kIters * 8 * warps_per_smsp * (OP == 6 ? 3 : 1);
    printf("%-22s %d warps/SMSP: %6.2f cycles per warp-instruction\n", name, warps_per_smsp, h / n);
  }
  cudaFree(d);
}

int main() {
  run<0>("MUFU.EX2"); run<1>("FFMA2 (f32x2)"); run<2>("FADD2 (f32x2)"); run<3>("F2FP bf16x2 pack"); run<4>("FMNMX3"); run<5>("FFMA");
  run<6>("mix ex2+add2+f2fp");
  return 0;
}
