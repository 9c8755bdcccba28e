// Developer microbenchmark: throughput of the attention softmax inner loop ALONE - 8 warps per SM (two per scheduler,
// as in attn_fwd_sm100_kernel), each thread one score row of 128 columns: tcgen05.ld S row -> [round to bf16] -> row max
// -> exp2 (MUFU and/or FMA-pipe polynomial) -> sums -> pack -> tcgen05.st P.  No MMA, no barriers: what is measured is
// the issue / pipe cost of a (tile 0, tile 1) step, to compare with the 2048 cycles the tensor pipe needs for it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I mojo_opset_b200/csrc -I include \
//        tools/microbench/softmax_rate.cu -o tools/microbench/softmax_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "tcgen05.cuh"

using namespace mojo;

constexpr int kIters = 512;

__device__ __forceinline__ float fma_sat(float a, float b, float c) {
  float d;
  asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  float2 d;
  asm("sub.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<uint64_t&>(d))
      : "l"(reinterpret_cast<const uint64_t&>(a)), "l"(reinterpret_cast<const uint64_t&>(b)));
  return d;
}

// exp2 of a pair on the FMA pipe.  y = sat((x + 125) / 256) is computed by the caller straight from the score
// (one FFMA.SAT: the clamp to x >= -125 is free and masked (-inf) scores land on 2^-125).
//   t = 256 y + K           (K = 1.5 * 2^23 - 125: t = magic + n, n = round(x) in the low mantissa bits)
//   f = 256 y - (n + 125)   in [-0.5, 0.5]
//   2^f ~ degree-3 polynomial, 2^n by adding n << 23 to the bit pattern
template <int DEG>
__device__ __forceinline__ float2 ex2_fma_pipe(float2 y) {
  constexpr float kK = 12582912.f - 125.f;
  const float2 t = fma2(y, make_float2(256.f, 256.f), make_float2(kK, kK));
  const float2 mneg = sub2(make_float2(kK, kK), t);
  const float2 f = fma2(y, make_float2(256.f, 256.f), mneg);
  float2 q;
  if (DEG == 3) {
    q = fma2(f, make_float2(0.0551716648f, 0.0551716648f), make_float2(0.2426111251f, 0.2426111251f));
    q = fma2(q, f, make_float2(0.6932609677f, 0.6932609677f));
    q = fma2(q, f, make_float2(0.9999280572f, 0.9999280572f));
  } else {
    q = fma2(f, make_float2(0.23563174f, 0.23563174f), make_float2(0.69786238f, 0.69786238f));
    q = fma2(q, f, make_float2(1.00020933f, 1.00020933f));
  }
  float2 out;
  out.x = __uint_as_float(__float_as_uint(q.x) + (__float_as_uint(t.x) << 23));
  out.y = __uint_as_float(__float_as_uint(q.y) + (__float_as_uint(t.y) << 23));
  return out;
}

// V = 0: the round-1 loop (MUFU only, optional old emulation share OLD_EMU of 4)
// V = 1: new loop: E of every 8 pairs on the FMA pipe (ex2_fma_pipe), the rest on the MUFU
template <int V, bool ROUND_S, int E, int DEG>
__global__ void __launch_bounds__(256, 1) softmax_rate(long long* out, float scale_log2, float base_in, uint32_t* sink) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const int t = warp >> 2;
  const uint32_t tS = tmem + lane_base;
  const uint32_t tP = tmem + lane_base + 128u + (uint32_t)t * 64u;
  {  // scores ~ N(0, 11): what Q K^T of unit-variance 128-wide rows looks like
    uint32_t init[128];
    uint32_t h = threadIdx.x * 2654435761u + blockIdx.x;
#pragma unroll
    for (int c = 0; c < 128; ++c) {
      h = h * 1664525u + 1013904223u;
      const float u = (float)(h >> 8) * (1.f / 16777216.f) - 0.5f;
      init[c] = __float_as_uint(u * 40.f);
    }
    if (t == 0) {
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) tmem_st_x32(tS + q4 * 32, init + q4 * 32);
      tmem_wait_st();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  float l = 0.f, m_ref = base_in;
  uint32_t acc = 0;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < kIters; ++it) {
    uint32_t sr[128];
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) tmem_ld_x32(tS + q4 * 32, sr + q4 * 32);
    tmem_wait_ld();
    if (ROUND_S) {
#pragma unroll
      for (int c = 0; c < 128; ++c)
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(sr[c]) : "f"(__uint_as_float(sr[c])), "f"(0.f));
    }
    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
    for (int c = 0; c < 128; c += 4) {
      mx0 = fmaxf(mx0, __uint_as_float(sr[c]));
      mx1 = fmaxf(mx1, __uint_as_float(sr[c + 1]));
      mx2 = fmaxf(mx2, __uint_as_float(sr[c + 2]));
      mx3 = fmaxf(mx3, __uint_as_float(sr[c + 3]));
    }
    const float mt = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2;
    if (mt > m_ref + 8.f) m_ref = mt;  // (never taken with the inputs above; keeps the dependence on the max)
    const float base = m_ref;
    const float2 scale2 = make_float2(scale_log2, scale_log2), nbase2 = make_float2(-base, -base);
    const float sc256 = scale_log2 * (1.f / 256.f), b256 = (125.f - base) * (1.f / 256.f);
    float2 sum_a = make_float2(0.f, 0.f), sum_b = make_float2(0.f, 0.f);
    float2 sum_c = make_float2(0.f, 0.f), sum_d = make_float2(0.f, 0.f);
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 8) {
      float2 x[8], e[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float s0 = __uint_as_float(sr[2 * (c0 + i)]), s1 = __uint_as_float(sr[2 * (c0 + i) + 1]);
        if (V == 1 && i < E) x[i] = make_float2(fma_sat(s0, sc256, b256), fma_sat(s1, sc256, b256));
        else x[i] = fma2(make_float2(s0, s1), scale2, nbase2);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (V == 1 && i < E) {
          e[i] = ex2_fma_pipe<DEG>(x[i]);
        } else if (V == 0 && (i & 3) < E) {
          e[i] = ex2_emulated2(x[i]);
        } else {
          e[i].x = ex2_approx(x[i].x);
          e[i].y = ex2_approx(x[i].y);
        }
      }
      sum_a = add2(sum_a, e[0]); sum_b = add2(sum_b, e[1]); sum_c = add2(sum_c, e[2]); sum_d = add2(sum_d, e[3]);
      sum_a = add2(sum_a, e[4]); sum_b = add2(sum_b, e[5]); sum_c = add2(sum_c, e[6]); sum_d = add2(sum_d, e[7]);
#pragma unroll
      for (int i = 0; i < 8; ++i) sr[c0 + i] = pack2<__nv_bfloat16>(e[i].x, e[i].y);
    }
    sum_a = add2(sum_a, sum_c);
    sum_b = add2(sum_b, sum_d);
    l += (sum_a.x + sum_a.y) + (sum_b.x + sum_b.y);
    tmem_st_x32(tP, sr);
    tmem_st_x32(tP + 32, sr + 32);
    tmem_wait_st();
    acc ^= sr[0] ^ sr[63];
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (l == 12345.f || acc == 0x12345u) sink[threadIdx.x] = acc;
  if (blockIdx.x == 0 && threadIdx.x < 4) {  // accuracy probe of the FMA-pipe exponential against exp2f
    float worst = 0.f;
    for (int k = 0; k < 20000; ++k) {
      const float xx = -60.f + 68.f * (float)k / 20000.f + 0.001f * threadIdx.x;
      const float y = fma_sat(xx, 1.f / 256.f, 125.f / 256.f);
      const float2 r = ex2_fma_pipe<DEG>(make_float2(y, y));
      const float ref = exp2f(xx);
      worst = fmaxf(worst, fabsf(r.x - ref) / ref);
    }
    reinterpret_cast<float*>(out)[4 + threadIdx.x] = worst;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int V, bool ROUND_S, int E, int DEG>
void run(const char* name) {
  long long* d;
  uint32_t* sink;
  cudaMalloc(&d, 64);
  cudaMalloc(&sink, 4096);
  cudaMemset(d, 0, 64);
  for (int rep = 0; rep < 2; ++rep) softmax_rate<V, ROUND_S, E, DEG><<<148, 256>>>(d, 0.1275f, 2.0f, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[8];
  cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  printf("%-44s %8.1f cycles per (tile0, tile1) step   [fma-pipe exp2 max rel err %.2e]  %s\n", name,
         (double)h[0] / kIters, reinterpret_cast<float*>(h)[4], e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
  cudaFree(sink);
}

int main() {
  printf("tensor-pipe budget per step: 2048 cycles (4 x M128 N128 K128)\n");
  run<0, false, 0, 3>("r1 loop, no rounding, MUFU only");
  run<0, true, 0, 3>("r1 loop, bf16 rounding, MUFU only");
  run<0, true, 1, 3>("r1 loop, rounding, old emulation 1/4");
  run<1, false, 2, 3>("new, no rounding, 2/8 pairs FMA deg3");
  run<1, false, 3, 3>("new, no rounding, 3/8 pairs FMA deg3");
  run<1, false, 4, 3>("new, no rounding, 4/8 pairs FMA deg3");
  run<1, false, 5, 3>("new, no rounding, 5/8 pairs FMA deg3");
  run<1, true, 2, 3>("new, rounding, 2/8 pairs FMA deg3");
  run<1, true, 3, 3>("new, rounding, 3/8 pairs FMA deg3");
  run<1, true, 4, 3>("new, rounding, 4/8 pairs FMA deg3");
  run<1, true, 5, 3>("new, rounding, 5/8 pairs FMA deg3");
  run<1, true, 4, 2>("new, rounding, 4/8 pairs FMA deg2");
  run<1, true, 8, 3>("new, rounding, 8/8 pairs FMA deg3");
  return 0;
}
