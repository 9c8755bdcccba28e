// Developer micro-benchmark: which launch shape / load flavour streams a 3-stream bf16 elementwise op
// (out = silu(g) * u) fastest on B200.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a stream3.cu -o stream3
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

struct alignas(16) V { __nv_bfloat16 v[8]; };

template <int MODE> __device__ __forceinline__ V ldv(const V* p) {
  int4 r;
  if (MODE == 0) r = *reinterpret_cast<const int4*>(p);
  else if (MODE == 1) asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  else asm volatile("ld.global.cs.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return *reinterpret_cast<V*>(&r);
}
template <int MODE> __device__ __forceinline__ void stv(V* p, const V& v) {
  const int4 r = *reinterpret_cast<const int4*>(&v);
  if (MODE == 0) *reinterpret_cast<int4*>(p) = r;
  else asm volatile("st.global.cs.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w));
}
__device__ __forceinline__ V compute(const V& g, const V& u, bool heavy) {
  V o;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float gf = __bfloat162float(g.v[e]), uf = __bfloat162float(u.v[e]);
    float s;
    if (heavy) {
      float ex, r;
      asm("ex2.approx.f32 %0, %1;" : "=f"(ex) : "f"(gf * -1.4426950408889634f));
      asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(1.0f + ex));
      s = __bfloat162float(__float2bfloat16_rn(gf * r));
    } else {
      s = gf;
    }
    o.v[e] = __float2bfloat16_rn(s * uf);
  }
  return o;
}

// each thread: UN vectors at tid + k*blockDim inside the CTA's contiguous chunk; CTAs loop grid-stride over chunks
template <int UN, int LD, int ST>
__global__ void k_chunk(const V* __restrict__ g, const V* __restrict__ u, V* __restrict__ o, long n, int heavy) {
  const long chunk = (long)blockDim.x * UN;
  for (long base = (long)blockIdx.x * chunk; base < n; base += (long)gridDim.x * chunk) {
    V gv[UN], uv[UN];
#pragma unroll
    for (int k = 0; k < UN; ++k) {
      const long v = base + threadIdx.x + (long)k * blockDim.x;
      if (v < n) { gv[k] = ldv<LD>(g + v); uv[k] = ldv<LD>(u + v); }
    }
#pragma unroll
    for (int k = 0; k < UN; ++k) {
      const long v = base + threadIdx.x + (long)k * blockDim.x;
      if (v < n) stv<ST>(o + v, compute(gv[k], uv[k], heavy));
    }
  }
}

int main() {
  const long n = 8192L * 12288 / 8;  // vectors
  const int sets = 3;
  std::vector<V*> G(sets), U(sets), O(sets);
  for (int i = 0; i < sets; ++i) {
    cudaMalloc(&G[i], n * 16); cudaMalloc(&U[i], n * 16); cudaMalloc(&O[i], n * 16);
    cudaMemset(G[i], 0x3c, n * 16); cudaMemset(U[i], 0x3c, n * 16);
  }
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  auto run = [&](const char* name, auto launch) {
    for (int i = 0; i < 3; ++i) launch(i % sets);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    const int iters = 12;
    for (int i = 0; i < iters; ++i) launch(i % sets);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double us = ms * 1e3 / iters;
    printf("%-44s %8.1f us %7.0f GB/s  %s\n", name, us, 3.0 * n * 16 / us / 1e3, cudaGetErrorString(cudaGetLastError()));
  };
#define RUN(UN, LD, ST, TH, GRIDEXPR, HEAVY)                                                                  \
  {                                                                                                           \
    char nm[128];                                                                                             \
    const long full = (n + (long)TH * UN - 1) / ((long)TH * UN);                                              \
    const long grid = (GRIDEXPR) > 0 ? ((GRIDEXPR) < full ? (GRIDEXPR) : full) : full;                        \
    snprintf(nm, sizeof nm, "un%d ld%d st%d th%d grid%ld heavy%d", UN, LD, ST, TH, grid, HEAVY);              \
    run(nm, [&](int s) { k_chunk<UN, LD, ST><<<(unsigned)grid, TH>>>(G[s], U[s], O[s], n, HEAVY); });          \
  }
  // copy-like memcpy baseline
  run("cudaMemcpyAsync d2d (2 streams of bytes)", [&](int s) { cudaMemcpyAsync(O[s], G[s], n * 16, cudaMemcpyDeviceToDevice); });
  for (int heavy = 0; heavy < 2; ++heavy) {
    RUN(1, 0, 0, 256, 0, heavy) RUN(2, 0, 0, 256, 0, heavy) RUN(4, 0, 0, 256, 0, heavy) RUN(8, 0, 0, 256, 0, heavy)
    RUN(4, 0, 0, 128, 0, heavy) RUN(4, 0, 0, 512, 0, heavy)
    RUN(4, 1, 0, 256, 0, heavy) RUN(4, 2, 0, 256, 0, heavy) RUN(4, 1, 1, 256, 0, heavy) RUN(4, 0, 1, 256, 0, heavy)
    RUN(2, 0, 0, 256, 148 * 8, heavy) RUN(4, 0, 0, 256, 148 * 4, heavy) RUN(4, 0, 0, 256, 148 * 8, heavy)
    RUN(4, 0, 0, 256, 148 * 16, heavy) RUN(2, 1, 0, 256, 148 * 8, heavy) RUN(4, 0, 0, 512, 148 * 2, heavy)
    RUN(4, 0, 0, 512, 148 * 4, heavy) RUN(8, 0, 0, 256, 148 * 4, heavy)
  }
  return 0;
}
