// Developer microbenchmark: signalling latencies the attention pipeline is made of.
//   (a) mbarrier ping-pong between two warps of a CTA (arrive -> waiter resumes), try_wait vs test_wait polling
//   (b) one tcgen05.mma (M128 N128 K16) + tcgen05.commit -> the issuing thread's wait returns
//   (c) 8 MMAs + commit -> another warp's wait returns -> it arrives -> the MMA warp's wait returns (the S hand-off)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I mojo_opset_b200/csrc -I include tools/microbench/sync_latency.cu -o tools/microbench/sync_latency
#include <cstdio>
#include <cuda_runtime.h>
#include "tcgen05.cuh"
using namespace mojo;

__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
template <int MODE> __device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
  if (MODE == 0) { while (!mbar_try_wait(bar, parity)) {} }
  else { while (!mbar_test_wait(bar, parity)) {} }
}

constexpr int kRounds = 256;

template <int MODE, bool DYN>
__global__ void __launch_bounds__(256, 1) latency_kernel(long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  fence_async_smem();
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); mbar_fence_init(); }
  if (warp == 1) tmem_alloc(slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(slot);
  constexpr uint32_t idesc = umma_idesc_f16(1, 128, 128, 0, 0);
  const uint32_t a = smem_u32(smem), b = a + 32768;
  // (a) ping-pong: warp 0 <-> warp 4 (different sub-partitions: 0 and 0; use warp 5 -> SMSP 1)
  if (warp == 0) {
    long long t0 = clock64();
    for (int r = 0; r < kRounds; ++r) {
      if (lane == 0) mbar_arrive(&bar[0]);
      wait<MODE>(&bar[1], r & 1);
    }
    long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / (2 * kRounds);
  } else if (warp == 5) {
    for (int r = 0; r < kRounds; ++r) {
      wait<MODE>(&bar[0], r & 1);
      if (lane == 0) mbar_arrive(&bar[1]);
    }
  }
  __syncthreads();
  // (b) one MMA + commit -> own wait
  if (warp == 0) {
    long long acc = 0;
    for (int r = 0; r < kRounds; ++r) {
      long long t0 = clock64();
      umma_ss(tmem, umma_desc_sw128(a, 16, 1024), umma_desc_sw128(b, 16, 1024), idesc, 0);
      umma_commit(&bar[2]);
      wait<MODE>(&bar[2], r & 1);
      acc += clock64() - t0;
    }
    if (lane == 0 && blockIdx.x == 0) out[1] = acc / kRounds;
    // (b') 8 MMAs + commit -> own wait
    acc = 0;
    for (int r = 0; r < kRounds; ++r) {
      long long t0 = clock64();
      for (int ks = 0; ks < 8; ++ks)
        umma_ss(tmem, umma_desc_sw128(a + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                umma_desc_sw128(b + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024), idesc, ks > 0);
      umma_commit(&bar[3]);
      wait<MODE>(&bar[3], r & 1);
      acc += clock64() - t0;
    }
    if (lane == 0 && blockIdx.x == 0) out[2] = acc / kRounds;
  }
  __syncthreads();
  // (c) S hand-off loop: MMA warp: 8 MMAs + commit(bar4); softmax warp 5: wait bar4, tcgen05.ld 128 cols, fence, arrive bar5;
  //     MMA warp waits bar5 then issues the next 8.  Period = the hand-off chain.
  if (warp == 0) {
    long long t0 = clock64();
    volatile int* dyn = reinterpret_cast<volatile int*>(out + 8);  // zeros the compiler cannot see through
    for (int r = 0; r < kRounds; ++r) {
      // DYN: operand bases and the accumulate flag come from memory each round, as the ring stage / tile index do in
      // the attention kernel: the descriptors cannot be hoisted or kept in uniform registers
      const uint32_t ad = DYN ? a + dyn[r & 1] : a, bd = DYN ? b + dyn[(r + 1) & 1] : b, td = DYN ? tmem + dyn[r & 1] : tmem;
      const bool accd = DYN ? dyn[r & 1] != 0 : false;
      for (int ks = 0; ks < 8; ++ks)
        umma_ss(td, umma_desc_sw128(ad + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                umma_desc_sw128(bd + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024), idesc, accd || ks > 0);
      umma_commit(&bar[4]);
      wait<MODE>(&bar[5], r & 1);
      tc_fence_after();
    }
    long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[3] = (t1 - t0) / kRounds;
  } else if (warp >= 4) {  // a "softmax" warpgroup: warps 4..7 own TMEM lanes 0..127
    for (int r = 0; r < kRounds; ++r) {
      wait<MODE>(&bar[4], r & 1);
      tc_fence_after();
      uint32_t sr[128];
      const uint32_t tS = tmem + ((uint32_t)((warp & 3) * 32) << 16);
      for (int q4 = 0; q4 < 4; ++q4) tmem_ld_x32(tS + q4 * 32, sr + q4 * 32);
      tmem_wait_ld();
      tc_fence_before();
      if ((sr[0] ^ sr[37] ^ sr[127]) == 0x12345) out[7] = 1;
      if (warp == 5) { __syncwarp(); if (lane == 0) mbar_arrive(&bar[5]); }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int MODE, bool DYN> void run(const char* name) {
  long long* d; cudaMalloc(&d, 128); cudaMemset(d, 0, 128);
  auto kern = latency_kernel<MODE, DYN>;
  const int smem = 1024 + 65536 + 128;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int r = 0; r < 2; ++r) kern<<<2, 256, smem>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
  printf("%s: arrive->waiter resumes %lld | 1 MMA+commit->wait %lld | 8 MMA+commit->wait %lld | S hand-off period (8 MMA, commit, ld 128 cols, arrive) %lld  (%s)\n",
         name, h[0], h[1], h[2], h[3], cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<0, false>("try_wait, static operands ");
  run<0, true>("try_wait, dynamic operands");
  return 0;
}
