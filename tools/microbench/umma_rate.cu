// Developer microbenchmark: issue-to-completion rate of tcgen05.mma M128/M256(pair) N128 K16 bf16 in the operand forms the
// attention kernel uses (SS K-major x K-major = QK; TS with MN-major B = PV; SS with MN-major B = PV with P in smem),
// single CTA and cta_group::2.  One CTA (pair) per launch slot; prints cycles per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I mojo_opset_b200/csrc -I include tools/microbench/umma_rate.cu -o tools/microbench/umma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "tcgen05.cuh"
using namespace mojo;

constexpr int kIters = 64;  // groups of 8 MMAs

template <bool PAIR, int FORM, int SPIN = 0>  // SPIN: 1 = every lane of the extra warps polls an mbarrier, 2 = lane 0 only
// FORM 0: SS QK-like, 1: TS PV-like, 2: SS PV-like (A K-major smem, B MN-major)
__global__ void __launch_bounds__(384, 1) rate_kernel(long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;            // 32 KB [half][128 rows][128 B]
  uint8_t* sB = smem + 32768;    // 32 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  fence_async_smem();
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
  if (warp == 1) { if (PAIR) tmem_alloc_pair(slot, 512); else tmem_alloc(slot, 512); }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(slot);
  volatile uint32_t* done = slot + 1;
  if (threadIdx.x == 0) *done = 0;
  __syncthreads();
  if (warp >= 4) {  // spinner warps (only launched when SPIN != 0)
    if (SPIN == 1 || (threadIdx.x & 31) == 0) {
      while (!*done) { if (mbar_try_wait(&bar[1], 1)) break; }
    }
    __syncwarp();
  }
  if (warp == 0 && rank == 0) {
    constexpr int kM = PAIR ? 256 : 128;
    constexpr uint32_t idesc_qk = umma_idesc_f16(1, kM, 128, 0, 0);
    constexpr uint32_t idesc_pv = umma_idesc_f16(1, kM, 128, 0, 1);
    constexpr uint32_t kKHalf = PAIR ? 8192 : 16384;
    const uint32_t a = smem_u32(sA), b = smem_u32(sB);
    long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
      if (FORM == 6) tc_fence_after();
      if (FORM == 7) { mbar_try_wait(&bar[1], 1); tc_fence_after(); }
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint32_t oq = (uint32_t)(ks >> 2) * 16384u + (uint32_t)(ks & 3) * 32u;
        const uint32_t ok = (uint32_t)(ks >> 2) * kKHalf + (uint32_t)(ks & 3) * 32u;
        const uint32_t d = tmem + (it & 1) * 128;
        if (FORM == 0) {
          if (PAIR) umma_ss_pair(d, umma_desc_sw128(a + oq, 16, 1024), umma_desc_sw128(b + ok, 16, 1024), idesc_qk, ks > 0);
          else umma_ss(d, umma_desc_sw128(a + oq, 16, 1024), umma_desc_sw128(b + ok, 16, 1024), idesc_qk, ks > 0);
        } else if (FORM == 1) {
          if (PAIR) umma_ts_pair(d, tmem + 256 + ks * 8, umma_desc_sw128(b + ks * 2048u, 16384, 1024), idesc_pv, ks > 0);
          else umma_ts(d, tmem + 256 + ks * 8, umma_desc_sw128(b + ks * 2048u, 16384, 1024), idesc_pv, ks > 0);
        } else if (FORM == 2) {
          if (PAIR) umma_ss_pair(d, umma_desc_sw128(a + oq, 16, 1024), umma_desc_sw128(b + ks * 2048u, 16384, 1024), idesc_pv, ks > 0);
          else umma_ss(d, umma_desc_sw128(a + oq, 16, 1024), umma_desc_sw128(b + ks * 2048u, 16384, 1024), idesc_pv, ks > 0);
        } else {  // FORM >= 3: attention-like: even groups QK -> S (col 0), odd groups PV: A = P_t (col 128 + 64 t), D = O_t (col 256 + 128 t)
          const int t = (it >> 1) & 1;
          if ((it & 1) == 0) {
            if (PAIR) umma_ss_pair(tmem, umma_desc_sw128(a + oq, 16, 1024), umma_desc_sw128(b + ok, 16, 1024), idesc_qk, ks > 0);
            else umma_ss(tmem, umma_desc_sw128(a + oq, 16, 1024), umma_desc_sw128(b + ok, 16, 1024), idesc_qk, ks > 0);
          } else {
            if (PAIR) umma_ts_pair(tmem + 256 + t * 128, tmem + 128 + t * 64 + ks * 8, umma_desc_sw128(b + ks * 2048u, 16384, 1024), idesc_pv, 1);
            else umma_ts(tmem + 256 + t * 128, tmem + 128 + t * 64 + ks * 8, umma_desc_sw128(b + ks * 2048u, 16384, 1024), idesc_pv, 1);
          }
        }
      }
      if (FORM == 4) { if (PAIR) umma_commit_pair(&bar[1]); else umma_commit(&bar[1]); }
      if (FORM == 5) { if (PAIR) { umma_commit_pair(&bar[1]); umma_commit_pair(&bar[1]); } else { umma_commit(&bar[1]); umma_commit(&bar[1]); } }
    }
    if (PAIR) umma_commit_pair(&bar[0]); else umma_commit(&bar[0]);
    mbar_wait_bounded(&bar[0], 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  if (warp == 0) { __syncwarp(); *done = 1; }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  if (warp == 1) { tc_fence_after(); if (PAIR) tmem_dealloc_pair(tmem, 512); else tmem_dealloc(tmem, 512); }
}

template <bool PAIR, int FORM, int SPIN = 0> void run(const char* name, int ctas) {
  long long* d; cudaMalloc(&d, 8); cudaMemset(d, 0, 8);
  auto kern = rate_kernel<PAIR, FORM, SPIN>;
  const int smem = 1024 + 65536 + 64;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(SPIN ? 384 : 128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = PAIR ? 1 : 0;
  for (int r = 0; r < 3; ++r) cudaLaunchKernelEx(&cfg, kern, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-28s ctas %3d: %7.1f cycles / MMA (%s)\n", name, ctas, (double)h / (kIters * 8), cudaGetErrorString(e));
  cudaFree(d);
}

static void cluster_occupancy() {
  // how many 2-CTA clusters of an attention-sized CTA (1 per SM by shared memory) can be resident at once
  auto kern = rate_kernel<true, 0>;
  const int smem = 198144;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(296); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = -1;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  printf("max active 2-CTA clusters at 198 KB smem/CTA: %d (%s); SMs %d\n", n, cudaGetErrorString(e), prop.multiProcessorCount);
}

int main() {
  cluster_occupancy();
  for (int ctas : {2}) {
    run<false, 0>("SS  QK  single M128", ctas);
    run<true, 0>("SS  QK  pair   M256", ctas);
    run<false, 1>("TS  PV  single M128", ctas);
    run<true, 1>("TS  PV  pair   M256", ctas);
    run<false, 2>("SS  PV  single M128", ctas);
    run<true, 2>("SS  PV  pair   M256", ctas);
    run<false, 3>("QK/PV alternate single", ctas);
    run<true, 3>("QK/PV alternate pair", ctas);
    run<false, 4>("alt + 1 commit/8 single", ctas);
    run<true, 4>("alt + 1 commit/8 pair", ctas);
    run<false, 5>("alt + 2 commits/8 single", ctas);
    run<true, 5>("alt + 2 commits/8 pair", ctas);
    run<false, 6>("alt + fence::after /8 single", ctas);
    run<true, 6>("alt + fence::after /8 pair", ctas);
    run<false, 7>("alt + try_wait+fence /8", ctas);
    run<false, 3, 1>("alt, 8 warps all lanes poll", ctas);
    run<true, 3, 1>("alt pair, 8 warps all poll", ctas);
    run<false, 3, 2>("alt, 8 warps lane0 polls", ctas);
    run<true, 3, 2>("alt pair, 8 warps lane0 poll", ctas);
  }
  return 0;
}
