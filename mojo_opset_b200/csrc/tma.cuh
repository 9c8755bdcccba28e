// mbarrier / TMA (cp.async.bulk.tensor) / ldmatrix / mma.sync PTX wrappers and the host-side tensor-map
// cache.  sm_100a only.
#pragma once

#include <cuda.h>  // CUtensorMap + enums only; the driver entry point is resolved through the runtime

#include "common.cuh"

namespace mojo {

// ---- shared-memory addresses -----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier --------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---- TMA -------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2,
                                             int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (TMA store / UMMA reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 128-byte swizzle (Swizzle<3,4,3>): a tile is a stack of 128-byte lines; 16-byte chunk c of line l lives at
// chunk (c ^ (l & 7)).  Tile bases must be 1024-byte aligned.
__device__ __forceinline__ uint32_t swz128(uint32_t line, uint32_t chunk) {
  return line * 128u + ((chunk ^ (line & 7u)) << 4);
}

// ---- ldmatrix / mma.sync (legacy tensor path: right-sized for the M=16 decode tile) -----------------
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
template <typename T> struct Mma16816;
template <> struct Mma16816<__nv_bfloat16> {
  __device__ __forceinline__ static void run(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  __device__ __forceinline__ static uint32_t pack(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
};
template <> struct Mma16816<__half> {
  __device__ __forceinline__ static void run(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  __device__ __forceinline__ static uint32_t pack(float lo, float hi) {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
};

// ---- host: tensor-map encode + cache ------------------------------------------------------------------
struct TensorMapKey {
  const void* base;
  int rank;
  uint64_t dims[5];
  uint64_t strides[4];  // bytes, dims 1..rank-1
  uint32_t box[5];
  int dtype;
  int swizzle;  // CUtensorMapSwizzle
  bool operator==(const TensorMapKey& o) const;
};
// Returns 0 and fills *out (a copy of a cached 128-byte descriptor) or an error code (+ last_error).
int get_tensor_map(const TensorMapKey& key, CUtensorMap* out);

// KV tiles are staged 64 tokens at a time by every attention kernel.
constexpr int kTile = 64;

// 5-D tensor map over a [blocks, heads, tokens, D] cache (element strides s_b, s_h, s_t; D contiguous, 16-bit
// elements, D a multiple of 64) whose box is `box_rows` token rows of one (block, head); rows past the end
// of a block are zero filled by the TMA unit:
//   split_halves = false: dims [64 | D/64 | token | head | block] -> a box lands as [token][half][128 B]
//   split_halves = true : dims [64 | token | D/64 | head | block] -> a box lands as [half][token][128 B]
// Both use the 128-byte swizzle; the second keeps ldmatrix conflict free for D = 128.
// A dense [B, H, S, D] tensor is the same thing with block = batch and block_size = S.
int build_cache_map(const void* base, int dtype, int head_dim, int64_t block_size, int num_kv_heads,
                    int64_t num_blocks, int64_t s_b, int64_t s_h, int64_t s_t, bool split_halves, int box_rows,
                    CUtensorMap* out);


// 5-D tensor map for the tcgen05 attention kernel (head_dim 64 * halves, 16-bit elements) over a [blocks, heads, rows, D]
// tensor (element strides s_b, s_h, s_t): dims [64 | row | half | head | block], box = `box_rows` rows of ONE
// 64-column half, 128-byte swizzle -> a box lands as [row][128 B], the K-major line layout UMMA reads.
int build_tile_map(const void* base, int dtype, int64_t rows_per_block, int num_heads, int64_t num_blocks, int64_t s_b,
                   int64_t s_h, int64_t s_t, int box_rows, CUtensorMap* out, int halves = 2);

}  // namespace mojo
