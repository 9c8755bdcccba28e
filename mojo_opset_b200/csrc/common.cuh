// Shared helpers for the sm_100a kernels behind include/mojo_b200.h.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <utility>

#include "../../include/mojo_b200.h"

namespace mojo {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// ---- error plumbing ------------------------------------------------------------------------------
char* last_error_buffer();  // thread-local, defined in api.cu
int fail(int code, const char* fmt, ...);
int* error_word();  // the current device's registered error word (api.cu), or null
int* decode_tickets(int64_t* count);  // the current device's registered split-KV arrival counters (api.cu), or null

#define MOJO_REQUIRE(cond, code, ...)                 \
  do {                                                \
    if (!(cond)) return ::mojo::fail((code), __VA_ARGS__); \
  } while (0)

#define MOJO_CUDA_OK(expr)                                                              \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess)                                                             \
      return ::mojo::fail((int)e__, "%s failed: %s", #expr, cudaGetErrorString(e__));   \
  } while (0)

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, "%s launch failed: %s", what, cudaGetErrorString(e));
  return 0;
}

// ---- dtype traits ----------------------------------------------------------------------------------
template <typename T> struct DType;
template <> struct DType<__nv_bfloat16> {
  static constexpr int id = MOJO_B200_BF16;
  __device__ __forceinline__ static float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ __forceinline__ static __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};
template <> struct DType<__half> {
  static constexpr int id = MOJO_B200_F16;
  __device__ __forceinline__ static float to_f(__half v) { return __half2float(v); }
  __device__ __forceinline__ static __half from_f(float v) { return __float2half_rn(v); }
};
template <> struct DType<float> {
  static constexpr int id = MOJO_B200_F32;
  __device__ __forceinline__ static float to_f(float v) { return v; }
  __device__ __forceinline__ static float from_f(float v) { return v; }
};

inline int dtype_bytes(int dtype) { return dtype == MOJO_B200_F32 ? 4 : 2; }

// Round an fp32 value through T and back (emulates an eager op that materialises its result in T).
template <typename T> __device__ __forceinline__ float round_through(float v) {
  return DType<T>::to_f(DType<T>::from_f(v));
}

// ---- 128-bit vectors -------------------------------------------------------------------------------
template <typename T> struct alignas(16) Vec16 {
  static constexpr int N = 16 / sizeof(T);
  T v[N];
};

template <typename T> __device__ __forceinline__ Vec16<T> ld_vec(const T* p) {
  return *reinterpret_cast<const Vec16<T>*>(p);
}
// streaming (read-once) load: bypass L1 allocation
template <typename T> __device__ __forceinline__ Vec16<T> ld_vec_stream(const T* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return *reinterpret_cast<Vec16<T>*>(&r);
}
template <typename T> __device__ __forceinline__ void st_vec(T* p, const Vec16<T>& v) {
  *reinterpret_cast<Vec16<T>*>(p) = v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------
// The path is a chain of short dependent kernels (norm -> RoPE -> store -> decode -> SwiGLU / o_proj): with a plain
// launch every link pays the full drain + launch latency of its predecessor.  Kernels launched through launch_pdl()
// carry cudaLaunchAttributeProgrammaticStreamSerialization: their CTAs may become resident while the previous kernel's
// last wave drains and run their prologue (barrier init, descriptor / weight prefetch); pdl_wait() blocks until the
// previous kernel has COMPLETED and its writes are visible, so every such kernel calls it before it touches any
// memory another kernel may have produced (completion is transitive along the chain: each link waits itself).
// pdl_trigger() lets the NEXT kernel's CTAs be scheduled; it sits after the wait so at most one successor is staged.
// MOJO_B200_PDL=0 turns the attribute off (plain stream order; the device-side calls are then no-ops).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();  // api.cu

template <typename... P, typename... A>
inline cudaError_t launch_pdl(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, A&&... args) {
  cudaLaunchConfig_t cfg;
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<A>(args)...);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename F> inline int dispatch_dtype(int dtype, F&& f) {
  switch (dtype) {
    case MOJO_B200_BF16: return f(__nv_bfloat16{});
    case MOJO_B200_F16: return f(__half{});
    case MOJO_B200_F32: return f(float{});
    default: return fail(MOJO_B200_EINVAL, "unknown dtype id %d", dtype);
  }
}

}  // namespace mojo
