// MojoSwiGLU / MojoSilu: flat element-wise, 128-bit accesses, one HBM pass (3 resp. 2 streams).
//
//   silu(g) = round_T(g / (1 + exp(-g)))            (fp32 math, IEEE division, accurate expf)
//   swiglu  = round_T(silu(g) * u)                  (second rounding, as the eager golden: F.silu(g) * u)
//
// Rows may be strided (gate / up are commonly the two halves of one fused projection).  Grid: 2-D,
// x over 16-byte vectors of a row, y over rows; one vector per stream per thread at 32 registers (full occupancy).
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace mojo {

// fp32 / fp16 tensors: IEEE division and accurate expf.  bf16 tensors: the result is rounded to 8 significand bits
// right away, so the MUFU forms (ex2.approx, rcp.approx; ~1e-6 relative) give the same rounded value except on
// a ~5e-4 fraction of rounding boundaries (then 1 ulp of bf16) - and keep the kernel HBM-bound instead of ALU-bound
// (the accurate form costs ~40 instructions per element: 110 us of issue time at 8192 x 12288).
template <typename T> __device__ __forceinline__ float silu_f(float g) {
  if constexpr (!std::is_same<T, __nv_bfloat16>::value) {
    return __fdiv_rn(g, __fadd_rn(1.0f, expf(-g)));
  } else {
    float e, r;
    // .ftz: one MUFU each, no denormal range fix-ups (only |g| > 87 differs: -0 instead of ~-1e-37)
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(g * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return g * r;
  }
}

// torch.clamp in the input dtype: NaN propagates, and a bound that is not representable in T is rounded
// when it replaces a value (the eager op materialises its result in T).
template <typename T> __device__ __forceinline__ void clamp_pair(float& gf, float& uf, float limit) {
  if (uf == uf) uf = round_through<T>(fminf(fmaxf(uf, -limit), limit));
  if (gf == gf) gf = round_through<T>(fminf(gf, limit));
}

// exact (erf) GELU, the default of torch.nn.functional.gelu: 0.5 * x * (1 + erf(x / sqrt(2))) in fp32
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// bf16 tensors: 1 - erf(z) = poly(t) * exp(-z^2), t = 1 / (1 + p z)  (Abramowitz-Stegun 7.1.26, |error| <= 1.5e-7 -
// far inside bf16's 8 significand bits) with the two transcendental pieces on the MUFU.  The complementary form has no
// cancellation in the negative tail, and the kernel stays HBM-bound (libm's erff made it ALU-bound: 143 us instead of
// ~75 us at 8192 x 12288).
__device__ __forceinline__ float gelu_fast_f(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  float q = fmaf(1.061405429f, t, -1.453152027f);
  q = fmaf(q, t, 1.421413741f);
  q = fmaf(q, t, -0.284496736f);
  q = fmaf(q, t, 0.254829592f);
  q = q * t * e;  // 1 - erf(z) = erfc(z)
  return 0.5f * x * (x >= 0.f ? 2.0f - q : q);
}

template <typename T, bool GELU> __device__ __forceinline__ float unary_f(float x) {
  if constexpr (GELU) {
    if constexpr (std::is_same<T, __nv_bfloat16>::value) return gelu_fast_f(x);
    else return gelu_f(x);
  } else {
    return silu_f<T>(x);
  }
}

template <typename T, bool GATED, bool CLAMP, bool GELU = false>
__global__ void __launch_bounds__(256, 4) act_kernel(const T* __restrict__ gate, const T* __restrict__ up,
                                                  T* __restrict__ out, int64_t rows, int64_t cols, int64_t g_rs,
                                                  int64_t u_rs, int64_t o_rs, float limit) {
  constexpr int N = Vec16<T>::N;
  const int64_t vecs = cols / N;
  pdl_wait();
  pdl_trigger();
  for (int64_t row = blockIdx.y; row < rows; row += gridDim.y) {
    const T* g = gate + row * g_rs;
    const T* u = GATED ? up + row * u_rs : nullptr;
    T* o = out + row * o_rs;
    // Grid-stride over the row's vectors, TWO vectors in flight per thread and trip.  The math of these ops (two MUFU
    // + ~10-18 FP instructions per element) makes them ISSUE-bound, not HBM-bound: with one vector per thread the
    // ~100 instructions of per-thread set-up (64-bit index arithmetic) were 38 % of all instructions (ncu: 33.8 issued
    // per element for GELU, 84 % issue utilisation, 0.58 of the HBM peak); a capped grid amortises them over ~20 vectors.
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    auto compute = [&](const Vec16<T>& gv, const Vec16<T>& uv) {
      Vec16<T> ov;
#pragma unroll
      for (int e = 0; e < N; ++e) {
        float gf = DType<T>::to_f(gv.v[e]);
        if (GATED) {
          float uf = DType<T>::to_f(uv.v[e]);
          if (CLAMP) clamp_pair<T>(gf, uf, limit);
          ov.v[e] = DType<T>::from_f(__fmul_rn(round_through<T>(silu_f<T>(gf)), uf));
        } else {
          ov.v[e] = DType<T>::from_f(unary_f<T, GELU>(gf));
        }
      }
      return ov;
    };
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; v + stride < vecs; v += 2 * stride) {
      const Vec16<T> g0 = ld_vec(g + v * N), g1 = ld_vec(g + (v + stride) * N);
      Vec16<T> u0, u1;
      if (GATED) {
        u0 = ld_vec(u + v * N);
        u1 = ld_vec(u + (v + stride) * N);
      }
      st_vec(o + v * N, compute(g0, u0));
      st_vec(o + (v + stride) * N, compute(g1, u1));
    }
    if (v < vecs) {
      const Vec16<T> g0 = ld_vec(g + v * N);
      Vec16<T> u0;
      if (GATED) u0 = ld_vec(u + v * N);
      st_vec(o + v * N, compute(g0, u0));
    }
    // scalar tail when cols is not a multiple of the vector width
    for (int64_t c = vecs * N + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < cols;
         c += (int64_t)gridDim.x * blockDim.x) {
      float gf = DType<T>::to_f(g[c]);
      if (GATED) {
        float uf = DType<T>::to_f(u[c]);
        if (CLAMP) clamp_pair<T>(gf, uf, limit);
        o[c] = DType<T>::from_f(__fmul_rn(round_through<T>(silu_f<T>(gf)), uf));
      } else {
        o[c] = DType<T>::from_f(unary_f<T, GELU>(gf));
      }
    }
  }
}

// scalar variant for rows whose strides / base pointers are not 16-byte aligned
template <typename T, bool GATED, bool GELU = false>
__global__ void __launch_bounds__(256) act_scalar_kernel(const T* __restrict__ gate, const T* __restrict__ up,
                                                         T* __restrict__ out, int64_t rows, int64_t cols, int64_t g_rs,
                                                         int64_t u_rs, int64_t o_rs, float limit) {
  for (int64_t row = blockIdx.y; row < rows; row += gridDim.y) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += (int64_t)gridDim.x * blockDim.x) {
      float gf = DType<T>::to_f(gate[row * g_rs + c]);
      if (GATED) {
        float uf = DType<T>::to_f(up[row * u_rs + c]);
        if (limit > 0.f) clamp_pair<T>(gf, uf, limit);
        out[row * o_rs + c] = DType<T>::from_f(__fmul_rn(round_through<T>(silu_f<T>(gf)), uf));
      } else {
        out[row * o_rs + c] = DType<T>::from_f(unary_f<T, GELU>(gf));
      }
    }
  }
}

static int act_entry(const void* gate, const void* up, void* out, int64_t rows, int64_t cols, int64_t g_rs,
                     int64_t u_rs, int64_t o_rs, float limit, int dtype, void* stream, bool gated,
                     bool gelu = false) {
  MOJO_REQUIRE(rows >= 0 && cols >= 0, MOJO_B200_EINVAL, "activation: bad sizes");
  if (rows == 0 || cols == 0) return 0;
  MOJO_REQUIRE(gate && out && (!gated || up), MOJO_B200_EINVAL, "activation: null tensor pointer");
  const int eb = dtype_bytes(dtype);
  const int n = 16 / eb;
  const bool vec_ok = aligned16(gate) && aligned16(out) && (!gated || aligned16(up)) &&
                      (rows == 1 || (g_rs % n == 0 && o_rs % n == 0 && (!gated || u_rs % n == 0)));
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t per_row = vec_ok ? (cols + n - 1) / n : cols;
  // one vector per thread: x covers a row (or the whole flat tensor), y the rows; the loops only wrap for tensors
  // beyond the grid limits
  int64_t gx = (per_row + 255) / 256;
  gx = gx < 1 ? 1 : (gx > 0x7fffffffLL ? 0x7fffffffLL : gx);
  int64_t gy = rows < 1 ? 1 : (rows > 65535 ? 65535 : rows);
  // the vector kernel strides over its row: at most ~2 resident waves of blocks in all (8 blocks of 256 threads per SM)
  if (vec_ok) {
    const int64_t cap = (int64_t)kNumSMs * 4 * 4;  // 4 resident blocks per SM, 4 waves
    const int64_t gx_cap = (cap + gy - 1) / gy;
    if (gx > gx_cap) gx = gx_cap < 1 ? 1 : gx_cap;
  }
  dim3 grid((unsigned)gx, (unsigned)gy);
  return dispatch_dtype(dtype, [&](auto tag) {
    using T = decltype(tag);
    if (vec_ok) {
      const int64_t zero = 0;
      if (gated && limit > 0.f) launch_pdl(act_kernel<T, true, true>, grid, dim3(256), 0, s, (const T*)gate, (const T*)up, (T*)out, rows, cols, g_rs, u_rs, o_rs, limit);
      else if (gated) launch_pdl(act_kernel<T, true, false>, grid, dim3(256), 0, s, (const T*)gate, (const T*)up, (T*)out, rows, cols, g_rs, u_rs, o_rs, limit);
      else if (gelu) launch_pdl(act_kernel<T, false, false, true>, grid, dim3(256), 0, s, (const T*)gate, (const T*)nullptr, (T*)out, rows, cols, g_rs, zero, o_rs, 0.f);
      else launch_pdl(act_kernel<T, false, false>, grid, dim3(256), 0, s, (const T*)gate, (const T*)nullptr, (T*)out, rows, cols, g_rs, zero, o_rs, 0.f);
    } else {
      if (gated) act_scalar_kernel<T, true><<<grid, 256, 0, s>>>((const T*)gate, (const T*)up, (T*)out, rows, cols, g_rs, u_rs, o_rs, limit);
      else if (gelu) act_scalar_kernel<T, false, true><<<grid, 256, 0, s>>>((const T*)gate, nullptr, (T*)out, rows, cols, g_rs, 0, o_rs, 0.f);
      else act_scalar_kernel<T, false><<<grid, 256, 0, s>>>((const T*)gate, nullptr, (T*)out, rows, cols, g_rs, 0, o_rs, 0.f);
    }
    return check_launch("act_kernel");
  });
}

}  // namespace mojo

extern "C" int mojo_b200_swiglu(const void* gate, const void* up, void* out, int64_t rows, int64_t cols,
                                int64_t gate_row_stride, int64_t up_row_stride, int64_t out_row_stride,
                                float swiglu_limit, int dtype, void* stream) {
  return mojo::act_entry(gate, up, out, rows, cols, gate_row_stride, up_row_stride, out_row_stride, swiglu_limit, dtype,
                         stream, true);
}

extern "C" int mojo_b200_silu(const void* x, void* out, int64_t rows, int64_t cols, int64_t x_row_stride,
                              int64_t out_row_stride, int dtype, void* stream) {
  return mojo::act_entry(x, nullptr, out, rows, cols, x_row_stride, 0, out_row_stride, 0.f, dtype, stream, false);
}

extern "C" int mojo_b200_gelu(const void* x, void* out, int64_t rows, int64_t cols, int64_t x_row_stride,
                              int64_t out_row_stride, int dtype, void* stream) {
  return mojo::act_entry(x, nullptr, out, rows, cols, x_row_stride, 0, out_row_stride, 0.f, dtype, stream, false, true);
}
