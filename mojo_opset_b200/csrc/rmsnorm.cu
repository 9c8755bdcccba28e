// MojoRMSNorm / MojoResidualAddRMSNorm: one HBM pass, 128-bit accesses, fp32 math, one rounding.
//
//   sum = round_T(x + residual)                      (eager add materialises in the input dtype)
//   y   = round_T(sum * (1 / sqrt(mean(sum^2) + eps)) * w)
//
// A row is owned by a group of TPR threads - the narrowest power of two that keeps the row register-resident
// at 4 x 16 B per thread (4 threads for head_dim-sized q/k-norm rows, 128 for hidden 4096) - and a 256-thread
// CTA carries 256 / TPR rows, so every thread has 4 (8 with a residual) independent 16-byte loads in flight
// before the reduction.  Each thread keeps its slice of the row in registers between the reduction and the
// scaling, so x / residual are read exactly once: bytes = (2 reads + 2 writes) * rows * H * sizeof(T) + H.
#include <type_traits>

#include "tcgen05.cuh"  // packed fp32x2 arithmetic (mul.rn.f32x2 / add.rn.f32x2 round each lane like the scalar op)

namespace mojo {

template <typename T> __device__ __forceinline__ float2 rn_unpack_pair(uint32_t w) {
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
  } else {
    return __half22float2(*reinterpret_cast<const __half2*>(&w));
  }
}

constexpr int kMaxVecsPerThread = 4;  // register-resident slice: up to 4 x 16 B per thread

// VEC elements per access: 16 bytes when the hidden size, strides and pointers allow, narrower otherwise
// (e.g. hidden = 7338 or 734 in the reference's tests is only 4-byte divisible in 16-bit dtypes).
template <typename T, int VEC> struct alignas(sizeof(T) * VEC) PackN { T v[VEC]; };
template <typename T, int VEC> __device__ __forceinline__ PackN<T, VEC> ld_pack(const T* p) {
  return *reinterpret_cast<const PackN<T, VEC>*>(p);
}
template <typename T, int VEC> __device__ __forceinline__ void st_pack(T* p, const PackN<T, VEC>& v) {
  *reinterpret_cast<PackN<T, VEC>*>(p) = v;
}

constexpr int kNormCta = 256;

template <typename T, int VEC, int TPR, bool HAS_RES>
__global__ void __launch_bounds__(TPR > kNormCta ? TPR : kNormCta) rmsnorm_kernel(
    const T* __restrict__ x, const T* __restrict__ res, const T* __restrict__ w, T* __restrict__ y,
    T* __restrict__ sum_out, int64_t rows, int hidden, int64_t x_rs, int64_t res_rs, int64_t y_rs, int64_t sum_rs,
    float eps) {
  constexpr int N = VEC;
  constexpr int ROWS_PER_CTA = (TPR >= kNormCta ? 1 : kNormCta / TPR);
  pdl_wait();
  pdl_trigger();
  const int lane_in_row = threadIdx.x % TPR;
  const int64_t row = (int64_t)blockIdx.x * ROWS_PER_CTA + threadIdx.x / TPR;
  const bool active = row < rows;
  const int vecs = hidden / N;

  constexpr bool kPacked = sizeof(T) == 2 && VEC % 2 == 0;
  PackN<T, VEC> keep[kMaxVecsPerThread];
  float ss = 0.f;
  float2 ss2 = make_float2(0.f, 0.f);
  if (active) {
    const T* xr = x + row * x_rs;
    const T* rr = HAS_RES ? res + row * res_rs : nullptr;
#pragma unroll
    for (int i = 0; i < kMaxVecsPerThread; ++i) {
      const int v = lane_in_row + i * TPR;
      if (v < vecs) {
        PackN<T, VEC> a = ld_pack<T, VEC>(xr + (int64_t)v * N);
        if constexpr (kPacked) {
          // two elements per issue slot (FADD2 / FFMA2): the same roundings as the scalar form, lane by lane
          uint32_t* aw = reinterpret_cast<uint32_t*>(&a);
          PackN<T, VEC> b;
          if (HAS_RES) b = ld_pack<T, VEC>(rr + (int64_t)v * N);
          const uint32_t* bw = reinterpret_cast<const uint32_t*>(&b);
#pragma unroll
          for (int e = 0; e < N / 2; ++e) {
            float2 f = rn_unpack_pair<T>(aw[e]);
            if (HAS_RES) {
              f = add2(f, rn_unpack_pair<T>(bw[e]));
              aw[e] = pack2<T>(f.x, f.y);
              f = rn_unpack_pair<T>(aw[e]);
            }
            ss2 = fma2(f, f, ss2);
          }
          keep[i] = a;
        } else {
        if (HAS_RES) {
          const PackN<T, VEC> b = ld_pack<T, VEC>(rr + (int64_t)v * N);
#pragma unroll
          for (int e = 0; e < N; ++e) a.v[e] = DType<T>::from_f(__fadd_rn(DType<T>::to_f(a.v[e]), DType<T>::to_f(b.v[e])));
        }
        keep[i] = a;  // the rounded sum is stored together with y: a pure read phase, then a pure write phase
#pragma unroll
        for (int e = 0; e < N; ++e) {
          const float f = DType<T>::to_f(a.v[e]);
          ss = fmaf(f, f, ss);
        }
        }
      }
    }
    // rows wider than the register slice: stream the remainder (re-read in the second phase)
    for (int v = lane_in_row + kMaxVecsPerThread * TPR; v < vecs; v += TPR) {
      PackN<T, VEC> a = ld_pack<T, VEC>(xr + (int64_t)v * N);
      if (HAS_RES) {
        const PackN<T, VEC> b = ld_pack<T, VEC>(rr + (int64_t)v * N);
#pragma unroll
        for (int e = 0; e < N; ++e) a.v[e] = DType<T>::from_f(__fadd_rn(DType<T>::to_f(a.v[e]), DType<T>::to_f(b.v[e])));
        if (sum_out) st_pack<T, VEC>(sum_out + row * sum_rs + (int64_t)v * N, a);
      }
#pragma unroll
      for (int e = 0; e < N; ++e) {
        const float f = DType<T>::to_f(a.v[e]);
        ss = fmaf(f, f, ss);
      }
    }
  }

  if constexpr (kPacked) ss += ss2.x + ss2.y;
  // reduce ss over the TPR threads of the row (fixed order: deterministic)
  if constexpr (TPR <= 32) {
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  } else {
    __shared__ float part[32];
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = ss;
    __syncthreads();
    constexpr int WPR = TPR / 32;  // warps per row
    const int w0 = (threadIdx.x / TPR) * WPR;
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < WPR; ++i) t += part[w0 + i];
    ss = t;
  }
  if (!active) return;

  const float inv = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fdiv_rn(ss, (float)hidden), eps)));
  T* yr = y + row * y_rs;
#pragma unroll
  for (int i = 0; i < kMaxVecsPerThread; ++i) {
    const int v = lane_in_row + i * TPR;
    if (v < vecs) {
      const PackN<T, VEC> g = ld_pack<T, VEC>(w + (int64_t)v * N);
      PackN<T, VEC> o;
      if constexpr (kPacked) {
        const uint32_t* kw = reinterpret_cast<const uint32_t*>(&keep[i]);
        const uint32_t* gw = reinterpret_cast<const uint32_t*>(&g);
        uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
        const float2 inv2 = make_float2(inv, inv);
#pragma unroll
        for (int e = 0; e < N / 2; ++e) {
          const float2 t = mul2(mul2(rn_unpack_pair<T>(kw[e]), inv2), rn_unpack_pair<T>(gw[e]));
          ow[e] = pack2<T>(t.x, t.y);
        }
      } else {
#pragma unroll
      for (int e = 0; e < N; ++e)
        o.v[e] = DType<T>::from_f(__fmul_rn(__fmul_rn(DType<T>::to_f(keep[i].v[e]), inv), DType<T>::to_f(g.v[e])));
      }
      st_pack<T, VEC>(yr + (int64_t)v * N, o);
      if (HAS_RES && sum_out) st_pack<T, VEC>(sum_out + row * sum_rs + (int64_t)v * N, keep[i]);
    }
  }
  for (int v = lane_in_row + kMaxVecsPerThread * TPR; v < vecs; v += TPR) {
    // the summed row was either written to sum_out (re-read it) or must be recomputed from x (+ residual)
    PackN<T, VEC> a;
    if (HAS_RES && sum_out) {
      a = ld_pack<T, VEC>(sum_out + row * sum_rs + (int64_t)v * N);
    } else {
      a = ld_pack<T, VEC>(x + row * x_rs + (int64_t)v * N);
      if (HAS_RES) {
        const PackN<T, VEC> b = ld_pack<T, VEC>(res + row * res_rs + (int64_t)v * N);
#pragma unroll
        for (int e = 0; e < N; ++e) a.v[e] = DType<T>::from_f(__fadd_rn(DType<T>::to_f(a.v[e]), DType<T>::to_f(b.v[e])));
      }
    }
    const PackN<T, VEC> g = ld_pack<T, VEC>(w + (int64_t)v * N);
    PackN<T, VEC> o;
#pragma unroll
    for (int e = 0; e < N; ++e)
      o.v[e] = DType<T>::from_f(__fmul_rn(__fmul_rn(DType<T>::to_f(a.v[e]), inv), DType<T>::to_f(g.v[e])));
    st_pack<T, VEC>(yr + (int64_t)v * N, o);
  }
}

template <typename T, int VEC, bool HAS_RES>
static int launch_rmsnorm(const void* x, const void* res, const void* w, void* y, void* sum_out, int64_t rows,
                          int hidden, int64_t x_rs, int64_t res_rs, int64_t y_rs, int64_t sum_rs, float eps,
                          cudaStream_t s) {
  const int vecs = hidden / VEC;
#define RUN(TPR)                                                                                              \
  do {                                                                                                        \
    constexpr int RPC = (TPR >= kNormCta ? 1 : kNormCta / TPR);                                               \
    const int64_t ctas = (rows + RPC - 1) / RPC;                                                              \
    launch_pdl(rmsnorm_kernel<T, VEC, TPR, HAS_RES>, dim3((unsigned)ctas), dim3(TPR > kNormCta ? TPR : kNormCta), 0, s, \
               (const T*)x, (const T*)res, (const T*)w, (T*)y, (T*)sum_out, rows, hidden, x_rs, res_rs, y_rs, sum_rs,  \
               eps);                                                                                          \
  } while (0)
  // the narrowest group that keeps the row register-resident with FOUR 16-byte loads in flight per thread
  // (4 packs of x, or 2 of x + 2 of the residual): measured best on B200 - wider groups pay for the CTA-level
  // reduction, narrower ones lose occupancy to registers (sweep in profiles/README.md)
  const int packs = HAS_RES ? 2 : kMaxVecsPerThread;
  int tpr = 4;
  while (tpr < 1024 && vecs > tpr * packs) tpr *= 2;
  // ... widened while the grid would leave SMs idle (decode-sized row counts) and a thread still has >= 2 packs
  auto ctas_for = [&](int t) { return (rows + (t >= kNormCta ? 1 : kNormCta / t) - 1) / (t >= kNormCta ? 1 : kNormCta / t); };
  while (tpr < 1024 && vecs >= tpr * 2 && ctas_for(tpr) < 2 * kNumSMs) tpr *= 2;
  switch (tpr) {
    case 4: RUN(4); break;
    case 8: RUN(8); break;
    case 16: RUN(16); break;
    case 32: RUN(32); break;
    case 64: RUN(64); break;
    case 128: RUN(128); break;
    case 256: RUN(256); break;
    case 512: RUN(512); break;
    default: RUN(1024); break;
  }
#undef RUN
  return check_launch("rmsnorm_kernel");
}

static int rmsnorm_entry(const void* x, const void* res, const void* w, void* y, void* sum_out, int64_t rows,
                         int hidden, int64_t x_rs, int64_t res_rs, int64_t y_rs, int64_t sum_rs, float eps, int dtype,
                         void* stream, bool has_res) {
  MOJO_REQUIRE(rows >= 0 && hidden > 0, MOJO_B200_EINVAL, "rms_norm: bad sizes rows=%lld hidden=%d", (long long)rows,
               hidden);
  if (rows == 0) return 0;
  MOJO_REQUIRE(x && w && y && (!has_res || res), MOJO_B200_EINVAL, "rms_norm: null tensor pointer");
  MOJO_REQUIRE(rows <= 0x7fffffffLL, MOJO_B200_EUNSUPPORTED, "rms_norm: too many rows");
  const int eb = dtype_bytes(dtype);
  // widest power-of-two pack (<= 16 bytes) that the hidden size, every row stride and base pointer allow
  int vec = 16 / eb;
  auto fits = [&](int v) {
    if (hidden % v || x_rs % v || y_rs % v || (has_res && res_rs % v) || (sum_out && sum_rs % v)) return false;
    const uintptr_t m = (uintptr_t)v * eb - 1;
    uintptr_t bits = (uintptr_t)x | (uintptr_t)w | (uintptr_t)y;
    if (has_res) bits |= (uintptr_t)res;
    if (sum_out) bits |= (uintptr_t)sum_out;
    return (bits & m) == 0;
  };
  while (vec > 1 && !fits(vec)) vec >>= 1;
  cudaStream_t s = (cudaStream_t)stream;
  return dispatch_dtype(dtype, [&](auto tag) {
    using T = decltype(tag);
    constexpr int kMaxVec = 16 / (int)sizeof(T);
#define GO(V)                                                                                                      \
  (has_res ? launch_rmsnorm<T, V, true>(x, res, w, y, sum_out, rows, hidden, x_rs, res_rs, y_rs, sum_rs, eps, s)   \
           : launch_rmsnorm<T, V, false>(x, res, w, y, sum_out, rows, hidden, x_rs, res_rs, y_rs, sum_rs, eps, s))
    if constexpr (kMaxVec >= 8) {
      if (vec == 8) return GO(8);
    }
    if (vec >= 4) return GO(4);
    if (vec == 2) return GO(2);
    return GO(1);
#undef GO
  });
}

}  // namespace mojo

extern "C" int mojo_b200_rms_norm(const void* x, const void* weight, void* y, int64_t rows, int hidden,
                                  int64_t x_row_stride, int64_t y_row_stride, float eps, int dtype, void* stream) {
  return mojo::rmsnorm_entry(x, nullptr, weight, y, nullptr, rows, hidden, x_row_stride, 0, y_row_stride, 0, eps, dtype,
                             stream, false);
}

extern "C" int mojo_b200_residual_add_rms_norm(const void* x, const void* residual, const void* weight, void* y,
                                               void* sum_out, int64_t rows, int hidden, int64_t x_row_stride,
                                               int64_t res_row_stride, int64_t y_row_stride, int64_t sum_row_stride,
                                               float eps, int dtype, void* stream) {
  return mojo::rmsnorm_entry(x, residual, weight, y, sum_out, rows, hidden, x_row_stride, res_row_stride, y_row_stride,
                             sum_row_stride, eps, dtype, stream, true);
}
