// The non-GEMM ops around MojoSdpa in the DiT block (modeling/wan2_2/mojo_wan_model.py):
//
//   MojoLayerNorm  core/operators/normalization.py:19-66     F.layer_norm over the last dim, optional affine
//   MojoGridRoPE   experimental/operators/position_embedding.py:80-118   3-D grid RoPE: interleaved (re, im) pairs of
//                  x[b, :seq_len] times a per-sample complex phase table, padding tokens passed through
//
// Both are one HBM pass.  LayerNorm follows rmsnorm.cu's shape: a row is owned by a group of threads that keeps it in
// registers (four 16-byte loads in flight per thread), two-pass moments from registers (mean, then the centred sum
// of squares).  GridRoPE is a flat one-vector-per-thread streaming kernel (full occupancy).
#include <type_traits>

#include "tcgen05.cuh"  // packed fp32x2 arithmetic

namespace mojo {

// a packed pair of 16-bit values -> two fp32 (bf16: one shift / one mask)
template <typename T> __device__ __forceinline__ float2 unpack_pair(uint32_t w) {
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
  } else {
    return __half22float2(*reinterpret_cast<const __half2*>(&w));
  }
}

constexpr int kLnCta = 256;
constexpr int kLnPacks = 4;

template <typename T, int VEC> struct alignas(sizeof(T) * VEC) LnPack { T v[VEC]; };

template <typename T, int VEC, int TPR>
__global__ void __launch_bounds__(TPR > kLnCta ? TPR : kLnCta) layernorm_kernel(
    const T* __restrict__ x, const T* __restrict__ w, const T* __restrict__ b, T* __restrict__ y, int64_t rows,
    int hidden, int64_t x_rs, int64_t y_rs, float eps) {
  constexpr int RPC = TPR >= kLnCta ? 1 : kLnCta / TPR;
  const int lane_in_row = threadIdx.x % TPR;
  const int64_t row = (int64_t)blockIdx.x * RPC + threadIdx.x / TPR;
  const bool active = row < rows;
  const int vecs = hidden / VEC;
  __shared__ float part[2][32];

  // both moments in ONE reduction round: the sums are taken of d = x - x0 (x0 = the row's first element), so
  // var = E[d^2] - E[d]^2 does not cancel for rows that sit far from zero, and the second pass over the registers with
  // its second block-level reduction (a barrier + shared-memory round trip: the kernel is latency-, not bandwidth-bound
  // there) is gone
  auto group_sum2 = [&](float2 v) -> float2 {
    if constexpr (TPR <= 32) {
#pragma unroll
      for (int o = TPR / 2; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
      }
      return v;
    } else {
      v.x = warp_sum(v.x);
      v.y = warp_sum(v.y);
      if ((threadIdx.x & 31) == 0) {
        part[0][threadIdx.x >> 5] = v.x;
        part[1][threadIdx.x >> 5] = v.y;
      }
      __syncthreads();
      constexpr int WPR = TPR / 32;
      const int w0 = (threadIdx.x / TPR) * WPR;
      float2 t = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < WPR; ++i) {
        t.x += part[0][w0 + i];
        t.y += part[1][w0 + i];
      }
      return t;
    }
  };

  // 16-bit tensors with whole 16-byte packs: both passes run on packed fp32 pairs (sm_100 FADD2 / FFMA2: one issue
  // slot for two elements; the kernel was ~11.5 instructions per element against RMSNorm's ~6.5 and 0.64 vs 0.82 of
  // the HBM peak) and the output is two FMAs per pair: (x * rstd - mean * rstd) * w + b
  constexpr bool kPacked = sizeof(T) == 2 && VEC % 2 == 0;
  LnPack<T, VEC> keep[kLnPacks];
  float s = 0.f, ss = 0.f, x0 = 0.f;
  if (active) {
    const T* xr = x + row * x_rs;
    x0 = DType<T>::to_f(xr[0]);
#pragma unroll
    for (int i = 0; i < kLnPacks; ++i) {
      const int v = lane_in_row + i * TPR;
      if (v < vecs) keep[i] = *reinterpret_cast<const LnPack<T, VEC>*>(xr + (int64_t)v * VEC);
    }
    if constexpr (kPacked) {
      float2 s2 = make_float2(0.f, 0.f), ss2 = s2;
      const float2 x02 = make_float2(x0, x0);
#pragma unroll
      for (int i = 0; i < kLnPacks; ++i) {
        const int v = lane_in_row + i * TPR;
        if (v < vecs) {
          const uint32_t* wds = reinterpret_cast<const uint32_t*>(&keep[i]);
#pragma unroll
          for (int e = 0; e < VEC / 2; ++e) {
            const float2 d = sub2(unpack_pair<T>(wds[e]), x02);
            s2 = add2(s2, d);
            ss2 = fma2(d, d, ss2);
          }
        }
      }
      s = s2.x + s2.y;
      ss = ss2.x + ss2.y;
    } else {
#pragma unroll
    for (int i = 0; i < kLnPacks; ++i) {
      const int v = lane_in_row + i * TPR;
      if (v < vecs) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const float d = DType<T>::to_f(keep[i].v[e]) - x0;
          s += d;
          ss = fmaf(d, d, ss);
        }
      }
    }
    }
    for (int v = lane_in_row + kLnPacks * TPR; v < vecs; v += TPR) {  // rows wider than the register slice
      const LnPack<T, VEC> a = *reinterpret_cast<const LnPack<T, VEC>*>(xr + (int64_t)v * VEC);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float d = DType<T>::to_f(a.v[e]) - x0;
        s += d;
        ss = fmaf(d, d, ss);
      }
    }
  }
  const float2 mom = group_sum2(make_float2(s, ss));
  const float dmean = mom.x / (float)hidden;
  const float mean = x0 + dmean;
  const float var = fmaxf(mom.y / (float)hidden - dmean * dmean, 0.f);
  if (!active) return;
  const float inv = __fdiv_rn(1.0f, __fsqrt_rn(var + eps));
  T* yr = y + row * y_rs;
  const float nmi = -mean * inv;
  auto emit = [&](const LnPack<T, VEC>& a, int v) {
    LnPack<T, VEC> o;
    LnPack<T, VEC> g, h;
    if (w) g = *reinterpret_cast<const LnPack<T, VEC>*>(w + (int64_t)v * VEC);
    if (b) h = *reinterpret_cast<const LnPack<T, VEC>*>(b + (int64_t)v * VEC);
    if constexpr (kPacked) {
      const uint32_t* aw = reinterpret_cast<const uint32_t*>(&a);
      const uint32_t* gw = reinterpret_cast<const uint32_t*>(&g);
      const uint32_t* hw = reinterpret_cast<const uint32_t*>(&h);
      uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
      const float2 inv2 = make_float2(inv, inv), nmi2 = make_float2(nmi, nmi);
#pragma unroll
      for (int e = 0; e < VEC / 2; ++e) {
        float2 t = fma2(unpack_pair<T>(aw[e]), inv2, nmi2);
        if (w && b) t = fma2(t, unpack_pair<T>(gw[e]), unpack_pair<T>(hw[e]));
        else if (w) t = mul2(t, unpack_pair<T>(gw[e]));
        else if (b) t = add2(t, unpack_pair<T>(hw[e]));
        ow[e] = pack2<T>(t.x, t.y);
      }
      *reinterpret_cast<LnPack<T, VEC>*>(yr + (int64_t)v * VEC) = o;
      return;
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      float t = (DType<T>::to_f(a.v[e]) - mean) * inv;
      if (w) t *= DType<T>::to_f(g.v[e]);
      if (b) t += DType<T>::to_f(h.v[e]);
      o.v[e] = DType<T>::from_f(t);
    }
    *reinterpret_cast<LnPack<T, VEC>*>(yr + (int64_t)v * VEC) = o;
  };
#pragma unroll
  for (int i = 0; i < kLnPacks; ++i) {
    const int v = lane_in_row + i * TPR;
    if (v < vecs) emit(keep[i], v);
  }
  for (int v = lane_in_row + kLnPacks * TPR; v < vecs; v += TPR)
    emit(*reinterpret_cast<const LnPack<T, VEC>*>(x + row * x_rs + (int64_t)v * VEC), v);
}

template <typename T, int VEC>
static int launch_layernorm(const void* x, const void* w, const void* b, void* y, int64_t rows, int hidden, int64_t x_rs,
                            int64_t y_rs, float eps, cudaStream_t s) {
  const int vecs = hidden / VEC;
  int tpr = 4;
  while (tpr < 1024 && vecs > tpr * kLnPacks) tpr *= 2;
  auto ctas_for = [&](int t) { const int rpc = t >= kLnCta ? 1 : kLnCta / t; return (rows + rpc - 1) / rpc; };
  while (tpr < 1024 && vecs >= tpr * 2 && ctas_for(tpr) < 2 * kNumSMs) tpr *= 2;
#define LN_RUN(TPR)                                                                                           \
  layernorm_kernel<T, VEC, TPR><<<(unsigned)ctas_for(TPR), (TPR > kLnCta ? TPR : kLnCta), 0, s>>>(            \
      (const T*)x, (const T*)w, (const T*)b, (T*)y, rows, hidden, x_rs, y_rs, eps)
  switch (tpr) {
    case 4: LN_RUN(4); break;
    case 8: LN_RUN(8); break;
    case 16: LN_RUN(16); break;
    case 32: LN_RUN(32); break;
    case 64: LN_RUN(64); break;
    case 128: LN_RUN(128); break;
    case 256: LN_RUN(256); break;
    case 512: LN_RUN(512); break;
    default: LN_RUN(1024); break;
  }
#undef LN_RUN
  return check_launch("layernorm_kernel");
}

// ---- grid RoPE: x [L, N, D] of one sample, phase [seq_len, D/2] complex64 as interleaved (cos, sin) fp32 ----------
template <typename T, int VEC>  // VEC elements (= VEC/2 complex pairs) per thread
__global__ void __launch_bounds__(256) grid_rope_kernel(const T* __restrict__ x, const float* __restrict__ phase,
                                                        T* __restrict__ out, int64_t seq_len, int64_t tokens, int heads,
                                                        int head_dim, int64_t x_st, int64_t x_sh, int64_t o_st,
                                                        int64_t o_sh, int64_t phase_st) {
  const int vecs_per_head = head_dim / VEC;
  const int64_t total = tokens * heads * vecs_per_head;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs_per_head);
    const int64_t th = i / vecs_per_head;
    const int h = (int)(th % heads);
    const int64_t t = th / heads;
    LnPack<T, VEC> a = *reinterpret_cast<const LnPack<T, VEC>*>(x + t * x_st + (int64_t)h * x_sh + v * VEC);
    if (t < seq_len) {
      const float* ph = phase + t * phase_st + v * VEC;  // (cos, sin) of pair j at [2j, 2j+1]
      float c[VEC];
#pragma unroll
      for (int e = 0; e < VEC; e += 4) *reinterpret_cast<float4*>(c + e) = *reinterpret_cast<const float4*>(ph + e);
#pragma unroll
      for (int e = 0; e < VEC; e += 2) {
        const float re = DType<T>::to_f(a.v[e]), im = DType<T>::to_f(a.v[e + 1]);
        // complex product exactly as ATen evaluates it: (ac - bd) + i(ad + bc), each product rounded (no FMA)
        a.v[e] = DType<T>::from_f(__fsub_rn(__fmul_rn(re, c[e]), __fmul_rn(im, c[e + 1])));
        a.v[e + 1] = DType<T>::from_f(__fadd_rn(__fmul_rn(re, c[e + 1]), __fmul_rn(im, c[e])));
      }
    }
    *reinterpret_cast<LnPack<T, VEC>*>(out + t * o_st + (int64_t)h * o_sh + v * VEC) = a;
  }
}

}  // namespace mojo

extern "C" int mojo_b200_layer_norm(const void* x, const void* weight, const void* bias, void* y, int64_t rows,
                                    int hidden, int64_t x_row_stride, int64_t y_row_stride, float eps, int dtype,
                                    void* stream) {
  using namespace mojo;
  MOJO_REQUIRE(rows >= 0 && hidden > 0, MOJO_B200_EINVAL, "layer_norm: bad sizes");
  if (rows == 0) return 0;
  MOJO_REQUIRE(x && y, MOJO_B200_EINVAL, "layer_norm: null tensor pointer");
  MOJO_REQUIRE(rows <= 0x7fffffffLL, MOJO_B200_EUNSUPPORTED, "layer_norm: too many rows");
  const int eb = dtype_bytes(dtype);
  const int full = 16 / eb;
  uintptr_t bits = (uintptr_t)x | (uintptr_t)y | (uintptr_t)(x_row_stride * eb) | (uintptr_t)(y_row_stride * eb) |
                   (uintptr_t)weight | (uintptr_t)bias;
  const bool wide = hidden % full == 0 && (bits & 15) == 0;
  cudaStream_t s = (cudaStream_t)stream;
  return dispatch_dtype(dtype, [&](auto tag) {
    using T = decltype(tag);
    if (wide) return launch_layernorm<T, 16 / (int)sizeof(T)>(x, weight, bias, y, rows, hidden, x_row_stride, y_row_stride, eps, s);
    return launch_layernorm<T, 1>(x, weight, bias, y, rows, hidden, x_row_stride, y_row_stride, eps, s);
  });
}

extern "C" int mojo_b200_grid_rope(const void* x, const float* phase, void* out, int64_t seq_len, int64_t tokens,
                                   int heads, int head_dim, int64_t x_stride_t, int64_t x_stride_h, int64_t o_stride_t,
                                   int64_t o_stride_h, int64_t phase_stride_t, int dtype, void* stream) {
  using namespace mojo;
  MOJO_REQUIRE(tokens >= 0 && seq_len >= 0 && seq_len <= tokens && heads > 0 && head_dim > 0 && head_dim % 2 == 0,
               MOJO_B200_EINVAL, "grid_rope: bad sizes");
  if (tokens == 0) return 0;
  MOJO_REQUIRE(x && out && (phase || seq_len == 0), MOJO_B200_EINVAL, "grid_rope: null tensor pointer");
  const int eb = dtype_bytes(dtype);
  const int full = 16 / eb;
  const int64_t strides[] = {x_stride_t, x_stride_h, o_stride_t, o_stride_h};
  bool wide = head_dim % full == 0 && aligned16(x) && aligned16(out) && aligned16(phase) && phase_stride_t % 4 == 0;
  for (int64_t st : strides) wide = wide && st % full == 0;
  cudaStream_t s = (cudaStream_t)stream;
  MOJO_REQUIRE(wide, MOJO_B200_EUNSUPPORTED, "grid_rope: head_dim, strides and pointers must allow 16-byte vectors");
  const int64_t total = tokens * heads * (head_dim / full);
  int64_t grid = (total + 255) / 256;
  grid = grid < 1 ? 1 : (grid > 0x7fffffffLL ? 0x7fffffffLL : grid);
  return dispatch_dtype(dtype, [&](auto tag) {
    using T = decltype(tag);
    grid_rope_kernel<T, 16 / (int)sizeof(T)><<<(unsigned)grid, 256, 0, s>>>(
        (const T*)x, phase, (T*)out, seq_len, tokens, heads, head_dim, x_stride_t, x_stride_h, o_stride_t, o_stride_h,
        phase_stride_t);
    return check_launch("grid_rope_kernel");
  });
}
