// MojoApplyRoPE (rotate-half on the last rope_dim features of q and k, fused in one launch) and
// MojoRotaryEmbedding (cos/sin rows for a batch of positions).
//
// apply_rope: q/k are addressed as [batch, seq, heads, D] through element strides, so token-first,
// head-first and transposed-view layouts all take the same path.  One CTA per (batch, seq) token; its
// threads sweep the (head, vector) space of q then k.  A work item is one VEC-wide slice of the first
// rotary half together with its partner in the second half (both outputs produced by one thread), or one
// pass-through slice of the leading non-rotary features.  bytes = 2 * T * (Nq+Nk) * D * sizeof(T) + 2*T*d*sizeof(cos).
//
// Rounding follows the eager golden: math in promote(T, cos dtype); each product and the sum are rounded
// separately (no FMA contraction); when cos has the same 16-bit dtype as q the intermediates are rounded
// to that dtype too.
#include <type_traits>

#include "tcgen05.cuh"  // packed fp32x2 arithmetic: mul.rn.f32x2 / add.rn.f32x2 round each lane like the scalar op

namespace mojo {

template <typename T, int VEC> struct alignas(sizeof(T) * VEC) Pack { T v[VEC]; };

template <typename T, typename C, int VEC, bool ROUND_T>
__device__ __forceinline__ void rope_rows(const T* __restrict__ src, T* __restrict__ dst, int heads, int64_t s_h,
                                          int64_t d_h, const C* __restrict__ cos, const C* __restrict__ sin,
                                          int head_dim, int rope_dim) {
  const int nope = head_dim - rope_dim;
  const int half = rope_dim / 2;
  const int nope_items = nope / VEC;
  const int items = nope_items + half / VEC;
  const int total = heads * items;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int h = i / items;
    const int it = i - h * items;
    const T* s = src + h * s_h;
    T* d = dst + h * d_h;
    if (it < nope_items) {
      *reinterpret_cast<Pack<T, VEC>*>(d + it * VEC) = *reinterpret_cast<const Pack<T, VEC>*>(s + it * VEC);
      continue;
    }
    const int r = (it - nope_items) * VEC;  // offset inside the first rotary half
    const Pack<T, VEC> x1 = *reinterpret_cast<const Pack<T, VEC>*>(s + nope + r);
    const Pack<T, VEC> x2 = *reinterpret_cast<const Pack<T, VEC>*>(s + nope + half + r);
    const Pack<C, VEC> c1 = *reinterpret_cast<const Pack<C, VEC>*>(cos + r);
    const Pack<C, VEC> c2 = *reinterpret_cast<const Pack<C, VEC>*>(cos + half + r);
    const Pack<C, VEC> s1 = *reinterpret_cast<const Pack<C, VEC>*>(sin + r);
    const Pack<C, VEC> s2 = *reinterpret_cast<const Pack<C, VEC>*>(sin + half + r);
    Pack<T, VEC> o1, o2;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const float a = DType<T>::to_f(x1.v[e]), b = DType<T>::to_f(x2.v[e]);
      // first half:  x1*cos + (-x2)*sin ; second half: x2*cos + x1*sin
      float p1 = __fmul_rn(a, DType<C>::to_f(c1.v[e])), q1 = __fmul_rn(-b, DType<C>::to_f(s1.v[e]));
      float p2 = __fmul_rn(b, DType<C>::to_f(c2.v[e])), q2 = __fmul_rn(a, DType<C>::to_f(s2.v[e]));
      if (ROUND_T) {
        p1 = round_through<T>(p1); q1 = round_through<T>(q1);
        p2 = round_through<T>(p2); q2 = round_through<T>(q2);
      }
      o1.v[e] = DType<T>::from_f(__fadd_rn(p1, q1));
      o2.v[e] = DType<T>::from_f(__fadd_rn(p2, q2));
    }
    *reinterpret_cast<Pack<T, VEC>*>(d + nope + r) = o1;
    *reinterpret_cast<Pack<T, VEC>*>(d + nope + half + r) = o2;
  }
}

struct RopeArgs {
  const void *q, *k, *cos, *sin;
  void *qo, *ko;
  int64_t seq;
  int64_t tokens;  // batch * seq (the slice kernel strides over them)
  int q_heads, k_heads, head_dim, rope_dim;
  int64_t q_b, q_s, q_h, k_b, k_s, k_h, qo_b, qo_s, qo_h, ko_b, ko_s, ko_h, cos_b, cos_s;
};

template <typename T, typename C, int VEC, bool ROUND_T>
__global__ void __launch_bounds__(256) apply_rope_kernel(const RopeArgs a) {
  const int64_t tok = blockIdx.x;
  const int64_t b = tok / a.seq;
  const int64_t s = tok - b * a.seq;
  const C* cos = (const C*)a.cos + b * a.cos_b + s * a.cos_s;
  const C* sin = (const C*)a.sin + b * a.cos_b + s * a.cos_s;
  rope_rows<T, C, VEC, ROUND_T>((const T*)a.q + b * a.q_b + s * a.q_s, (T*)a.qo + b * a.qo_b + s * a.qo_s, a.q_heads,
                                a.q_h, a.qo_h, cos, sin, a.head_dim, a.rope_dim);
  rope_rows<T, C, VEC, ROUND_T>((const T*)a.k + b * a.k_b + s * a.k_s, (T*)a.ko + b * a.ko_b + s * a.ko_s, a.k_heads,
                                a.k_h, a.ko_h, cos, sin, a.head_dim, a.rope_dim);
}


// Fast path: one CTA per token, a thread owns ONE output slice position of a head - a pass-through slice, a
// first-half slice (x1*cos1 - x2*sin1) or a second-half slice (x2*cos2 + x1*sin2) - so it needs just one cos
// and one sin slice (read once per thread, not once per head) and walks the q then k heads in steps of
// 256 / SLICES.  Kept deliberately lean (<= 40 registers, two 16-byte loads per trip): on B200 a streaming
// kernel is fastest at FULL occupancy with one vector per thread in flight - unrolling more loads per thread
// costs registers, hence resident warps, and measured slower (tools/microbench/stream3.cu).  The partner slice
// is loaded by two threads of the same warp instruction (one L1 request).  No integer division in the loop.
// Requires SLICES = head_dim / VEC to be a power of two <= 32.
// THREADS / TRIPS: 256 threads with three head rows in flight per thread for decode-sized launches (latency: 2.0 us at
// 64 tokens); 128 threads with five for prefill-sized ones - ncu: ~21 instructions per element at 256 threads, most of
// them per-token set-up (cos / sin conversion, 64-bit address arithmetic) amortised over only 24 elements per thread;
// at 128 threads a thread owns 40 (T = 8192: 35.2 -> 32.6 us = 0.83 of the HBM peak; 64 threads: no better)
template <typename T, typename C, int VEC, bool ROUND_T, int SLICES, int THREADS, int TRIPS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) apply_rope_slice_kernel(const RopeArgs a) {
  constexpr int HSLOTS = THREADS / SLICES;
  const int sl = threadIdx.x % SLICES, hs = threadIdx.x / SLICES;
  pdl_wait();
  pdl_trigger();
  const int nope = a.head_dim - a.rope_dim, half = a.rope_dim / 2;
  const int off = sl * VEC;  // this thread's output slice inside a head
  const bool rotary = off >= nope;
  const bool second = off >= nope + half;
  const int partner = second ? off - half : off + half;
  // a CTA strides over tokens (the grid is capped at a few resident waves: thousands of 10 KB CTAs spent their lives
  // in launch and first-load latency)
#pragma unroll 1
  for (uint32_t tok = blockIdx.x; tok < (uint32_t)a.tokens; tok += gridDim.x) {
  const uint32_t b = tok / (uint32_t)a.seq;
  const uint32_t s = tok - b * (uint32_t)a.seq;
  const T* qs = (const T*)a.q + (int64_t)b * a.q_b + (int64_t)s * a.q_s;
  const T* ks = (const T*)a.k + (int64_t)b * a.k_b + (int64_t)s * a.k_s - (int64_t)a.q_heads * a.k_h;
  T* qd = (T*)a.qo + (int64_t)b * a.qo_b + (int64_t)s * a.qo_s;
  T* kd = (T*)a.ko + (int64_t)b * a.ko_b + (int64_t)s * a.ko_s - (int64_t)a.q_heads * a.ko_h;
  const int total = a.q_heads + a.k_heads;
  Pack<C, VEC> cs, sn;
  if (rotary) {
    const int64_t row = (int64_t)b * a.cos_b + (int64_t)s * a.cos_s + (off - nope);
    cs = *reinterpret_cast<const Pack<C, VEC>*>((const C*)a.cos + row);
    sn = *reinterpret_cast<const Pack<C, VEC>*>((const C*)a.sin + row);
  }
  // The first kRopeTrips head rows of a thread are all requested before any is rotated (one dependent load per trip made
  // the kernel latency-bound: 0.71 of the HBM peak at T = 8192), and the partner slice comes from the partner lane
  // (same head, rope_dim/2 elements away = a fixed lane distance inside the warp) by shuffle instead of a second load.
  constexpr int kRopeTrips = TRIPS;
  constexpr int kWords = VEC * (int)sizeof(T) / 4;
  static_assert(kWords >= 1, "slices are at least one word");
  const int partner_lane = (int)(threadIdx.x & 31u) + (second ? -(half / VEC) : (half / VEC));
  // cos / sin of this thread's slice as fp32, the sign of rotate_half folded into sin once (rotate_half(x) =
  // cat(-x2, x1): the partner enters negated in the first half only; (-y) * s == y * (-s) exactly)
  float csf[VEC], snf[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    csf[e] = rotary ? DType<C>::to_f(cs.v[e]) : 0.f;
    snf[e] = rotary ? (second ? DType<C>::to_f(sn.v[e]) : -DType<C>::to_f(sn.v[e])) : 0.f;
  }
  // 16-bit pairs are unpacked word-wise (bf16: one shift / one mask per element; element-wise conversions of the packed
  // struct compiled to a PRMT extract + a shift each - ncu: 22 instructions issued per element, ALU pipe 56 %)
  auto unpack = [&](const Pack<T, VEC>& pk, float (&f)[VEC]) {
    if constexpr (std::is_same<T, __nv_bfloat16>::value && VEC % 2 == 0) {
      const uint32_t* w = reinterpret_cast<const uint32_t*>(&pk);
#pragma unroll
      for (int i = 0; i < VEC / 2; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
      }
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) f[e] = DType<T>::to_f(pk.v[e]);
    }
  };
  auto rotate = [&](Pack<T, VEC>& o, const Pack<T, VEC>& xb) {
    float x[VEC], y[VEC], r[VEC];
    unpack(o, x);
    unpack(xb, y);
    if constexpr (VEC % 2 == 0) {  // two elements per issue slot (FMUL2 / FADD2), the scalar form's roundings lane by lane
#pragma unroll
      for (int e = 0; e < VEC; e += 2) {
        float2 p = mul2(make_float2(x[e], x[e + 1]), make_float2(csf[e], csf[e + 1]));
        float2 q = mul2(make_float2(y[e], y[e + 1]), make_float2(snf[e], snf[e + 1]));
        if (ROUND_T) {
          p = make_float2(round_through<T>(p.x), round_through<T>(p.y));
          q = make_float2(round_through<T>(q.x), round_through<T>(q.y));
        }
        // scalar adds: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (seen in the SASS; the products must be
        // rounded before the sum to reproduce the reference bit for bit)
        r[e] = __fadd_rn(p.x, q.x);
        r[e + 1] = __fadd_rn(p.y, q.y);
      }
    } else {
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      float p = __fmul_rn(x[e], csf[e]), q = __fmul_rn(y[e], snf[e]);
      if (ROUND_T) {
        p = round_through<T>(p);
        q = round_through<T>(q);
      }
      r[e] = __fadd_rn(p, q);
    }
    }
    if constexpr (sizeof(T) == 2 && VEC % 2 == 0) {
      uint32_t* w = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
      for (int i = 0; i < VEC / 2; ++i) {
        if constexpr (std::is_same<T, __nv_bfloat16>::value) {
          const __nv_bfloat162 t = __floats2bfloat162_rn(r[2 * i], r[2 * i + 1]);
          w[i] = *reinterpret_cast<const uint32_t*>(&t);
        } else {
          const __half2 t = __floats2half2_rn(r[2 * i], r[2 * i + 1]);
          w[i] = *reinterpret_cast<const uint32_t*>(&t);
        }
      }
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) o.v[e] = DType<T>::from_f(r[e]);
    }
  };
  Pack<T, VEC> pre[kRopeTrips];
#pragma unroll
  for (int i = 0; i < kRopeTrips; ++i) {
    const int h = hs + i * HSLOTS;
    const T* src = h < a.q_heads ? qs + (int64_t)h * a.q_h : ks + (int64_t)h * a.k_h;
    if (h < total) pre[i] = *reinterpret_cast<const Pack<T, VEC>*>(src + off);
    else memset(&pre[i], 0, sizeof(pre[i]));
  }
#pragma unroll
  for (int i = 0; i < kRopeTrips; ++i) {
    const int h = hs + i * HSLOTS;
    Pack<T, VEC> xb;
    {  // every lane of the warp takes part in the exchange (heads of one warp may straddle `total`)
      const uint32_t* w = reinterpret_cast<const uint32_t*>(&pre[i]);
      uint32_t* d = reinterpret_cast<uint32_t*>(&xb);
#pragma unroll
      for (int k = 0; k < kWords; ++k) d[k] = __shfl_sync(0xffffffffu, w[k], rotary ? partner_lane : (int)(threadIdx.x & 31u));
    }
    if (h < total) {
      T* dst = h < a.q_heads ? qd + (int64_t)h * a.qo_h : kd + (int64_t)h * a.ko_h;
      if (rotary) rotate(pre[i], xb);
      *reinterpret_cast<Pack<T, VEC>*>(dst + off) = pre[i];
    }
  }
#pragma unroll 1
  for (int h = hs + kRopeTrips * HSLOTS; h < total; h += HSLOTS) {
    const T* src = h < a.q_heads ? qs + (int64_t)h * a.q_h : ks + (int64_t)h * a.k_h;
    T* dst = h < a.q_heads ? qd + (int64_t)h * a.qo_h : kd + (int64_t)h * a.ko_h;
    Pack<T, VEC> o = *reinterpret_cast<const Pack<T, VEC>*>(src + off);
    if (rotary) {
      const Pack<T, VEC> xb = *reinterpret_cast<const Pack<T, VEC>*>(src + partner);
      rotate(o, xb);
    }
    *reinterpret_cast<Pack<T, VEC>*>(dst + off) = o;
  }
  }  // token loop
}

template <typename T, typename C, int VEC, bool ROUND_T>
static bool launch_rope_fast(const RopeArgs& a, int64_t tokens, cudaStream_t s) {
  if (a.head_dim % VEC) return false;
  const bool big = tokens >= 1024;  // prefill-sized: fewer, busier threads per token
  const int threads = big ? 128 : 256;
  const int64_t cap = (int64_t)kNumSMs * (1024 / threads) * 3;
  const unsigned grid = (unsigned)(tokens < cap ? tokens : cap);
#define ROPE_FAST(SL)                                                                                                  \
  do {                                                                                                                 \
    if (big) launch_pdl(apply_rope_slice_kernel<T, C, VEC, ROUND_T, SL, 128, 5>, dim3(grid), dim3(128), 0, s, a);      \
    else launch_pdl(apply_rope_slice_kernel<T, C, VEC, ROUND_T, SL, 256, 3>, dim3(grid), dim3(256), 0, s, a);          \
    return true;                                                                                                       \
  } while (0)
  switch (a.head_dim / VEC) {
    case 8: ROPE_FAST(8);
    case 16: ROPE_FAST(16);
    case 32: ROPE_FAST(32);
    default: return false;
  }
#undef ROPE_FAST
}

template <typename T, typename C, bool ROUND_T>
static int launch_rope(const RopeArgs& a, int64_t tokens, int vec, cudaStream_t s) {
  const int64_t work = (int64_t)(a.q_heads + a.k_heads) * ((a.head_dim - a.rope_dim / 2) / vec);
  int threads = work >= 256 ? 256 : (work >= 128 ? 128 : 64);
  // widest vectors and a power-of-two slice count per head (every production shape): slice-per-thread kernel
  if (vec == 16 / (int)sizeof(T) && work >= 192) {
    if (launch_rope_fast<T, C, 16 / (int)sizeof(T), ROUND_T>(a, tokens, s)) return check_launch("apply_rope_slice_kernel");
  }
  switch (vec) {
    case 8: apply_rope_kernel<T, C, (sizeof(T) == 2 ? 8 : 4), ROUND_T><<<(unsigned)tokens, threads, 0, s>>>(a); break;
    case 4: apply_rope_kernel<T, C, 4, ROUND_T><<<(unsigned)tokens, threads, 0, s>>>(a); break;
    case 2: apply_rope_kernel<T, C, 2, ROUND_T><<<(unsigned)tokens, threads, 0, s>>>(a); break;
    default: apply_rope_kernel<T, C, 1, ROUND_T><<<(unsigned)tokens, threads, 0, s>>>(a); break;
  }
  return check_launch("apply_rope_kernel");
}

// positions -> cos/sin rows
__global__ void __launch_bounds__(128) rotary_cos_sin_kernel(
    float* __restrict__ cos_out, float* __restrict__ sin_out, int64_t num_tokens, int rope_dim,
    const float* __restrict__ inv_freq, float scaling, const int32_t* __restrict__ position_ids,
    const int32_t* __restrict__ cu_q, const int32_t* __restrict__ total_lens, int num_seqs, int64_t period,
    const float* __restrict__ table_cos, const float* __restrict__ table_sin, int64_t table_rows) {
  const int64_t tok = blockIdx.x;
  int64_t pos;
  if (position_ids) {
    pos = position_ids[tok];
  } else if (cu_q) {
    int lo = 0, hi = num_seqs + 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cu_q[mid] <= tok) lo = mid + 1; else hi = mid;
    }
    const int seq = lo - 1;
    if (seq < 0 || seq >= num_seqs) {
      pos = -1;  // token outside every sequence keeps the reference's -1 position
    } else {
      const int q_len = cu_q[seq + 1] - cu_q[seq];
      const int ctx = total_lens ? total_lens[seq] - q_len : 0;
      pos = ctx + (tok - cu_q[seq]);
    }
  } else {
    pos = tok % period;
  }
  const int half = rope_dim / 2;
  float* c = cos_out + tok * rope_dim;
  float* s = sin_out + tok * rope_dim;
  if (table_cos) {
    // python-style negative index wraps once; anything still outside the table yields zeros
    int64_t row = pos < 0 ? pos + table_rows : pos;
    const bool ok = row >= 0 && row < table_rows;
    for (int i = threadIdx.x; i < rope_dim; i += blockDim.x) {
      c[i] = ok ? table_cos[row * rope_dim + i] : 0.f;
      s[i] = ok ? table_sin[row * rope_dim + i] : 0.f;
    }
    return;
  }
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float ang = __fmul_rn((float)pos, inv_freq[i]);
    float sv, cv;
    sincosf(ang, &sv, &cv);
    cv = __fmul_rn(cv, scaling);
    sv = __fmul_rn(sv, scaling);
    c[i] = cv; c[i + half] = cv;
    s[i] = sv; s[i + half] = sv;
  }
}

}  // namespace mojo

extern "C" int mojo_b200_apply_rope(const void* q, const void* k, const void* cos, const void* sin, void* q_out,
                                    void* k_out, int64_t batch, int64_t seq, int q_heads, int k_heads, int head_dim,
                                    int rope_dim, int64_t q_stride_b, int64_t q_stride_s, int64_t q_stride_h,
                                    int64_t k_stride_b, int64_t k_stride_s, int64_t k_stride_h, int64_t qo_stride_b,
                                    int64_t qo_stride_s, int64_t qo_stride_h, int64_t ko_stride_b, int64_t ko_stride_s,
                                    int64_t ko_stride_h, int64_t cos_stride_b, int64_t cos_stride_s, int dtype,
                                    int cos_dtype, void* stream) {
  using namespace mojo;
  MOJO_REQUIRE(batch >= 0 && seq >= 0 && q_heads >= 0 && k_heads >= 0 && head_dim > 0, MOJO_B200_EINVAL,
               "apply_rope: bad sizes");
  MOJO_REQUIRE(rope_dim >= 0 && rope_dim <= head_dim && rope_dim % 2 == 0, MOJO_B200_EINVAL,
               "apply_rope: rope_dim %d must be even and <= head_dim %d", rope_dim, head_dim);
  const int64_t tokens = batch * seq;
  if (tokens == 0 || q_heads + k_heads == 0) return 0;
  MOJO_REQUIRE(q && k && cos && sin && q_out && k_out, MOJO_B200_EINVAL, "apply_rope: null tensor pointer");
  MOJO_REQUIRE(tokens <= 0x7fffffffLL, MOJO_B200_EUNSUPPORTED, "apply_rope: too many tokens");
  MOJO_REQUIRE(dtype >= 0 && dtype <= 2 && cos_dtype >= 0 && cos_dtype <= 2, MOJO_B200_EINVAL, "apply_rope: bad dtype");

  RopeArgs a{q, k, cos, sin, q_out, k_out, seq, tokens, q_heads, k_heads, head_dim, rope_dim,
             q_stride_b, q_stride_s, q_stride_h, k_stride_b, k_stride_s, k_stride_h,
             qo_stride_b, qo_stride_s, qo_stride_h, ko_stride_b, ko_stride_s, ko_stride_h, cos_stride_b, cos_stride_s};

  // widest power-of-two element vector every offset, stride and base pointer allows (16 bytes max for T)
  const int eb = dtype_bytes(dtype), cb = dtype_bytes(cos_dtype);
  const int half = rope_dim / 2, nope = head_dim - rope_dim;
  int vec = 16 / eb;
  auto fits = [&](int v) {
    if (half % v || nope % v) return false;
    const int64_t strides[] = {q_stride_b, q_stride_s, q_stride_h, k_stride_b, k_stride_s, k_stride_h,
                               qo_stride_b, qo_stride_s, qo_stride_h, ko_stride_b, ko_stride_s, ko_stride_h,
                               cos_stride_b, cos_stride_s};
    for (int64_t st : strides) if (st % v) return false;
    const uintptr_t tb = (uintptr_t)v * eb - 1, cbm = (uintptr_t)v * cb - 1;
    if (((uintptr_t)q | (uintptr_t)k | (uintptr_t)q_out | (uintptr_t)k_out) & tb) return false;
    if (((uintptr_t)cos | (uintptr_t)sin) & cbm) return false;
    return true;
  };
  while (vec > 1 && !fits(vec)) vec >>= 1;

  cudaStream_t s = (cudaStream_t)stream;
  const bool round_t = (cos_dtype == dtype) && dtype != MOJO_B200_F32;
  return dispatch_dtype(dtype, [&](auto tt) {
    using T = decltype(tt);
    return dispatch_dtype(cos_dtype, [&](auto ct) {
      using C = decltype(ct);
      if constexpr (sizeof(T) == 2 && std::is_same<T, C>::value) {
        return launch_rope<T, C, true>(a, tokens, vec, s);
      } else {
        (void)round_t;
        return launch_rope<T, C, false>(a, tokens, vec, s);
      }
    });
  });
}

extern "C" int mojo_b200_rotary_cos_sin(float* cos_out, float* sin_out, int64_t num_tokens, int rope_dim,
                                        const float* inv_freq, float attention_scaling, const int32_t* position_ids,
                                        const int32_t* cu_q_lens, const int32_t* total_seq_lens, int num_seqs,
                                        int64_t period, const float* table_cos, const float* table_sin,
                                        int64_t table_rows, void* stream) {
  using namespace mojo;
  MOJO_REQUIRE(num_tokens >= 0 && rope_dim > 0 && rope_dim % 2 == 0, MOJO_B200_EINVAL, "rotary: bad sizes");
  if (num_tokens == 0) return 0;
  MOJO_REQUIRE(cos_out && sin_out, MOJO_B200_EINVAL, "rotary: null output");
  MOJO_REQUIRE((table_cos && table_sin && table_rows > 0) || inv_freq, MOJO_B200_EINVAL,
               "rotary: need inv_freq or a cos/sin table");
  MOJO_REQUIRE(!(position_ids && cu_q_lens), MOJO_B200_EINVAL, "rotary: position_ids and cu_q_lens are exclusive");
  MOJO_REQUIRE(position_ids || cu_q_lens || period > 0, MOJO_B200_EINVAL, "rotary: period must be > 0");
  MOJO_REQUIRE(num_tokens <= 0x7fffffffLL, MOJO_B200_EUNSUPPORTED, "rotary: too many tokens");
  rotary_cos_sin_kernel<<<(unsigned)num_tokens, 128, 0, (cudaStream_t)stream>>>(
      cos_out, sin_out, num_tokens, rope_dim, inv_freq, attention_scaling, position_ids, cu_q_lens, total_seq_lens,
      num_seqs, period, table_cos, table_sin, table_rows);
  return check_launch("rotary_cos_sin_kernel");
}
