// Fused pre-attention pass of a decoder layer (the reference lists the op shells MojoRoPEStoreKV /
// MojoNormRoPEStoreKV in README.md:128-130 and composes them today from four launches in
// modeling/qwen3/mojo_qwen3_dense.py:229-234 + PagedDummyCache.update :99-109):
//
//   q' = RoPE(RMSNorm_D(q) * wq)   -> q_out [T, Hq, D]
//   k' = RoPE(RMSNorm_D(k) * wk)   -> key_cache page slot of the token   (and k_out if asked for)
//   v                              -> value_cache page slot of the token
//
// in ONE pass over HBM: q, k, v are read once and written once - the unfused chain (2 norms, RoPE, store) moves
// q and k three times and needs a host-built chunk plan.  The slot of a token is found on the device from
// (block_table, cu_q_lens | NULL, context_kv_lens) exactly as build_paged_kv_chunk_metadata does
// (core/operators/kv_cache.py:33-101): tokens with a negative context, a logical block past the table or a
// negative / out-of-range block id are not stored (their q' is still produced).
//
// Rounding points are those of the unfused ops: the normalised row is rounded to T (normalization.py:93-108),
// RoPE multiplies in promote(T, cos dtype) with each product and the sum rounded separately
// (position_embedding.py:128-129), K/V are moved bit-exactly.
//
// One CTA (128 threads) per token (ncu, T = 8192 at Qwen3 shapes: the kernel is ISSUE-bound - 3.15 instructions per
// cycle, 32 issued per element, two thirds of them per-thread set-up and per-trip bookkeeping - so a thread now owns
// six head rows instead of three, one thread looks the page slot up for the CTA, and the norm weights are unpacked once).  A head row (D elements) is owned by D/8 consecutive lanes, 8 elements
// (16 bytes) per lane: the sum of squares is a shuffle reduction inside the lane group, the RoPE partner half sits
// rope_dim/16 lanes away (one shuffle of the packed row).  cos/sin slices are loaded once per thread.
#include <type_traits>

#include "common.cuh"

namespace mojo {

struct NrsArgs {
  const void *q, *k, *v, *wq, *wk, *cos, *sin;
  void *q_out, *k_out, *kc, *vc;
  const int32_t *table, *cu_q, *ctx_lens;
  int64_t table_stride;
  int max_blocks, num_seqs;
  int64_t num_tokens, num_blocks;
  int hq, hkv, head_dim, rope_dim, block_size;
  int64_t q_t, q_h, k_t, k_h, v_t, v_h, qo_t, qo_h, ko_t, ko_h, cos_t;
  int64_t kc_b, kc_h, kc_t, vc_b, vc_h, vc_t;
  float eps;
};

struct alignas(16) Row8 { uint32_t w[4]; };  // 8 sixteen-bit elements

template <typename T> __device__ __forceinline__ void unpack8(const Row8& r, float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (std::is_same<T, __nv_bfloat16>::value) {
      f[2 * i] = __uint_as_float(r.w[i] << 16);
      f[2 * i + 1] = __uint_as_float(r.w[i] & 0xffff0000u);
    } else {
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&r.w[i]));
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
}
template <typename T> __device__ __forceinline__ Row8 pack8(const float (&f)[8]) {
  Row8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (std::is_same<T, __nv_bfloat16>::value) {
      __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      r.w[i] = *reinterpret_cast<uint32_t*>(&t);
    } else {
      __half2 t = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      r.w[i] = *reinterpret_cast<uint32_t*>(&t);
    }
  }
  return r;
}

// LPH = lanes per head = D / 8; C = cos/sin element type (float or T); NORM: apply the per-head RMSNorm
constexpr int kNrsThreads = 128;
#ifndef MOJO_NRS_MIN_CTAS
#define MOJO_NRS_MIN_CTAS 6  // 80 registers, no spill: T = 8192 44.9 -> 42.5 us (7 and 8 spill and lose at decode sizes)
#endif

template <typename T, typename C, int LPH, bool NORM>
__global__ void __launch_bounds__(kNrsThreads, MOJO_NRS_MIN_CTAS) norm_rope_store_kernel(const NrsArgs a) {
  constexpr int D = LPH * 8;
  constexpr int HSLOTS = kNrsThreads / LPH;
  __shared__ int64_t s_slot[2];
  constexpr bool ROUND_T = std::is_same<T, C>::value;  // intermediates of a same-dtype RoPE are rounded to T
  const int64_t tok = blockIdx.x;
  const int sl = threadIdx.x % LPH, hs = threadIdx.x / LPH;
  pdl_wait();
  pdl_trigger();

  // ---- this lane's role inside a head and its cos / sin / norm-weight slices
  const int nope = D - a.rope_dim, half = a.rope_dim / 2;
  const int e0 = sl * 8;
  const bool rotary = e0 >= nope;
  const bool second = e0 >= nope + half;
  const int partner_lane = second ? sl - half / 8 : sl + half / 8;
  float cs[8], sn[8];
  if (rotary) {
    // 8-element slices are 8 * sizeof(C) aligned (checked by the entry point): vector loads
    struct alignas(8 * sizeof(C)) Slice { C v[8]; };
    const Slice cv = *reinterpret_cast<const Slice*>((const C*)a.cos + tok * a.cos_t + (e0 - nope));
    const Slice sv = *reinterpret_cast<const Slice*>((const C*)a.sin + tok * a.cos_t + (e0 - nope));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      cs[i] = DType<C>::to_f(cv.v[i]);
      sn[i] = DType<C>::to_f(sv.v[i]);
    }
  }
  float wqf[8], wkf[8];  // this lane's slices of the two norm weights, unpacked once
  if (NORM) {
    unpack8<T>(*reinterpret_cast<const Row8*>((const T*)a.wq + e0), wqf);
    unpack8<T>(*reinterpret_cast<const Row8*>((const T*)a.wk + e0), wkf);
  }

  // ---- three plain loops - q heads, k heads, v heads - each with its own bumped pointers (the first version walked
  // one mixed head list: per trip ~300 instructions of kind selects and 64-bit address arithmetic for 8 elements per
  // thread - ncu: 32 instructions issued per element, the kernel issue-bound at 0.59 of the HBM peak).  The rows of
  // the first trips are requested BEFORE the page-slot lookup: that lookup is a chain of four to eight dependent loads.
  const int64_t hstep = HSLOTS;
  const T* qp = (const T*)a.q + tok * a.q_t + (int64_t)hs * a.q_h + e0;
  const T* kp = (const T*)a.k + tok * a.k_t + (int64_t)hs * a.k_h + e0;
  const T* vp = (const T*)a.v + tok * a.v_t + (int64_t)hs * a.v_h + e0;
  constexpr int kPreQ = 4, kPreKV = 1;  // Qwen3 / Llama shapes at head_dim 128: the whole token (12 KB) in flight
  Row8 pre_q[kPreQ], pre_k[kPreKV], pre_v[kPreKV];
#pragma unroll
  for (int i = 0; i < kPreQ; ++i) {
    pre_q[i] = Row8{{0u, 0u, 0u, 0u}};
    if (i * HSLOTS + hs < a.hq) pre_q[i] = *reinterpret_cast<const Row8*>(qp + i * hstep * a.q_h);
  }
#pragma unroll
  for (int i = 0; i < kPreKV; ++i) {
    pre_k[i] = Row8{{0u, 0u, 0u, 0u}};
    pre_v[i] = pre_k[i];
    if (i * HSLOTS + hs < a.hkv) {
      pre_k[i] = *reinterpret_cast<const Row8*>(kp + i * hstep * a.k_h);
      pre_v[i] = *reinterpret_cast<const Row8*>(vp + i * hstep * a.v_h);
    }
  }

  // ---- page slot of this token: ONE thread walks cu_q_lens / context lengths / block table while the others' row
  // loads are in flight and the q heads are processed; the barrier that publishes it sits in front of the k heads
  if (threadIdx.x == 0) {
    int64_t kc_off = -1, vc_off = -1;
    int seq = -1, pos = -1;
    if (a.cu_q == nullptr) {  // decode: token i is sequence i at position context_i
      if (tok < a.num_seqs) {
        seq = (int)tok;
        pos = a.ctx_lens[seq];
      }
    } else {
      int lo = 0, hi = a.num_seqs + 1;  // largest seq with cu_q[seq] <= tok; empty sequences skip themselves
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a.cu_q[mid] <= tok) lo = mid + 1; else hi = mid;
      }
      seq = lo - 1;
      if (seq >= 0 && seq < a.num_seqs) {
        const int ctx = a.ctx_lens[seq];
        pos = ctx < 0 ? -1 : ctx + (int)(tok - a.cu_q[seq]);
      } else {
        seq = -1;
      }
    }
    if (seq >= 0 && pos >= 0) {
      const int logical = pos / a.block_size;
      if (logical < a.max_blocks) {
        const int blk = a.table[seq * a.table_stride + logical];
        if (blk >= 0 && blk < a.num_blocks) {
          const int off = pos - logical * a.block_size;
          kc_off = (int64_t)blk * a.kc_b + (int64_t)off * a.kc_t;
          vc_off = (int64_t)blk * a.vc_b + (int64_t)off * a.vc_t;
        }
      }
    }
    s_slot[0] = kc_off;
    s_slot[1] = vc_off;
  }

  const int lane = threadIdx.x & 31;
  const int src_lane = lane - sl + (rotary ? partner_lane : sl);
  // RMSNorm over the head (lane group) + RoPE of one packed row; executed by every lane of the warp (shuffles)
  auto norm_rope = [&](Row8 row, const float (&wf)[8]) -> Row8 {
    float x[8];
    unpack8<T>(row, x);
    if (NORM) {
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) ss = fmaf(x[i], x[i], ss);
#pragma unroll
      for (int o = LPH / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      // MUFU reciprocal square root (<= 2 ulp of fp32; D is a power of two, so the mean is exact)
      const float inv = rsqrtf(__fmaf_rn(ss, 1.0f / (float)D, a.eps));
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = __fmul_rn(__fmul_rn(x[i], inv), wf[i]);
      row = pack8<T>(x);  // the norm's output is materialised in T
      unpack8<T>(row, x);
    }
    Row8 other;  // partner half of the (normalised) row: lane (sl +- half/8) of the same head
#pragma unroll
    for (int i = 0; i < 4; ++i) other.w[i] = __shfl_sync(0xffffffffu, row.w[i], src_lane);
    if (rotary) {
      float y[8], o[8];
      unpack8<T>(other, y);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        // rotate_half(x) = cat(-x2, x1): the partner enters negated in the first half only
        float p = __fmul_rn(x[i], cs[i]), r = __fmul_rn(second ? y[i] : -y[i], sn[i]);
        if (ROUND_T) {
          p = round_through<T>(p);
          r = round_through<T>(r);
        }
        o[i] = __fadd_rn(p, r);
      }
      row = pack8<T>(o);
    }
    return row;
  };

  // ---- q heads
  T* qo = (T*)a.q_out + tok * a.qo_t + (int64_t)hs * a.qo_h + e0;
#pragma unroll
  for (int i = 0; i < kPreQ; ++i) {
    if (i * HSLOTS < a.hq) {  // (uniform over the CTA; every lane of a warp runs the shuffles)
      const Row8 r = norm_rope(pre_q[i], wqf);
      if (i * HSLOTS + hs < a.hq) *reinterpret_cast<Row8*>(qo + i * hstep * a.qo_h) = r;
    }
  }
  for (int h0 = kPreQ * HSLOTS; h0 < a.hq; h0 += HSLOTS) {
    const bool active = h0 + hs < a.hq;
    Row8 row = {{0u, 0u, 0u, 0u}};
    if (active) row = *reinterpret_cast<const Row8*>(qp + (int64_t)h0 * a.q_h);
    const Row8 r = norm_rope(row, wqf);
    if (active) *reinterpret_cast<Row8*>(qo + (int64_t)h0 * a.qo_h) = r;
  }

  __syncthreads();
  const int64_t kc_off = s_slot[0], vc_off = s_slot[1];

  // ---- k heads -> key_cache slot (and k_out), v heads -> value_cache slot (pure copy)
  T* kc_p = (T*)a.kc + kc_off + (int64_t)hs * a.kc_h + e0;
  T* vc_p = (T*)a.vc + vc_off + (int64_t)hs * a.vc_h + e0;
  T* ko = a.k_out ? (T*)a.k_out + tok * a.ko_t + (int64_t)hs * a.ko_h + e0 : nullptr;
  for (int h0 = 0; h0 < a.hkv; h0 += HSLOTS) {
    const bool active = h0 + hs < a.hkv;
    Row8 krow = {{0u, 0u, 0u, 0u}}, vrow = krow;
    if (h0 < kPreKV * HSLOTS) {
      krow = pre_k[0];
      vrow = pre_v[0];
    } else if (active) {
      krow = *reinterpret_cast<const Row8*>(kp + (int64_t)h0 * a.k_h);
      vrow = *reinterpret_cast<const Row8*>(vp + (int64_t)h0 * a.v_h);
    }
    const Row8 r = norm_rope(krow, wkf);
    if (active) {
      if (kc_off >= 0) *reinterpret_cast<Row8*>(kc_p + (int64_t)h0 * a.kc_h) = r;
      if (ko) *reinterpret_cast<Row8*>(ko + (int64_t)h0 * a.ko_h) = r;
      if (vc_off >= 0) *reinterpret_cast<Row8*>(vc_p + (int64_t)h0 * a.vc_h) = vrow;
    }
  }
}

}  // namespace mojo

extern "C" int mojo_b200_norm_rope_store_kv(
    const void* q, const void* k, const void* v, const void* q_norm_weight, const void* k_norm_weight, float eps,
    const void* cos, const void* sin, void* q_out, void* k_out, void* key_cache, void* value_cache,
    const int32_t* block_table, int64_t table_stride, int max_blocks_per_seq, const int32_t* cu_q_lens,
    const int32_t* context_kv_lens, int num_seqs, int64_t num_tokens, int num_q_heads, int num_kv_heads, int head_dim,
    int rope_dim, int64_t num_blocks, int block_size, int64_t q_stride_t, int64_t q_stride_h, int64_t k_stride_t,
    int64_t k_stride_h, int64_t v_stride_t, int64_t v_stride_h, int64_t qo_stride_t, int64_t qo_stride_h,
    int64_t ko_stride_t, int64_t ko_stride_h, int64_t cos_stride_t, int64_t kc_stride_b, int64_t kc_stride_h,
    int64_t kc_stride_t, int64_t vc_stride_b, int64_t vc_stride_h, int64_t vc_stride_t, int dtype, int cos_dtype,
    void* stream) {
  using namespace mojo;
  MOJO_REQUIRE(num_tokens >= 0 && num_q_heads >= 0 && num_kv_heads > 0 && head_dim > 0 && block_size > 0 &&
                   num_blocks >= 0 && num_seqs >= 0 && max_blocks_per_seq >= 0,
               MOJO_B200_EINVAL, "norm_rope_store_kv: bad sizes");
  if (num_tokens == 0) return 0;
  MOJO_REQUIRE(q && k && v && cos && sin && q_out && key_cache && value_cache && context_kv_lens &&
                   (block_table || max_blocks_per_seq == 0),
               MOJO_B200_EINVAL, "norm_rope_store_kv: null tensor pointer");
  MOJO_REQUIRE((q_norm_weight == nullptr) == (k_norm_weight == nullptr), MOJO_B200_EINVAL,
               "norm_rope_store_kv: give both norm weights or neither");
  MOJO_REQUIRE(dtype == MOJO_B200_BF16 || dtype == MOJO_B200_F16, MOJO_B200_EUNSUPPORTED,
               "norm_rope_store_kv: bf16 / fp16 tensors only");
  MOJO_REQUIRE(cos_dtype == MOJO_B200_F32 || cos_dtype == dtype, MOJO_B200_EUNSUPPORTED,
               "norm_rope_store_kv: cos/sin must be fp32 or the tensors' dtype");
  MOJO_REQUIRE(head_dim == 64 || head_dim == 128 || head_dim == 256, MOJO_B200_EUNSUPPORTED,
               "norm_rope_store_kv: head_dim %d not in {64, 128, 256}", head_dim);
  MOJO_REQUIRE(rope_dim >= 0 && rope_dim <= head_dim && rope_dim % 16 == 0 && (head_dim - rope_dim) % 8 == 0,
               MOJO_B200_EUNSUPPORTED, "norm_rope_store_kv: rope_dim %d must be a multiple of 16 (<= head_dim)", rope_dim);
  MOJO_REQUIRE(num_tokens <= 0x7fffffffLL, MOJO_B200_EUNSUPPORTED, "norm_rope_store_kv: too many tokens");
  const int64_t strides[] = {q_stride_t, q_stride_h, k_stride_t, k_stride_h, v_stride_t, v_stride_h, qo_stride_t,
                             qo_stride_h, kc_stride_b, kc_stride_h, kc_stride_t, vc_stride_b, vc_stride_h, vc_stride_t};
  for (int64_t st : strides)
    MOJO_REQUIRE(st % 8 == 0, MOJO_B200_EUNSUPPORTED, "norm_rope_store_kv: strides must be multiples of 8 elements");
  MOJO_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(q_out) && aligned16(key_cache) &&
                   aligned16(value_cache) && (!k_out || (aligned16(k_out) && ko_stride_t % 8 == 0 && ko_stride_h % 8 == 0)) &&
                   (!q_norm_weight || (aligned16(q_norm_weight) && aligned16(k_norm_weight))),
               MOJO_B200_EUNSUPPORTED, "norm_rope_store_kv: tensors must be 16-byte aligned");
  const int cb = dtype_bytes(cos_dtype);
  MOJO_REQUIRE(((uintptr_t)cos % (8 * cb)) == 0 && ((uintptr_t)sin % (8 * cb)) == 0 && cos_stride_t % 8 == 0,
               MOJO_B200_EUNSUPPORTED, "norm_rope_store_kv: cos/sin rows must be aligned to 8 elements");

  NrsArgs a;
  a.q = q; a.k = k; a.v = v; a.wq = q_norm_weight; a.wk = k_norm_weight; a.cos = cos; a.sin = sin;
  a.q_out = q_out; a.k_out = k_out; a.kc = key_cache; a.vc = value_cache;
  a.table = block_table; a.cu_q = cu_q_lens; a.ctx_lens = context_kv_lens; a.table_stride = table_stride;
  a.max_blocks = max_blocks_per_seq; a.num_seqs = num_seqs; a.num_tokens = num_tokens; a.num_blocks = num_blocks;
  a.hq = num_q_heads; a.hkv = num_kv_heads; a.head_dim = head_dim; a.rope_dim = rope_dim; a.block_size = block_size;
  a.q_t = q_stride_t; a.q_h = q_stride_h; a.k_t = k_stride_t; a.k_h = k_stride_h; a.v_t = v_stride_t; a.v_h = v_stride_h;
  a.qo_t = qo_stride_t; a.qo_h = qo_stride_h; a.ko_t = ko_stride_t; a.ko_h = ko_stride_h; a.cos_t = cos_stride_t;
  a.kc_b = kc_stride_b; a.kc_h = kc_stride_h; a.kc_t = kc_stride_t;
  a.vc_b = vc_stride_b; a.vc_h = vc_stride_h; a.vc_t = vc_stride_t;
  a.eps = eps;

  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = (unsigned)num_tokens;
  const bool norm = q_norm_weight != nullptr;
  const bool cos_f32 = cos_dtype == MOJO_B200_F32;
#define NRS_LAUNCH(TT, CC, LL)                                                                \
  do {                                                                                        \
    if (norm) launch_pdl(norm_rope_store_kernel<TT, CC, LL, true>, dim3(grid), dim3(kNrsThreads), 0, s, a);  \
    else launch_pdl(norm_rope_store_kernel<TT, CC, LL, false>, dim3(grid), dim3(kNrsThreads), 0, s, a);      \
  } while (0)
#define NRS_DIM(TT, CC)                                        \
  do {                                                         \
    if (head_dim == 64) NRS_LAUNCH(TT, CC, 8);                 \
    else if (head_dim == 128) NRS_LAUNCH(TT, CC, 16);          \
    else NRS_LAUNCH(TT, CC, 32);                               \
  } while (0)
  if (dtype == MOJO_B200_BF16) {
    if (cos_f32) NRS_DIM(__nv_bfloat16, float); else NRS_DIM(__nv_bfloat16, __nv_bfloat16);
  } else {
    if (cos_f32) NRS_DIM(__half, float); else NRS_DIM(__half, __half);
  }
#undef NRS_DIM
#undef NRS_LAUNCH
  return check_launch("norm_rope_store_kernel");
}
