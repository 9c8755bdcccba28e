// MojoStorePagedKVCache: scatter new K/V tokens [T,Hkv,D] into pages [NB,Hkv,bs,D].
// Pure byte movement (bit exact); HBM-bound: bytes = 2 * (read + write) of the new tokens + the plan.
//
//  * chunk-plan path  : one CTA column per chunk row (src_token_start, block, offset, len); the (token,
//                       head, 16-byte vector) space of the chunk is spread over the CTA's threads, K and V
//                       moved by the same thread so two independent 128-bit loads are in flight.
//  * table path       : one CTA per new token; the token finds its sequence (binary search in cu_q_lens),
//                       its logical block and slot on the device - no plan tensor, no host sync.
#include "common.cuh"

namespace mojo {

struct StoreStrides {
  int64_t ks_t, ks_h, vs_t, vs_h;
  int64_t kc_b, kc_h, kc_t, vc_b, vc_h, vc_t;
};

template <int VB> struct BytesVec;
template <> struct BytesVec<16> { using type = int4; };
template <> struct BytesVec<8> { using type = int2; };
template <> struct BytesVec<4> { using type = int; };
template <> struct BytesVec<2> { using type = short; };

// Copies one token's heads (K and V) for vector indices [first, first+step, ...) of the Hkv*vecs_per_row space.
template <int VB>
__device__ __forceinline__ void copy_token_heads(const char* __restrict__ ks, const char* __restrict__ vs,
                                                 char* __restrict__ kc, char* __restrict__ vc, int num_kv_heads,
                                                 int vecs_per_row, int64_t ks_h, int64_t vs_h, int64_t kc_h,
                                                 int64_t vc_h, int first, int step) {
  using V = typename BytesVec<VB>::type;
  const int total = num_kv_heads * vecs_per_row;
  for (int i = first; i < total; i += step) {
    const int h = i / vecs_per_row;
    const int c = i - h * vecs_per_row;
    const V kv = *reinterpret_cast<const V*>(ks + h * ks_h + (int64_t)c * VB);
    const V vv = *reinterpret_cast<const V*>(vs + h * vs_h + (int64_t)c * VB);
    *reinterpret_cast<V*>(kc + h * kc_h + (int64_t)c * VB) = kv;
    *reinterpret_cast<V*>(vc + h * vc_h + (int64_t)c * VB) = vv;
  }
}

// strides below are in BYTES
template <int VB>
__global__ void __launch_bounds__(256) store_chunks_kernel(
    const char* __restrict__ ks, const char* __restrict__ vs, char* __restrict__ kc, char* __restrict__ vc,
    const int32_t* __restrict__ plan, int64_t num_tokens, int num_kv_heads, int row_bytes, int64_t num_blocks,
    int block_size, StoreStrides st) {
  pdl_wait();
  pdl_trigger();
  const int4 row = reinterpret_cast<const int4*>(plan)[blockIdx.x];
  const int src = row.x, blk = row.y, off = row.z, len = row.w;
  if (len <= 0 || src < 0 || (int64_t)src + len > num_tokens) return;
  if (blk < 0 || blk >= num_blocks || off < 0 || off + len > block_size) return;

  const int vecs_per_row = row_bytes / VB;
  const int per_token = num_kv_heads * vecs_per_row;
  const int total = len * per_token;
  const char* ks0 = ks + (int64_t)src * st.ks_t;
  const char* vs0 = vs + (int64_t)src * st.vs_t;
  char* kc0 = kc + (int64_t)blk * st.kc_b + (int64_t)off * st.kc_t;
  char* vc0 = vc + (int64_t)blk * st.vc_b + (int64_t)off * st.vc_t;
  using V = typename BytesVec<VB>::type;
  // four (K, V) vector pairs requested per thread before the first is stored: a prefill chunk is one page per CTA
  // column (512 CTAs at T = 8192), and with one pair in flight per thread an SM held ~28 KB of loads - too little to
  // cover the HBM latency (0.74 of the peak)
  constexpr int kUnroll = 4;
  const int stride = gridDim.y * blockDim.x;
  for (int i0 = blockIdx.y * blockDim.x + threadIdx.x; i0 < total; i0 += kUnroll * stride) {
    V kv[kUnroll], vv[kUnroll];
    int64_t ko[kUnroll], vo[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int i = i0 + u * stride;
      if (i < total) {
        const int j = i / per_token;
        const int r = i - j * per_token;
        const int h = r / vecs_per_row;
        const int c = r - h * vecs_per_row;
        kv[u] = *reinterpret_cast<const V*>(ks0 + j * st.ks_t + h * st.ks_h + (int64_t)c * VB);
        vv[u] = *reinterpret_cast<const V*>(vs0 + j * st.vs_t + h * st.vs_h + (int64_t)c * VB);
        ko[u] = j * st.kc_t + h * st.kc_h + (int64_t)c * VB;
        vo[u] = j * st.vc_t + h * st.vc_h + (int64_t)c * VB;
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (i0 + u * stride < total) {
        *reinterpret_cast<V*>(kc0 + ko[u]) = kv[u];
        *reinterpret_cast<V*>(vc0 + vo[u]) = vv[u];
      }
    }
  }
}

template <int VB>
__global__ void __launch_bounds__(128) store_table_kernel(
    const char* __restrict__ ks, const char* __restrict__ vs, char* __restrict__ kc, char* __restrict__ vc,
    const int32_t* __restrict__ table, int64_t table_stride, int max_blocks, const int32_t* __restrict__ cu_q,
    const int32_t* __restrict__ ctx_lens, int num_seqs, int64_t num_tokens, int num_kv_heads, int row_bytes,
    int64_t num_blocks, int block_size, StoreStrides st) {
  const int64_t tok = blockIdx.x;
  pdl_wait();
  pdl_trigger();
  // The token's rows are requested BEFORE its page slot is looked up: the lookup is a chain of four to six dependent
  // loads (binary search over cu_q_lens, context length, block table) and a CTA moves only 4 KB, so with the rows
  // issued after it the CTA's life was almost all lookup latency (0.70 of the HBM peak at T = 8192).
  using V = typename BytesVec<VB>::type;
  const int vecs_per_row = row_bytes / VB;
  const int total = num_kv_heads * vecs_per_row;
  constexpr int kPre = 2;
  V pk[kPre], pv[kPre];
  const char* ks0 = ks + tok * st.ks_t;
  const char* vs0 = vs + tok * st.vs_t;
#pragma unroll
  for (int t = 0; t < kPre; ++t) {
    const int i = threadIdx.x + t * blockDim.x;
    if (i < total) {
      const int h = i / vecs_per_row, c = i - h * vecs_per_row;
      pk[t] = *reinterpret_cast<const V*>(ks0 + h * st.ks_h + (int64_t)c * VB);
      pv[t] = *reinterpret_cast<const V*>(vs0 + h * st.vs_h + (int64_t)c * VB);
    }
  }
  int seq, pos;
  if (cu_q == nullptr) {  // decode: token i belongs to sequence i at position context_i
    if (tok >= num_seqs) return;
    seq = (int)tok;
    pos = ctx_lens[seq];
    if (pos < 0) return;
  } else {
    // largest seq with cu_q[seq] <= tok (upper_bound - 1); empty sequences are skipped automatically
    int lo = 0, hi = num_seqs + 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cu_q[mid] <= tok) lo = mid + 1; else hi = mid;
    }
    seq = lo - 1;
    if (seq < 0 || seq >= num_seqs) return;
    const int ctx = ctx_lens[seq];
    if (ctx < 0) return;
    pos = ctx + (int)(tok - cu_q[seq]);
  }
  const int logical = pos / block_size;
  if (logical >= max_blocks) return;
  const int blk = table[seq * table_stride + logical];
  if (blk < 0 || blk >= num_blocks) return;
  const int off = pos - logical * block_size;
  char* kc0 = kc + (int64_t)blk * st.kc_b + (int64_t)off * st.kc_t;
  char* vc0 = vc + (int64_t)blk * st.vc_b + (int64_t)off * st.vc_t;
#pragma unroll
  for (int t = 0; t < kPre; ++t) {
    const int i = threadIdx.x + t * blockDim.x;
    if (i < total) {
      const int h = i / vecs_per_row, c = i - h * vecs_per_row;
      *reinterpret_cast<V*>(kc0 + h * st.kc_h + (int64_t)c * VB) = pk[t];
      *reinterpret_cast<V*>(vc0 + h * st.vc_h + (int64_t)c * VB) = pv[t];
    }
  }
  copy_token_heads<VB>(ks0, vs0, kc0, vc0, num_kv_heads, vecs_per_row, st.ks_h, st.vs_h, st.kc_h, st.vc_h,
                       threadIdx.x + kPre * blockDim.x, blockDim.x);
}

static int pick_vec_bytes(int row_bytes, const StoreStrides& s, const void* a, const void* b, const void* c,
                          const void* d) {
  uintptr_t bits = (uintptr_t)row_bytes | (uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d;
  const int64_t all[] = {s.ks_t, s.ks_h, s.vs_t, s.vs_h, s.kc_b, s.kc_h, s.kc_t, s.vc_b, s.vc_h, s.vc_t};
  for (int64_t v : all) bits |= (uintptr_t)v;
  if ((bits & 15) == 0) return 16;
  if ((bits & 7) == 0) return 8;
  if ((bits & 3) == 0) return 4;
  return 2;
}

static int validate_store(const void* ks, const void* vs, void* kc, void* vc, int64_t num_tokens, int num_kv_heads,
                          int head_dim, int64_t num_blocks, int block_size, int dtype) {
  MOJO_REQUIRE(dtype >= 0 && dtype <= 2, MOJO_B200_EINVAL, "store_paged_kv: bad dtype %d", dtype);
  MOJO_REQUIRE(num_tokens >= 0 && num_kv_heads > 0 && head_dim > 0 && num_blocks >= 0 && block_size > 0,
               MOJO_B200_EINVAL, "store_paged_kv: bad sizes");
  if (num_tokens > 0 && num_blocks > 0)
    MOJO_REQUIRE(ks && vs && kc && vc, MOJO_B200_EINVAL, "store_paged_kv: null tensor pointer");
  return 0;
}

static StoreStrides to_bytes(int eb, int64_t ks_t, int64_t ks_h, int64_t vs_t, int64_t vs_h, int64_t kc_b,
                             int64_t kc_h, int64_t kc_t, int64_t vc_b, int64_t vc_h, int64_t vc_t) {
  return StoreStrides{ks_t * eb, ks_h * eb, vs_t * eb, vs_h * eb, kc_b * eb, kc_h * eb, kc_t * eb,
                      vc_b * eb, vc_h * eb, vc_t * eb};
}

}  // namespace mojo

extern "C" int mojo_b200_store_paged_kv_chunks(
    const void* key_states, const void* value_states, void* key_cache, void* value_cache,
    const int32_t* chunk_metadata, int64_t num_chunks, int64_t num_tokens, int num_kv_heads, int head_dim,
    int64_t num_blocks, int block_size, int64_t ks_stride_t, int64_t ks_stride_h, int64_t vs_stride_t,
    int64_t vs_stride_h, int64_t kc_stride_b, int64_t kc_stride_h, int64_t kc_stride_t, int64_t vc_stride_b,
    int64_t vc_stride_h, int64_t vc_stride_t, int dtype, void* stream) {
  using namespace mojo;
  if (int rc = validate_store(key_states, value_states, key_cache, value_cache, num_tokens, num_kv_heads, head_dim,
                              num_blocks, block_size, dtype))
    return rc;
  MOJO_REQUIRE(num_chunks >= 0, MOJO_B200_EINVAL, "store_paged_kv: negative chunk count");
  if (num_chunks == 0 || num_tokens == 0 || num_blocks == 0) return 0;
  MOJO_REQUIRE(chunk_metadata != nullptr && aligned16(chunk_metadata), MOJO_B200_EINVAL,
               "store_paged_kv: chunk_metadata must be a 16-byte aligned contiguous int32 [C,4]");
  MOJO_REQUIRE(num_chunks <= 0x7fffffffLL, MOJO_B200_EUNSUPPORTED, "store_paged_kv: too many chunks");
  const int eb = dtype_bytes(dtype);
  const StoreStrides st = to_bytes(eb, ks_stride_t, ks_stride_h, vs_stride_t, vs_stride_h, kc_stride_b, kc_stride_h,
                                   kc_stride_t, vc_stride_b, vc_stride_h, vc_stride_t);
  const int row_bytes = head_dim * eb;
  const int vb = pick_vec_bytes(row_bytes, st, key_states, value_states, key_cache, value_cache);
  // a chunk holds at most block_size tokens; spread big pages over several CTAs (surplus CTAs exit at once)
  const int64_t max_vecs = (int64_t)block_size * num_kv_heads * (row_bytes / vb);
  int ysplit = (int)((max_vecs + 256 * 8 - 1) / (256 * 8));
  ysplit = ysplit < 1 ? 1 : (ysplit > 64 ? 64 : ysplit);
  dim3 grid((unsigned)num_chunks, (unsigned)ysplit);
  cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH(VB)                                                                                              \
  launch_pdl(store_chunks_kernel<VB>, grid, dim3(256), 0, s, (const char*)key_states, (const char*)value_states, \
             (char*)key_cache, (char*)value_cache, chunk_metadata, num_tokens, num_kv_heads, row_bytes, num_blocks, \
             block_size, st)
  switch (vb) {
    case 16: LAUNCH(16); break;
    case 8: LAUNCH(8); break;
    case 4: LAUNCH(4); break;
    default: LAUNCH(2); break;
  }
#undef LAUNCH
  return check_launch("store_chunks_kernel");
}

extern "C" int mojo_b200_store_paged_kv_table(
    const void* key_states, const void* value_states, void* key_cache, void* value_cache,
    const int32_t* block_table, int64_t table_stride, int max_blocks_per_seq, const int32_t* cu_q_lens,
    const int32_t* context_kv_lens, int num_seqs, int64_t num_tokens, int num_kv_heads, int head_dim,
    int64_t num_blocks, int block_size, int64_t ks_stride_t, int64_t ks_stride_h, int64_t vs_stride_t,
    int64_t vs_stride_h, int64_t kc_stride_b, int64_t kc_stride_h, int64_t kc_stride_t, int64_t vc_stride_b,
    int64_t vc_stride_h, int64_t vc_stride_t, int dtype, void* stream) {
  using namespace mojo;
  if (int rc = validate_store(key_states, value_states, key_cache, value_cache, num_tokens, num_kv_heads, head_dim,
                              num_blocks, block_size, dtype))
    return rc;
  MOJO_REQUIRE(num_seqs >= 0 && max_blocks_per_seq >= 0, MOJO_B200_EINVAL, "store_paged_kv: bad table sizes");
  if (num_tokens == 0 || num_seqs == 0 || num_blocks == 0 || max_blocks_per_seq == 0) return 0;
  MOJO_REQUIRE(block_table && context_kv_lens, MOJO_B200_EINVAL, "store_paged_kv: null table / context lens");
  MOJO_REQUIRE(num_tokens <= 0x7fffffffLL, MOJO_B200_EUNSUPPORTED, "store_paged_kv: too many tokens");
  const int eb = dtype_bytes(dtype);
  const StoreStrides st = to_bytes(eb, ks_stride_t, ks_stride_h, vs_stride_t, vs_stride_h, kc_stride_b, kc_stride_h,
                                   kc_stride_t, vc_stride_b, vc_stride_h, vc_stride_t);
  const int row_bytes = head_dim * eb;
  const int vb = pick_vec_bytes(row_bytes, st, key_states, value_states, key_cache, value_cache);
  cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH(VB)                                                                                               \
  launch_pdl(store_table_kernel<VB>, dim3((unsigned)num_tokens), dim3(128), 0, s, (const char*)key_states,         \
             (const char*)value_states, (char*)key_cache, (char*)value_cache, block_table, table_stride,           \
             max_blocks_per_seq, cu_q_lens, context_kv_lens, num_seqs, num_tokens, num_kv_heads, row_bytes,         \
             num_blocks, block_size, st)
  switch (vb) {
    case 16: LAUNCH(16); break;
    case 8: LAUNCH(8); break;
    case 4: LAUNCH(4); break;
    default: LAUNCH(2); break;
  }
#undef LAUNCH
  return check_launch("store_table_kernel");
}
