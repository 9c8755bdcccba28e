// Host-side interface of the tcgen05/TMEM attention forward (attention_fwd_sm100.cu), shared with the entry
// points in attention_fwd.cu that choose between it and the mma.sync general path.
#pragma once

#include <climits>

#include "common.cuh"

namespace mojo {

constexpr int kAttnNotEligible = INT_MIN;  // shape not covered by the tcgen05 kernel: take the general path

struct AttnSm100Args {
  // query / out addressed as [batch, rows, heads, D] through element strides (paged: batch = 1, rows = T)
  const void* q;
  void* out;
  int64_t q_rows, q_sb, q_st, q_sh, o_sb, o_st, o_sh;
  // K/V as [blocks, heads, tokens, D] (paged cache) or [batch, heads, kv_len, D] (dense)
  const void* k;
  const void* v;
  int64_t k_b, k_h, k_t, v_b, v_h, v_t;
  int64_t rows_per_block, num_blocks;
  const int32_t* cu_q;    // paged only
  const int32_t* cu_kv;   // paged only, may be null
  const int32_t* tables;  // paged only
  int64_t table_stride;
  int max_blocks;
  int batch, num_q_heads, num_kv_heads, head_dim;
  int64_t max_q_len, q_len_dense, kv_len_dense;
  float softmax_scale;
  int interleave, causal, dense, round_scores, dtype;
  int win_local, win_global;  // sliding window (causal only); -1 = not set
  // MojoSdpa attn_mask (dense, non-causal only): bool bytes, 1 = the key takes part; element strides per batch / head /
  // query row (0 = broadcast), keys contiguous; null = no mask
  const uint8_t* mask;
  int64_t mask_sb, mask_sh, mask_sq;
  // packed K/V (non-paged MojoSWA): key / value are [total_kv_tokens, heads, D]; sequence b owns rows
  // cu_kv[b] .. cu_kv[b+1] (k_t = token stride, k_h = head stride, rows_per_block = total_kv_tokens, num_blocks = 1)
  int packed;
};

// 0 = launched; kAttnNotEligible = not covered (nothing was launched); anything else = error code.
int launch_attn_sm100(const AttnSm100Args& a, cudaStream_t stream);

}  // namespace mojo
