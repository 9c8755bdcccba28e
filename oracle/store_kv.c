/* TEST INFRASTRUCTURE - plain-C restatement of the reference's integer/byte work on the path:
 *   build_paged_kv_chunk_metadata   mojo_opset/core/operators/kv_cache.py:33-101
 *   MojoStorePagedKVCache.forward   mojo_opset/core/operators/kv_cache.py:161-169
 * Scalar, single threaded, contiguous tensors only.  Checked against tests/golden/store_paged_kv.pt (vectors
 * produced by the unmodified reference) in tests/test_oracle_golden.py.  Never linked into the product. */
#include <stdint.h>
#include <string.h>

/* plan rows (src_token_start, block, offset_in_block, len); returns the number of rows written (<= cap). */
int64_t oracle_build_chunk_plan(const int32_t* block_table, int64_t num_seqs, int64_t max_blocks,
                                const int32_t* cu_q_lens /* NULL = decode */, const int32_t* context_kv_lens,
                                int32_t block_size, int32_t* plan, int64_t cap) {
  int64_t n = 0;
  if (num_seqs == 0 || max_blocks == 0) return 0;
  for (int64_t i = 0; i < num_seqs; ++i) {
    const int32_t ctx = context_kv_lens[i];
    const int32_t* row = block_table + i * max_blocks;
    if (cu_q_lens == NULL) {
      if (ctx < 0) continue;
      const int64_t logical = ctx / block_size;
      if (logical >= max_blocks || row[logical] < 0) continue;
      if (n < cap) {
        plan[4 * n + 0] = (int32_t)i;
        plan[4 * n + 1] = row[logical];
        plan[4 * n + 2] = ctx % block_size;
        plan[4 * n + 3] = 1;
      }
      ++n;
      continue;
    }
    const int32_t q_len = cu_q_lens[i + 1] - cu_q_lens[i];
    if (q_len <= 0 || ctx < 0) continue;
    for (int64_t j = 0; j < max_blocks; ++j) {
      const int64_t blk_lo = j * block_size, blk_hi = blk_lo + block_size;
      const int64_t lo = ctx > blk_lo ? ctx : blk_lo;
      const int64_t hi = (int64_t)ctx + q_len < blk_hi ? (int64_t)ctx + q_len : blk_hi;
      if (hi - lo <= 0 || row[j] < 0) continue;
      if (n < cap) {
        plan[4 * n + 0] = (int32_t)(cu_q_lens[i] + (lo - ctx));
        plan[4 * n + 1] = row[j];
        plan[4 * n + 2] = (int32_t)(lo - blk_lo);
        plan[4 * n + 3] = (int32_t)(hi - lo);
      }
      ++n;
    }
  }
  return n;
}

/* states [T, Hkv, D] -> cache [NB, Hkv, bs, D], elem_bytes per element, rows in plan order. */
void oracle_store_paged_kv(const uint8_t* states, uint8_t* cache, const int32_t* plan, int64_t num_chunks,
                           int64_t num_kv_heads, int64_t head_dim, int64_t block_size, int64_t elem_bytes) {
  const int64_t row = head_dim * elem_bytes;
  for (int64_t c = 0; c < num_chunks; ++c) {
    const int64_t src = plan[4 * c], blk = plan[4 * c + 1], off = plan[4 * c + 2], len = plan[4 * c + 3];
    for (int64_t t = 0; t < len; ++t)
      for (int64_t h = 0; h < num_kv_heads; ++h)
        memcpy(cache + (((blk * num_kv_heads + h) * block_size) + off + t) * row,
               states + ((src + t) * num_kv_heads + h) * row, (size_t)row);
  }
}
