/*
 * mojo_b200.h - C ABI of libmojo_b200.so, the B200 (sm_100a) kernels behind Mojo Opset's
 * paged-attention decoder hot path.
 *
 * The reference (XPU-Forces/mojo_opset) is pure Python: a backend plugs in as a Python class per op
 * (mojo_opset/core/backend_registry.py:48-91) whose forward() hands tensors to a kernel launcher.  The
 * entry points below are what such a backend's launchers bind through ctypes - one per reference op
 * forward, cited on each declaration.  INTEGRATION.md shows the Python-side stub.
 *
 * Conventions
 *   - plain C: raw DEVICE pointers, explicit sizes and element strides, no torch types;
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*): no allocation, no
 *     synchronisation, no host read of device data -> CUDA-graph capturable;
 *   - return 0 on success, a negative MOJO_B200_E* code for a rejected request, or a positive
 *     cudaError_t; mojo_b200_last_error() gives the message (thread local);
 *   - `dtype` is a mojo_b200_dtype; strides are in ELEMENTS; the innermost (feature) dim is always
 *     contiguous;
 *   - there is no CPU path: without an sm_100 device the calls fail with the CUDA error.
 */
#ifndef MOJO_B200_H_
#define MOJO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOJO_B200_ABI_VERSION 1

#if defined(__GNUC__)
#define MOJO_B200_API __attribute__((visibility("default")))
#else
#define MOJO_B200_API
#endif

typedef enum {
  MOJO_B200_BF16 = 0,
  MOJO_B200_F16 = 1,
  MOJO_B200_F32 = 2
} mojo_b200_dtype;

enum {
  MOJO_B200_OK = 0,
  MOJO_B200_EINVAL = -1,       /* malformed argument (null pointer, negative size, misaligned stride) */
  MOJO_B200_EUNSUPPORTED = -2, /* valid request this build has no kernel for -> NotImplementedError */
  MOJO_B200_EWORKSPACE = -3    /* workspace too small */
};

MOJO_B200_API int mojo_b200_abi_version(void);
MOJO_B200_API const char* mojo_b200_last_error(void);
/* 1 if the current device is compute capability 10.x, 0 otherwise, <0 / >0 on error. */
MOJO_B200_API int mojo_b200_device_ok(void);

/* Device error word.  The reference raises ValueError when a sequence with kv_len > 0 has no first block
 * (mojo_opset/core/operators/attention.py:186-187 decode, :396-397 prefill) - a host read of the block table.  Here the
 * caller registers one int32 of DEVICE memory per device (null = none); the attention kernels OR a bit into it when they
 * meet such a row (and treat the unmapped keys as zeros); the caller reads and clears it when convenient. */
#define MOJO_B200_ERR_DECODE_UNMAPPED_BLOCK 1
#define MOJO_B200_ERR_PREFILL_UNMAPPED_BLOCK 2
MOJO_B200_API int mojo_b200_set_error_word(int* device_word);

/* Split-KV decode (mojo_b200_paged_decode_gqa / _swa with more than one split): `count` ints of ZERO-INITIALISED device
 * memory on the current device, used as arrival counters so that the last split of a (sequence, kv head) group folds
 * the group's partials inside the main kernel (no second launch).  The kernels leave the words zero; a launch uses
 * batch * num_kv_heads * ceil(group / 16) of them and consecutive launches rotate over up to 16 disjoint slices.
 * Without a registration (or with too few words) the partials are folded by a second kernel.  NULL / 0 unregisters. */
MOJO_B200_API int mojo_b200_set_decode_tickets(int* device_words, int64_t count);

/* ---------------------------------------------------------------------------------------------------
 * MojoStorePagedKVCache.forward                     mojo_opset/core/operators/kv_cache.py:110-171
 *
 * key_states/value_states [T, Hkv, D]  ->  key_cache/value_cache [NB, Hkv, bs, D]  (in place, bit exact)
 * chunk plan rows: (src_token_start, dst_block_id, dst_block_offset, chunk_len), int32 [C, 4].
 * Rows whose block id is outside [0, NB) or whose token range leaves [0, T) are skipped.
 * ------------------------------------------------------------------------------------------------- */
MOJO_B200_API int mojo_b200_store_paged_kv_chunks(
    const void* key_states, const void* value_states, void* key_cache, void* value_cache,
    const int32_t* chunk_metadata, int64_t num_chunks,
    int64_t num_tokens, int num_kv_heads, int head_dim, int64_t num_blocks, int block_size,
    int64_t ks_stride_t, int64_t ks_stride_h, int64_t vs_stride_t, int64_t vs_stride_h,
    int64_t kc_stride_b, int64_t kc_stride_h, int64_t kc_stride_t,
    int64_t vc_stride_b, int64_t vc_stride_h, int64_t vc_stride_t,
    int dtype, void* stream);

/* Same op through the legacy (block_table, cu_q_lens | NULL, context_kv_lens) triple
 * (kv_cache.py:142-150 + build_paged_kv_chunk_metadata :33-101) WITHOUT materialising the plan: one
 * thread group per new token finds its sequence and slot on the device.  cu_q_lens == NULL is decode
 * mode (one token per sequence, T == num_seqs). */
MOJO_B200_API int mojo_b200_store_paged_kv_table(
    const void* key_states, const void* value_states, void* key_cache, void* value_cache,
    const int32_t* block_table, int64_t table_stride, int max_blocks_per_seq,
    const int32_t* cu_q_lens, const int32_t* context_kv_lens, int num_seqs,
    int64_t num_tokens, int num_kv_heads, int head_dim, int64_t num_blocks, int block_size,
    int64_t ks_stride_t, int64_t ks_stride_h, int64_t vs_stride_t, int64_t vs_stride_h,
    int64_t kc_stride_b, int64_t kc_stride_h, int64_t kc_stride_t,
    int64_t vc_stride_b, int64_t vc_stride_h, int64_t vc_stride_t,
    int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * MojoRMSNorm.forward                               mojo_opset/core/operators/normalization.py:93-108
 * MojoResidualAddRMSNorm.forward                    mojo_opset/core/operators/normalization.py:340-359
 *
 * rows x hidden, fp32 math, one rounding.  residual_add: sum = x + residual rounded to dtype first;
 * sum_out may be NULL (norm_pos="post": only y is returned).  weight has the same dtype as x.
 * ------------------------------------------------------------------------------------------------- */
MOJO_B200_API int mojo_b200_rms_norm(const void* x, const void* weight, void* y, int64_t rows, int hidden,
                       int64_t x_row_stride, int64_t y_row_stride, float eps, int dtype, void* stream);
MOJO_B200_API int mojo_b200_residual_add_rms_norm(const void* x, const void* residual, const void* weight, void* y,
                                    void* sum_out, int64_t rows, int hidden, int64_t x_row_stride,
                                    int64_t res_row_stride, int64_t y_row_stride, int64_t sum_row_stride,
                                    float eps, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * MojoApplyRoPE.forward                             mojo_opset/core/operators/position_embedding.py:137-175
 *
 * Rotate-half on the LAST rope_dim features of every (batch, position, head) row of q and k.
 * q/k are addressed as [batch, seq, heads, D] through strides, which covers [T,N,D], [N,T,D],
 * [B,S,N,D] and [B,N,S,D] views alike; cos/sin as [batch, seq, rope_dim] (batch stride 0 to broadcast).
 * cos_dtype == F32 with 16-bit q/k: fp32 math; cos_dtype == dtype: math in that dtype with every
 * product and the sum rounded, as the eager golden does.
 * ------------------------------------------------------------------------------------------------- */
MOJO_B200_API int mojo_b200_apply_rope(const void* q, const void* k, const void* cos, const void* sin, void* q_out,
                         void* k_out, int64_t batch, int64_t seq, int q_heads, int k_heads, int head_dim,
                         int rope_dim,
                         int64_t q_stride_b, int64_t q_stride_s, int64_t q_stride_h,
                         int64_t k_stride_b, int64_t k_stride_s, int64_t k_stride_h,
                         int64_t qo_stride_b, int64_t qo_stride_s, int64_t qo_stride_h,
                         int64_t ko_stride_b, int64_t ko_stride_s, int64_t ko_stride_h,
                         int64_t cos_stride_b, int64_t cos_stride_s, int dtype, int cos_dtype, void* stream);

/* MojoRotaryEmbedding.forward                       mojo_opset/core/operators/position_embedding.py:43-95
 * cos/sin [T, rope_dim] fp32 for T tokens.  Positions come from exactly one of
 *   position_ids[T]                      (decode / explicit),
 *   cu_q_lens[num_seqs+1] (+ optional total_seq_lens[num_seqs]: position = total - q_len + t), or
 *   neither: position = token_index % period  (padded prefill, period = S).
 * If table_cos/table_sin are non-NULL they are [table_rows, rope_dim] fp32 tables to gather from,
 * otherwise angle = position * inv_freq[rope_dim/2] is evaluated in fp32 and scaled by attention_scaling. */
MOJO_B200_API int mojo_b200_rotary_cos_sin(float* cos_out, float* sin_out, int64_t num_tokens, int rope_dim,
                             const float* inv_freq, float attention_scaling,
                             const int32_t* position_ids, const int32_t* cu_q_lens,
                             const int32_t* total_seq_lens, int num_seqs, int64_t period,
                             const float* table_cos, const float* table_sin, int64_t table_rows, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * MojoSwiGLU.forward / MojoSilu.forward             mojo_opset/core/operators/activation.py:43-63, :21-35
 * [rows, cols] with per-tensor row strides (gate/up are often two halves of one fused projection).
 * swiglu_limit > 0: up clamped to [-limit, limit], gate to <= limit first.
 * ------------------------------------------------------------------------------------------------- */
MOJO_B200_API int mojo_b200_swiglu(const void* gate, const void* up, void* out, int64_t rows, int64_t cols,
                     int64_t gate_row_stride, int64_t up_row_stride, int64_t out_row_stride,
                     float swiglu_limit, int dtype, void* stream);
MOJO_B200_API int mojo_b200_silu(const void* x, void* out, int64_t rows, int64_t cols, int64_t x_row_stride,
                   int64_t out_row_stride, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * MojoPagedDecodeGQA.forward                        mojo_opset/core/operators/attention.py:141-229
 *
 * query [B, Hq, D] (strides b,h), caches [NB, Hkv, bs, D], total_seq_lens [B], block_tables [B, MB]
 * (row stride table_stride) -> out [B, Hq, D].  Rows with seq_len <= 0 get zeros.  gqa_interleave: 0 =
 * AABB (kv = h / G), 1 = ABAB (kv = h % Hkv).  Split-KV: `num_splits` partitions of each sequence's KV
 * are reduced by a second kernel; the partials live in `workspace` (mojo_b200_paged_decode_workspace_bytes).
 * num_splits <= 0 lets the library choose from (B, Hkv, max_seq_len); max_seq_len is a host-side hint
 * (<= MB * bs), never read from the device.
 * ------------------------------------------------------------------------------------------------- */
MOJO_B200_API int mojo_b200_paged_decode_num_splits(int batch, int num_q_heads, int num_kv_heads, int head_dim,
                                      int block_size, int64_t max_seq_len, int dtype);
MOJO_B200_API size_t mojo_b200_paged_decode_workspace_bytes(int batch, int num_q_heads, int head_dim, int num_splits);
MOJO_B200_API int mojo_b200_paged_decode_gqa(
    const void* query, const void* key_cache, const void* value_cache, const int32_t* total_seq_lens,
    const int32_t* block_tables, void* out, void* workspace, size_t workspace_bytes,
    int batch, int num_q_heads, int num_kv_heads, int head_dim, int64_t num_blocks, int block_size,
    int max_blocks_per_seq, int64_t table_stride, int64_t max_seq_len,
    int64_t q_stride_b, int64_t q_stride_h, int64_t o_stride_b, int64_t o_stride_h,
    int64_t kc_stride_b, int64_t kc_stride_h, int64_t kc_stride_t,
    int64_t vc_stride_b, int64_t vc_stride_h, int64_t vc_stride_t,
    float softmax_scale, int gqa_interleave, int num_splits, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * MojoPagedDecodeSWA.forward                        mojo_opset/core/operators/attention.py:645-745
 *
 * mojo_b200_paged_decode_gqa with the window rule of _generate_window_mask (:507-531) for the single query token at
 * position seq_len - 1: key k is visible iff k + local_window_size >= position or k < global_window_size (-1 = None).
 * Only the KV tiles of the global prefix and of the local window are read (O(window) traffic); the splits divide
 * those.  Built on the tensor-tile kernel (bf16/fp16, head_dim 64/128): other shapes return MOJO_B200_EUNSUPPORTED
 * (the host mirror then uses mojo_b200_paged_prefill_swa with one query row per sequence).  num_splits and the
 * workspace: as for mojo_b200_paged_decode_gqa, sized for max_seq_len.
 * ------------------------------------------------------------------------------------------------- */
MOJO_B200_API int mojo_b200_paged_decode_swa(
    const void* query, const void* key_cache, const void* value_cache, const int32_t* total_seq_lens,
    const int32_t* block_tables, void* out, void* workspace, size_t workspace_bytes,
    int batch, int num_q_heads, int num_kv_heads, int head_dim, int64_t num_blocks, int block_size,
    int max_blocks_per_seq, int64_t table_stride, int64_t max_seq_len,
    int64_t q_stride_b, int64_t q_stride_h, int64_t o_stride_b, int64_t o_stride_h,
    int64_t kc_stride_b, int64_t kc_stride_h, int64_t kc_stride_t,
    int64_t vc_stride_b, int64_t vc_stride_h, int64_t vc_stride_t,
    float softmax_scale, int gqa_interleave, int num_splits, int local_window_size, int global_window_size,
    int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * MojoPagedPrefillGQA.forward                       mojo_opset/core/operators/attention.py:335-447
 *
 * query [T, Hq, D], cu_q_lens [B+1], cu_total_seq_lens [B+1] or NULL (kv_len = q_len), causal with offset
 * kv_len - q_len.  max_q_len / max_kv_len are host-side hints used for grid sizing only.
 * ------------------------------------------------------------------------------------------------- */
MOJO_B200_API int mojo_b200_paged_prefill_gqa(
    const void* query, const void* key_cache, const void* value_cache, const int32_t* cu_q_lens,
    const int32_t* cu_total_seq_lens, const int32_t* block_tables, void* out,
    int64_t total_q_tokens, int batch, int num_q_heads, int num_kv_heads, int head_dim,
    int64_t num_blocks, int block_size, int max_blocks_per_seq, int64_t table_stride,
    int64_t max_q_len, int64_t max_kv_len,
    int64_t q_stride_t, int64_t q_stride_h, int64_t o_stride_t, int64_t o_stride_h,
    int64_t kc_stride_b, int64_t kc_stride_h, int64_t kc_stride_t,
    int64_t vc_stride_b, int64_t vc_stride_h, int64_t vc_stride_t,
    float softmax_scale, int gqa_interleave, int is_causal, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * MojoPagedPrefillSWA.forward                       mojo_opset/core/operators/attention.py:533-640
 * MojoPagedDecodeSWA.forward                        mojo_opset/core/operators/attention.py:645-742
 *   (decode = one query token per sequence: cu_q_lens = 0..B, cu_total_seq_lens = cumulative total_seq_lens)
 *
 * Same tensors as mojo_b200_paged_prefill_gqa.  On top of the causal limit a key is visible iff
 * key + local_window_size >= position  or  key < global_window_size  (_generate_window_mask, attention.py:507-531);
 * -1 stands for None; both -1 = plain causal attention.  KV tiles no row of a query block can see are never loaded.
 * ------------------------------------------------------------------------------------------------- */
MOJO_B200_API int mojo_b200_paged_prefill_swa(
    const void* query, const void* key_cache, const void* value_cache, const int32_t* cu_q_lens,
    const int32_t* cu_total_seq_lens, const int32_t* block_tables, void* out,
    int64_t total_q_tokens, int batch, int num_q_heads, int num_kv_heads, int head_dim,
    int64_t num_blocks, int block_size, int max_blocks_per_seq, int64_t table_stride,
    int64_t max_q_len, int64_t max_kv_len,
    int64_t q_stride_t, int64_t q_stride_h, int64_t o_stride_t, int64_t o_stride_h,
    int64_t kc_stride_b, int64_t kc_stride_h, int64_t kc_stride_t,
    int64_t vc_stride_b, int64_t vc_stride_h, int64_t vc_stride_t,
    float softmax_scale, int gqa_interleave, int is_causal, int local_window_size, int global_window_size,
    int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * MojoSWA.forward (non-paged)                       mojo_opset/core/operators/attention.py:747-838
 *
 * query [Tq, Hq, D]; key / value PACKED [Tk, Hkv, D] (token, head strides): sequence b owns query rows
 * cu_q_lens[b] .. cu_q_lens[b+1] and key rows cu_total_seq_lens[b] .. cu_total_seq_lens[b+1]; causal with offset
 * kv_len - q_len and the window rule of the paged SWA ops (-1 = window not set).  is_causal = 0 is not built.
 * ------------------------------------------------------------------------------------------------- */
MOJO_B200_API int mojo_b200_swa(
    const void* query, const void* key, const void* value, const int32_t* cu_q_lens, const int32_t* cu_total_seq_lens,
    void* out, int64_t total_q_tokens, int64_t total_kv_tokens, int batch, int num_q_heads, int num_kv_heads, int head_dim,
    int64_t max_q_len, int64_t max_kv_len, int64_t q_stride_t, int64_t q_stride_h, int64_t o_stride_t, int64_t o_stride_h,
    int64_t k_stride_t, int64_t k_stride_h, int64_t v_stride_t, int64_t v_stride_h, float softmax_scale,
    int gqa_interleave, int is_causal, int local_window_size, int global_window_size, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * MojoSdpa.forward                                  mojo_opset/core/operators/attention.py:466-501
 *
 * query [B, Hq, Sq, D], key/value [B, Hkv, Skv, D] through (b, h, s) strides (D contiguous), no mask,
 * non-causal, q head h -> kv head h / (Hq / Hkv).  out through its own strides.
 * ------------------------------------------------------------------------------------------------- */
MOJO_B200_API int mojo_b200_sdpa(const void* query, const void* key, const void* value, void* out,
                   int batch, int num_q_heads, int num_kv_heads, int64_t q_len, int64_t kv_len, int head_dim,
                   int64_t q_stride_b, int64_t q_stride_h, int64_t q_stride_s,
                   int64_t k_stride_b, int64_t k_stride_h, int64_t k_stride_s,
                   int64_t v_stride_b, int64_t v_stride_h, int64_t v_stride_s,
                   int64_t o_stride_b, int64_t o_stride_h, int64_t o_stride_s,
                   float softmax_scale, int dtype, void* stream);

/* The same with `attn_mask` (reference attention.py:466-501 passes it to F.scaled_dot_product_attention; reference test
 * tests/accuracy/operators/test_attention.py:899-922: a [S, S] bool block-diffusion mask, 5 q / 1 kv heads): `mask` is
 * bool bytes (1 = the key takes part) addressed [b, h, q, k] through byte strides (0 = broadcast over that dimension), keys
 * contiguous.  A query row without any visible key reads as zeros (ATen returns NaN there). */
MOJO_B200_API int mojo_b200_sdpa_masked(const void* query, const void* key, const void* value, void* out,
                   int batch, int num_q_heads, int num_kv_heads, int64_t q_len, int64_t kv_len, int head_dim,
                   int64_t q_stride_b, int64_t q_stride_h, int64_t q_stride_s,
                   int64_t k_stride_b, int64_t k_stride_h, int64_t k_stride_s,
                   int64_t v_stride_b, int64_t v_stride_h, int64_t v_stride_s,
                   int64_t o_stride_b, int64_t o_stride_h, int64_t o_stride_s,
                   float softmax_scale, const void* mask, int64_t mask_stride_b, int64_t mask_stride_h,
                   int64_t mask_stride_q, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * MojoGemmAllReduce.forward           mojo_opset/core/operators/compute_with_comm.py:57-117
 *                                     (fused precedent: backends/ttx/operators/compute_with_comm.py:102-167)
 *
 * out[m, n] = all_reduce_sum over `world` ranks of ( x[m, k] @ weight[n, k]^T + bias[n] )   (bias on every rank,
 * as the reference adds it before the all-reduce).  ONE persistent kernel: tcgen05 GEMM (CTA pairs, cta_group::2,
 * once there are two 128-row blocks), partial tile rows pushed over NVLink into the tile owner's workspace as
 * self-validating 16-byte lines (3 payload words + the call's epoch, one 128-bit system-scope store each), fp32
 * reduction in rank order (bit-identical on all ranks), reduced lines broadcast to every rank.  world == 1 is a plain
 * GEMM (no workspace).  Launched with programmatic stream serialization (MOJO_B200_PDL=0 turns it off): the kernel
 * requests its first tiles of `weight` while the previous kernel of the stream is still draining and waits for that
 * kernel before it reads x / bias - so `weight` must not be written by the kernel launched immediately before this
 * call (model weights never are; a just-in-time dequantisation into `weight` needs MOJO_B200_PDL=0 or any kernel in
 * between).
 *
 * The workspace is "symmetric": every rank allocates the same number of bytes with mojo_b200_symm_alloc
 * (cudaMalloc, zero-filled), exports it (64-byte CUDA IPC handle), exchanges handles out of band (the host side
 * uses torch.distributed) and opens its peers' handles; peer_workspaces[r] is rank r's workspace as mapped in
 * THIS process (peer_workspaces[rank] = the local allocation).  The call counter (epoch) that versions the lines lives in
 * the workspace itself (device memory), so the launch is CUDA-graph replayable; all ranks must make the same
 * sequence of calls with the same m.  world must be 1, 2, 4 or 8.  x / weight rows must be 16-byte aligned; dtype bf16 / fp16.
 * ------------------------------------------------------------------------------------------------- */
MOJO_B200_API int mojo_b200_symm_alloc(size_t bytes, void** ptr);
MOJO_B200_API int mojo_b200_symm_free(void* ptr);
MOJO_B200_API int mojo_b200_symm_export(void* ptr, void* handle64);
MOJO_B200_API int mojo_b200_symm_open(const void* handle64, void** peer_ptr);
MOJO_B200_API int mojo_b200_symm_close(void* peer_ptr);
MOJO_B200_API size_t mojo_b200_gemm_allreduce_workspace_bytes(int64_t max_m, int64_t n, int world);
MOJO_B200_API int mojo_b200_gemm_allreduce(
    const void* x, const void* weight, const void* bias, void* out, int64_t m, int64_t n, int64_t k,
    int64_t x_row_stride, int64_t w_row_stride, int64_t out_row_stride,
    void* const* peer_workspaces, size_t workspace_bytes, int64_t workspace_max_m,
    int world, int rank, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Fused pre-attention pass: per-head RMSNorm (optional) -> RoPE -> paged KV store, one HBM pass.
 * Op shells MojoRoPEStoreKV / MojoNormRoPEStoreKV (reference README.md:128-130); today's composition is
 * modeling/qwen3/mojo_qwen3_dense.py:229-234 (q_norm, k_norm, rope) + PagedDummyCache.update :99-109
 * (MojoStorePagedKVCache) = four launches and a host-built plan.
 *
 * q [T, Hq, D], k / v [T, Hkv, D] (strides t, h; e.g. views of one fused QKV projection), norm weights [D] or
 * both NULL (no norm), cos / sin [T, rope_dim] (fp32 or the tensors' dtype, row stride cos_stride_t; the
 * rotation covers the LAST rope_dim features) -> q_out [T, Hq, D]; k' and v go to the page slot of their token,
 * found from (block_table, cu_q_lens | NULL = decode, context_kv_lens) as in store_paged_kv_table; k_out
 * (optional, [T, Hkv, D]) also receives k'.  bf16 / fp16, D in {64, 128, 256}.
 * ------------------------------------------------------------------------------------------------- */
MOJO_B200_API int mojo_b200_norm_rope_store_kv(
    const void* q, const void* k, const void* v, const void* q_norm_weight, const void* k_norm_weight, float eps,
    const void* cos, const void* sin, void* q_out, void* k_out, void* key_cache, void* value_cache,
    const int32_t* block_table, int64_t table_stride, int max_blocks_per_seq,
    const int32_t* cu_q_lens, const int32_t* context_kv_lens, int num_seqs,
    int64_t num_tokens, int num_q_heads, int num_kv_heads, int head_dim, int rope_dim,
    int64_t num_blocks, int block_size,
    int64_t q_stride_t, int64_t q_stride_h, int64_t k_stride_t, int64_t k_stride_h,
    int64_t v_stride_t, int64_t v_stride_h, int64_t qo_stride_t, int64_t qo_stride_h,
    int64_t ko_stride_t, int64_t ko_stride_h, int64_t cos_stride_t,
    int64_t kc_stride_b, int64_t kc_stride_h, int64_t kc_stride_t,
    int64_t vc_stride_b, int64_t vc_stride_h, int64_t vc_stride_t,
    int dtype, int cos_dtype, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * DiT block, non-GEMM ops around MojoSdpa (modeling/wan2_2/mojo_wan_model.py):
 *   MojoGelu.forward        mojo_opset/core/operators/activation.py:6-17      exact (erf) GELU
 *   MojoLayerNorm.forward   mojo_opset/core/operators/normalization.py:19-66  F.layer_norm, weight / bias may be NULL
 *   MojoGridRoPE.forward    mojo_opset/experimental/operators/position_embedding.py:80-118, ONE sample per call:
 *       x [tokens, heads, D] as interleaved (re, im) pairs; phase [seq_len, D/2] complex64 viewed as fp32
 *       (cos, sin) pairs with row stride phase_stride_t (floats); tokens >= seq_len are copied unchanged.
 * ------------------------------------------------------------------------------------------------- */
MOJO_B200_API int mojo_b200_gelu(const void* x, void* out, int64_t rows, int64_t cols, int64_t x_row_stride,
                   int64_t out_row_stride, int dtype, void* stream);
MOJO_B200_API int mojo_b200_layer_norm(const void* x, const void* weight, const void* bias, void* y, int64_t rows,
                         int hidden, int64_t x_row_stride, int64_t y_row_stride, float eps, int dtype, void* stream);
MOJO_B200_API int mojo_b200_grid_rope(const void* x, const float* phase, void* out, int64_t seq_len, int64_t tokens,
                        int heads, int head_dim, int64_t x_stride_t, int64_t x_stride_h,
                        int64_t o_stride_t, int64_t o_stride_h, int64_t phase_stride_t, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Paged-KV bookkeeping on the device: PagedAttentionRuntimeState._reserve / _allocate_blocks /
 * _build_positions (mojo_opset/runtime/runtime.py:112-158) without the per-sequence .item() host syncs.
 *
 * paged_reserve: for every sequence b (batch order) append q_lens[b] tokens (NULL = 1 each): new logical blocks
 * take the top entries of the free stack free_blocks[0 .. *num_free) exactly in the reference's order, are written
 * to block_tables[b, old_blocks ...], *num_free shrinks, total_seq_lens[b] += q_lens[b], and
 * context_lens_out[b] receives the length BEFORE the append.  If the request cannot be served nothing changes
 * and *error_flag is set (1 = out of blocks, 2 = a sequence would exceed max_blocks_per_seq).
 * paged_positions: positions[t] = context_lens[seq(t)] + (t - cu_q_lens[seq(t)]) for prefill tokens.
 * ------------------------------------------------------------------------------------------------- */
MOJO_B200_API int mojo_b200_paged_reserve(int32_t* block_tables, int64_t table_stride, int max_blocks_per_seq,
                            int32_t* total_seq_lens, const int32_t* q_lens, const int32_t* free_blocks,
                            int32_t* num_free, int32_t* context_lens_out, int batch, int block_size,
                            int32_t* error_flag, void* stream);
MOJO_B200_API int mojo_b200_paged_positions(int64_t* positions, const int32_t* cu_q_lens, const int32_t* context_lens,
                              int batch, int64_t num_tokens, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MOJO_B200_H_ */
