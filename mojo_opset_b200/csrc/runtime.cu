// Device-side bookkeeping of the paged KV cache: what PagedAttentionRuntimeState._reserve /
// _allocate_blocks / _build_positions do on the HOST with one .item() per sequence
// (mojo_opset/runtime/runtime.py:112-158) runs here as one small kernel each, so a decode step has no host sync and
// the whole step - bookkeeping included - can be captured in a CUDA graph.
//
// paged_reserve is deterministic and reproduces the reference's allocation order exactly: sequences are served in
// batch order, sequence b takes the top n_b entries of the free stack below those taken by sequences < b
// (free_blocks[num_free - prefix_b - n_b : num_free - prefix_b], in stack order).  One CTA: the per-sequence block
// needs are prefix-summed in shared memory.
#include "common.cuh"

namespace mojo {

constexpr int kReserveThreads = 1024;

__global__ void __launch_bounds__(kReserveThreads) paged_reserve_kernel(
    int32_t* __restrict__ block_tables, int64_t table_stride, int max_blocks, int32_t* __restrict__ total_seq_lens,
    const int32_t* __restrict__ q_lens, const int32_t* __restrict__ free_blocks, int32_t* __restrict__ num_free,
    int32_t* __restrict__ context_lens_out, int batch, int block_size, int32_t* __restrict__ error_flag) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s, fail_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    carry_s = 0;
    fail_s = 0;
  }
  __syncthreads();
  const int free_now = *num_free;
  // pass 0: total demand (so that an out-of-memory request changes nothing, like the reference's ValueError)
  int demand = 0;
  for (int b = threadIdx.x; b < batch; b += blockDim.x) {
    const int ctx = total_seq_lens[b], app = q_lens ? q_lens[b] : 1;
    const int old_n = (ctx + block_size - 1) / block_size, new_n = (ctx + app + block_size - 1) / block_size;
    if (new_n > max_blocks) atomicExch(&fail_s, 2);
    demand += max(new_n - old_n, 0);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) demand += __shfl_xor_sync(0xffffffffu, demand, o);
  if (lane == 0) warp_tot[warp] = demand;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += warp_tot[w];
    if (tot > free_now && fail_s == 0) fail_s = 1;
  }
  __syncthreads();
  if (fail_s != 0) {
    if (threadIdx.x == 0 && error_flag) *error_flag = fail_s;  // 1: out of blocks, 2: sequence longer than the table
    for (int b = threadIdx.x; b < batch; b += blockDim.x) context_lens_out[b] = total_seq_lens[b];
    return;
  }
  __syncthreads();
  // pass 1: batch-ordered exclusive prefix of the needs, chunk by chunk
  for (int base = 0; base < batch; base += blockDim.x) {
    const int b = base + threadIdx.x;
    int ctx = 0, app = 0, old_n = 0, need = 0;
    if (b < batch) {
      ctx = total_seq_lens[b];
      app = q_lens ? q_lens[b] : 1;
      old_n = (ctx + block_size - 1) / block_size;
      need = max((ctx + app + block_size - 1) / block_size - old_n, 0);
    }
    int incl = need;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int warp_off = 0;
    for (int w = 0; w < warp; ++w) warp_off += warp_tot[w];
    const int excl = carry_s + warp_off + incl - need;
    if (b < batch) {
      // the reference slices free_blocks[num_free - need : num_free] AFTER the earlier sequences shrank num_free
      const int top = free_now - excl;
      for (int j = 0; j < need; ++j) block_tables[b * table_stride + old_n + j] = free_blocks[top - need + j];
      context_lens_out[b] = ctx;
      total_seq_lens[b] = ctx + app;
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = excl + need;
    __syncthreads();
  }
  if (threadIdx.x == 0) *num_free = free_now - carry_s;
}

__global__ void __launch_bounds__(256) paged_positions_kernel(int64_t* __restrict__ positions,
                                                              const int32_t* __restrict__ cu_q,
                                                              const int32_t* __restrict__ context_lens, int batch,
                                                              int64_t num_tokens) {
  const int64_t tok = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tok >= num_tokens) return;
  int lo = 0, hi = batch + 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cu_q[mid] <= tok) lo = mid + 1; else hi = mid;
  }
  const int seq = lo - 1;
  positions[tok] = (seq >= 0 && seq < batch) ? (int64_t)context_lens[seq] + (tok - cu_q[seq]) : -1;
}

}  // namespace mojo

extern "C" int mojo_b200_paged_reserve(int32_t* block_tables, int64_t table_stride, int max_blocks_per_seq,
                                       int32_t* total_seq_lens, const int32_t* q_lens, const int32_t* free_blocks,
                                       int32_t* num_free, int32_t* context_lens_out, int batch, int block_size,
                                       int32_t* error_flag, void* stream) {
  using namespace mojo;
  MOJO_REQUIRE(batch >= 0 && block_size > 0 && max_blocks_per_seq >= 0, MOJO_B200_EINVAL, "paged_reserve: bad sizes");
  if (batch == 0) return 0;
  MOJO_REQUIRE(block_tables && total_seq_lens && free_blocks && num_free && context_lens_out, MOJO_B200_EINVAL,
               "paged_reserve: null pointer");
  paged_reserve_kernel<<<1, kReserveThreads, 0, (cudaStream_t)stream>>>(block_tables, table_stride, max_blocks_per_seq,
                                                                      total_seq_lens, q_lens, free_blocks, num_free,
                                                                      context_lens_out, batch, block_size, error_flag);
  return check_launch("paged_reserve_kernel");
}

extern "C" int mojo_b200_paged_positions(int64_t* positions, const int32_t* cu_q_lens, const int32_t* context_lens,
                                         int batch, int64_t num_tokens, void* stream) {
  using namespace mojo;
  MOJO_REQUIRE(batch >= 0 && num_tokens >= 0, MOJO_B200_EINVAL, "paged_positions: bad sizes");
  if (num_tokens == 0) return 0;
  MOJO_REQUIRE(positions && cu_q_lens && context_lens, MOJO_B200_EINVAL, "paged_positions: null pointer");
  paged_positions_kernel<<<(unsigned)((num_tokens + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      positions, cu_q_lens, context_lens, batch, num_tokens);
  return check_launch("paged_positions_kernel");
}
