// MojoPagedDecodeGQA: split-KV flash-decode over a paged KV cache.
//
// HBM-bound: every K/V byte of a sequence is read exactly once per KV head and serves all G query heads of
// the group (packed into the M dimension of one MMA tile).  bytes = 2 * sum_b(len_b) * Hkv * D * sizeof(T)
// + q/out + table.
//
// Fast path (`paged_decode_mma_kernel`, 16-bit dtypes, D in {64,128}, power-of-two pages >= 8):
//   grid  (num_splits, Hkv * head_tiles, B), 160 threads = 4 consumer warps + 1 producer warp, 2 CTAs/SM.
//   producer : walks the block table and issues one TMA tensor copy (cp.async.bulk.tensor, 128B swizzle) per
//              (page, K|V) into a ring of `stages` 64-token tiles guarded by full/empty mbarriers.
//   consumers: each warp owns 16 of the tile's 64 tokens.  S = Q K^T with mma.sync m16n8k16 (M = the group's
//              query heads, zero padded to 16), online softmax in registers (quad shuffles), P V accumulated
//              in fp32 registers.  The four warps' (m, l, O) are merged through shared memory at the end.
//   output   : num_splits == 1 -> normalised rows straight to `out`; otherwise un-normalised fp32 partials +
//              (m, l) per split, folded by `paged_decode_reduce_kernel`.
// Generic path (`paged_decode_simt_kernel`): any dtype (fp32 included), any D <= 256, any page size; one
// warp per token, dot products by warp shuffle.  Same partial format, same reduce kernel.
//
// MojoPagedDecodeSWA (reference attention.py:645-745) is the same kernel over the VISIBLE KV tiles only: the tiles of
// the global prefix [0, g) followed by the tiles of the local window [len - 1 - local, len); the splits divide that
// list, the tiles in between are never loaded, and the two edge tiles are masked per key (fast path only).
//
// Golden rounding points (reference attention.py:217-228) reproduced: for 16-bit inputs the score is rounded
// to the input dtype after the dot product and again after scaling; probabilities are rounded to the input
// dtype before the PV product; everything else is fp32.
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "tma.cuh"

namespace mojo {

constexpr int kConsumerWarps = 4;
constexpr int kDecodeThreads = 32 * (kConsumerWarps + 1);
constexpr float kLog2e = 1.4426950408889634f;

struct DecodeParams {
  const void* q;
  void* out;
  float* part_o;    // [B, Hq, splits, D]
  float2* part_ml;  // [B, Hq, splits] (running max in log2 units, sum)
  const int32_t* seq_lens;
  const int32_t* tables;
  const void* kc;
  const void* vc;
  int64_t table_stride;
  int64_t num_blocks;
  int num_q_heads, num_kv_heads, group, head_tiles, head_dim;
  int block_size, log2_bs, max_blocks;
  int64_t q_sb, q_sh, o_sb, o_sh;
  int64_t kc_b, kc_h, kc_t, vc_b, vc_h, vc_t;
  float scale;
  int interleave, num_splits, stages;
  // MojoPagedDecodeSWA: the query token (position seq_len - 1) sees key k iff k + win_local >= position or
  // k < win_global; -1 = that window is not set (both -1: every key, MojoPagedDecodeGQA)
  int win_local, win_global;
  int* err;  // device error word or null (include/mojo_b200.h)
  // split-KV: arrival counters (one per (sequence, kv head tile), zero between launches) - the last split of a group
  // to arrive folds the group's partials itself; null = fold by paged_decode_reduce_kernel
  int* tickets;
  unsigned long long* trace;  // developer timeline (MOJO_DECODE_TRACE builds only, tools/decode_trace.py)
};

#ifdef MOJO_DECODE_TRACE
// per CTA: [0] entry, [1] past pdl_wait, [2] first tile landed, [3] last tile consumed, [4] partial written,
// [5] fold done (the group's last CTA only)  (globaltimer, ns)
#define DTRACE(ev)                                                                                             \
  do {                                                                                                         \
    if (p.trace) {                                                                                             \
      unsigned long long t__;                                                                                  \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));                                                  \
      p.trace[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 + (ev)] = t__;      \
    }                                                                                                          \
  } while (0)
#else
#define DTRACE(ev) do {} while (0)
#endif

__device__ __forceinline__ void split_tile_range(int seq_len, int num_splits, int split, int& tile_begin,
                                                 int& tile_end) {
  const int tiles = seq_len > 0 ? (seq_len + kTile - 1) / kTile : 0;
  const int per = (tiles + num_splits - 1) / num_splits;
  tile_begin = split * per;
  tile_end = min(tiles, tile_begin + per);
}

// The KV tiles a windowed decode has to read: the tiles of the global prefix [0, g) followed by the tiles of the local
// window [lo, seq_len).  `visible` tiles are numbered 0 .. n_vis-1 (what the splits divide); tile(v) is the real one.
// Tiles in between are never loaded, so the traffic is O(window) instead of O(context).
struct DecodeWindow {
  int lo, g;        // key k is visible iff k < g or k >= lo (and k < seq_len)
  int t_g, t_lo;    // visible tile v is real tile v (v < t_g) or t_lo + (v - t_g)
  int n_vis;
  __device__ __forceinline__ int tile(int v) const { return v < t_g ? v : t_lo + (v - t_g); }
  __device__ __forceinline__ bool sees(int key) const { return key < g || key >= lo; }
};
__device__ __forceinline__ DecodeWindow decode_window(int seq_len, int win_local, int win_global) {
  DecodeWindow w;
  const int tiles = seq_len > 0 ? (seq_len + kTile - 1) / kTile : 0;
  if (win_local < 0 && win_global < 0) {
    w.lo = 0; w.g = 0; w.t_g = 0; w.t_lo = 0; w.n_vis = tiles;
    return w;
  }
  w.lo = win_local >= 0 ? max(0, seq_len - 1 - win_local) : seq_len;  // seq_len: no local window at all
  w.g = win_global >= 0 ? min(win_global, seq_len) : 0;
  w.t_g = (w.g + kTile - 1) / kTile;
  w.t_lo = w.lo >= seq_len ? tiles : w.lo / kTile;
  if (w.t_g >= w.t_lo) {  // the two ranges touch or overlap: every tile (the mask still applies inside a tile)
    w.t_g = 0; w.t_lo = 0; w.n_vis = tiles;
  } else {
    w.n_vis = w.t_g + (tiles - w.t_lo);
  }
  return w;
}
__device__ __forceinline__ void split_visible_range(const DecodeWindow& w, int num_splits, int split, int& v_begin,
                                                    int& v_end) {
  const int per = (w.n_vis + num_splits - 1) / num_splits;
  v_begin = split * per;
  v_end = min(w.n_vis, v_begin + per);
}

__device__ __forceinline__ int q_head_of(const DecodeParams& p, int kvh, int j) {
  return p.interleave ? j * p.num_kv_heads + kvh : kvh * p.group + j;
}

// ======================================================================================================
// split-KV fold inside the main kernel
// ======================================================================================================
// Timeline of the batch-1 / 32k-context launch (tools/decode_trace.py, 296 CTAs): the KV stream itself runs at 7.0 TB/s
// (19 of the 31.7 us), but 5.6 us passed between the last partial and the next launch getting past its wait - the fold
// kernel's launch hand-offs around ~3 us of dependent L2 round trips.  So the LAST split of a (sequence, kv head tile)
// group to arrive (an arrival counter per group, reset by that CTA: zero between launches) folds the group's partials
// itself: a warp per query head, every lane four features of the row, 16 partial rows requested per trip.
template <typename T, int D>
__device__ __forceinline__ void fold_group(const DecodeParams& p, int b, int kvh, int ht, int rows_valid, int tid) {
  constexpr int F4 = D / 4;   // float4 chunks of a row: a lane takes chunk `lane`
  constexpr int kBatch = 20;  // partial rows requested per trip (37 splits at batch 1 / 32k context: two trips)
  const int warp = tid >> 5, lane = tid & 31;
  const bool has = lane < F4;
  const int n = p.num_splits;  // <= 64 (checked by the launcher): a lane holds the (m, l) of splits lane, lane + 32
  for (int r = warp; r < rows_valid; r += kConsumerWarps) {
    const int hq = q_head_of(p, kvh, ht * 16 + r);
    const int64_t base = ((int64_t)b * p.num_q_heads + hq) * n;
    const float4* rows = reinterpret_cast<const float4*>(p.part_o + base * D) + lane;
    // the (m, l) pairs and the first batch of rows are requested together: the rows do not depend on the weights
    float2 ml0 = make_float2(-INFINITY, 0.f), ml1 = ml0;
    if (lane < n) ml0 = __ldcg(&p.part_ml[base + lane]);
    if (lane + 32 < n) ml1 = __ldcg(&p.part_ml[base + lane + 32]);
    float4 v[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (u < n && has) v[u] = __ldcg(rows + (size_t)u * F4);
    }
    const float m = warp_max(fmaxf(ml0.x, ml1.x));
    // an empty split (m = -inf) contributes nothing - its row is uninitialised memory and is never multiplied
    const float f0 = ml0.x != -INFINITY ? exp2f(ml0.x - m) : 0.f;
    const float f1 = ml1.x != -INFINITY ? exp2f(ml1.x - m) : 0.f;
    const float l = warp_sum(f0 * ml0.y + f1 * ml1.y);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s0 = 0; s0 < n; s0 += kBatch) {
      if (s0 > 0) {
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (s0 + u < n && has) v[u] = __ldcg(rows + (size_t)(s0 + u) * F4);
        }
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int si = s0 + u;
        const float fs = __shfl_sync(0xffffffffu, si < 32 ? f0 : f1, si & 31);
        if (si < n && fs != 0.f) {
          acc.x = fmaf(fs, v[u].x, acc.x);
          acc.y = fmaf(fs, v[u].y, acc.y);
          acc.z = fmaf(fs, v[u].z, acc.z);
          acc.w = fmaf(fs, v[u].w, acc.w);
        }
      }
    }
    if (has) {
      const float inv = l > 0.f ? 1.0f / l : 0.f;
      T* o = reinterpret_cast<T*>(p.out) + b * p.o_sb + hq * p.o_sh + lane * 4;
      o[0] = DType<T>::from_f(acc.x * inv);
      o[1] = DType<T>::from_f(acc.y * inv);
      o[2] = DType<T>::from_f(acc.z * inv);
      o[3] = DType<T>::from_f(acc.w * inv);
    }
  }
}

// Called by the CTA's first 128 threads once its partial is written: count the arrival, and fold if it was the last.
template <typename T, int D>
__device__ __forceinline__ void arrive_and_fold(const DecodeParams& p, int b, int kvh, int ht, int rows_valid) {
  __shared__ int s_last;
  __threadfence();  // this thread's partial stores are visible device-wide before the arrival is counted
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory");
  if (threadIdx.x == 0) {
    int* tk = p.tickets + ((int64_t)b * gridDim.y + blockIdx.y);
    const int last = atomicAdd(tk, 1) == p.num_splits - 1;
    if (last) *tk = 0;  // every split of the group has arrived: the counter is ready for the next launch
    s_last = last;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory");
  if (s_last) {
    __threadfence();
    fold_group<T, D>(p, b, kvh, ht, rows_valid, threadIdx.x);
  }
}

// ======================================================================================================
// fast path
// ======================================================================================================
// FOLD: the split-KV instantiation with the in-kernel fold above (the single-split instantiation - the cfg2 / cfg4
// serving shapes at 1.00 of the HBM peak - carries none of that code: with it inlined the kernel measured 1 % slower)
template <typename T, int D, bool SPLIT_HALVES, bool FOLD>
__global__ void __launch_bounds__(kDecodeThreads, 2)
paged_decode_mma_kernel(const __grid_constant__ CUtensorMap k_map, const __grid_constant__ CUtensorMap v_map,
                        const DecodeParams p) {
  constexpr int NH = D / 64;                 // 128-byte lines per token row
  constexpr int KS = D / 16;                 // k-steps of the QK product / n-tile pairs of the PV product
  constexpr int TILE_BYTES = kTile * D * 2;  // one tensor, one stage
  constexpr int O_STRIDE = D + 8;            // padded fp32 row of the merge buffer

  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled TMA tiles need 1024-byte aligned bases; the launch adds 1 KiB of slack for this round-up
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stages = p.stages;
  uint8_t* tiles = smem;                                                        // [stages][K | V]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * 2 * TILE_BYTES);
  uint64_t* empty = full + stages;

  const int split = blockIdx.x;
  const int kvh = blockIdx.y / p.head_tiles;
  const int ht = blockIdx.y - kvh * p.head_tiles;
  const int b = blockIdx.z;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) DTRACE(0);
  // prologue that touches no other kernel's data (PDL: runs while the previous kernel drains)
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kConsumerWarps);
    }
    mbar_fence_init();
  }
  if (warp == kConsumerWarps && lane == 0) {
    tma_prefetch_desc(&k_map);
    tma_prefetch_desc(&v_map);
  }
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x == 0) DTRACE(1);
  const int seq_len = p.seq_lens[b];
  const int rows_valid = min(16, p.group - ht * 16);  // query heads in this tile
  // the reference raises ValueError for a row with keys but no first block (attention.py:186-187): flag it
  if (p.err && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && seq_len > 0 && p.max_blocks > 0 &&
      p.tables[(int64_t)b * p.table_stride] < 0)
    atomicOr(p.err, 1);

  const DecodeWindow win = decode_window(seq_len, p.win_local, p.win_global);
  const bool swa = p.win_local >= 0 || p.win_global >= 0;
  int tile_begin, tile_end;  // in visible-tile numbering
  split_visible_range(win, p.num_splits, split, tile_begin, tile_end);
  const int n_tiles = tile_end - tile_begin;

  if (n_tiles <= 0) {
    // nothing to attend to: zero rows (single split) or an empty partial (m = -inf, l = 0)
    for (int i = threadIdx.x; i < rows_valid * D; i += blockDim.x) {
      const int r = i / D, d = i - r * D;
      const int hq = q_head_of(p, kvh, ht * 16 + r);
      if (p.num_splits == 1) {
        reinterpret_cast<T*>(p.out)[b * p.o_sb + hq * p.o_sh + d] = DType<T>::from_f(0.f);
      } else if (d == 0) {
        p.part_ml[((int64_t)b * p.num_q_heads + hq) * p.num_splits + split] = make_float2(-INFINITY, 0.f);
      }
    }
    if constexpr (FOLD) {
      // every thread of the CTA (the producer warp included) may have written one of the (m, l) pairs above: all of
      // them are ordered before the arrival is counted by the first 128
      __threadfence();
      __syncthreads();
      if (threadIdx.x < kConsumerWarps * 32) arrive_and_fold<T, D>(p, b, kvh, ht, rows_valid);
    }
    return;
  }

  __syncthreads();

  const int log2_bs = p.log2_bs;
  const int box_rows = min(p.block_size, kTile);  // R
  const int boxes_per_tile = kTile / box_rows;

  if (warp == kConsumerWarps) {
    // ------------------------------------------------------------------ producer warp
    const int32_t* table = p.tables + (int64_t)b * p.table_stride;
    const uint32_t box_bytes = (uint32_t)box_rows * D * 2;
    for (int it = 0; it < n_tiles; ++it) {
      const int stage = it % stages;
      const uint32_t phase = (uint32_t)(it / stages) & 1u;
      const int tok0 = win.tile(tile_begin + it) * kTile;
      // boxes that hold at least one valid token
      const int want = min(boxes_per_tile, (seq_len - tok0 + box_rows - 1) / box_rows);
      int blk = 0, row_in_page = 0;
      if (lane < want) {
        const int tok = tok0 + lane * box_rows;
        const int page = tok >> log2_bs;
        row_in_page = tok & (p.block_size - 1);
        blk = page < p.max_blocks ? table[page] : -1;  // -1 -> out of bounds -> TMA zero fill
      }
      if (lane == 0) {
        mbar_wait(&empty[stage], phase ^ 1u);
        mbar_expect_tx(&full[stage], 2u * box_bytes * (uint32_t)want);
      }
      __syncwarp();
      if (lane < want) {
        uint8_t* kdst = tiles + (size_t)stage * 2 * TILE_BYTES + (size_t)lane * box_bytes;
        uint8_t* vdst = kdst + TILE_BYTES;
        if (SPLIT_HALVES) {
          tma_load_5d(kdst, &k_map, &full[stage], 0, row_in_page, 0, kvh, blk);
          tma_load_5d(vdst, &v_map, &full[stage], 0, row_in_page, 0, kvh, blk);
        } else {
          tma_load_5d(kdst, &k_map, &full[stage], 0, 0, row_in_page, kvh, blk);
          tma_load_5d(vdst, &v_map, &full[stage], 0, 0, row_in_page, kvh, blk);
        }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumer warps
  const int g = lane >> 2;  // fragment row (and row + 8)
  const int c = lane & 3;
  const bool hi_rows = rows_valid > 8;

  // Q fragments (A operand, row-major 16 x D), rows beyond the group are zero
  uint32_t qf[KS][4];
  {
    const T* qb = reinterpret_cast<const T*>(p.q) + (int64_t)b * p.q_sb;
    const T* r0 = g < rows_valid ? qb + (int64_t)q_head_of(p, kvh, ht * 16 + g) * p.q_sh : nullptr;
    const T* r1 = g + 8 < rows_valid ? qb + (int64_t)q_head_of(p, kvh, ht * 16 + g + 8) * p.q_sh : nullptr;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int d0 = ks * 16 + 2 * c;
      qf[ks][0] = r0 ? *reinterpret_cast<const uint32_t*>(r0 + d0) : 0u;
      qf[ks][1] = r1 ? *reinterpret_cast<const uint32_t*>(r1 + d0) : 0u;
      qf[ks][2] = r0 ? *reinterpret_cast<const uint32_t*>(r0 + d0 + 8) : 0u;
      qf[ks][3] = r1 ? *reinterpret_cast<const uint32_t*>(r1 + d0 + 8) : 0u;
    }
  }

  float o[2 * KS][4];
#pragma unroll
  for (int i = 0; i < 2 * KS; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;

  // ldmatrix lane roles
  const int mat = lane >> 3, mr = lane & 7;
  const int k_row = warp * 16 + (mat >> 1) * 8 + mr;  // K: matrices (tok 0-7 | 8-15) x (d lo | d hi)
  const int v_row = warp * 16 + (mat & 1) * 8 + mr;   // V (transposed): (tok 0-7 | 8-15) per d chunk
  auto line_of = [&](int row, int half) -> uint32_t {
    if (SPLIT_HALVES) {
      const int box = row / box_rows, r = row - box * box_rows;
      return (uint32_t)(box * box_rows * NH + half * box_rows + r);
    }
    return (uint32_t)(row * NH + half);
  };
  uint32_t k_line[NH], v_line[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    k_line[h] = line_of(k_row, h);
    v_line[h] = line_of(v_row, h);
  }
  const float scale = p.scale;

  for (int it = 0; it < n_tiles; ++it) {
    const int stage = it % stages;
    const uint32_t phase = (uint32_t)(it / stages) & 1u;
    const int tok0 = win.tile(tile_begin + it) * kTile;
    const int valid = seq_len - tok0;  // tokens of this tile that exist (may exceed kTile)
    uint8_t* sk = tiles + (size_t)stage * 2 * TILE_BYTES;
    uint8_t* sv = sk + TILE_BYTES;
    const uint32_t sk_a = smem_u32(sk), sv_a = smem_u32(sv);

    mbar_wait(&full[stage], phase);
    if (it == 0 && threadIdx.x == 0) DTRACE(2);

    if (valid < kTile) {
      // Tail tile: slots past the end of the sequence hold stale shared memory or uninitialised cache
      // contents.  Zero this warp's V rows there so 0 * garbage cannot produce NaN in the PV product.
      for (int r = lane >> 1; r < 16; r += 16) {
        const int row = warp * 16 + r;
        if (row >= valid) {
#pragma unroll
          for (int h = 0; h < NH; ++h) {
            uint4* dst = reinterpret_cast<uint4*>(sv + line_of(row, h) * 128u + (lane & 1) * 64);
            dst[0] = dst[1] = dst[2] = dst[3] = make_uint4(0, 0, 0, 0);
          }
        }
      }
      __syncwarp();
    }

    // ---- S = Q K^T for this warp's 16 tokens (two n8 tiles)
    float s[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t b0, b1, b2, b3;  // 16-byte chunk (2 ks + {0,1}) of the row: half = ks / 4
      ldsm_x4(sk_a + swz128(k_line[ks >> 2], ((ks * 2) & 7) + (mat & 1)), b0, b1, b2, b3);
      Mma16816<T>::run(s[0], qf[ks], b0, b1);
      Mma16816<T>::run(s[1], qf[ks], b2, b3);
    }

    // ---- golden rounding of the scores, masking, online softmax (log2 domain)
    float tile_lo = -INFINITY, tile_hi = -INFINITY;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int tok = warp * 16 + j * 8 + 2 * c + (e & 1);
        // MojoPagedDecodeGQA rounds the scores again after scaling (attention.py:217); the SWA op scales in fp32 (:700)
        const float sc = __fmul_rn(round_through<T>(s[j][e]), scale);
        float v = (swa ? sc : round_through<T>(sc)) * kLog2e;
        v = tok < valid && win.sees(tok0 + tok) ? v : -INFINITY;
        s[j][e] = v;
        if (e < 2) tile_lo = fmaxf(tile_lo, v); else tile_hi = fmaxf(tile_hi, v);
      }
    }
    tile_lo = fmaxf(tile_lo, __shfl_xor_sync(0xffffffffu, tile_lo, 1));
    tile_lo = fmaxf(tile_lo, __shfl_xor_sync(0xffffffffu, tile_lo, 2));
    tile_hi = fmaxf(tile_hi, __shfl_xor_sync(0xffffffffu, tile_hi, 1));
    tile_hi = fmaxf(tile_hi, __shfl_xor_sync(0xffffffffu, tile_hi, 2));
    const float new_lo = fmaxf(m_lo, tile_lo), new_hi = fmaxf(m_hi, tile_hi);
    const float base_lo = new_lo == -INFINITY ? 0.f : new_lo, base_hi = new_hi == -INFINITY ? 0.f : new_hi;
    const float a_lo = exp2f(m_lo - base_lo), a_hi = exp2f(m_hi - base_hi);
    m_lo = new_lo;
    m_hi = new_hi;
    uint32_t pa[4];
    {
      const float p00 = exp2f(s[0][0] - base_lo), p01 = exp2f(s[0][1] - base_lo);
      const float p10 = exp2f(s[1][0] - base_lo), p11 = exp2f(s[1][1] - base_lo);
      l_lo = l_lo * a_lo + (p00 + p01) + (p10 + p11);
      pa[0] = Mma16816<T>::pack(p00, p01);
      pa[2] = Mma16816<T>::pack(p10, p11);
      if (hi_rows) {
        const float q00 = exp2f(s[0][2] - base_hi), q01 = exp2f(s[0][3] - base_hi);
        const float q10 = exp2f(s[1][2] - base_hi), q11 = exp2f(s[1][3] - base_hi);
        l_hi = l_hi * a_hi + (q00 + q01) + (q10 + q11);
        pa[1] = Mma16816<T>::pack(q00, q01);
        pa[3] = Mma16816<T>::pack(q10, q11);
      } else {
        pa[1] = pa[3] = 0u;
      }
    }
#pragma unroll
    for (int i = 0; i < 2 * KS; ++i) {
      o[i][0] *= a_lo;
      o[i][1] *= a_lo;
      o[i][2] *= a_hi;
      o[i][3] *= a_hi;
    }

    // ---- O += P V
#pragma unroll
    for (int dp = 0; dp < KS; ++dp) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_trans(sv_a + swz128(v_line[dp >> 2], ((dp * 2) & 7) + (mat >> 1)), b0, b1, b2, b3);
      Mma16816<T>::run(o[2 * dp], pa, b0, b1);
      Mma16816<T>::run(o[2 * dp + 1], pa, b2, b3);
    }

    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
  }

  if (threadIdx.x == 0) DTRACE(3);
  // -------------------------------------------------------------------- merge the four warps
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);

  float* s_o = reinterpret_cast<float*>(tiles);                       // [4][16][O_STRIDE]
  float2* s_ml = reinterpret_cast<float2*>(s_o + kConsumerWarps * 16 * O_STRIDE);  // [4][16]
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory");  // every warp is done reading tiles
  if (c == 0) {
    s_ml[warp * 16 + g] = make_float2(m_lo, l_lo);
    s_ml[warp * 16 + g + 8] = make_float2(m_hi, l_hi);
  }
#pragma unroll
  for (int i = 0; i < 2 * KS; ++i) {
    const int d = i * 8 + 2 * c;
    if (g < rows_valid) *reinterpret_cast<float2*>(&s_o[(warp * 16 + g) * O_STRIDE + d]) = make_float2(o[i][0], o[i][1]);
    if (g + 8 < rows_valid)
      *reinterpret_cast<float2*>(&s_o[(warp * 16 + g + 8) * O_STRIDE + d]) = make_float2(o[i][2], o[i][3]);
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory");

  for (int i = threadIdx.x; i < rows_valid * D; i += kConsumerWarps * 32) {
    const int r = i / D, d = i - r * D;
    float mw[kConsumerWarps], m = -INFINITY;
#pragma unroll
    for (int w = 0; w < kConsumerWarps; ++w) {
      mw[w] = s_ml[w * 16 + r].x;
      m = fmaxf(m, mw[w]);
    }
    float l = 0.f, acc = 0.f;
#pragma unroll
    for (int w = 0; w < kConsumerWarps; ++w) {
      if (mw[w] != -INFINITY) {  // a warp whose 16 tokens were all masked contributes nothing
        const float f = exp2f(mw[w] - m);
        l += f * s_ml[w * 16 + r].y;
        acc += f * s_o[(w * 16 + r) * O_STRIDE + d];
      }
    }
    const int hq = q_head_of(p, kvh, ht * 16 + r);
    if (p.num_splits == 1) {
      reinterpret_cast<T*>(p.out)[b * p.o_sb + hq * p.o_sh + d] = DType<T>::from_f(l > 0.f ? acc / l : 0.f);
    } else {
      const int64_t slot = ((int64_t)b * p.num_q_heads + hq) * p.num_splits + split;
      p.part_o[slot * D + d] = acc;
      if (d == 0) p.part_ml[slot] = make_float2(m, l);
    }
  }
  if (threadIdx.x == 0) DTRACE(4);
  if constexpr (FOLD) {
    arrive_and_fold<T, D>(p, b, kvh, ht, rows_valid);
    if (threadIdx.x == 0) DTRACE(5);
  }
}

// ======================================================================================================
// generic path: one warp per token
// ======================================================================================================
template <typename T, int GH, int DPL>
__global__ void __launch_bounds__(256) paged_decode_simt_kernel(const DecodeParams p) {
  constexpr int WARPS = 8;
  extern __shared__ __align__(16) float sm[];
  const int D = p.head_dim;
  const int split = blockIdx.x;
  const int kvh = blockIdx.y / p.head_tiles;
  const int ht = blockIdx.y - kvh * p.head_tiles;
  const int b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_wait();
  pdl_trigger();
  const int seq_len = p.seq_lens[b];
  const int rows_valid = min(GH, p.group - ht * GH);
  if (p.err && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && seq_len > 0 && p.max_blocks > 0 &&
      p.tables[(int64_t)b * p.table_stride] < 0)
    atomicOr(p.err, 1);

  int tile_begin, tile_end;
  split_tile_range(seq_len, p.num_splits, split, tile_begin, tile_end);
  const int tok_begin = tile_begin * kTile;
  const int tok_end = min(seq_len, tile_end * kTile);

  float qv[GH][DPL], acc[GH][DPL], m[GH], l[GH];
#pragma unroll
  for (int r = 0; r < GH; ++r) {
    m[r] = -INFINITY;
    l[r] = 0.f;
    const T* qr = r < rows_valid ? reinterpret_cast<const T*>(p.q) + (int64_t)b * p.q_sb +
                                       (int64_t)q_head_of(p, kvh, ht * GH + r) * p.q_sh
                                 : nullptr;
#pragma unroll
    for (int i = 0; i < DPL; ++i) {
      const int d = lane + 32 * i;
      qv[r][i] = (qr && d < D) ? DType<T>::to_f(qr[d]) : 0.f;
      acc[r][i] = 0.f;
    }
  }

  const int32_t* table = p.tables + (int64_t)b * p.table_stride;
  for (int t = tok_begin + warp; t < tok_end; t += WARPS) {
    const int page = t / p.block_size;
    const int slot = t - page * p.block_size;
    const int blk = page < p.max_blocks ? table[page] : -1;
    float kf[DPL], vf[DPL];
    const bool ok = blk >= 0 && blk < p.num_blocks;  // unmapped page: keys/values read as zeros
    const T* kp = reinterpret_cast<const T*>(p.kc) + (int64_t)blk * p.kc_b + (int64_t)kvh * p.kc_h + (int64_t)slot * p.kc_t;
    const T* vp = reinterpret_cast<const T*>(p.vc) + (int64_t)blk * p.vc_b + (int64_t)kvh * p.vc_h + (int64_t)slot * p.vc_t;
#pragma unroll
    for (int i = 0; i < DPL; ++i) {
      const int d = lane + 32 * i;
      kf[i] = (ok && d < D) ? DType<T>::to_f(kp[d]) : 0.f;
      vf[i] = (ok && d < D) ? DType<T>::to_f(vp[d]) : 0.f;
    }
#pragma unroll
    for (int r = 0; r < GH; ++r) {
      if (r < rows_valid) {
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < DPL; ++i) dot = fmaf(qv[r][i], kf[i], dot);
        dot = warp_sum(dot);
        const float s2 = round_through<T>(__fmul_rn(round_through<T>(dot), p.scale)) * kLog2e;
        const float m_new = fmaxf(m[r], s2);
        const float alpha = exp2f(m[r] - m_new);
        const float pr = exp2f(s2 - m_new);
        l[r] = l[r] * alpha + pr;
        m[r] = m_new;
        const float pv = round_through<T>(pr);
#pragma unroll
        for (int i = 0; i < DPL; ++i) acc[r][i] = fmaf(pv, vf[i], acc[r][i] * alpha);
      }
    }
  }

  // merge the 8 warps: sm = [WARPS][GH][D] accumulators then [WARPS][GH] (m, l)
  float* s_o = sm;
  float2* s_ml = reinterpret_cast<float2*>(sm + WARPS * GH * D);
#pragma unroll
  for (int r = 0; r < GH; ++r) {
    if (r < rows_valid) {
      if (lane == 0) s_ml[warp * GH + r] = make_float2(m[r], l[r]);
#pragma unroll
      for (int i = 0; i < DPL; ++i) {
        const int d = lane + 32 * i;
        if (d < D) s_o[(warp * GH + r) * D + d] = acc[r][i];
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < rows_valid * D; i += blockDim.x) {
    const int r = i / D, d = i - r * D;
    float mm = -INFINITY;
    for (int w = 0; w < WARPS; ++w) mm = fmaxf(mm, s_ml[w * GH + r].x);
    float ll = 0.f, a = 0.f;
    for (int w = 0; w < WARPS; ++w) {
      const float2 ml = s_ml[w * GH + r];
      if (ml.x != -INFINITY) {
        const float f = exp2f(ml.x - mm);
        ll += f * ml.y;
        a += f * s_o[(w * GH + r) * D + d];
      }
    }
    const int hq = q_head_of(p, kvh, ht * GH + r);
    if (p.num_splits == 1) {
      reinterpret_cast<T*>(p.out)[b * p.o_sb + hq * p.o_sh + d] = DType<T>::from_f(ll > 0.f ? a / ll : 0.f);
    } else {
      const int64_t slot = ((int64_t)b * p.num_q_heads + hq) * p.num_splits + split;
      p.part_o[slot * D + d] = a;
      if (d == 0) p.part_ml[slot] = make_float2(mm, ll);
    }
  }
}

// ======================================================================================================
// second pass: fold the per-split partials
// ======================================================================================================
// One CTA (4 warps) per (query head, sequence).  Every warp finds the global maximum and the normaliser from the
// (m, l) pairs (a lane per split), then accumulates the partial rows of the splits s = warp, warp + 4, ... - four
// rows requested per trip, so a 37-split fold (batch 1, 32k context) is ~3 load round trips instead of 37 - and the
// four warps' sums meet in shared memory.
template <typename T>
__global__ void __launch_bounds__(128) paged_decode_reduce_kernel(const float* __restrict__ part_o,
                                                                  const float2* __restrict__ part_ml,
                                                                  T* __restrict__ out, int num_q_heads, int head_dim,
                                                                  int num_splits, int64_t o_sb, int64_t o_sh) {
  constexpr int kWarps = 4, kMaxDpl = 8;  // head_dim <= 256
  __shared__ float s_acc[kWarps][256];
  const int hq = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t base = ((int64_t)b * num_q_heads + hq) * num_splits;
  const int dpl = (head_dim + 31) >> 5;
  pdl_wait();
  pdl_trigger();
  float m = -INFINITY;
  for (int s0 = 0; s0 < num_splits; s0 += 32)
    if (s0 + lane < num_splits) m = fmaxf(m, part_ml[base + s0 + lane].x);
  m = warp_max(m);
  float l = 0.f;
  float acc[kMaxDpl];
#pragma unroll
  for (int j = 0; j < kMaxDpl; ++j) acc[j] = 0.f;
  for (int s0 = 0; s0 < num_splits; s0 += 32) {
    float f = 0.f;
    if (s0 + lane < num_splits) {
      const float2 ml = part_ml[base + s0 + lane];
      if (ml.x != -INFINITY) {  // an empty split contributes nothing
        f = exp2f(ml.x - m);
        l += f * ml.y;
      }
    }
    const int n = min(32, num_splits - s0);
    for (int i = warp; i < n; i += 4 * kWarps) {  // four of this warp's splits per trip
      float v[4][kMaxDpl], fs[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int si = i + u * kWarps;
        fs[u] = __shfl_sync(0xffffffffu, f, si & 31);
        const float* row = part_o + (base + s0 + si) * head_dim;
#pragma unroll
        for (int j = 0; j < kMaxDpl; ++j) v[u][j] = (si < n && j < dpl && lane + 32 * j < head_dim) ? row[lane + 32 * j] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (i + u * kWarps < n && fs[u] != 0.f) {  // (an empty split's row is uninitialised memory)
#pragma unroll
          for (int j = 0; j < kMaxDpl; ++j) acc[j] = fmaf(fs[u], v[u][j], acc[j]);
        }
      }
    }
  }
  l = warp_sum(l);
#pragma unroll
  for (int j = 0; j < kMaxDpl; ++j)
    if (j < dpl && lane + 32 * j < head_dim) s_acc[warp][lane + 32 * j] = acc[j];
  __syncthreads();
  for (int d = threadIdx.x; d < head_dim; d += blockDim.x) {
    const float a = (s_acc[0][d] + s_acc[1][d]) + (s_acc[2][d] + s_acc[3][d]);
    out[b * o_sb + hq * o_sh + d] = DType<T>::from_f(l > 0.f ? a / l : 0.f);
  }
}

// ======================================================================================================
// host side
// ======================================================================================================
static int env_int(const char* name, int fallback) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : fallback;
}

static bool fast_path_ok(int dtype, int head_dim, int block_size, const void* q, const void* kc, const void* vc,
                         int64_t q_sb, int64_t q_sh, int64_t kc_b, int64_t kc_h, int64_t kc_t, int64_t vc_b,
                         int64_t vc_h, int64_t vc_t) {
  if (dtype == MOJO_B200_F32) return false;
  if (head_dim != 64 && head_dim != 128) return false;
  if (block_size < 8 || (block_size & (block_size - 1))) return false;
  const int64_t strides[] = {kc_b, kc_h, kc_t, vc_b, vc_h, vc_t};
  for (int64_t s : strides)
    if (s <= 0 || s % 8) return false;  // TMA: byte strides multiple of 16
  if (!aligned16(kc) || !aligned16(vc)) return false;
  if ((q_sb | q_sh) % 2 || ((uintptr_t)q & 3)) return false;  // 32-bit Q fragment loads
  return env_int("MOJO_B200_DECODE_FORCE_SIMT", 0) == 0;
}

static int choose_splits(int batch, int num_kv_heads, int head_tiles, int64_t max_seq_len) {
  const int forced = env_int("MOJO_B200_DECODE_SPLITS", 0);
  if (forced > 0) return forced;
  const int64_t tiles = (max_seq_len + kTile - 1) / kTile;
  if (tiles <= 1) return 1;
  const int64_t base = (int64_t)batch * num_kv_heads * head_tiles;
  const int64_t slots = (int64_t)kNumSMs * 2;  // two resident CTAs per SM
  int best = 1;
  double best_cost = 1e30;
  const int max_splits = (int)(tiles < 64 ? tiles : 64);
  for (int s = 1; s <= max_splits; ++s) {
    const int64_t per = (tiles + s - 1) / s;
    const int64_t waves = (base * s + slots - 1) / slots;
    // cost in tile-times: every wave streams `per` tiles plus a fixed start-up/merge overhead; the second
    // pass costs a little more with every extra split.
    const double cost = (double)waves * ((double)per + 3.0) + (s > 1 ? 0.05 * s + 1.0 : 0.0);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = s;
    }
  }
  return best;
}

}  // namespace mojo

extern "C" int mojo_b200_paged_decode_num_splits(int batch, int num_q_heads, int num_kv_heads, int head_dim,
                                                 int block_size, int64_t max_seq_len, int dtype) {
  (void)head_dim; (void)block_size; (void)dtype;
  if (batch <= 0 || num_kv_heads <= 0 || num_q_heads <= 0) return 1;
  const int group = num_q_heads / num_kv_heads;
  const int head_tiles = (group + 15) / 16;
  return mojo::choose_splits(batch, num_kv_heads, head_tiles, max_seq_len);
}

extern "C" size_t mojo_b200_paged_decode_workspace_bytes(int batch, int num_q_heads, int head_dim, int num_splits) {
  if (num_splits <= 1 || batch <= 0) return 0;
  const size_t slots = (size_t)batch * num_q_heads * num_splits;
  return slots * head_dim * sizeof(float) + slots * sizeof(float2) + 256;
}

static int paged_decode_impl(
    const void* query, const void* key_cache, const void* value_cache, const int32_t* total_seq_lens,
    const int32_t* block_tables, void* out, void* workspace, size_t workspace_bytes, int batch, int num_q_heads,
    int num_kv_heads, int head_dim, int64_t num_blocks, int block_size, int max_blocks_per_seq, int64_t table_stride,
    int64_t max_seq_len, int64_t q_stride_b, int64_t q_stride_h, int64_t o_stride_b, int64_t o_stride_h,
    int64_t kc_stride_b, int64_t kc_stride_h, int64_t kc_stride_t, int64_t vc_stride_b, int64_t vc_stride_h,
    int64_t vc_stride_t, float softmax_scale, int gqa_interleave, int num_splits, int win_local, int win_global,
    int dtype, void* stream) {
  using namespace mojo;
  MOJO_REQUIRE(batch >= 0 && num_q_heads > 0 && num_kv_heads > 0 && head_dim > 0 && block_size > 0 &&
                   max_blocks_per_seq >= 0 && num_blocks >= 0,
               MOJO_B200_EINVAL, "paged_decode: bad sizes");
  MOJO_REQUIRE(num_q_heads % num_kv_heads == 0, MOJO_B200_EINVAL, "paged_decode: Hq %d not a multiple of Hkv %d",
               num_q_heads, num_kv_heads);
  MOJO_REQUIRE(dtype >= 0 && dtype <= 2, MOJO_B200_EINVAL, "paged_decode: bad dtype %d", dtype);
  if (batch == 0) return 0;
  MOJO_REQUIRE(query && out && total_seq_lens, MOJO_B200_EINVAL, "paged_decode: null tensor pointer");
  MOJO_REQUIRE(max_blocks_per_seq == 0 || (block_tables && key_cache && value_cache) || num_blocks == 0,
               MOJO_B200_EINVAL, "paged_decode: null cache / table pointer");
  MOJO_REQUIRE(head_dim <= 256, MOJO_B200_EUNSUPPORTED, "paged_decode: head_dim %d > 256", head_dim);
  MOJO_REQUIRE(batch <= 65535, MOJO_B200_EUNSUPPORTED, "paged_decode: batch %d > 65535", batch);

  const int64_t table_cap = (int64_t)max_blocks_per_seq * block_size;
  if (max_seq_len <= 0 || max_seq_len > table_cap) max_seq_len = table_cap;
  const int group = num_q_heads / num_kv_heads;
  cudaStream_t s = (cudaStream_t)stream;

  const bool fast = fast_path_ok(dtype, head_dim, block_size, query, key_cache, value_cache, q_stride_b, q_stride_h,
                                 kc_stride_b, kc_stride_h, kc_stride_t, vc_stride_b, vc_stride_h, vc_stride_t) &&
                    num_blocks > 0 && max_blocks_per_seq > 0;
  const bool windowed = win_local >= 0 || win_global >= 0;
  MOJO_REQUIRE(fast || !windowed, MOJO_B200_EUNSUPPORTED,
               "paged_decode_swa: windows are built on the tensor-tile kernel only (bf16/fp16, head_dim 64/128, "
               "power-of-two pages, aligned strides)");
  if (windowed) {  // the splits divide the VISIBLE tiles: size them for the window, not the context
    const int64_t vis = (win_local >= 0 ? (int64_t)win_local + 1 : 0) + (win_global >= 0 ? win_global : 0) + 2 * kTile;
    if (vis < max_seq_len) max_seq_len = vis;
  }
  const int gh = fast ? 16 : (group >= 8 ? 8 : 4);
  const int head_tiles = (group + gh - 1) / gh;
  MOJO_REQUIRE((int64_t)num_kv_heads * head_tiles <= 65535, MOJO_B200_EUNSUPPORTED, "paged_decode: too many heads");
  if (num_splits <= 0) num_splits = choose_splits(batch, num_kv_heads, head_tiles, max_seq_len);
  {
    const int64_t tiles = (max_seq_len + kTile - 1) / kTile;
    if (num_splits > tiles) num_splits = (int)(tiles > 0 ? tiles : 1);
  }

  DecodeParams p;
  memset(&p, 0, sizeof(p));
  p.q = query; p.out = out; p.seq_lens = total_seq_lens; p.tables = block_tables; p.kc = key_cache; p.vc = value_cache;
  p.table_stride = table_stride; p.num_blocks = num_blocks;
  p.num_q_heads = num_q_heads; p.num_kv_heads = num_kv_heads; p.group = group; p.head_tiles = head_tiles;
  p.head_dim = head_dim; p.block_size = block_size; p.max_blocks = max_blocks_per_seq;
  p.log2_bs = 0;
  while ((1 << p.log2_bs) < block_size) ++p.log2_bs;
  p.q_sb = q_stride_b; p.q_sh = q_stride_h; p.o_sb = o_stride_b; p.o_sh = o_stride_h;
  p.kc_b = kc_stride_b; p.kc_h = kc_stride_h; p.kc_t = kc_stride_t;
  p.vc_b = vc_stride_b; p.vc_h = vc_stride_h; p.vc_t = vc_stride_t;
  p.scale = softmax_scale; p.interleave = gqa_interleave ? 1 : 0; p.num_splits = num_splits;
  p.win_local = win_local; p.win_global = win_global;
  p.err = error_word();
#ifdef MOJO_DECODE_TRACE
  if (const char* tp = getenv("MOJO_B200_DECODE_TRACE_PTR")) p.trace = reinterpret_cast<unsigned long long*>(strtoull(tp, nullptr, 0));
#endif
  if (fast && num_splits > 1 && num_splits <= 64 && env_int("MOJO_B200_DECODE_FOLD", 1) != 0) {
    // in-kernel fold: needs the registered, zero-between-launches arrival counters (mojo_b200_set_decode_tickets);
    // consecutive launches rotate over disjoint slices, so launches in flight on different streams do not share one
    int64_t count = 0;
    int* words = decode_tickets(&count);
    const int64_t groups = (int64_t)batch * num_kv_heads * head_tiles;
    if (words && count >= groups) {
      static std::atomic<unsigned> launch_seq{0};  // (host threads may launch concurrently)
      const int64_t slices = count / groups < 16 ? count / groups : 16;
      p.tickets = words + (int64_t)(launch_seq.fetch_add(1, std::memory_order_relaxed) % (unsigned)slices) * groups;
    }
  }

  if (num_splits > 1) {
    const size_t need = mojo_b200_paged_decode_workspace_bytes(batch, num_q_heads, head_dim, num_splits);
    MOJO_REQUIRE(workspace && workspace_bytes >= need, MOJO_B200_EWORKSPACE,
                 "paged_decode: workspace of %zu bytes needed for %d splits, got %zu", need, num_splits, workspace_bytes);
    uintptr_t w = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;
    const size_t slots = (size_t)batch * num_q_heads * num_splits;
    p.part_o = reinterpret_cast<float*>(w);
    p.part_ml = reinterpret_cast<float2*>(w + slots * head_dim * sizeof(float));
  }

  dim3 grid((unsigned)num_splits, (unsigned)(num_kv_heads * head_tiles), (unsigned)batch);

  if (fast) {
    const int tile_bytes = kTile * head_dim * 2;
    int stages = env_int("MOJO_B200_DECODE_STAGES", 0);
    if (stages <= 0) stages = head_dim == 128 ? 3 : 6;  // ~96 KB per CTA -> two CTAs per SM
    if (stages < 2) stages = 2;
    while ((size_t)stages * 2 * tile_bytes + 2 * stages * 8 + 1024 > 224 * 1024) --stages;
    p.stages = stages;
    const size_t smem = (size_t)stages * 2 * tile_bytes + 2 * stages * sizeof(uint64_t) + 1024;

    const char* layout = getenv("MOJO_B200_DECODE_LAYOUT");
    bool split_halves = head_dim == 128 && !(layout && !strcmp(layout, "natural"));
    CUtensorMap k_map, v_map;
    const int box_rows = block_size < kTile ? block_size : kTile;
    int rc = build_cache_map(key_cache, dtype, head_dim, block_size, num_kv_heads, num_blocks, kc_stride_b, kc_stride_h,
                             kc_stride_t, split_halves, box_rows, &k_map);
    if (rc != 0 && split_halves && !(layout && !strcmp(layout, "split"))) {
      split_halves = false;  // driver rejected the permuted strides: natural order (2-way ldmatrix conflicts)
      rc = build_cache_map(key_cache, dtype, head_dim, block_size, num_kv_heads, num_blocks, kc_stride_b, kc_stride_h,
                           kc_stride_t, false, box_rows, &k_map);
    }
    if (rc != 0) return rc;
    rc = build_cache_map(value_cache, dtype, head_dim, block_size, num_kv_heads, num_blocks, vc_stride_b, vc_stride_h,
                         vc_stride_t, split_halves, box_rows, &v_map);
    if (rc != 0) return rc;

#define LAUNCH_FAST(TT, DD, SH)                                                                          \
  do {                                                                                                   \
    auto kern = p.tickets ? paged_decode_mma_kernel<TT, DD, SH, true> : paged_decode_mma_kernel<TT, DD, SH, false>; \
    MOJO_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    MOJO_CUDA_OK(launch_pdl(kern, grid, dim3(kDecodeThreads), smem, s, k_map, v_map, p));                \
  } while (0)
    if (dtype == MOJO_B200_BF16) {
      if (head_dim == 128) { if (split_halves) LAUNCH_FAST(__nv_bfloat16, 128, true); else LAUNCH_FAST(__nv_bfloat16, 128, false); }
      else LAUNCH_FAST(__nv_bfloat16, 64, false);
    } else {
      if (head_dim == 128) { if (split_halves) LAUNCH_FAST(__half, 128, true); else LAUNCH_FAST(__half, 128, false); }
      else LAUNCH_FAST(__half, 64, false);
    }
#undef LAUNCH_FAST
    if (int rc2 = check_launch("paged_decode_mma_kernel")) return rc2;
  } else {
    const int dpl = (head_dim + 31) / 32;
    const size_t smem = (size_t)8 * gh * head_dim * sizeof(float) + (size_t)8 * gh * sizeof(float2);
#define LAUNCH_SIMT(TT, GG, DP)                                                                         \
  do {                                                                                                  \
    auto kern = paged_decode_simt_kernel<TT, GG, DP>;                                                   \
    MOJO_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    MOJO_CUDA_OK(launch_pdl(kern, grid, dim3(256), smem, s, p));                                        \
  } while (0)
#define SIMT_BY_D(TT, GG)                              \
  do {                                                 \
    if (dpl <= 2) LAUNCH_SIMT(TT, GG, 2);              \
    else if (dpl <= 4) LAUNCH_SIMT(TT, GG, 4);         \
    else LAUNCH_SIMT(TT, GG, 8);                       \
  } while (0)
    int rc = dispatch_dtype(dtype, [&](auto tag) {
      using TT = decltype(tag);
      if (gh == 8) SIMT_BY_D(TT, 8); else SIMT_BY_D(TT, 4);
      return 0;
    });
#undef SIMT_BY_D
#undef LAUNCH_SIMT
    if (rc) return rc;
    if (int rc2 = check_launch("paged_decode_simt_kernel")) return rc2;
  }

  if (num_splits > 1 && !p.tickets) {
    dim3 rgrid((unsigned)num_q_heads, (unsigned)batch);
    int rc = dispatch_dtype(dtype, [&](auto tag) {
      using TT = decltype(tag);
      MOJO_CUDA_OK(launch_pdl(paged_decode_reduce_kernel<TT>, rgrid, dim3(128), 0, s, (const float*)p.part_o,
                              (const float2*)p.part_ml, (TT*)out, num_q_heads, head_dim, num_splits, o_stride_b,
                              o_stride_h));
      return 0;
    });
    if (rc) return rc;
    return check_launch("paged_decode_reduce_kernel");
  }
  return 0;
}

extern "C" int mojo_b200_paged_decode_gqa(
    const void* query, const void* key_cache, const void* value_cache, const int32_t* total_seq_lens,
    const int32_t* block_tables, void* out, void* workspace, size_t workspace_bytes, int batch, int num_q_heads,
    int num_kv_heads, int head_dim, int64_t num_blocks, int block_size, int max_blocks_per_seq, int64_t table_stride,
    int64_t max_seq_len, int64_t q_stride_b, int64_t q_stride_h, int64_t o_stride_b, int64_t o_stride_h,
    int64_t kc_stride_b, int64_t kc_stride_h, int64_t kc_stride_t, int64_t vc_stride_b, int64_t vc_stride_h,
    int64_t vc_stride_t, float softmax_scale, int gqa_interleave, int num_splits, int dtype, void* stream) {
  return paged_decode_impl(query, key_cache, value_cache, total_seq_lens, block_tables, out, workspace, workspace_bytes,
                           batch, num_q_heads, num_kv_heads, head_dim, num_blocks, block_size, max_blocks_per_seq,
                           table_stride, max_seq_len, q_stride_b, q_stride_h, o_stride_b, o_stride_h, kc_stride_b,
                           kc_stride_h, kc_stride_t, vc_stride_b, vc_stride_h, vc_stride_t, softmax_scale, gqa_interleave,
                           num_splits, -1, -1, dtype, stream);
}

extern "C" int mojo_b200_paged_decode_swa(
    const void* query, const void* key_cache, const void* value_cache, const int32_t* total_seq_lens,
    const int32_t* block_tables, void* out, void* workspace, size_t workspace_bytes, int batch, int num_q_heads,
    int num_kv_heads, int head_dim, int64_t num_blocks, int block_size, int max_blocks_per_seq, int64_t table_stride,
    int64_t max_seq_len, int64_t q_stride_b, int64_t q_stride_h, int64_t o_stride_b, int64_t o_stride_h,
    int64_t kc_stride_b, int64_t kc_stride_h, int64_t kc_stride_t, int64_t vc_stride_b, int64_t vc_stride_h,
    int64_t vc_stride_t, float softmax_scale, int gqa_interleave, int num_splits, int local_window_size,
    int global_window_size, int dtype, void* stream) {
  using namespace mojo;
  MOJO_REQUIRE(local_window_size >= -1 && global_window_size >= -1, MOJO_B200_EINVAL,
               "paged_decode_swa: window sizes must be >= 0, or -1 for None");
  return paged_decode_impl(query, key_cache, value_cache, total_seq_lens, block_tables, out, workspace, workspace_bytes,
                           batch, num_q_heads, num_kv_heads, head_dim, num_blocks, block_size, max_blocks_per_seq,
                           table_stride, max_seq_len, q_stride_b, q_stride_h, o_stride_b, o_stride_h, kc_stride_b,
                           kc_stride_h, kc_stride_t, vc_stride_b, vc_stride_h, vc_stride_t, softmax_scale, gqa_interleave,
                           num_splits, local_window_size, global_window_size, dtype, stream);
}
