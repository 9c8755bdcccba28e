// MojoPagedPrefillGQA (var-len causal attention over a paged KV cache) and MojoSdpa (dense non-causal
// attention over strided [B,H,S,D] views): FlashAttention-2 style forward on mma.sync tensor-core tiles.
//
// This is the general path of the two ops: any 16-bit dtype, D in {64,128}, any power-of-two page size >= 8,
// ragged sequences, cached prefixes, GQA in both layouts.  (The tcgen05/TMEM kernel in attention_fwd_sm100.cu
// takes the shapes it is specialised for - head_dim 128, query chunks of >= 192 rows; the entry points below pick.)
//
//   grid  (q tiles of 64 rows [heaviest first], Hq, B); 160 threads = 4 consumer warps (16 query rows each) +
//         1 TMA producer warp; 2 CTAs/SM.
//   K/V   64-token tiles through a full/empty mbarrier ring, one TMA tensor copy per (page, K|V); a dense
//         tensor is addressed as "one page per batch element" by the same 5-D tensor map.
//   math  S = Q K^T (fp32 accumulate) -> [paged prefill: rounded to the input dtype, as the golden's einsum]
//         -> * scale -> causal / length mask -> online softmax (log2 domain) -> P rounded to the input dtype
//         -> O += P V (fp32) -> O / l.
// Sliding windows (MojoPagedPrefillSWA, reference attention.py:533-643; one-row-per-sequence decode fallback): on top
// of the causal limit a key is visible iff key + local >= position or key < global; KV tiles between the global prefix
// and the first row's window are skipped by producer and consumers alike.
// FLOPs = 4 * D * (number of unmasked (q, k) pairs) per query head.
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "attention_sm100.cuh"
#include "tma.cuh"

namespace mojo {

constexpr int kAttnConsumerWarps = 4;
constexpr int kAttnThreads = 32 * (kAttnConsumerWarps + 1);
constexpr int kBM = 16 * kAttnConsumerWarps;  // query rows per CTA
constexpr float kLog2eF = 1.4426950408889634f;

struct AttnParams {
  const void* q;
  void* out;
  const int32_t* cu_q;    // paged: [B+1]
  const int32_t* cu_kv;   // paged: [B+1] or null
  const int32_t* tables;  // paged: [B, MB]
  int64_t table_stride;
  int max_blocks, block_size, log2_bs, box_rows;
  int num_q_heads, num_kv_heads, group;
  int64_t q_len_dense, kv_len_dense;
  int64_t q_st, q_sh, q_sb, o_st, o_sh, o_sb;
  float scale_log2;  // softmax_scale * log2(e)
  int interleave, causal, dense, stages, round_scores;
  int packed;  // K/V are [total tokens, heads, D]: sequence b at rows cu_kv[b].. (non-paged MojoSWA)
  // sliding-window attention (MojoPagedPrefillSWA / MojoPagedDecodeSWA): on top of the causal limit a key is visible
  // iff key + win_local >= position or key < win_global; -1 = that window is not set (both -1: plain causal)
  int win_local, win_global;
  // MojoSdpa attn_mask (dense, non-causal): bool bytes, 1 = visible, strides in bytes per batch / head / row; null = none
  const uint8_t* mask;
  int64_t mask_sb, mask_sh, mask_sq;
};

template <typename T, int D, bool SPLIT_HALVES>
__global__ void __launch_bounds__(kAttnThreads, 2)
attn_fwd_mma_kernel(const __grid_constant__ CUtensorMap k_map, const __grid_constant__ CUtensorMap v_map,
                    const AttnParams p) {
  constexpr int NH = D / 64;
  constexpr int KS = D / 16;
  constexpr int TILE_BYTES = kTile * D * 2;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stages = p.stages;
  uint8_t* tiles = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * 2 * TILE_BYTES);
  uint64_t* empty = full + stages;

  const int b = blockIdx.z;
  const int hq = blockIdx.y;
  const int q_tile = gridDim.x - 1 - blockIdx.x;  // causal: the longest rows start first
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int64_t q_start;
  int q_len, kv_len, kv_start = 0;
  if (p.dense) {
    q_start = 0;
    q_len = (int)p.q_len_dense;
    kv_len = (int)p.kv_len_dense;
  } else {
    q_start = p.cu_q[b];
    q_len = p.cu_q[b + 1] - (int)q_start;
    kv_len = p.cu_kv ? p.cu_kv[b + 1] - p.cu_kv[b] : q_len;
    kv_start = p.cu_kv ? p.cu_kv[b] : (int)q_start;
  }
  const int m0 = q_tile * kBM;
  if (m0 >= q_len || kv_len <= 0) return;
  const int off = kv_len - q_len;  // query row t sees keys 0 .. off + t
  const int last_row = min(m0 + kBM, q_len) - 1;
  const int n_end = p.causal ? min(kv_len, off + last_row + 1) : kv_len;
  const int n_tiles = n_end > 0 ? (n_end + kTile - 1) / kTile : 0;
  if (n_tiles == 0) return;  // rows that see no key keep the zeros the output was initialised with
  const int kvh = p.interleave ? hq % p.num_kv_heads : hq / p.group;
  // windows: row at position pos sees keys >= pos - win_local (if set) and keys < win_global (if set)
  const bool has_win = p.causal && (p.win_local >= 0 || p.win_global >= 0);
  const int win_g = has_win && p.win_global >= 0 ? p.win_global : 0;
  auto win_lo = [&](int row) -> int {  // first key of the row's local window
    if (!has_win) return INT_MIN;
    return p.win_local >= 0 ? off + row - p.win_local : INT_MAX;
  };
  // KV tiles between the global prefix and the first row's local window are invisible to the whole CTA: producer and
  // consumers skip them alike (the same predicate on both sides keeps the ring in step)
  const int cta_lo = win_lo(m0);
  auto tile_live = [&](int it) -> bool { return it * kTile < win_g || it * kTile + kTile - 1 >= cta_lo; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kAttnConsumerWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const int box_rows = p.box_rows;
  const int boxes_per_tile = kTile / box_rows;

  if (warp == kAttnConsumerWarps) {
    // ------------------------------------------------------------------ producer warp
    if (lane == 0) {
      tma_prefetch_desc(&k_map);
      tma_prefetch_desc(&v_map);
    }
    const int32_t* table = (p.dense || p.packed) ? nullptr : p.tables + (int64_t)b * p.table_stride;
    const uint32_t box_bytes = (uint32_t)box_rows * D * 2;
    for (int it = 0, use = 0; it < n_tiles; ++it) {
      if (!tile_live(it)) continue;
      const int stage = use % stages;
      const uint32_t phase = (uint32_t)(use / stages) & 1u;
      ++use;
      const int tok0 = it * kTile;
      const int want = min(boxes_per_tile, (kv_len - tok0 + box_rows - 1) / box_rows);
      int blk = 0, row_in_page = 0;
      if (lane < want) {
        const int tok = tok0 + lane * box_rows;
        if (p.dense) {
          blk = b;
          row_in_page = tok;
        } else if (p.packed) {
          blk = 0;
          row_in_page = kv_start + tok;
        } else {
          const int page = tok >> p.log2_bs;
          row_in_page = tok & (p.block_size - 1);
          blk = page < p.max_blocks ? table[page] : -1;
        }
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
  const int g = lane >> 2, c = lane & 3;
  const int row_lo = m0 + warp * 16 + g, row_hi = row_lo + 8;  // rows inside the sequence

  uint32_t qf[KS][4];
  {
    const T* qb = reinterpret_cast<const T*>(p.q) + (int64_t)b * p.q_sb + (int64_t)hq * p.q_sh;
    const T* r0 = row_lo < q_len ? qb + (q_start + row_lo) * p.q_st : nullptr;
    const T* r1 = row_hi < q_len ? qb + (q_start + row_hi) * p.q_st : nullptr;
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

  const int mat = lane >> 3, mr = lane & 7;
  auto line_of = [&](int row, int half) -> uint32_t {
    if (SPLIT_HALVES) {
      const int box = row / box_rows, r = row - box * box_rows;
      return (uint32_t)(box * box_rows * NH + half * box_rows + r);
    }
    return (uint32_t)(row * NH + half);
  };
  // last key each of this thread's two rows may see
  const int lim_lo = p.causal ? min(kv_len - 1, off + row_lo) : kv_len - 1;
  const int lim_hi = p.causal ? min(kv_len - 1, off + row_hi) : kv_len - 1;
  const int warp_first_lim = p.causal ? off + m0 + warp * 16 : kv_len - 1;       // most restrictive row
  const int warp_last_lim = p.causal ? off + m0 + warp * 16 + 15 : kv_len - 1;   // least restrictive row
  const int lo_lo = win_lo(row_lo), lo_hi = win_lo(row_hi);                      // this thread's rows' local windows
  const int warp_min_lo = win_lo(m0 + warp * 16), warp_max_lo = win_lo(m0 + warp * 16 + 15);
  const float scale_log2 = p.scale_log2;
  const bool round_scores = p.round_scores != 0;

  for (int it = 0, use = 0; it < n_tiles; ++it) {
    if (!tile_live(it)) continue;
    const int stage = use % stages;
    const uint32_t phase = (uint32_t)(use / stages) & 1u;
    ++use;
    const int n0 = it * kTile;
    uint8_t* sk = tiles + (size_t)stage * 2 * TILE_BYTES;
    uint8_t* sv = sk + TILE_BYTES;
    const uint32_t sk_a = smem_u32(sk), sv_a = smem_u32(sv);

    mbar_wait(&full[stage], phase);

    // otherwise none of this warp's rows sees the tile (causal limit / between the global prefix and the windows)
    if (n0 <= warp_last_lim && (n0 < win_g || n0 + kTile - 1 >= warp_min_lo)) {
      const int valid = kv_len - n0;
      if (valid < kTile) {
        // tail tile: zero the V rows past the end of the sequence (every warp reads all 64 rows, and each
        // zeroes them itself, so no cross-warp ordering is needed - all writers store zeros)
        for (int r = valid + (lane >> 1); r < kTile; r += 16) {
#pragma unroll
          for (int h = 0; h < NH; ++h) {
            uint4* dst = reinterpret_cast<uint4*>(sv + line_of(r, h) * 128u + (lane & 1) * 64);
            dst[0] = dst[1] = dst[2] = dst[3] = make_uint4(0, 0, 0, 0);
          }
        }
        __syncwarp();
      }

      // ---- S = Q K^T : 16 rows x 64 keys
      float s[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          const int row = np * 16 + (mat >> 1) * 8 + mr;
          uint32_t b0, b1, b2, b3;
          ldsm_x4(sk_a + swz128(line_of(row, ks >> 2), ((ks * 2) & 7) + (mat & 1)), b0, b1, b2, b3);
          Mma16816<T>::run(s[2 * np], qf[ks], b0, b1);
          Mma16816<T>::run(s[2 * np + 1], qf[ks], b2, b3);
        }
      }

      // ---- rounding, scale, mask
      const bool need_mask = n0 + kTile - 1 > warp_first_lim || valid < kTile || p.mask != nullptr ||
                             (has_win && !(n0 >= warp_max_lo || n0 + kTile - 1 < win_g));
      const uint8_t* mrow_lo = nullptr;
      const uint8_t* mrow_hi = nullptr;
      if (p.mask) {
        const uint8_t* mb = p.mask + (int64_t)b * p.mask_sb + (int64_t)hq * p.mask_sh;
        mrow_lo = mb + (int64_t)min(row_lo, q_len - 1) * p.mask_sq;
        mrow_hi = mb + (int64_t)min(row_hi, q_len - 1) * p.mask_sq;
      }
      float tile_lo = -INFINITY, tile_hi = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float v = round_scores ? round_through<T>(s[j][e]) : s[j][e];
          v *= scale_log2;
          if (need_mask) {
            const int key = n0 + j * 8 + 2 * c + (e & 1);
            if (key > (e < 2 ? lim_lo : lim_hi) || (key < (e < 2 ? lo_lo : lo_hi) && key >= win_g)) v = -INFINITY;
            else if (p.mask && !(e < 2 ? mrow_lo : mrow_hi)[key]) v = -INFINITY;
          }
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
      float sum_lo = 0.f, sum_hi = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j][0] = exp2f(s[j][0] - base_lo);
        s[j][1] = exp2f(s[j][1] - base_lo);
        s[j][2] = exp2f(s[j][2] - base_hi);
        s[j][3] = exp2f(s[j][3] - base_hi);
        sum_lo += s[j][0] + s[j][1];
        sum_hi += s[j][2] + s[j][3];
      }
      l_lo = l_lo * a_lo + sum_lo;
      l_hi = l_hi * a_hi + sum_hi;
#pragma unroll
      for (int i = 0; i < 2 * KS; ++i) {
        o[i][0] *= a_lo;
        o[i][1] *= a_lo;
        o[i][2] *= a_hi;
        o[i][3] *= a_hi;
      }

      // ---- O += P V
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t pa[4];
        pa[0] = Mma16816<T>::pack(s[2 * kk][0], s[2 * kk][1]);
        pa[1] = Mma16816<T>::pack(s[2 * kk][2], s[2 * kk][3]);
        pa[2] = Mma16816<T>::pack(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[3] = Mma16816<T>::pack(s[2 * kk + 1][2], s[2 * kk + 1][3]);
        const int row = kk * 16 + (mat & 1) * 8 + mr;
#pragma unroll
        for (int dp = 0; dp < KS; ++dp) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_trans(sv_a + swz128(line_of(row, dp >> 2), ((dp * 2) & 7) + (mat >> 1)), b0, b1, b2, b3);
          Mma16816<T>::run(o[2 * dp], pa, b0, b1);
          Mma16816<T>::run(o[2 * dp + 1], pa, b2, b3);
        }
      }
    }

    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
  }

  // -------------------------------------------------------------------- epilogue: O / l -> out
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
  const float inv_lo = l_lo > 0.f ? 1.f / l_lo : 0.f, inv_hi = l_hi > 0.f ? 1.f / l_hi : 0.f;
  T* ob = reinterpret_cast<T*>(p.out) + (int64_t)b * p.o_sb + (int64_t)hq * p.o_sh;
  if (row_lo < q_len) {
    T* dst = ob + (q_start + row_lo) * p.o_st;
#pragma unroll
    for (int i = 0; i < 2 * KS; ++i)
      *reinterpret_cast<uint32_t*>(dst + i * 8 + 2 * c) = Mma16816<T>::pack(o[i][0] * inv_lo, o[i][1] * inv_lo);
  }
  if (row_hi < q_len) {
    T* dst = ob + (q_start + row_hi) * p.o_st;
#pragma unroll
    for (int i = 0; i < 2 * KS; ++i)
      *reinterpret_cast<uint32_t*>(dst + i * 8 + 2 * c) = Mma16816<T>::pack(o[i][2] * inv_hi, o[i][3] * inv_hi);
  }
}

// ------------------------------------------------------------------------------------------------------
static int attn_env_int(const char* name, int fallback) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : fallback;
}

struct KvDesc {
  const void* k;
  const void* v;
  int64_t k_b, k_h, k_t, v_b, v_h, v_t;  // element strides: block|batch, head, token
  int64_t rows_per_block;                // page size, or kv_len for a dense tensor
  int64_t num_blocks;                    // pages, or batch for a dense tensor
  int num_kv_heads;
};

static int launch_attn_mma(const KvDesc& kv, AttnParams& p, int head_dim, int dtype, dim3 grid, cudaStream_t s) {
  const int64_t strides[] = {kv.k_b, kv.k_h, kv.k_t, kv.v_b, kv.v_h, kv.v_t};
  for (int64_t st : strides)
    MOJO_REQUIRE(st > 0 && st % 8 == 0, MOJO_B200_EUNSUPPORTED,
                 "attention: K/V strides must be positive multiples of 8 elements (TMA), got %lld", (long long)st);
  MOJO_REQUIRE(aligned16(kv.k) && aligned16(kv.v), MOJO_B200_EUNSUPPORTED, "attention: K/V base must be 16-byte aligned");
  MOJO_REQUIRE(((p.q_st | p.q_sh | p.q_sb | p.o_st | p.o_sh | p.o_sb) % 2) == 0 && ((uintptr_t)p.q & 3) == 0 &&
                   ((uintptr_t)p.out & 3) == 0,
               MOJO_B200_EUNSUPPORTED, "attention: q/out strides must be even and bases 4-byte aligned");

  const int tile_bytes = kTile * head_dim * 2;
  int stages = attn_env_int("MOJO_B200_ATTN_STAGES", 0);
  if (stages <= 0) stages = head_dim == 128 ? 3 : 6;
  if (stages < 2) stages = 2;
  p.stages = stages;
  const size_t smem = (size_t)stages * 2 * tile_bytes + 2 * stages * sizeof(uint64_t) + 1024;

  const char* layout = getenv("MOJO_B200_DECODE_LAYOUT");
  bool split_halves = head_dim == 128 && !(layout && !strcmp(layout, "natural"));
  CUtensorMap k_map, v_map;
  int rc = build_cache_map(kv.k, dtype, head_dim, kv.rows_per_block, kv.num_kv_heads, kv.num_blocks, kv.k_b, kv.k_h,
                           kv.k_t, split_halves, p.box_rows, &k_map);
  if (rc != 0 && split_halves && !(layout && !strcmp(layout, "split"))) {
    split_halves = false;
    rc = build_cache_map(kv.k, dtype, head_dim, kv.rows_per_block, kv.num_kv_heads, kv.num_blocks, kv.k_b, kv.k_h,
                         kv.k_t, false, p.box_rows, &k_map);
  }
  if (rc != 0) return rc;
  rc = build_cache_map(kv.v, dtype, head_dim, kv.rows_per_block, kv.num_kv_heads, kv.num_blocks, kv.v_b, kv.v_h, kv.v_t,
                       split_halves, p.box_rows, &v_map);
  if (rc != 0) return rc;

#define LAUNCH_ATTN(TT, DD, SH)                                                                          \
  do {                                                                                                   \
    auto kern = attn_fwd_mma_kernel<TT, DD, SH>;                                                         \
    MOJO_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    kern<<<grid, kAttnThreads, smem, s>>>(k_map, v_map, p);                                              \
  } while (0)
  if (dtype == MOJO_B200_BF16) {
    if (head_dim == 128) { if (split_halves) LAUNCH_ATTN(__nv_bfloat16, 128, true); else LAUNCH_ATTN(__nv_bfloat16, 128, false); }
    else LAUNCH_ATTN(__nv_bfloat16, 64, false);
  } else {
    if (head_dim == 128) { if (split_halves) LAUNCH_ATTN(__half, 128, true); else LAUNCH_ATTN(__half, 128, false); }
    else LAUNCH_ATTN(__half, 64, false);
  }
#undef LAUNCH_ATTN
  return check_launch("attn_fwd_mma_kernel");
}

// Output rows no attention CTA writes must read as zeros (golden: `torch.zeros_like(query)`, attention.py:384):
// tokens past cu_q_lens[batch], every row of a sequence without keys, and - when kv_len < q_len - the leading rows
// whose causal window is empty.  Usually none exist and every CTA exits at once; the host no longer memsets `out`.
__global__ void prefill_zero_unseen_rows_kernel(void* out, const int32_t* cu_q, const int32_t* cu_kv, int batch,
                                                int64_t total_q, int num_q_heads, int head_dim, int64_t o_st,
                                                int64_t o_sh, int no_kv, const int32_t* tables, int64_t table_stride,
                                                int* err) {
  const int b = blockIdx.x;  // b == batch: the tail after the last sequence
  int64_t r0, r1;
  if (b == batch) {
    r0 = cu_q[batch];
    r1 = total_q;
  } else {
    const int64_t q_start = cu_q[b];
    const int64_t q_len = cu_q[b + 1] - q_start;
    const int64_t kv_len = no_kv ? 0 : (cu_kv ? (int64_t)cu_kv[b + 1] - cu_kv[b] : q_len);
    // the reference raises ValueError for a sequence with queries and keys but no first block (attention.py:396-397)
    if (err && tables && blockIdx.y == 0 && threadIdx.x == 0 && q_len > 0 && kv_len > 0 && tables[b * table_stride] < 0)
      atomicOr(err, 2);
    r0 = q_start;
    r1 = q_start + (kv_len <= 0 ? q_len : min(q_len, max((int64_t)0, q_len - kv_len)));
  }
  r0 = max(r0, (int64_t)0);
  r1 = min(r1, total_q);
  if (r1 <= r0) return;
  const int words = head_dim / 2;  // 2-byte elements, rows 4-byte aligned (checked by the launcher)
  const int64_t per_row = (int64_t)num_q_heads * words;
  uint16_t* base = reinterpret_cast<uint16_t*>(out);
  for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < (r1 - r0) * per_row;
       i += (int64_t)gridDim.y * blockDim.x) {
    const int64_t r = r0 + i / per_row;
    const int rem = (int)(i % per_row);
    const int h = rem / words, w = rem % words;
    *reinterpret_cast<uint32_t*>(base + r * o_st + (int64_t)h * o_sh + 2 * w) = 0u;
  }
}

}  // namespace mojo

static int paged_prefill_impl(
    const void* query, const void* key_cache, const void* value_cache, const int32_t* cu_q_lens,
    const int32_t* cu_total_seq_lens, const int32_t* block_tables, void* out, int64_t total_q_tokens, int batch,
    int num_q_heads, int num_kv_heads, int head_dim, int64_t num_blocks, int block_size, int max_blocks_per_seq,
    int64_t table_stride, int64_t max_q_len, int64_t max_kv_len, int64_t q_stride_t, int64_t q_stride_h,
    int64_t o_stride_t, int64_t o_stride_h, int64_t kc_stride_b, int64_t kc_stride_h, int64_t kc_stride_t,
    int64_t vc_stride_b, int64_t vc_stride_h, int64_t vc_stride_t, float softmax_scale, int gqa_interleave,
    int is_causal, int win_local, int win_global, int dtype, void* stream, int packed = 0) {
  using namespace mojo;
  (void)max_kv_len;
  MOJO_REQUIRE(total_q_tokens >= 0 && batch >= 0 && num_q_heads > 0 && num_kv_heads > 0 && head_dim > 0 &&
                   block_size > 0 && max_blocks_per_seq >= 0 && num_blocks >= 0,
               MOJO_B200_EINVAL, "paged_prefill: bad sizes");
  MOJO_REQUIRE(num_q_heads % num_kv_heads == 0, MOJO_B200_EINVAL, "paged_prefill: Hq %d not a multiple of Hkv %d",
               num_q_heads, num_kv_heads);
  if (total_q_tokens == 0) return 0;
  MOJO_REQUIRE(query && key_cache && value_cache && cu_q_lens && (block_tables || packed) && out, MOJO_B200_EINVAL,
               "paged_prefill: null tensor pointer");
  MOJO_REQUIRE(is_causal, MOJO_B200_EUNSUPPORTED, "paged_prefill: only causal attention is built");
  MOJO_REQUIRE(dtype == MOJO_B200_BF16 || dtype == MOJO_B200_F16, MOJO_B200_EUNSUPPORTED,
               "paged_prefill: bf16/fp16 only (tensor-core path)");
  MOJO_REQUIRE(head_dim == 64 || head_dim == 128, MOJO_B200_EUNSUPPORTED, "paged_prefill: head_dim %d not in {64,128}",
               head_dim);
  MOJO_REQUIRE(packed || (block_size >= 8 && (block_size & (block_size - 1)) == 0), MOJO_B200_EUNSUPPORTED,
               "paged_prefill: block_size %d must be a power of two >= 8", block_size);
  MOJO_REQUIRE(batch <= 65535 && num_q_heads <= 65535, MOJO_B200_EUNSUPPORTED, "paged_prefill: grid too large");
  MOJO_REQUIRE(head_dim % 2 == 0 && ((o_stride_t | o_stride_h) % 2) == 0 && ((uintptr_t)out & 3) == 0,
               MOJO_B200_EUNSUPPORTED, "paged_prefill: out strides must be even and the base 4-byte aligned");
  {
    // a global-only window of size 0 leaves no key visible to any row: every row reads as zero (both kernels)
    const int no_kv = max_blocks_per_seq == 0 || num_blocks == 0 || (win_local < 0 && win_global == 0);
    prefill_zero_unseen_rows_kernel<<<dim3((unsigned)batch + 1, 8), 256, 0, (cudaStream_t)stream>>>(
        out, cu_q_lens, cu_total_seq_lens, batch, total_q_tokens, num_q_heads, head_dim, o_stride_t, o_stride_h, no_kv,
        max_blocks_per_seq > 0 && !packed ? block_tables : nullptr, table_stride, error_word());
    const int rc = check_launch("prefill_zero_unseen_rows_kernel");
    if (rc != 0 || batch == 0 || no_kv) return rc;
  }
  if (max_q_len <= 0 || max_q_len > total_q_tokens) max_q_len = total_q_tokens;

  {  // tcgen05/TMEM kernel when the shape is covered
    AttnSm100Args a;
    memset(&a, 0, sizeof(a));
    a.win_local = win_local; a.win_global = win_global;
    a.q = query; a.out = out; a.q_rows = total_q_tokens; a.q_st = q_stride_t; a.q_sh = q_stride_h;
    a.o_st = o_stride_t; a.o_sh = o_stride_h;
    a.k = key_cache; a.v = value_cache; a.k_b = kc_stride_b; a.k_h = kc_stride_h; a.k_t = kc_stride_t;
    a.v_b = vc_stride_b; a.v_h = vc_stride_h; a.v_t = vc_stride_t;
    a.rows_per_block = block_size; a.num_blocks = num_blocks;
    a.cu_q = cu_q_lens; a.cu_kv = cu_total_seq_lens; a.tables = block_tables; a.table_stride = table_stride;
    a.max_blocks = max_blocks_per_seq; a.batch = batch; a.num_q_heads = num_q_heads; a.num_kv_heads = num_kv_heads;
    a.head_dim = head_dim; a.max_q_len = max_q_len; a.softmax_scale = softmax_scale;
    a.interleave = gqa_interleave ? 1 : 0; a.causal = 1; a.dense = 0; a.round_scores = 1; a.dtype = dtype;
    a.packed = packed;
    const int rc = launch_attn_sm100(a, (cudaStream_t)stream);
    if (rc != kAttnNotEligible) return rc;
  }

  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.q = query; p.out = out; p.cu_q = cu_q_lens; p.cu_kv = cu_total_seq_lens; p.tables = block_tables;
  p.table_stride = table_stride; p.max_blocks = max_blocks_per_seq; p.block_size = block_size;
  while (!packed && (1 << p.log2_bs) < block_size) ++p.log2_bs;
  p.box_rows = (packed || block_size >= kTile) ? kTile : block_size;
  p.packed = packed;
  p.num_q_heads = num_q_heads; p.num_kv_heads = num_kv_heads; p.group = num_q_heads / num_kv_heads;
  p.q_st = q_stride_t; p.q_sh = q_stride_h; p.q_sb = 0; p.o_st = o_stride_t; p.o_sh = o_stride_h; p.o_sb = 0;
  p.scale_log2 = softmax_scale * kLog2eF;
  p.interleave = gqa_interleave ? 1 : 0; p.causal = 1; p.dense = 0; p.round_scores = 1;
  p.win_local = win_local; p.win_global = win_global;

  KvDesc kv{key_cache, value_cache, kc_stride_b, kc_stride_h, kc_stride_t, vc_stride_b, vc_stride_h, vc_stride_t,
            block_size, num_blocks, num_kv_heads};
  dim3 grid((unsigned)((max_q_len + kBM - 1) / kBM), (unsigned)num_q_heads, (unsigned)batch);
  return launch_attn_mma(kv, p, head_dim, dtype, grid, (cudaStream_t)stream);
}

extern "C" int mojo_b200_paged_prefill_gqa(
    const void* query, const void* key_cache, const void* value_cache, const int32_t* cu_q_lens,
    const int32_t* cu_total_seq_lens, const int32_t* block_tables, void* out, int64_t total_q_tokens, int batch,
    int num_q_heads, int num_kv_heads, int head_dim, int64_t num_blocks, int block_size, int max_blocks_per_seq,
    int64_t table_stride, int64_t max_q_len, int64_t max_kv_len, int64_t q_stride_t, int64_t q_stride_h,
    int64_t o_stride_t, int64_t o_stride_h, int64_t kc_stride_b, int64_t kc_stride_h, int64_t kc_stride_t,
    int64_t vc_stride_b, int64_t vc_stride_h, int64_t vc_stride_t, float softmax_scale, int gqa_interleave,
    int is_causal, int dtype, void* stream) {
  return paged_prefill_impl(query, key_cache, value_cache, cu_q_lens, cu_total_seq_lens, block_tables, out,
                            total_q_tokens, batch, num_q_heads, num_kv_heads, head_dim, num_blocks, block_size,
                            max_blocks_per_seq, table_stride, max_q_len, max_kv_len, q_stride_t, q_stride_h, o_stride_t,
                            o_stride_h, kc_stride_b, kc_stride_h, kc_stride_t, vc_stride_b, vc_stride_h, vc_stride_t,
                            softmax_scale, gqa_interleave, is_causal, -1, -1, dtype, stream);
}

extern "C" int mojo_b200_paged_prefill_swa(
    const void* query, const void* key_cache, const void* value_cache, const int32_t* cu_q_lens,
    const int32_t* cu_total_seq_lens, const int32_t* block_tables, void* out, int64_t total_q_tokens, int batch,
    int num_q_heads, int num_kv_heads, int head_dim, int64_t num_blocks, int block_size, int max_blocks_per_seq,
    int64_t table_stride, int64_t max_q_len, int64_t max_kv_len, int64_t q_stride_t, int64_t q_stride_h,
    int64_t o_stride_t, int64_t o_stride_h, int64_t kc_stride_b, int64_t kc_stride_h, int64_t kc_stride_t,
    int64_t vc_stride_b, int64_t vc_stride_h, int64_t vc_stride_t, float softmax_scale, int gqa_interleave,
    int is_causal, int local_window_size, int global_window_size, int dtype, void* stream) {
  using namespace mojo;
  MOJO_REQUIRE(local_window_size >= -1 && global_window_size >= -1, MOJO_B200_EINVAL,
               "paged_prefill_swa: window sizes must be >= 0, or -1 for None");
  return paged_prefill_impl(query, key_cache, value_cache, cu_q_lens, cu_total_seq_lens, block_tables, out,
                            total_q_tokens, batch, num_q_heads, num_kv_heads, head_dim, num_blocks, block_size,
                            max_blocks_per_seq, table_stride, max_q_len, max_kv_len, q_stride_t, q_stride_h, o_stride_t,
                            o_stride_h, kc_stride_b, kc_stride_h, kc_stride_t, vc_stride_b, vc_stride_h, vc_stride_t,
                            softmax_scale, gqa_interleave, is_causal, local_window_size, global_window_size, dtype,
                            stream);
}

static int sdpa_impl(const void* query, const void* key, const void* value, void* out, int batch,
                     int num_q_heads, int num_kv_heads, int64_t q_len, int64_t kv_len, int head_dim,
                     int64_t q_stride_b, int64_t q_stride_h, int64_t q_stride_s, int64_t k_stride_b,
                     int64_t k_stride_h, int64_t k_stride_s, int64_t v_stride_b, int64_t v_stride_h,
                     int64_t v_stride_s, int64_t o_stride_b, int64_t o_stride_h, int64_t o_stride_s,
                     float softmax_scale, const uint8_t* mask, int64_t mask_stride_b, int64_t mask_stride_h,
                     int64_t mask_stride_q, int dtype, void* stream) {
  using namespace mojo;
  MOJO_REQUIRE(mask == nullptr || (mask_stride_b >= 0 && mask_stride_h >= 0 && mask_stride_q >= 0), MOJO_B200_EINVAL,
               "sdpa: negative mask stride");
  MOJO_REQUIRE(batch >= 0 && num_q_heads > 0 && num_kv_heads > 0 && q_len >= 0 && kv_len >= 0 && head_dim > 0,
               MOJO_B200_EINVAL, "sdpa: bad sizes");
  MOJO_REQUIRE(num_q_heads % num_kv_heads == 0, MOJO_B200_EINVAL, "sdpa: Hq %d not a multiple of Hkv %d", num_q_heads,
               num_kv_heads);
  if (batch == 0 || q_len == 0) return 0;
  MOJO_REQUIRE(kv_len > 0, MOJO_B200_EINVAL, "sdpa: kv_len must be > 0");
  MOJO_REQUIRE(query && key && value && out, MOJO_B200_EINVAL, "sdpa: null tensor pointer");
  MOJO_REQUIRE(dtype == MOJO_B200_BF16 || dtype == MOJO_B200_F16, MOJO_B200_EUNSUPPORTED,
               "sdpa: bf16/fp16 only (tensor-core path)");
  MOJO_REQUIRE(head_dim == 64 || head_dim == 128, MOJO_B200_EUNSUPPORTED, "sdpa: head_dim %d not in {64,128}", head_dim);
  MOJO_REQUIRE(batch <= 65535 && num_q_heads <= 65535 && q_len < (1LL << 31) && kv_len < (1LL << 31),
               MOJO_B200_EUNSUPPORTED, "sdpa: shape too large");

  {  // tcgen05/TMEM kernel when the shape is covered
    AttnSm100Args a;
    memset(&a, 0, sizeof(a));
    a.q = query; a.out = out; a.q_rows = q_len; a.q_sb = q_stride_b; a.q_st = q_stride_s; a.q_sh = q_stride_h;
    a.o_sb = o_stride_b; a.o_st = o_stride_s; a.o_sh = o_stride_h;
    a.k = key; a.v = value; a.k_b = k_stride_b; a.k_h = k_stride_h; a.k_t = k_stride_s;
    a.v_b = v_stride_b; a.v_h = v_stride_h; a.v_t = v_stride_s;
    a.rows_per_block = kv_len; a.num_blocks = batch;
    a.batch = batch; a.num_q_heads = num_q_heads; a.num_kv_heads = num_kv_heads; a.head_dim = head_dim;
    a.max_q_len = q_len; a.q_len_dense = q_len; a.kv_len_dense = kv_len; a.softmax_scale = softmax_scale;
    a.interleave = 0; a.causal = 0; a.dense = 1; a.round_scores = 0; a.dtype = dtype;
    a.win_local = a.win_global = -1;
    a.mask = mask; a.mask_sb = mask_stride_b; a.mask_sh = mask_stride_h; a.mask_sq = mask_stride_q;
    const int rc = launch_attn_sm100(a, (cudaStream_t)stream);
    if (rc != kAttnNotEligible) return rc;
  }

  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.q = query; p.out = out;
  p.mask = mask; p.mask_sb = mask_stride_b; p.mask_sh = mask_stride_h; p.mask_sq = mask_stride_q;
  p.box_rows = kTile;
  p.num_q_heads = num_q_heads; p.num_kv_heads = num_kv_heads; p.group = num_q_heads / num_kv_heads;
  p.q_len_dense = q_len; p.kv_len_dense = kv_len;
  p.q_st = q_stride_s; p.q_sh = q_stride_h; p.q_sb = q_stride_b;
  p.o_st = o_stride_s; p.o_sh = o_stride_h; p.o_sb = o_stride_b;
  p.scale_log2 = softmax_scale * kLog2eF;
  p.interleave = 0; p.causal = 0; p.dense = 1; p.round_scores = 0;
  p.win_local = p.win_global = -1;

  KvDesc kv{key, value, k_stride_b, k_stride_h, k_stride_s, v_stride_b, v_stride_h, v_stride_s, kv_len, batch,
            num_kv_heads};
  dim3 grid((unsigned)((q_len + kBM - 1) / kBM), (unsigned)num_q_heads, (unsigned)batch);
  return launch_attn_mma(kv, p, head_dim, dtype, grid, (cudaStream_t)stream);
}

extern "C" int mojo_b200_sdpa(const void* query, const void* key, const void* value, void* out, int batch,
                              int num_q_heads, int num_kv_heads, int64_t q_len, int64_t kv_len, int head_dim,
                              int64_t q_stride_b, int64_t q_stride_h, int64_t q_stride_s, int64_t k_stride_b,
                              int64_t k_stride_h, int64_t k_stride_s, int64_t v_stride_b, int64_t v_stride_h,
                              int64_t v_stride_s, int64_t o_stride_b, int64_t o_stride_h, int64_t o_stride_s,
                              float softmax_scale, int dtype, void* stream) {
  return sdpa_impl(query, key, value, out, batch, num_q_heads, num_kv_heads, q_len, kv_len, head_dim, q_stride_b,
                   q_stride_h, q_stride_s, k_stride_b, k_stride_h, k_stride_s, v_stride_b, v_stride_h, v_stride_s,
                   o_stride_b, o_stride_h, o_stride_s, softmax_scale, nullptr, 0, 0, 0, dtype, stream);
}

extern "C" int mojo_b200_sdpa_masked(const void* query, const void* key, const void* value, void* out, int batch,
                                     int num_q_heads, int num_kv_heads, int64_t q_len, int64_t kv_len, int head_dim,
                                     int64_t q_stride_b, int64_t q_stride_h, int64_t q_stride_s, int64_t k_stride_b,
                                     int64_t k_stride_h, int64_t k_stride_s, int64_t v_stride_b, int64_t v_stride_h,
                                     int64_t v_stride_s, int64_t o_stride_b, int64_t o_stride_h, int64_t o_stride_s,
                                     float softmax_scale, const void* mask, int64_t mask_stride_b,
                                     int64_t mask_stride_h, int64_t mask_stride_q, int dtype, void* stream) {
  MOJO_REQUIRE(mask != nullptr, MOJO_B200_EINVAL, "sdpa_masked: null mask");
  return sdpa_impl(query, key, value, out, batch, num_q_heads, num_kv_heads, q_len, kv_len, head_dim, q_stride_b,
                   q_stride_h, q_stride_s, k_stride_b, k_stride_h, k_stride_s, v_stride_b, v_stride_h, v_stride_s,
                   o_stride_b, o_stride_h, o_stride_s, softmax_scale, reinterpret_cast<const uint8_t*>(mask),
                   mask_stride_b, mask_stride_h, mask_stride_q, dtype, stream);
}

// MojoSWA (non-paged): the same kernels over PACKED key / value tensors [total_kv_tokens, Hkv, D] - sequence b owns rows
// cu_total_seq_lens[b] .. cu_total_seq_lens[b+1] - instead of a paged cache.
extern "C" int mojo_b200_swa(
    const void* query, const void* key, const void* value, const int32_t* cu_q_lens, const int32_t* cu_total_seq_lens,
    void* out, int64_t total_q_tokens, int64_t total_kv_tokens, int batch, int num_q_heads, int num_kv_heads, int head_dim,
    int64_t max_q_len, int64_t max_kv_len, int64_t q_stride_t, int64_t q_stride_h, int64_t o_stride_t, int64_t o_stride_h,
    int64_t k_stride_t, int64_t k_stride_h, int64_t v_stride_t, int64_t v_stride_h, float softmax_scale,
    int gqa_interleave, int is_causal, int local_window_size, int global_window_size, int dtype, void* stream) {
  using namespace mojo;
  MOJO_REQUIRE(local_window_size >= -1 && global_window_size >= -1, MOJO_B200_EINVAL,
               "swa: window sizes must be >= 0, or -1 for None");
  MOJO_REQUIRE(cu_total_seq_lens != nullptr && total_kv_tokens > 0 && total_kv_tokens < (1LL << 31), MOJO_B200_EINVAL,
               "swa: cu_total_seq_lens and a non-empty key tensor are required");
  return paged_prefill_impl(query, key, value, cu_q_lens, cu_total_seq_lens, nullptr, out, total_q_tokens, batch,
                            num_q_heads, num_kv_heads, head_dim, /*num_blocks=*/1, /*block_size=*/(int)total_kv_tokens,
                            /*max_blocks_per_seq=*/1, 0, max_q_len, max_kv_len, q_stride_t, q_stride_h, o_stride_t,
                            o_stride_h, total_kv_tokens * k_stride_t, k_stride_h, k_stride_t,
                            total_kv_tokens * v_stride_t, v_stride_h, v_stride_t, softmax_scale, gqa_interleave, is_causal,
                            local_window_size, global_window_size, dtype, stream, /*packed=*/1);
}
