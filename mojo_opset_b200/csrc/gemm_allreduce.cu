// MojoGemmAllReduce (row-parallel o_proj: out = all_reduce(x @ W^T + bias)) as ONE persistent kernel: a tcgen05
// GEMM whose epilogue pushes each partial tile over NVLink into the tile owner's memory, a reduce step the owner
// runs as soon as a tile's partials have landed, and a broadcast of the reduced tile into every rank's memory.
// Reference semantics: mojo_opset/core/operators/compute_with_comm.py:57-117 (F.linear then all_reduce(sum)).
//
// Roles of a CTA (192 threads, one CTA per SM, persistent over 128 x 128 output tiles; with >= 2 row blocks the CTAs
// run as pairs, see the kernel):
//   warp 0     TMA producer: A [128 x 64] and B [128 | 64 x 64] slabs (K-major, 128B swizzle) through a 5/7-stage ring
//   warp 1     MMA issuer: tcgen05.mma M128/256 N128 K16, fp32 accumulators in TMEM, two accumulator buffers so the
//              epilogue of tile i overlaps the main loop of tile i+1; TMEM allocation
//   warps 2-5  epilogue, one thread per tile row: TMEM -> (+bias) -> bf16 -> the row's 64 packed words, then
//                world == 1: swizzled tile image in shared memory -> out (coalesced);
//                world  > 1: the row goes out as 22 self-validating 16-byte LINES (3 payload words + the call's
//                            epoch, ONE 128-bit store each - st.relaxed.sys.b128 is single-copy atomic, so a line is
//                            seen whole or not at all) straight from registers into PEER memory: two-shot = the tile
//                            owner's partial slot [tile][src rank]; one-shot = every rank's slot.  No bulk copy to
//                            wait for, no fence, no separate flag: a reader polls the lines themselves.  (The first
//                            version moved 32 KB images with cp.async.bulk + wait + system fence + flag per hop:
//                            ~15 us of synchronisation around ~9 us of NVLink time at TP8.)  Lines cost 4/3 of the
//                            bytes; rows past m are neither written nor read.
//              after the CTA's last GEMM tile the same warps run (every CTA takes part):
//                two-shot: reduce units (owned tile, slab of lines): poll the `world` partial lines, sum them in fp32 in
//                  rank order (deterministic, bit-identical on every rank), store the result line into the result slot
//                  of EVERY rank; copy units (tile): poll the result lines of a row, image -> out (coalesced);
//                one-shot: poll the `world` partial lines of a row, sum in rank order, image -> out.
//
// No wait ever blocks a GEMM tile (all of a CTA's tiles are computed and pushed before its first poll), so the
// protocol cannot deadlock however the ranks' CTAs are scheduled; every poll is time-bounded (trap instead of a hung
// GPU).  Lines carry the call's epoch and the slots alternate with its parity, so nothing is ever reset: rank A can
// only start call n+2 after every rank finished the reduce step of call n+1, hence after every rank's kernel of call
// n - the last reader of the parity it is about to overwrite - has completed.
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "tcgen05.cuh"

namespace mojo {
namespace gar {

constexpr int kBM = 128, kBN = 128, kBK = 64;
constexpr int kStages = 5;
#ifndef MOJO_GAR_STAGES_PAIR
#define MOJO_GAR_STAGES_PAIR 7
#endif
constexpr int kStagesPair = MOJO_GAR_STAGES_PAIR;                // CTA pairs stage only half of B: 24 KB per stage
constexpr int kSlabBytes = kBM * kBK * 2;     // 16 KB: one operand slab of one stage
constexpr int kStageBytes = 2 * kSlabBytes;   // A + B
constexpr int kImageBytes = kBM * kBN * 2;    // 32 KB: one output tile, 16-bit elements
constexpr int kThreads = 192;
constexpr int kEpiThreads = 128;
constexpr int kMaxWorld = 8;
constexpr size_t kSmemBytes = 1024 + (size_t)kStagesPair * (kSlabBytes + kSlabBytes / 2) + kImageBytes + 256;
static_assert((size_t)kStagesPair * (kSlabBytes + kSlabBytes / 2) >= (size_t)kStages * kStageBytes, "ring sized by the pair mode");
// a waiting rank gives up (trap: launch failure rather than a hung GPU) only after MOJO_B200_GAR_TIMEOUT_S seconds (default
// 600, 0 = wait for ever like NCCL): ranks of an eager serving loop may skew by seconds (GC pause, first-call module load)
constexpr unsigned long long kDefaultWaitTimeoutNs = 600ull * 1000000000ull;
constexpr size_t kHeaderBytes = 256;
// a tile row (128 columns = 64 packed words) as 16-byte lines of 3 words + epoch; line (j, row) sits at index
// j * 128 + row of the tile's slot, so the 32 rows of a warp store / load 512 contiguous bytes
constexpr int kRowLines = 22;
constexpr int kTileLines = kRowLines * kBM;
constexpr int kTileLLBytes = kTileLines * 16;   // 45056
constexpr int kOneShotMaxTiles = 128;                       // one-shot mode: every rank holds every rank's tile
// ... so it is used only while (world-1) * m * n * 2 B is small.  Measured with the line protocol (8 GPUs, n 8192,
// k 1024): m 16: two-shot 20.1 us / one-shot 24.4; m 64: 19.2 / 35.1; m 256: 30.2 / 88.2; (2 GPUs, k 4096) m 256: 37.0 / 39.8
constexpr size_t kOneShotMaxPeerBytes = 512u << 10;

struct Params {
  void* out;
  const void* bias;
  int64_t m, n, k, out_rs;
  int tiles_m, tiles_n, n_tiles;
  int world, rank;
  int owned_cap;    // partial slots per parity = ceil(max tiles / world)
  int tiles_cap;    // result images per parity = max tiles
  uint8_t* ws[kMaxWorld];  // workspace base of every rank as mapped HERE (ws[rank] is local memory)
  size_t off_partial, off_result;
  int n_pair_tiles;  // PAIR: work items of a 2-CTA cluster = (256-row block, 128-column block)
  int one_shot, one_cap;  // one-shot mode: slots [parity][tile < one_cap][src]
  size_t off_one;
  unsigned long long timeout_ns;  // 0 = never give up
  int prefetch;       // B slabs pulled into L2 ahead of the ring (MOJO_B200_GAR_PREFETCH, k-blocks; 0 = off)
  long long* trace;   // developer timeline (MOJO_GAR_TRACE builds only, tools/gar_trace.py)
};

#ifdef MOJO_GAR_TRACE
// role 0 producer / 1 MMA or relay / 2 epilogue warp 2; CTAs 0 and 1; 96 slots each
#define GTRACE(role, j)                                                                                  \
  do {                                                                                                   \
    if (p.trace && blockIdx.x < 2 && (j) < 96 && (j) >= 0) p.trace[((role) + 3 * blockIdx.x) * 96 + (j)] = clock64(); \
  } while (0)
#else
#define GTRACE(role, j) do {} while (0)
#endif

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Layout {
  size_t off_partial, off_result, off_one, total;
  int owned_cap, tiles_cap, one_cap;
};

inline Layout make_layout(int64_t max_m, int64_t n, int world) {
  Layout l;
  const int64_t tiles = ((max_m + kBM - 1) / kBM) * ((n + kBN - 1) / kBN);
  l.tiles_cap = (int)tiles;
  l.owned_cap = (int)((tiles + world - 1) / world);
  l.off_partial = kHeaderBytes;  // [0] epoch counter, [4] finished-CTA counter
  l.off_result = l.off_partial + (size_t)2 * l.owned_cap * world * kTileLLBytes;
  l.one_cap = l.tiles_cap < kOneShotMaxTiles ? l.tiles_cap : kOneShotMaxTiles;
  l.off_one = l.off_result + (size_t)2 * l.tiles_cap * kTileLLBytes;
  l.total = align_up(l.off_one + (size_t)2 * l.one_cap * world * kTileLLBytes, 256);
  return l;
}

__device__ __forceinline__ size_t partial_off(const Params& p, uint32_t par, int local_tile, int src) {
  return p.off_partial + (((size_t)par * p.owned_cap + local_tile) * p.world + src) * kTileLLBytes;
}
__device__ __forceinline__ size_t result_off(const Params& p, uint32_t par, int tile) {
  return p.off_result + ((size_t)par * p.tiles_cap + tile) * kTileLLBytes;
}
__device__ __forceinline__ size_t one_off(const Params& p, uint32_t par, int tile, int src) {
  return p.off_one + (((size_t)par * p.one_cap + tile) * p.world + src) * kTileLLBytes;
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// L2 prefetch of a tensor-map box (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0),
               "r"(c1)
               : "memory");
}
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }
// one line = one 128-bit single-copy-atomic store / load at system scope (local or peer memory)
__device__ __forceinline__ void st_line(void* dst, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t epoch) {
  asm volatile("{\n\t.reg .b128 t;\n\tmov.b128 t, {%1, %2};\n\tst.relaxed.sys.global.b128 [%0], t;\n\t}" ::"l"(dst),
               "l"(((uint64_t)w1 << 32) | w0), "l"(((uint64_t)epoch << 32) | w2)
               : "memory");
}
__device__ __forceinline__ uint4 ld_line(const void* src) {
  uint64_t lo, hi;
  asm volatile("{\n\t.reg .b128 t;\n\tld.relaxed.sys.global.b128 t, [%2];\n\tmov.b128 {%0, %1}, t;\n\t}"
               : "=l"(lo), "=l"(hi)
               : "l"(src)
               : "memory");
  return make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));  // .w = epoch
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// bookkeeping of a polling loop: backs off a little, and a peer that never arrives becomes a launch failure (trap)
// after timeout_ns instead of a hung GPU
struct PollGuard {
  unsigned long long t0 = 0;
  uint32_t spins = 0;
  __device__ __forceinline__ void miss(unsigned long long timeout_ns) {
    __nanosleep(20);
    if ((++spins & 255u) == 0 && timeout_ns) {
      const unsigned long long now = global_ns();
      if (t0 == 0) t0 = now;
      if (now - t0 > timeout_ns) __trap();
    }
  }
};
// byte offset of 16-byte chunk `c` (8 elements) of row `r` inside a tile image: rows of 256 B, chunk index XOR-ed
// with the row so that the per-row epilogue stores are bank-conflict free
__device__ __forceinline__ uint32_t image_off(uint32_t r, uint32_t c) { return r * 256u + ((c ^ (r & 7u)) << 4); }

// the three payload words of a line (6 sixteen-bit values) added to six fp32 sums
template <typename T> __device__ __forceinline__ void acc_line(float (&a)[6], const uint4& v) {
  const uint32_t w[3] = {v.x, v.y, v.z};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (std::is_same<T, __nv_bfloat16>::value) {
      a[2 * i] += __uint_as_float(w[i] << 16);
      a[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
    } else {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      a[2 * i] += f.x;
      a[2 * i + 1] += f.y;
    }
  }
}

// Two-shot reduce unit: lines [slab * LPS, +LPS) of owned tile `t` (LPS = 2816 / WORLD).  A thread takes 16 / WORLD
// lines per pass - all `WORLD` sources of all of them requested before the first is looked at - polls until every line
// carries this call's epoch, sums in rank order and stores the result line into the result slot of every rank.
template <typename T, int WORLD>
__device__ __forceinline__ void ll_reduce_unit(const Params& p, uint32_t par, uint32_t epoch, int local_tile, int t,
                                               int slab, int tid) {
  constexpr int LPS = kTileLines / WORLD;
  constexpr int LPB = 16 / WORLD;
  const int tm = t / p.tiles_n;
  const int rows_valid = (int)(p.m - (int64_t)tm * kBM < kBM ? p.m - (int64_t)tm * kBM : kBM);
  const uint8_t* src0 = p.ws[p.rank] + partial_off(p, par, local_tile, 0);
  const size_t dst_off = result_off(p, par, t);
  const int l_end = (slab + 1) * LPS;
#pragma unroll 1
  for (int l0 = slab * LPS + tid; l0 < l_end; l0 += kEpiThreads * LPB) {
    uint4 v[LPB][WORLD];
    bool act[LPB];
#pragma unroll
    for (int i = 0; i < LPB; ++i) {
      const int line = l0 + i * kEpiThreads;
      act[i] = line < l_end && (line & (kBM - 1)) < rows_valid;  // (line = j * 128 + row)
    }
    PollGuard guard;
    for (;;) {
      bool ok = true;
#pragma unroll
      for (int i = 0; i < LPB; ++i) {
        if (act[i]) {
#pragma unroll
          for (int sr = 0; sr < WORLD; ++sr)
            v[i][sr] = ld_line(src0 + (size_t)sr * kTileLLBytes + ((size_t)(l0 + i * kEpiThreads) << 4));
        }
      }
#pragma unroll
      for (int i = 0; i < LPB; ++i) {
        if (act[i]) {
#pragma unroll
          for (int sr = 0; sr < WORLD; ++sr) ok = ok && v[i][sr].w == epoch;
        }
      }
      if (ok) break;
      guard.miss(p.timeout_ns);
    }
#pragma unroll
    for (int i = 0; i < LPB; ++i) {
      if (act[i]) {
        float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int sr = 0; sr < WORLD; ++sr) acc_line<T>(acc, v[i][sr]);
        const uint32_t w0 = pack2<T>(acc[0], acc[1]), w1 = pack2<T>(acc[2], acc[3]), w2 = pack2<T>(acc[4], acc[5]);
        const size_t off = dst_off + ((size_t)(l0 + i * kEpiThreads) << 4);
#pragma unroll
        for (int d = 0; d < WORLD; ++d) st_line(p.ws[d] + off, w0, w1, w2, epoch);
      }
    }
  }
}

// The 22 lines of tile row `row` from NSRC slots (slot s at src0 + s * kTileLLBytes), summed in slot order (NSRC = 1:
// taken as they are), into the swizzled tile image in shared memory.  One-shot: NSRC = world partial slots; two-shot
// copy: NSRC = 1, the result slot.
template <typename T, int NSRC>
__device__ __forceinline__ void ll_gather_row(const Params& p, const uint8_t* src0, uint32_t epoch, int row, uint8_t* image) {
  constexpr int LPB = NSRC >= 8 ? 2 : (NSRC == 4 ? 4 : (NSRC == 2 ? 8 : 11));
#pragma unroll 1
  for (int j0 = 0; j0 < kRowLines; j0 += LPB) {
    uint4 v[LPB][NSRC];
    PollGuard guard;
    for (;;) {
      bool ok = true;
#pragma unroll
      for (int i = 0; i < LPB; ++i) {
        if (j0 + i < kRowLines) {
#pragma unroll
          for (int sr = 0; sr < NSRC; ++sr)
            v[i][sr] = ld_line(src0 + (size_t)sr * kTileLLBytes + ((size_t)((j0 + i) * kBM + row) << 4));
        }
      }
#pragma unroll
      for (int i = 0; i < LPB; ++i) {
        if (j0 + i < kRowLines) {
#pragma unroll
          for (int sr = 0; sr < NSRC; ++sr) ok = ok && v[i][sr].w == epoch;
        }
      }
      if (ok) break;
      guard.miss(p.timeout_ns);
    }
#pragma unroll
    for (int i = 0; i < LPB; ++i) {
      const int j = j0 + i;
      if (j < kRowLines) {
        uint32_t w[3];
        if (NSRC == 1) {
          w[0] = v[i][0].x; w[1] = v[i][0].y; w[2] = v[i][0].z;
        } else {
          float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int sr = 0; sr < NSRC; ++sr) acc_line<T>(acc, v[i][sr]);
          w[0] = pack2<T>(acc[0], acc[1]); w[1] = pack2<T>(acc[2], acc[3]); w[2] = pack2<T>(acc[4], acc[5]);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int widx = 3 * j + k;  // packed word of the row: columns 2 widx, 2 widx + 1
          if (widx < kBN / 2)
            *reinterpret_cast<uint32_t*>(image + image_off((uint32_t)row, (uint32_t)(widx >> 2)) + (widx & 3) * 4) = w[k];
        }
      }
    }
  }
}

// PAIR = true: the same kernel on 2-CTA clusters with cta_group::2 MMAs (M = 256 = the two CTAs' 128-row blocks of x
// against ONE 128-column block of W, of which each CTA stages 64 rows): per k-block a CTA takes 16 KB of A + 8 KB of B
// out of L2 instead of 16 + 16 - at 128 x 128 single-CTA tiles the GEMM sat on the chip's L2 -> SM throughput cap
// (~6300 B/clk).  The leader CTA issues the MMAs; the peer relays "my slabs landed" to the leader's ring barriers,
// commits multicast to both CTAs; each CTA runs the epilogue of its own 128 rows (its own TMEM lanes).
template <typename T, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
gemm_allreduce_kernel(const __grid_constant__ CUtensorMap a_map, const __grid_constant__ CUtensorMap b_map,
                      const Params p) {
  constexpr int kStg = PAIR ? kStagesPair : kStages;
  constexpr int kBBytes = PAIR ? kSlabBytes / 2 : kSlabBytes;
  constexpr int kStgBytes = kSlabBytes + kBBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = smem;
  uint8_t* image = smem + kStg * kStgBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(image + kImageBytes);
  uint64_t* full = bars;                    // [kStg]
  uint64_t* empty = bars + kStg;            // [kStg]
  uint64_t* acc_full = empty + kStg;        // [2]
  uint64_t* acc_empty = acc_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  uint32_t* epoch_slot = tmem_slot + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;  // 0 = leader (issues the MMAs)
  if (threadIdx.x == 64) GTRACE(2, 0);
  // GEMM work items of this CTA (PAIR: of its cluster): w = w0, w0 + wstep, ... < n_work
  const int n_work = PAIR ? p.n_pair_tiles : p.n_tiles;
  const int w0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int wstep = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int k_blocks = (int)((p.k + kBK - 1) / kBK);
  // PAIR: this CTA stages rows [128 (2 tm + rank), +128) of x and rows [128 tn + 64 rank, +64) of W
  const int b_row_off = PAIR ? (int)crank * (kBN / 2) : 0;
  if (threadIdx.x == 0) {
    // descriptor fetches first: a TMA instruction whose descriptor is not cached yet holds its thread for the whole
    // fetch (timeline: ~2000 cycles for the first load issued right behind the prefetch); the TMEM allocation, the CTA
    // barrier and the cluster handshake below cover it
    tma_prefetch_desc(&b_map);
    tma_prefetch_desc(&a_map);
    for (int s = 0; s < kStg; ++s) {
      mbar_init(&full[s], (PAIR && crank == 0) ? 2 : 1);  // leader: own TMA + the peer's relay
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], PAIR ? 8 : 4);  // one arrive per epilogue warp (of both CTAs)
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_pair(tmem_slot, 256); else tmem_alloc(tmem_slot, 256);
  }
  // The call's epoch lives in DEVICE memory (the workspace header), so a CUDA graph can replay the launch: every
  // CTA reads the counter at its start, the last CTA to finish publishes the new value.  A CTA can only finish
  // after every CTA of the grid has started (the finished-CTA count reaches gridDim.x), so all of them read the
  // same value.  1 .. 0xFFFFFFFE, never 0 (the flags' initial value), parity alternates across the wrap.
  // (read after pdl_wait() by the first epilogue thread: the previous call's last CTA publishes the counter at its end)
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them remotely
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    // ------------------------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      // The weights do not depend on the previous kernel (PDL): the B slabs of the first ring pass are requested (and
      // the next p.prefetch slabs of that column pulled into L2; the loop keeps that distance) BEFORE pdl_wait(), i.e.
      // while the previous kernel - the decode that just streamed gigabytes through L2 - drains; x (A) follows after it.
      int pre = 0;
      if (w0 < n_work) {
        const int tn0 = w0 % p.tiles_n;
        pre = k_blocks < kStg ? k_blocks : kStg;
        for (int kb = 0; kb < pre; ++kb) {
          mbar_expect_tx(&full[kb], kStgBytes);
          tma_load_2d(ring + kb * kStgBytes + kSlabBytes, &b_map, &full[kb], kb * kBK, tn0 * kBN + b_row_off);
        }
        const int pf_end = k_blocks < pre + p.prefetch ? k_blocks : pre + p.prefetch;
        for (int kb = pre; kb < pf_end; ++kb) tma_prefetch_l2_2d(&b_map, kb * kBK, tn0 * kBN + b_row_off);
      }
      GTRACE(0, 0);
      pdl_wait();
      uint32_t c = 0;
      for (int w = w0; w < n_work; w += wstep) {
        const int tmw = w / p.tiles_n, tn = w - tmw * p.tiles_n;
        const int a_row = (PAIR ? 2 * tmw + (int)crank : tmw) * kBM;
        for (int kb = 0; kb < k_blocks; ++kb, ++c) {
          const uint32_t s = c % kStg;
          if ((int)c < pre) {  // B of this stage is already on its way
            tma_load_2d(ring + s * kStgBytes, &a_map, &full[s], kb * kBK, a_row);
            continue;
          }
          mbar_wait_bounded(&empty[s], ((c / kStg) & 1u) ^ 1u);
          GTRACE(0, (int)c + 1);
#ifdef MOJO_GAR_DBG_NOTMA  // developer experiment: no operand traffic after the first ring pass (results garbage)
          mbar_arrive(&full[s]);
          continue;
#endif
          mbar_expect_tx(&full[s], kStgBytes);
          tma_load_2d(ring + s * kStgBytes, &a_map, &full[s], kb * kBK, a_row);
          tma_load_2d(ring + s * kStgBytes + kSlabBytes, &b_map, &full[s], kb * kBK, tn * kBN + b_row_off);
          if (kb + p.prefetch < k_blocks) tma_prefetch_l2_2d(&b_map, (kb + p.prefetch) * kBK, tn * kBN + b_row_off);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 && PAIR && crank != 0) {
    // ------------------------------------------------------------------------------------ peer relay: "my slabs landed"
    if (lane == 0) {
      const uint32_t lead_full = mapa_u32(smem_u32(full), 0);
      uint32_t c = 0;
      for (int w = w0; w < n_work; w += wstep) {
        for (int kb = 0; kb < k_blocks; ++kb, ++c) {
          const uint32_t s = c % kStg;
          mbar_wait_bounded(&full[s], (c / kStg) & 1u);
          GTRACE(1, (int)c + 1);
          mbar_arrive_cluster(lead_full + s * 8u);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------ MMA issuer (whole warp)
    constexpr int kFmt = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
    constexpr uint32_t idesc = umma_idesc_f16(kFmt, PAIR ? 2 * kBM : kBM, kBN, 0, 0);
    const uint64_t desc_a0 = umma_desc_sw128(smem_u32(ring), 16, 1024);  // ring offsets stay inside the 14-bit address field
    uint32_t c = 0;
    int it = 0;
    for (int w = w0; w < n_work; w += wstep, ++it) {
      const int a = it & 1;
      mbar_wait_bounded(&acc_empty[a], (((uint32_t)it >> 1) & 1u) ^ 1u);
      tc_fence_after();
      for (int kb = 0; kb < k_blocks; ++kb, ++c) {
        const uint32_t s = c % kStg;
        mbar_wait_bounded(&full[s], (c / kStg) & 1u);  // PAIR: in both CTAs (the relay arrives on the same barrier)
        if (lane == 0) GTRACE(1, (int)c + 1);
        tc_fence_after();
        // the four K-steps of the stage in ONE asm block (one elect, descriptors derived by 64-bit adds): issued one by
        // one, the descriptor / elect / R2UR chains of the single issuing warp took ~97 cycles per MMA against 65 of
        // tensor work (timeline with the operand traffic switched off: the same 388 cycles per k-block)
        const uint64_t da = desc_a0 + (uint64_t)((s * (uint32_t)kStgBytes) >> 4);
        const uint64_t db = da + (uint64_t)(kSlabBytes >> 4);
        static_assert(kBK == 64, "x4 issue = four K-steps of 16 inside one 128-byte swizzle row");
        if (PAIR) umma_ss_x4_pair(tmem + a * kBN, da, db, idesc, kb != 0);
        else umma_ss_x4(tmem + a * kBN, da, db, idesc, kb != 0);
        if (PAIR) umma_commit_pair(&empty[s]); else umma_commit(&empty[s]);
      }
      if (PAIR) umma_commit_pair(&acc_full[a]); else umma_commit(&acc_full[a]);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------------------------ epilogue / reduce / copy
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;          // row of the tile
    const int tid = threadIdx.x - 64;       // 0..127
    pdl_wait();
    pdl_trigger();
    if (tid == 0)
      *epoch_slot = p.world > 1 ? *reinterpret_cast<volatile uint32_t*>(p.ws[p.rank]) % 0xFFFFFFFEu + 1u : 1u;
    epi_barrier();
    const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(epoch_slot);
    const uint32_t par = epoch & 1u;
    T* out = reinterpret_cast<T*>(p.out);
    const T* bias = reinterpret_cast<const T*>(p.bias);

    // un-swizzle a tile image (shared or global) into out; 16 threads cover one 256-byte row: coalesced
    auto image_to_out = [&](auto load_chunk, int tm, int tn) {
#pragma unroll 4
      for (int idx = tid; idx < kBM * 16; idx += kEpiThreads) {
        const int r = idx >> 4, ch = idx & 15;
        const int64_t gr = (int64_t)tm * kBM + r, gc = (int64_t)tn * kBN + ch * 8;
        if (gr < p.m && gc < p.n) {
          const uint4 v = load_chunk(image_off(r, ch));
          T* dst = out + gr * p.out_rs + gc;
          if (gc + 8 <= p.n && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
            *reinterpret_cast<uint4*>(dst) = v;
          } else {
            const T* e = reinterpret_cast<const T*>(&v);
            for (int i = 0; i < 8 && gc + i < p.n; ++i) dst[i] = e[i];
          }
        }
      }
    };

    const uint32_t lead_acc_empty = PAIR ? mapa_u32(smem_u32(acc_empty), 0) : 0u;
    int it = 0;
    for (int w = w0; w < n_work; w += wstep, ++it) {
      const int a = it & 1;
      const int tmw = w / p.tiles_n, tn = w - tmw * p.tiles_n;
      const int tm = PAIR ? 2 * tmw + (int)crank : tmw;
      const int t = tm * p.tiles_n + tn;            // the 128 x 128 output tile this CTA holds
      const bool tile_valid = tm < p.tiles_m;       // PAIR: the peer's row block may lie past m (an odd number of them)
      mbar_wait_bounded(&acc_full[a], ((uint32_t)it >> 1) & 1u);
      if (tid == 0) GTRACE(2, 1 + 2 * it);
      tc_fence_after();
      if (!tile_valid) {  // nothing to keep: hand the accumulator back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(lead_acc_empty + (uint32_t)a * 8u);
        continue;
      }
      uint32_t words[kBN / 2];  // this thread's row of the tile, two 16-bit values per word
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        uint32_t r[32];
        tmem_ld_x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * kBN + c4 * 32), r);
        tmem_wait_ld();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[g * 8 + i]);
          if (bias) {
            const int64_t gc = (int64_t)tn * kBN + c4 * 32 + g * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (gc + i < p.n) f[i] += DType<T>::to_f(bias[gc + i]);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) words[c4 * 16 + g * 4 + i] = pack2<T>(f[2 * i], f[2 * i + 1]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {  // the MMA warp may start the tile after next
        if (PAIR) mbar_arrive_cluster(lead_acc_empty + (uint32_t)a * 8u); else mbar_arrive(&acc_empty[a]);
      }
      if (p.world == 1) {
#pragma unroll
        for (int ch = 0; ch < 16; ++ch)
          *reinterpret_cast<uint4*>(image + image_off(row, ch)) =
              make_uint4(words[4 * ch], words[4 * ch + 1], words[4 * ch + 2], words[4 * ch + 3]);
        epi_barrier();
        image_to_out([&](uint32_t off) { return *reinterpret_cast<const uint4*>(image + off); }, tm, tn);
        epi_barrier();
        if (tid == 0) GTRACE(2, 2 + 2 * it);
      } else if ((int64_t)tm * kBM + row < p.m) {
        // the row leaves as 22 lines, each ONE 128-bit store that carries its own validity (the epoch): nothing to
        // wait for, nothing to fence, no flag to raise
        const size_t line0 = (size_t)row << 4;
        if (p.one_shot) {
          const size_t off = one_off(p, par, t, p.rank) + line0;
          for (int d = 0; d < p.world; ++d) {
            uint8_t* dst = p.ws[d] + off;
#pragma unroll
            for (int j = 0; j < kRowLines; ++j)
              st_line(dst + (size_t)j * (kBM * 16), words[3 * j], 3 * j + 1 < kBN / 2 ? words[3 * j + 1] : 0u,
                      3 * j + 2 < kBN / 2 ? words[3 * j + 2] : 0u, epoch);
          }
        } else {
          uint8_t* dst = p.ws[t % p.world] + partial_off(p, par, t / p.world, p.rank) + line0;
#pragma unroll
          for (int j = 0; j < kRowLines; ++j)
            st_line(dst + (size_t)j * (kBM * 16), words[3 * j], 3 * j + 1 < kBN / 2 ? words[3 * j + 1] : 0u,
                    3 * j + 2 < kBN / 2 ? words[3 * j + 2] : 0u, epoch);
        }
      }
    }

    if (p.world > 1 && p.one_shot) {
      // ---- one-shot: every rank holds every rank's partial lines of every tile; tile t is summed straight into `out`
      // by CTA t % grid (a thread per row: poll the `world` slots' lines, sum in rank order, image -> out)
      const uint8_t* self = p.ws[p.rank];
      for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
        const int tm = t / p.tiles_n, tn = t - tm * p.tiles_n;
        if ((int64_t)tm * kBM + row < p.m) {
          const uint8_t* src0 = self + one_off(p, par, t, 0);
          switch (p.world) {
            case 2: ll_gather_row<T, 2>(p, src0, epoch, row, image); break;
            case 4: ll_gather_row<T, 4>(p, src0, epoch, row, image); break;
            default: ll_gather_row<T, 8>(p, src0, epoch, row, image); break;
          }
        }
        epi_barrier();
        image_to_out([&](uint32_t off) { return *reinterpret_cast<const uint4*>(image + off); }, tm, tn);
        epi_barrier();  // the image is free again
      }
    } else if (p.world > 1) {
      // ---- two-shot, reduce units: (owned tile, slab of lines); unit u of this rank goes to CTA u % grid
      const int owned = (p.n_tiles - p.rank + p.world - 1) / p.world;  // tiles t with t % world == rank
      for (int u = blockIdx.x; u < owned * p.world; u += gridDim.x) {
        const int local_tile = u / p.world, slab = u - local_tile * p.world;
        const int t = local_tile * p.world + p.rank;
        switch (p.world) {
          case 2: ll_reduce_unit<T, 2>(p, par, epoch, local_tile, t, slab, tid); break;
          case 4: ll_reduce_unit<T, 4>(p, par, epoch, local_tile, t, slab, tid); break;
          default: ll_reduce_unit<T, 8>(p, par, epoch, local_tile, t, slab, tid); break;
        }
      }
      // ---- copy units: tile t by CTA t % grid (a thread per row: poll the result lines, image -> out)
      const uint8_t* self = p.ws[p.rank];
      for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
        const int tm = t / p.tiles_n, tn = t - tm * p.tiles_n;
        if ((int64_t)tm * kBM + row < p.m) ll_gather_row<T, 1>(p, self + result_off(p, par, t), epoch, row, image);
        epi_barrier();
        image_to_out([&](uint32_t off) { return *reinterpret_cast<const uint4*>(image + off); }, tm, tn);
        epi_barrier();  // the image is free again
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the leader's MMAs read the peer's shared memory and write its TMEM until here
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem, 256); else tmem_dealloc(tmem, 256);
  }
  if (p.world > 1 && threadIdx.x == 0) {
    uint32_t* hdr = reinterpret_cast<uint32_t*>(p.ws[p.rank]);
    __threadfence();
    if (atomicAdd(hdr + 1, 1u) == gridDim.x - 1) {
      hdr[1] = 0;
      __threadfence();
      *reinterpret_cast<volatile uint32_t*>(hdr) = *reinterpret_cast<volatile uint32_t*>(epoch_slot);
    }
  }
}

static int build_2d_map(const void* base, int dtype, int64_t rows, int64_t cols, int64_t row_stride, int box_rows,
                        CUtensorMap* out) {
  TensorMapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base;
  key.rank = 2;
  key.dtype = dtype;
  key.swizzle = (int)CU_TENSOR_MAP_SWIZZLE_128B;
  key.dims[0] = (uint64_t)cols;  key.box[0] = kBK;
  key.dims[1] = (uint64_t)rows;  key.strides[0] = (uint64_t)row_stride * 2;  key.box[1] = (uint32_t)box_rows;
  return get_tensor_map(key, out);
}

static bool env_flag(const char* name, int fallback) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) != 0 : fallback != 0;
}

// Can this device / context co-schedule 2-CTA clusters of this kernel (one CTA per SM by shared memory)?  Asked once
// per device; on a partitioned GPU the answer can be "none" and the launcher stays with single CTAs.
static bool pair_clusters_fit() {
  static int cached[64];  // 0 = not asked, 1 = yes, 2 = no
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  if (cached[dev] == 0) {
    auto kern = gemm_allreduce_kernel<__nv_bfloat16, true>;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2, 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = kSmemBytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    if (e != cudaSuccess) {
      cudaGetLastError();  // clear: the question failed, the answer is "no"
      n = 0;
    }
    cached[dev] = n > 0 ? 1 : 2;
  }
  return cached[dev] == 1;
}

}  // namespace gar
}  // namespace mojo

using namespace mojo;

extern "C" int mojo_b200_symm_alloc(size_t bytes, void** ptr) {
  MOJO_REQUIRE(ptr && bytes > 0, MOJO_B200_EINVAL, "symm_alloc: bad arguments");
  MOJO_CUDA_OK(cudaMalloc(ptr, bytes));
  MOJO_CUDA_OK(cudaMemset(*ptr, 0, bytes));
  MOJO_CUDA_OK(cudaDeviceSynchronize());
  return 0;
}

extern "C" int mojo_b200_symm_free(void* ptr) {
  if (ptr) MOJO_CUDA_OK(cudaFree(ptr));
  return 0;
}

extern "C" int mojo_b200_symm_export(void* ptr, void* handle64) {
  MOJO_REQUIRE(ptr && handle64, MOJO_B200_EINVAL, "symm_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  MOJO_CUDA_OK(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle64, &h, sizeof(h));
  return 0;
}

extern "C" int mojo_b200_symm_open(const void* handle64, void** peer_ptr) {
  MOJO_REQUIRE(handle64 && peer_ptr, MOJO_B200_EINVAL, "symm_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  MOJO_CUDA_OK(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

extern "C" int mojo_b200_symm_close(void* peer_ptr) {
  if (peer_ptr) MOJO_CUDA_OK(cudaIpcCloseMemHandle(peer_ptr));
  return 0;
}

extern "C" size_t mojo_b200_gemm_allreduce_workspace_bytes(int64_t max_m, int64_t n, int world) {
  if (max_m <= 0 || n <= 0 || world < 1 || world > gar::kMaxWorld || (world & (world - 1))) return 0;
  if (world == 1) return 0;
  return gar::make_layout(max_m, n, world).total;
}

extern "C" int mojo_b200_gemm_allreduce(const void* x, const void* weight, const void* bias, void* out, int64_t m,
                                        int64_t n, int64_t k, int64_t x_row_stride, int64_t w_row_stride,
                                        int64_t out_row_stride, void* const* peer_workspaces, size_t workspace_bytes,
                                        int64_t workspace_max_m, int world, int rank, int dtype, void* stream) {
  using namespace gar;
  MOJO_REQUIRE(m >= 0 && n > 0 && k > 0, MOJO_B200_EINVAL, "gemm_allreduce: bad sizes");
  MOJO_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, MOJO_B200_EINVAL,
               "gemm_allreduce: bad world/rank %d/%d", rank, world);
  MOJO_REQUIRE((world & (world - 1)) == 0, MOJO_B200_EUNSUPPORTED, "gemm_allreduce: world %d must be 1, 2, 4 or 8", world);
  MOJO_REQUIRE(dtype == MOJO_B200_BF16 || dtype == MOJO_B200_F16, MOJO_B200_EUNSUPPORTED,
               "gemm_allreduce: bf16/fp16 only (tensor-core path)");
  if (m == 0) return 0;  // nothing to compute, and every rank sees the same m: nothing to exchange either
  MOJO_REQUIRE(x && weight && out, MOJO_B200_EINVAL, "gemm_allreduce: null tensor pointer");
  MOJO_REQUIRE(x_row_stride % 8 == 0 && w_row_stride % 8 == 0 && aligned16(x) && aligned16(weight), MOJO_B200_EUNSUPPORTED,
               "gemm_allreduce: x / weight rows must be 16-byte aligned (TMA)");
  MOJO_REQUIRE(x_row_stride >= k && w_row_stride >= k && out_row_stride >= n, MOJO_B200_EINVAL,
               "gemm_allreduce: row strides smaller than the row");
  MOJO_REQUIRE(m < (1LL << 31) && n < (1LL << 31) && k < (1LL << 31), MOJO_B200_EUNSUPPORTED, "gemm_allreduce: too large");

  Params p;
  memset(&p, 0, sizeof(p));
  p.out = out; p.bias = bias; p.m = m; p.n = n; p.k = k; p.out_rs = out_row_stride;
  p.tiles_m = (int)((m + kBM - 1) / kBM);
  p.tiles_n = (int)((n + kBN - 1) / kBN);
  const int64_t tiles = (int64_t)p.tiles_m * p.tiles_n;
  MOJO_REQUIRE(tiles < (1LL << 30), MOJO_B200_EUNSUPPORTED, "gemm_allreduce: too many tiles");
  p.n_tiles = (int)tiles;
  p.world = world; p.rank = rank;
  p.timeout_ns = kDefaultWaitTimeoutNs;
  p.prefetch = 0;
  if (const char* pf = getenv("MOJO_B200_GAR_PREFETCH")) p.prefetch = atoi(pf);
#ifdef MOJO_GAR_TRACE
  if (const char* tp = getenv("MOJO_B200_GAR_TRACE_PTR")) p.trace = reinterpret_cast<long long*>(strtoull(tp, nullptr, 0));
#endif
  if (const char* t = getenv("MOJO_B200_GAR_TIMEOUT_S")) p.timeout_ns = (unsigned long long)(atof(t) * 1e9);
  if (world > 1) {
    MOJO_REQUIRE(peer_workspaces, MOJO_B200_EINVAL, "gemm_allreduce: peer workspace table is null");
    MOJO_REQUIRE(workspace_max_m >= m, MOJO_B200_EWORKSPACE, "gemm_allreduce: m %lld exceeds the workspace's max_m %lld",
                 (long long)m, (long long)workspace_max_m);
    const Layout l = make_layout(workspace_max_m, n, world);
    MOJO_REQUIRE(workspace_bytes >= l.total, MOJO_B200_EWORKSPACE, "gemm_allreduce: workspace %zu < %zu bytes",
                 workspace_bytes, l.total);
    p.owned_cap = l.owned_cap; p.tiles_cap = l.tiles_cap;
    p.off_partial = l.off_partial; p.off_result = l.off_result;
    p.one_cap = l.one_cap; p.off_one = l.off_one;
    // one-shot (push to everyone, reduce locally: one NVLink hop) while the extra traffic is cheap; two-shot
    // (push to the owner, reduce, broadcast: two hops, 1/world of the bytes per hop) beyond
    const char* mode = getenv("MOJO_B200_GAR_MODE");  // "one" / "two" force a mode (both ranks alike!)
    p.one_shot = p.n_tiles <= l.one_cap && (size_t)(world - 1) * m * n * 2 <= kOneShotMaxPeerBytes;
    if (mode && !strcmp(mode, "two")) p.one_shot = 0;
    if (mode && !strcmp(mode, "one") && p.n_tiles <= l.one_cap) p.one_shot = 1;
    for (int r = 0; r < world; ++r) {
      MOJO_REQUIRE(peer_workspaces[r], MOJO_B200_EINVAL, "gemm_allreduce: workspace of rank %d is null", r);
      p.ws[r] = reinterpret_cast<uint8_t*>(peer_workspaces[r]);
    }
  }

  // CTA pairs (cta_group::2, 256 x 128 work items) once there are at least two 128-row blocks to pair up - except for
  // the shallow, decode-sized case (fewer than four row blocks and k <= 1024), where the cluster set-up and the slower
  // peer-served operand reads cost more than the lower L2 traffic saves.  Measured, n 8192 (single CTAs / pairs): own
  // GEMM m 256 k 1024 8.2 / 8.7 us; m 256 k 4096 22.3 / 22.2; m 512 k 4096 41.0 / 38.9; m 8192 k 1024 143 / 134; inside
  // the TP8 cfg4 layer step (k 1024) fused kernel 36.8 / 38.5 us.  MOJO_B200_GAR_PAIR = 0 / 1 forces a mode.
  const int k_blocks_host = (int)((k + kBK - 1) / kBK);
  const bool pair_pays = p.tiles_m >= 4 || k_blocks_host > 16;
  const char* pair_env = getenv("MOJO_B200_GAR_PAIR");
  const bool pair = p.tiles_m >= 2 && ((pair_env && *pair_env) ? atoi(pair_env) != 0 : pair_pays) && pair_clusters_fit();
  p.n_pair_tiles = ((p.tiles_m + 1) / 2) * p.tiles_n;

  CUtensorMap a_map, b_map;
  int rc = build_2d_map(x, dtype, m, k, x_row_stride, kBM, &a_map);
  if (rc != 0) return rc;
  rc = build_2d_map(weight, dtype, n, k, w_row_stride, pair ? kBN / 2 : kBN, &b_map);
  if (rc != 0) return rc;

  int grid = p.n_tiles < kNumSMs ? p.n_tiles : kNumSMs;
  if (pair) grid = 2 * (p.n_pair_tiles < kNumSMs / 2 ? p.n_pair_tiles : kNumSMs / 2);
  // several virtual ranks sharing ONE GPU (comm.LocalRanks) must all be resident at once: cap each rank's CTAs
  if (const char* cap = getenv("MOJO_B200_GAR_MAX_CTAS")) {
    int c = atoi(cap);
    if (pair) c &= ~1;
    if (c > 0 && c < grid) grid = c;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 2; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pair ? 2 : 1;
#define GAR_LAUNCH(TT, PP)                                                                                     \
  do {                                                                                                         \
    auto kern = gemm_allreduce_kernel<TT, PP>;                                                                 \
    MOJO_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));    \
    MOJO_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, a_map, b_map, p));                                             \
  } while (0)
  if (dtype == MOJO_B200_BF16) {
    if (pair) GAR_LAUNCH(__nv_bfloat16, true); else GAR_LAUNCH(__nv_bfloat16, false);
  } else {
    if (pair) GAR_LAUNCH(__half, true); else GAR_LAUNCH(__half, false);
  }
#undef GAR_LAUNCH
  return check_launch("gemm_allreduce_kernel");
}
