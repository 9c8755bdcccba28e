// tcgen05 (5th-generation tensor core) + TMEM PTX wrappers for sm_100a: TMEM allocation, UMMA shared-memory and
// instruction descriptors, tcgen05.mma (A from smem or from TMEM), tcgen05.commit -> mbarrier, tcgen05.ld/st.
#pragma once

#include "tma.cuh"

namespace mojo {

// ---- bounded mbarrier wait: a lost arrive becomes a trap (launch error) instead of a hung GPU ----------
// The wait SUSPENDS the warp (try_wait with a time hint: resumed as soon as the phase completes, else after the
// hint) instead of polling: a polling issuer / producer warp takes issue slots from the two softmax warps of its
// sub-partition - the attention timeline showed the softmax warps next to the QK issuer arriving up to 1000 cycles
// after their siblings, and a tile only moves on when its slowest warp has arrived.
constexpr uint32_t kWaitHintNs = 1000000u;
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  uint64_t t0 = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(kWaitHintNs)
        : "memory");
    if (ok) break;
    uint64_t now;  // only on the slow path: a wait that outlives 20 s of wall clock (time-slicing included) is a lost arrive
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (t0 == 0) t0 = now;
    if (now - t0 > 20000000000ull) __trap();  // surfaces as cudaErrorLaunchFailure on the host
  }
}

// ---- TMEM allocation (one warp, .sync.aligned) -----------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- descriptors ------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (sm_100 format, version 1), 128-byte swizzle:
//   bits [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4 |
//   [46,48) version = 1 | [49,52) base offset = 0 (tiles are 1024-byte aligned) | [61,64) layout = 2 (SWIZZLE_128B)
// K-major operand ([rows][64 elements = 128 B] lines, 8-row swizzle atoms of 1024 B):  LBO unused (1), SBO = 1024.
// MN-major operand ([k][64 elements = 128 B] lines, i.e. the MN dimension is contiguous): LBO = byte distance
//   between consecutive 64-element groups along MN, SBO = 1024 (distance between 8-line groups along K).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}

// Instruction descriptor for kind::f16 (fp16/bf16 inputs, fp32 accumulate):
//   [4,6) D format = 1 (f32) | [7,10) A format | [10,13) B format (0 = f16, 1 = bf16) | [15] A major | [16] B major
//   (0 = K, 1 = MN) | [17,23) N >> 3 | [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc_f16(int fmt_bf16, int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | ((uint32_t)fmt_bf16 << 7) | ((uint32_t)fmt_bf16 << 10) | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---- MMA issue ------------------------------------------------------------------------------------------------
// Called by ALL 32 lanes of the (converged) MMA warp with warp-uniform operands; elect.sync predicates the one
// issuing lane inside the asm block, so there is no divergent branch around the uniform-datapath instruction
// (a `if (lane == 0)` region makes ptxas wrap every UTCHMMA in a per-active-lane loop).  elect.sync on a full
// warp always picks the same lane, which tcgen05.commit relies on.
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// mbarrier arrive once every tcgen05 operation issued so far by the elected lane has completed
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(smem_u32(bar))
      : "memory");
}

// ---- CTA pairs (cta_group::2): one MMA spans the two SMs of a 2-CTA cluster --------------------------------------
// M = 256 = 128 rows of each CTA's A operand / TMEM accumulator; the B operand is split along N, each CTA stages
// N/2 of it at the same shared-memory offset.  Issued by the leader CTA (cluster rank 0) only; tcgen05.alloc /
// dealloc / relinquish are issued by the same warp of BOTH CTAs; commits multicast to the same barrier offset in
// both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// default semantics (release at CTA scope): what orders the payload here is the tcgen05 / async-proxy fence before
// it, and a .release.cluster arrive measured ~2000 cycles on the softmax -> MMA critical path
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
// wait on a LOCAL barrier whose arrivals may come from the peer CTA
__device__ __forceinline__ void mbar_wait_bounded_cluster(uint64_t* bar, uint32_t parity) { mbar_wait_bounded(bar, parity); }
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs once every tcgen05 operation issued so far has completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t.reg .b16 m;\n\t"
      "mov.b16 m, 3;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
      ::"r"(smem_u32(bar))
      : "memory");
}

// ---- grouped issue: the 8 K-steps of one 128-wide product in ONE asm block ------------------------------------
// One elect + one predicate set-up per group and the seven follow-up descriptors derived with 64-bit adds of
// constants (the address field of a descriptor is bits [0,14) = byte address >> 4, so a step inside the 128-byte
// swizzle atom is +2 and a jump to the second 64-column half is +half_bytes/16): the MMA warp shares its
// sub-partition's issue slots with two softmax warps, and the per-MMA elect / R2UR / VOTEU chains of the one-by-one
// form made the issue loop, not the tensor pipe, the bottleneck of the attention kernel.
//   QK: A and B K-major, K-steps 0..3 inside half 0 (+2 each), 4..7 inside half 1
#define MOJO_UMMA_SS_X8(CTA)                                                                                        \
  asm volatile(                                                                                                     \
      "{\n\t.reg .pred p, q, t;\n\t.reg .b64 a, b;\n\t"                                                             \
      "elect.sync _|q, 0xffffffff;\n\t"                                                                             \
      "setp.ne.b32 p, %4, 0;\n\t"                                                                                   \
      "setp.eq.b32 t, 0, 0;\n\t"                                                                                    \
      "@q tcgen05.mma.cta_group::" CTA ".kind::f16 [%0], %1, %2, %3, p;\n\t"                                        \
      "add.s64 a, %1, 2;\n\tadd.s64 b, %2, 2;\n\t"                                                                  \
      "@q tcgen05.mma.cta_group::" CTA ".kind::f16 [%0], a, b, %3, t;\n\t"                                          \
      "add.s64 a, %1, 4;\n\tadd.s64 b, %2, 4;\n\t"                                                                  \
      "@q tcgen05.mma.cta_group::" CTA ".kind::f16 [%0], a, b, %3, t;\n\t"                                          \
      "add.s64 a, %1, 6;\n\tadd.s64 b, %2, 6;\n\t"                                                                  \
      "@q tcgen05.mma.cta_group::" CTA ".kind::f16 [%0], a, b, %3, t;\n\t"                                          \
      "add.s64 a, %1, %5;\n\tadd.s64 b, %2, %6;\n\t"                                                                \
      "@q tcgen05.mma.cta_group::" CTA ".kind::f16 [%0], a, b, %3, t;\n\t"                                          \
      "add.s64 a, a, 2;\n\tadd.s64 b, b, 2;\n\t"                                                                    \
      "@q tcgen05.mma.cta_group::" CTA ".kind::f16 [%0], a, b, %3, t;\n\t"                                          \
      "add.s64 a, a, 2;\n\tadd.s64 b, b, 2;\n\t"                                                                    \
      "@q tcgen05.mma.cta_group::" CTA ".kind::f16 [%0], a, b, %3, t;\n\t"                                          \
      "add.s64 a, a, 2;\n\tadd.s64 b, b, 2;\n\t"                                                                    \
      "@q tcgen05.mma.cta_group::" CTA ".kind::f16 [%0], a, b, %3, t;\n\t}"                                         \
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc), "l"(a_half), "l"(b_half)                       \
      : "memory")
__device__ __forceinline__ void umma_ss_x8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint64_t a_half,
                                           uint64_t b_half, uint32_t idesc, uint32_t acc) {
  MOJO_UMMA_SS_X8("1");
}
__device__ __forceinline__ void umma_ss_x8_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint64_t a_half,
                                                uint64_t b_half, uint32_t idesc, uint32_t acc) {
  MOJO_UMMA_SS_X8("2");
}
#undef MOJO_UMMA_SS_X8
//   the four K-steps of ONE 64-column (128-byte) K-block: QK with head_dim 64, the GEMM main loop
#define MOJO_UMMA_SS_X4(CTA)                                                                                        \
  asm volatile(                                                                                                     \
      "{\n\t.reg .pred p, q, t;\n\t.reg .b64 a, b;\n\t"                                                             \
      "elect.sync _|q, 0xffffffff;\n\t"                                                                             \
      "setp.ne.b32 p, %4, 0;\n\t"                                                                                   \
      "setp.eq.b32 t, 0, 0;\n\t"                                                                                    \
      "@q tcgen05.mma.cta_group::" CTA ".kind::f16 [%0], %1, %2, %3, p;\n\t"                                        \
      "add.s64 a, %1, 2;\n\tadd.s64 b, %2, 2;\n\t"                                                                  \
      "@q tcgen05.mma.cta_group::" CTA ".kind::f16 [%0], a, b, %3, t;\n\t"                                          \
      "add.s64 a, %1, 4;\n\tadd.s64 b, %2, 4;\n\t"                                                                  \
      "@q tcgen05.mma.cta_group::" CTA ".kind::f16 [%0], a, b, %3, t;\n\t"                                          \
      "add.s64 a, %1, 6;\n\tadd.s64 b, %2, 6;\n\t"                                                                  \
      "@q tcgen05.mma.cta_group::" CTA ".kind::f16 [%0], a, b, %3, t;\n\t}"                                         \
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)                                                 \
      : "memory")
__device__ __forceinline__ void umma_ss_x4(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t acc) {
  MOJO_UMMA_SS_X4("1");
}
__device__ __forceinline__ void umma_ss_x4_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                uint32_t acc) {
  MOJO_UMMA_SS_X4("2");
}
#undef MOJO_UMMA_SS_X4
//   PV: A = 8 TMEM columns per K-step (+8), B MN-major (+b_step/16 per 16 keys)
#define MOJO_UMMA_TS_X8(CTA)                                                                                        \
  asm volatile(                                                                                                     \
      "{\n\t.reg .pred p, q, t;\n\t.reg .b64 b;\n\t.reg .b32 a;\n\t"                                                \
      "elect.sync _|q, 0xffffffff;\n\t"                                                                             \
      "setp.ne.b32 p, %4, 0;\n\t"                                                                                   \
      "setp.eq.b32 t, 0, 0;\n\t"                                                                                    \
      "@q tcgen05.mma.cta_group::" CTA ".kind::f16 [%0], [%1], %2, %3, p;\n\t"                                      \
      "add.u32 a, %1, 8;\n\tadd.s64 b, %2, %5;\n\t"                                                                 \
      "@q tcgen05.mma.cta_group::" CTA ".kind::f16 [%0], [a], b, %3, t;\n\t"                                        \
      "add.u32 a, a, 8;\n\tadd.s64 b, b, %5;\n\t"                                                                   \
      "@q tcgen05.mma.cta_group::" CTA ".kind::f16 [%0], [a], b, %3, t;\n\t"                                        \
      "add.u32 a, a, 8;\n\tadd.s64 b, b, %5;\n\t"                                                                   \
      "@q tcgen05.mma.cta_group::" CTA ".kind::f16 [%0], [a], b, %3, t;\n\t"                                        \
      "add.u32 a, a, 8;\n\tadd.s64 b, b, %5;\n\t"                                                                   \
      "@q tcgen05.mma.cta_group::" CTA ".kind::f16 [%0], [a], b, %3, t;\n\t"                                        \
      "add.u32 a, a, 8;\n\tadd.s64 b, b, %5;\n\t"                                                                   \
      "@q tcgen05.mma.cta_group::" CTA ".kind::f16 [%0], [a], b, %3, t;\n\t"                                        \
      "add.u32 a, a, 8;\n\tadd.s64 b, b, %5;\n\t"                                                                   \
      "@q tcgen05.mma.cta_group::" CTA ".kind::f16 [%0], [a], b, %3, t;\n\t"                                        \
      "add.u32 a, a, 8;\n\tadd.s64 b, b, %5;\n\t"                                                                   \
      "@q tcgen05.mma.cta_group::" CTA ".kind::f16 [%0], [a], b, %3, t;\n\t}"                                       \
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc), "l"(b_step)                                    \
      : "memory")
__device__ __forceinline__ void umma_ts_x8(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint64_t b_step,
                                           uint32_t idesc, uint32_t acc) {
  MOJO_UMMA_TS_X8("1");
}
__device__ __forceinline__ void umma_ts_x8_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint64_t b_step,
                                                uint32_t idesc, uint32_t acc) {
  MOJO_UMMA_TS_X8("2");
}
#undef MOJO_UMMA_TS_X8

// ---- TMEM <-> registers (warp w of a warpgroup owns lanes 32*(w%4) .. +31; thread = lane = matrix row) -------
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,"
      "%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
      "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
      "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- register redistribution between warpgroups -------------------------------------------------------------
template <int N> __device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N> __device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- packed fp32x2 arithmetic (sm_100: one FMA-pipe instruction for two lanes) -----------------------------------
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(reinterpret_cast<uint64_t&>(d))
      : "l"(reinterpret_cast<const uint64_t&>(a)), "l"(reinterpret_cast<const uint64_t&>(b)),
        "l"(reinterpret_cast<const uint64_t&>(c)));
  return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<uint64_t&>(d))
      : "l"(reinterpret_cast<const uint64_t&>(a)), "l"(reinterpret_cast<const uint64_t&>(b)));
  return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<uint64_t&>(d))
      : "l"(reinterpret_cast<const uint64_t&>(a)), "l"(reinterpret_cast<const uint64_t&>(b)));
  return d;
}

__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  float2 d;
  asm("sub.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<uint64_t&>(d))
      : "l"(reinterpret_cast<const uint64_t&>(a)), "l"(reinterpret_cast<const uint64_t&>(b)));
  return d;
}
__device__ __forceinline__ float fma_sat(float a, float b, float c) {
  float d;
  asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// exp2 of a pair on the FMA pipe (the MUFU runs 16 exponentials per clock and SM - as many cycles per attention tile
// as the tensor cores need for its two products).  The caller passes y = sat((x + 125) / 256), computed with ONE
// FFMA.SAT straight from the score: x is clamped to [-125, 131] for free and masked (-inf) scores become 2^-125.
//   t = 256 y + K          K = 1.5 * 2^23 - 125: t = magic + n, n = round(x) sits in the low mantissa bits
//   f = 256 y - (n + 125)  in [-0.5, 0.5]                      (exact: power-of-two scaling, integer n)
//   2^x = p(f) * 2^n       p = degree-3 minimax polynomial (relative error 8e-5: far below the bf16 / fp16 rounding of
//                          P), the exponent by adding n << 23 to the bit pattern (one LEA)
// Cost per pair: 2 FFMA.SAT + 5 FFMA2 / FADD2 + 2 LEA against 1 FFMA2 + 2 MUFU.EX2 (16 MUFU-pipe cycles).
__device__ __forceinline__ float2 ex2_fma_pipe2(float2 y) {
  constexpr float kK = 12582912.f - 125.f;
  const float2 t = fma2(y, make_float2(256.f, 256.f), make_float2(kK, kK));
  const float2 mneg = sub2(make_float2(kK, kK), t);
  const float2 f = fma2(y, make_float2(256.f, 256.f), mneg);
  float2 q = fma2(f, make_float2(0.0551716648f, 0.0551716648f), make_float2(0.2426111251f, 0.2426111251f));
  q = fma2(q, f, make_float2(0.6932609677f, 0.6932609677f));
  q = fma2(q, f, make_float2(0.9999280572f, 0.9999280572f));
  float2 out;
  out.x = __uint_as_float(__float_as_uint(q.x) + (__float_as_uint(t.x) << 23));
  out.y = __uint_as_float(__float_as_uint(q.y) + (__float_as_uint(t.y) << 23));
  return out;
}

// two fp32 -> packed 16-bit pair, `lo` in bits [0,16)
template <typename T> __device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <> __device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

}  // namespace mojo
