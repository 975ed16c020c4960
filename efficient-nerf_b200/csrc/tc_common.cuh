// Blackwell (sm_100a) primitives used by the fused MLP kernels: mbarrier, 1-D bulk
// copy (TMA engine, SASS UBLKCP), tcgen05 MMA / TMEM load-store / alloc, and the
// shared-memory matrix descriptors.  Everything is inline PTX; no library code.
//
// Operand layout convention used throughout this repo (both A and B operands):
//   K-major, SWIZZLE_NONE ("interleaved") canonical layout.  A [rows x K] 16-bit
//   operand is stored as 8x8-element "core matrices" of 128 contiguous bytes
//   (8 rows x 16 bytes).  Core matrices adjacent in the row (M/N) direction are
//   SBO bytes apart, core matrices adjacent in the K direction are LBO bytes apart.
//   We always use SBO = 128 (rows are dense) and LBO = rows*16, i.e. the operand
//   is "k-chunk major": byte offset of element (r, k) =
//        (k/8) * rows*16 + r*16 + (k%8)*2.
//   A warp writing one 16-byte chunk per row therefore stores 512 contiguous
//   bytes (bank-conflict free), and pre-packed weights can be moved with plain
//   1-D bulk copies — no tensor maps and no swizzle arithmetic anywhere.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace r2l {

// ---------------------------------------------------------------------------------
// debug / watchdog: a barrier wait that spins too long records who/where and traps,
// so a protocol bug shows up as a CUDA error instead of a hung GPU box.
// ---------------------------------------------------------------------------------
struct DebugBuf {
  unsigned int flag;       // 0 = ok
  unsigned int block;
  unsigned int thread;
  unsigned int barrier_id;
  unsigned int parity;
  unsigned int aux0;       // range flag: != 0 when a launch produced a non-finite output row (see note_nonfinite)
  unsigned int aux1;       // ... the low 32 bits of that row's index
  unsigned int aux2;
};

// fp16 operands hold |x| <= 65504.  The activation converts are NOT saturating on purpose: an activation (or a product
// sum) that leaves the range becomes +-inf, inf / NaN then reach every later layer (and stay in R2L's fp32 residual
// stream), so the kernel's OUTPUT row is non-finite — an overflow announces itself instead of silently rendering a
// clamped, wrong colour as a .satfinite convert would.  The output stage tests its few fp32 values per row (free) and
// records the event in the handle's mapped DebugBuf; the next call on the handle (and r2l_mlp_status) reports
// R2L_ERR_RANGE.  Weights are range-checked when they are packed.
// relu that PROPAGATES NaN (fmaxf(NaN, 0) is 0: it would turn an overflowed activation into a plausible-looking zero)
__device__ __forceinline__ float relu_nan(float v) {
  float r;
  asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(r) : "f"(v));
  return r;
}

__device__ __forceinline__ void note_nonfinite(DebugBuf* dbg, float probe, long long row) {
  if (!(fabsf(probe) <= 3.0e38f) && dbg != nullptr) {
    *reinterpret_cast<volatile unsigned int*>(&dbg->aux1) = static_cast<unsigned int>(row);
    *reinterpret_cast<volatile unsigned int*>(&dbg->aux0) = 1u;
  }
}

#ifndef R2L_WATCHDOG_SPINS
#define R2L_WATCHDOG_SPINS (1u << 24)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------- mbarrier ----------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Wait for the phase with the given parity to complete.  `id` only feeds the watchdog.
// `patience` multiplies the watchdog limit: producers wait with more patience than consumers, so that the record
// names the consumer that is really stuck rather than the producer starved behind it.
// The watchdog's slow path is ONE out-of-line function: inlined at every wait site it cost ~0.3 KB of SASS each, and
// the fused MLP kernels are sensitive to code size (a kernel that grows past ~128 KB slows down in every phase:
// measured A/B on one box, see mlp_nerf_pp.cu).
static __device__ __noinline__ void mbar_watchdog_fire(DebugBuf* dbg, uint32_t id, uint32_t parity) {
  if (dbg != nullptr && atomicCAS(&dbg->flag, 0u, 1u) == 0u) {
    dbg->block = blockIdx.x;
    dbg->thread = threadIdx.x;
    dbg->barrier_id = id;
    dbg->parity = parity;
    __threadfence_system();
  }
  __trap();
}
// non-blocking probe of a phase (mbarrier.test_wait returns at once; try_wait may suspend the thread)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, DebugBuf* dbg, uint32_t id,
                                          uint32_t patience = 1) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > R2L_WATCHDOG_SPINS * patience) mbar_watchdog_fire(dbg, id, parity);
  }
}

// ------------------------------- CTA pairs (cluster of 2) --------------------------
// (NOT volatile: the rank never changes, so the compiler may read the special register once and keep it — as a
// volatile asm it was re-read (S2UR, tens of cycles) in front of every operand-group arrive of the epilogue)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `bar` in the CTA with rank `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
// Arrive on an mbarrier of any CTA of the cluster (address from mapa_u32).  Default (.release.cta) semantics, as
// CUTLASS's ClusterBarrier::arrive does: a .release.cluster arrive costs a cluster-scope memory barrier (several
// hundred cycles, measured) on every signal, and what these signals order — this CTA's shared-memory operand
// writes (made visible to the async proxy by fence.proxy.async) and its TMEM reads (tcgen05.fence) — is consumed by
// this SM's own tensor core.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// ------------------------------- fences ------------------------------------------
// generic-proxy smem writes -> visible to the async proxy (UMMA / bulk copies)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ------------------------------- bulk copy (TMA engine, 1-D) ----------------------
// global -> shared, completion signalled on an mbarrier via complete_tx bytes.
// bytes must be a multiple of 16; both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// 2-D tiled tensor copy global -> this CTA's shared memory whose completion bytes are signalled on an mbarrier of
// EITHER CTA of the pair (.cta_group::2; address from mapa_u32).  tmap: CUtensorMap in kernel-parameter space.
__device__ __forceinline__ void tma2d_g2s_pair_bar(void* smem_dst, const void* tmap, int c0, int c1,
                                                   uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], "
      "[%4];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(bar_cluster_addr)
      : "memory");
}

// ------------------------------- TMEM --------------------------------------------
// Whole-warp collective.  Writes the TMEM base address to *smem_slot.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp receives lane
// (warp%4)*32+i, columns col..col+31.  taddr = (lane<<16) | col.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// CTA-pair variants: executed by one warp of EACH CTA of the pair
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------- UMMA --------------------------------------------
// Shared-memory matrix descriptor (sm_100 format, version 1), SWIZZLE_NONE, K-major.
//   bits [0,14)  start address >> 4
//   bits [16,30) leading-dimension byte offset >> 4  (distance between K-adjacent core matrices)
//   bits [32,46) stride-dimension byte offset >> 4   (distance between M/N-adjacent core matrices)
//   bits [46,48) version = 1;  bits [61,64) layout type = 0 (no swizzle)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

// Instruction descriptor for tcgen05.mma kind::f16, fp32 accumulate, both operands K-major.
//   bits [4,6) c_format = 1 (F32); [7,10) a_format; [10,13) b_format (0 = F16, 1 = BF16);
//   bit 15 a_major = 0 (K); bit 16 b_major = 0 (K); [17,23) N>>3; [24,29) M>>4.
__host__ __device__ constexpr uint32_t make_idesc_f16(bool bf16, int M, int N) {
  return (1u << 4) | ((bf16 ? 1u : 0u) << 7) | ((bf16 ? 1u : 0u) << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T.  Issued by ONE thread.
__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// CTA-pair MMA: M = 256 (128 rows per CTA), each CTA supplies its own A rows and HALF of B (N/2 rows) at the same
// shared-memory offsets; D lands in both CTAs' TMEM (128 lanes each).  Issued by ONE thread of the leader CTA.
__device__ __forceinline__ void umma_f16_ss_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit of the pair's MMAs: arrives on the mbarrier at this shared-memory offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}

// Make the mbarrier track completion of all prior tcgen05.mma issued by this thread.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ------------------------------- 16-bit packing ----------------------------------
// pack two fp32 -> one 32-bit word holding (lo, hi) 16-bit values, round-to-nearest-even.
template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t r;
  if (BF16) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  } else {
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  }
  return r;
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace r2l
