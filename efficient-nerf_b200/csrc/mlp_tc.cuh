// Shared pieces of the persistent fused-MLP kernels (NeRF W256xD8 and R2L W256xD88).
//
// Execution model (one CTA per SM, all 512 TMEM columns, warp-specialised):
//   warps 0-3  "WG0": epilogue warps, own the even 64-column groups (0,2) of a layer's output
//   warps 4-7  "WG1": epilogue warps, own the odd groups (1,3)                  (TMEM lane quarter = warp%4)
//   [NeRF only] warps 8-11 "WG2": encoder warps (ray -> point -> sin/cos features of the NEXT tile)
//   producer warp: streams pre-packed 16-bit weight stages global(L2) -> shared with 1-D bulk copies
//              (TMA engine, UBLKCP) through a ring of 32 KiB stages (K = 64) + a small ring of 8 KiB bias stages
//   MMA warp:  one thread issues tcgen05.mma (M=128, N=256|128, K=16), both operands in shared memory,
//              fp32 accumulators in TMEM.  mbarrier operations cost ~100 cycles each on the issuing thread
//              (measured, scratch/ubench), so a stage carries FOUR MMAs (512 tensor-pipe cycles) per pair of
//              waits and the two waits of a stage are issued together.
// A 128-row tile of activations never leaves the SM.  There is ONE activation buffer A [128 x 256]
// (16-bit, k-chunk major): layer l's accumulators are read from TMEM (tcgen05.ld) 32 columns at a time,
// rounded (+ReLU) to 16 bit by one cvt.rn[.relu].f16x2 per pair and written IN PLACE over A (all of layer
// l's MMAs have completed by then).  Each 64-column group has its own mbarrier, so layer l+1's MMAs for
// K-stage s start as soon as group s is written: the tensor pipe "chases" the epilogue and is idle only
// for the latency of the first group.  Accumulators alternate between TMEM columns [0,256) and [256,512),
// so the next layer writes its accumulator while the previous one is still being read.
// Biases are folded into the MMA: every layer STARTS with one K=16 step whose A operand is a constant
// "ones" block (columns 0 and 1 = 1.0) and whose weights hold bias_hi / bias_lo (16-bit split, exact to
// 2^-22 relative in fp16) — it needs no activations, so it also covers part of the epilogue latency, and
// the epilogue has no bias add at all.
//
// Synchronisation (mbarrier only on the critical path):
//   w_full[s]/w_empty[s]  weight ring (producer <-> MMA; tcgen05.commit frees a slot)
//   a_ready[g]            "64-column group g of A is written" (4 warps of the owning WG -> MMA), one phase / layer
//   d_full[dbuf]          "accumulator complete" (MMA -> WGs, tcgen05.commit)
#pragma once
#include "tc_common.cuh"

namespace r2l {

constexpr int kTileM = 128;              // rows (samples / rays) per tile = UMMA M
constexpr int kWidth = 256;              // hidden width = UMMA N
constexpr int kStageK = 64;              // K elements per weight stage (4 x UMMA K=16)
constexpr int kStageBytes = kWidth * kStageK * 2;      // 32 KiB
constexpr int kBiasStageBytes = kWidth * 16 * 2;       // 8 KiB: the K=16 bias step of a 256-wide layer
constexpr int kChunkBytes = kTileM * 16;               // one 8-element K-chunk of an A operand (128 rows x 16 B)
constexpr int kSubBytes = 4 * kChunkBytes;             // 32 columns of A (one tcgen05.ld.x32 worth)
constexpr int kGroupBytes = 8 * kChunkBytes;           // one 64-column group (= one weight stage's worth of K)
constexpr int kABufBytes = kTileM * kWidth * 2;        // 64 KiB: [128 x 256] 16-bit, k-chunk major
constexpr int kPBlockBytes = kTileM * 64 * 2;          // 16 KiB: one encoded 3-D point block (63 features + pad)
constexpr int kVBlockBytes = kTileM * 32 * 2;          // 8 KiB: encoded view direction (27 features, 1, 1, pad)
constexpr int kOnesBytes = 2 * kChunkBytes;            // 4 KiB: constant A operand of the bias step
constexpr uint32_t kLboA = kChunkBytes;  // 2048
constexpr uint32_t kSbo = 128;

// pack two fp32 into 16-bit x2 (lo in the low half), optionally clamping negatives to zero in the same instruction
template <bool BF16, bool RELU>
__device__ __forceinline__ uint32_t cvt2(float lo, float hi) {
  uint32_t r;
  if (BF16) {
    if (RELU)
      asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  } else {
    if (RELU)
      asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else
      asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  }
  return r;
}

// Write 32 accumulator columns of this thread's row as 4 chunks of the A operand.
//   a_sub = A base + (first column / 32)*kSubBytes + row*16.  A warp stores 512 contiguous bytes per chunk.
template <bool BF16, bool RELU>
__device__ __forceinline__ void store_sub(const uint32_t (&v)[32], uint8_t* a_grp) {
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    uint4 q;
    q.x = cvt2<BF16, RELU>(__uint_as_float(v[8 * ch + 0]), __uint_as_float(v[8 * ch + 1]));
    q.y = cvt2<BF16, RELU>(__uint_as_float(v[8 * ch + 2]), __uint_as_float(v[8 * ch + 3]));
    q.z = cvt2<BF16, RELU>(__uint_as_float(v[8 * ch + 4]), __uint_as_float(v[8 * ch + 5]));
    q.w = cvt2<BF16, RELU>(__uint_as_float(v[8 * ch + 6]), __uint_as_float(v[8 * ch + 7]));
    *reinterpret_cast<uint4*>(a_grp + ch * kChunkBytes) = q;
  }
}

// "My warp's part of this operand group is written": make the generic-proxy stores visible to the async
// proxy (UMMA reads shared memory through it), order this warp's TMEM accesses before the signal, then ONE
// lane arrives (barrier count = number of warps, not threads).
// PAIR: the barrier lives in the LEADER CTA of the pair (rank 0) and collects the warps of both CTAs.
template <bool PAIR = false>
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncwarp();
  if (lane == 0) {
    if (PAIR && cluster_ctarank() != 0)
      mbar_arrive_cluster(mapa_u32(bar, 0));
    else
      mbar_arrive(bar);
  }
}
// plain (no fences) one-lane arrival, `times` arrivals
template <bool PAIR = false>
__device__ __forceinline__ void lane_arrive(uint64_t* bar, int times = 1) {
  if (PAIR && cluster_ctarank() != 0) {
    const uint32_t a = mapa_u32(bar, 0);
    for (int i = 0; i < times; ++i) mbar_arrive_cluster(a);
  } else {
    for (int i = 0; i < times; ++i) mbar_arrive(bar);
  }
}

// mbarrier wait (same instruction in both modes: cta-scope acquire, see mbar_arrive_cluster)
template <bool PAIR = false>
__device__ __forceinline__ void mbar_wait_x(uint64_t* bar, uint32_t parity, DebugBuf* dbg, uint32_t id) {
  mbar_wait(bar, parity, dbg, id);
}

// sin/cos of (x, 2x, 4x, ..., 2^(L-1) x) for L <= 10: accurate sincosf at octaves 0 and 5, exact double-angle
// recurrence in between (4 doublings amplify the ~1 ulp seed error to < 2e-6, far below the 16-bit operand
// rounding).  The reference computes sin(x * 2^f) with the product exact in fp32 (helpers:41-56).
// (out of line: the accurate sincosf carries a Payne-Hanek slow path of ~1.5 KB of SASS per inlined copy)
static __device__ __noinline__ float2 sincos_accurate(float x) {
  float s, c;
  sincosf(x, &s, &c);
  return make_float2(s, c);
}
template <int L>
__device__ __forceinline__ void sincos_octaves(float x, float (&s)[L], float (&c)[L]) {
#pragma unroll
  for (int f = 0; f < L; ++f) {
    if (f == 0 || f == 5) {
      const float2 sc = sincos_accurate(x * static_cast<float>(1 << f));
      s[f] = sc.x, c[f] = sc.y;
    } else {
      const float t = 2.0f * s[f - 1];
      s[f] = t * c[f - 1];
      c[f] = fmaf(-t, s[f - 1], 1.0f);
    }
  }
}

// Encode one 3-D point into the 64-wide "point block" of row `row`:
//   k = 0..2 -> x,y,z ; k = 3+6f+c -> sin(2^f p_c) ; k = 3+6f+3+c -> cos(2^f p_c) ; k = 63 -> 0
// (the NeRF Embedder order, utils/run_nerf_raybased_helpers.py:24-56; weights are permuted to
// this order at pack time for the R2L head, model/nerf_raybased.py:198-208).
// dst = base of the block's 8 chunks (chunk stride kChunkBytes), written as 8 x 16-byte stores.
// (split in two so that a caller can compute the encoding BEFORE it is allowed to overwrite the destination)
template <bool BF16>
__device__ __forceinline__ void encode_point_packed(float px, float py, float pz, uint4 (&q)[8]) {
  float v[64];
  v[0] = px;
  v[1] = py;
  v[2] = pz;
  {
    float s[10], c[10];
    sincos_octaves<10>(px, s, c);
#pragma unroll
    for (int f = 0; f < 10; ++f) {
      v[3 + 6 * f + 0] = s[f];
      v[3 + 6 * f + 3] = c[f];
    }
    sincos_octaves<10>(py, s, c);
#pragma unroll
    for (int f = 0; f < 10; ++f) {
      v[3 + 6 * f + 1] = s[f];
      v[3 + 6 * f + 4] = c[f];
    }
    sincos_octaves<10>(pz, s, c);
#pragma unroll
    for (int f = 0; f < 10; ++f) {
      v[3 + 6 * f + 2] = s[f];
      v[3 + 6 * f + 5] = c[f];
    }
  }
  v[63] = 0.0f;
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    q[ch].x = pack2<BF16>(v[8 * ch + 0], v[8 * ch + 1]);
    q[ch].y = pack2<BF16>(v[8 * ch + 2], v[8 * ch + 3]);
    q[ch].z = pack2<BF16>(v[8 * ch + 4], v[8 * ch + 5]);
    q[ch].w = pack2<BF16>(v[8 * ch + 6], v[8 * ch + 7]);
  }
}
// the 8 packed chunks of a point block -> row `row` of the block at dst
__device__ __forceinline__ void store_point_block(uint8_t* dst, int row, const uint4 (&q)[8]) {
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) *reinterpret_cast<uint4*>(dst + ch * kChunkBytes + row * 16) = q[ch];
}
template <bool BF16>
__device__ __forceinline__ void encode_point_block(uint8_t* dst, int row, float px, float py, float pz) {
  uint4 q[8];
  encode_point_packed<BF16>(px, py, pz, q);
  store_point_block(dst, row, q);
}

// Write 32 fp32 values as the 32-wide block (4 chunks) of row `row`.
template <bool BF16>
__device__ __forceinline__ void store_block32(uint8_t* dst, int row, const float (&v)[32]) {
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    uint4 q;
    q.x = pack2<BF16>(v[8 * ch + 0], v[8 * ch + 1]);
    q.y = pack2<BF16>(v[8 * ch + 2], v[8 * ch + 3]);
    q.z = pack2<BF16>(v[8 * ch + 4], v[8 * ch + 5]);
    q.w = pack2<BF16>(v[8 * ch + 6], v[8 * ch + 7]);
    *reinterpret_cast<uint4*>(dst + ch * kChunkBytes + row * 16) = q;
  }
}

// Constant A operand of the bias step: column 0 and 1 are 1.0, the other 14 columns 0.
template <bool BF16>
__device__ __forceinline__ void write_ones_block(uint8_t* ones, int tid, int nthreads) {
  const uint32_t one2 = pack2<BF16>(1.0f, 1.0f);
  for (int r = tid; r < kTileM; r += nthreads) {
    *reinterpret_cast<uint4*>(ones + r * 16) = make_uint4(one2, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(ones + kChunkBytes + r * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
}

// Wait for two barriers at once: both try_waits are in flight together, so the ~100-cycle latency of an
// mbarrier operation is paid once per stage instead of twice.
template <bool PAIR = false>
__device__ __forceinline__ void mbar_wait2(uint64_t* bar_a, uint32_t par_a, uint64_t* bar_b, uint32_t par_b,
                                           DebugBuf* dbg, uint32_t id) {
  // ONE thread per CTA pair waits here (the MMA issuer), and what it waits for sits on the layer-to-layer critical path:
  // it polls with test_wait (returns at once) instead of try_wait (may suspend the thread): +0.5 % on the R2L
  // frame, A/B on one box.  (The same change for the 16 epilogue warps' accumulator wait was 4 % SLOWER: that many
  // spinning warps take issue slots from the warps that do the work.)
  auto tw = [](uint64_t* bar, uint32_t par) { return mbar_test_wait(bar, par); };
  bool a = tw(bar_a, par_a);
  bool b = tw(bar_b, par_b);
  uint32_t spins = 0;
  while (!(a && b)) {
    if (!a) a = tw(bar_a, par_a);
    if (!b) b = tw(bar_b, par_b);
    if (++spins > (R2L_WATCHDOG_SPINS << 4)) mbar_watchdog_fire(dbg, id + (a ? 1000u : 0u), par_a | (par_b << 1));
  }
}

// Issue the K=16 MMAs of one weight stage (one thread): N_MMA = 4 for a full K=64 stage, 2 for a K=32 stage.
//   a_addr : shared address of the A operand's first chunk for this stage (2*N_MMA chunks are consumed)
//   b_addr : shared address of the weight stage; lbo_b = N*16
// The MMA-issuing thread is a single thread: every instruction between two tcgen05.mma costs >= 4 cycles of
// dependent issue, and rebuilding both 64-bit descriptors per MMA (~15 instructions) made the issue rate ~74 cycles
// per MMA (measured) — slower than an N = 128 MMA executes.  So descriptors are assembled from a constant high word
// and a low word that advances by a constant per K-step: (lbo >> 4) << 16 | (addr >> 4); addresses stay below 256 KiB,
// so the 14-bit address field never carries.
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((lbo_bytes >> 4) << 16) | ((smem_addr >> 4) & 0x3FFFu);
}
constexpr uint32_t kDescHi = ((kSbo >> 4) & 0x3FFFu) | (1u << 14);   // SBO >> 4 at bits [32,46), version 1 at bit 46
__device__ __forceinline__ uint64_t desc_pack(uint32_t lo) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(kDescHi));
  return d;
}

// PAIR: tcgen05.mma.cta_group::2 — b_addr / lbo_b describe THIS CTA's N-half of the stage (N/2 rows).
template <int N_MMA = 4, bool PAIR = false>
__device__ __forceinline__ void issue_stage(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, uint32_t lbo_b,
                                            uint32_t idesc, bool fresh) {
  const uint32_t a_lo = desc_lo(a_addr, kLboA);
  const uint32_t b_lo = desc_lo(b_addr, lbo_b);
  const uint32_t b_step = (2 * lbo_b) >> 4;
#pragma unroll
  for (int j = 0; j < N_MMA; ++j) {
    const uint64_t ad = desc_pack(a_lo + j * ((2 * kLboA) >> 4));
    const uint64_t bd = desc_pack(b_lo + j * b_step);
    if (PAIR)
      umma_f16_ss_pair(d_tmem, ad, bd, idesc, (fresh && j == 0) ? 0u : 1u);
    else
      umma_f16_ss(d_tmem, ad, bd, idesc, (fresh && j == 0) ? 0u : 1u);
  }
}
// The single K=16 MMA of a bias stage (A = the constant ones block).
template <bool PAIR = false>
__device__ __forceinline__ void issue_bias_stage(uint32_t d_tmem, uint32_t ones_addr, uint32_t b_addr, uint32_t lbo_b,
                                                 uint32_t idesc, bool fresh) {
  const uint64_t ad = desc_pack(desc_lo(ones_addr, kLboA));
  const uint64_t bd = desc_pack(desc_lo(b_addr, lbo_b));
  if (PAIR)
    umma_f16_ss_pair(d_tmem, ad, bd, idesc, fresh ? 0u : 1u);
  else
    umma_f16_ss(d_tmem, ad, bd, idesc, fresh ? 0u : 1u);
}
// tcgen05.commit to this CTA's barrier (PAIR: to the barrier at the same offset in both CTAs of the pair)
template <bool PAIR = false>
__device__ __forceinline__ void umma_commit_x(uint64_t* bar) {
  if (PAIR)
    umma_commit_pair(bar);
  else
    umma_commit(bar);
}

}  // namespace r2l
