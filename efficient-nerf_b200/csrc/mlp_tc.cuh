// Shared pieces of the persistent fused-MLP kernels (NeRF W256xD8 and R2L W256xD88).
//
// Execution model (one CTA per SM, 320 threads, all 512 TMEM columns):
//   warps 0-3  "WG0": epilogue / encoder for output columns [0,128)   (TMEM lane quarter = warp%4)
//   warps 4-7  "WG1": epilogue / encoder for output columns [128,256)
//   warp  8    weight producer: streams pre-packed 16-bit weight stages global(L2) -> shared
//              with 1-D bulk copies (TMA engine) through a ring of kRing stages
//   warp  9    MMA issuer: one thread issues tcgen05.mma (M=128, N=256|128, K=16) with both
//              operands in shared memory and fp32 accumulators in TMEM
// A 128-row tile of activations never leaves the SM: layer l's accumulators are read from
// TMEM (tcgen05.ld), bias/activation is applied in fp32 registers, the result is rounded to
// 16 bit and written to the *other* shared-memory A buffer, which layer l+1's MMAs read.
// Accumulators alternate between TMEM columns [0,256) and [256,512).
//
// Synchronisation is mbarrier-only on the critical path:
//   w_full[s]/w_empty[s]   weight ring           (producer <-> MMA, tcgen05.commit frees a slot)
//   a_ready[buf][half]     "K-half of A buffer written and my TMEM reads are done" (WG -> MMA)
//   d_full[dbuf]           "accumulator complete" (MMA -> WGs, tcgen05.commit)
// The MMA warp starts layer l+1 on K-half 0 as soon as WG0 has produced it, so WG1's
// epilogue overlaps the first half of the next layer's MMAs.
#pragma once
#include "tc_common.cuh"

namespace r2l {

constexpr int kTileM = 128;              // rows (samples / rays) per tile = UMMA M
constexpr int kWidth = 256;              // hidden width = UMMA N
constexpr int kStageK = 32;              // K elements per weight stage (2 x UMMA K=16)
constexpr int kStageBytes = kWidth * kStageK * 2;      // 16 KiB
constexpr int kChunkBytes = kTileM * 16;               // one 8-element K-chunk of an A operand (128 rows x 16 B)
constexpr int kABufBytes = kTileM * kWidth * 2;        // 64 KiB: [128 x 256] 16-bit, k-chunk major
constexpr int kPBlockBytes = kTileM * 64 * 2;          // 16 KiB: one encoded 3-D point block (63 features + pad)
constexpr int kThreads = 320;
constexpr int kProducerWarp = 8;
constexpr int kMmaWarp = 9;
constexpr uint32_t kLboA = kChunkBytes;  // 2048
constexpr uint32_t kSbo = 128;

// Encode one 3-D point into the 64-wide "point block" of row `row`:
//   k = 0..2 -> x,y,z ; k = 3+6f+c -> sin(2^f p_c) ; k = 3+6f+3+c -> cos(2^f p_c) ; k = 63 -> 0
// (the NeRF Embedder order, utils/run_nerf_raybased_helpers.py:24-56; weights are permuted to
// this order at pack time for the R2L head, model/nerf_raybased.py:198-208).
// dst = base of the block's 8 chunks (chunk stride kChunkBytes), written as 8 x 16-byte stores.
template <bool BF16>
__device__ __forceinline__ void encode_point_block(uint8_t* dst, int row, float px, float py, float pz) {
  float v[64];
  v[0] = px;
  v[1] = py;
  v[2] = pz;
#pragma unroll
  for (int f = 0; f < 10; ++f) {
    const float sc = static_cast<float>(1 << f);
    float s, c;
    sincosf(px * sc, &s, &c);
    v[3 + 6 * f + 0] = s;
    v[3 + 6 * f + 3] = c;
    sincosf(py * sc, &s, &c);
    v[3 + 6 * f + 1] = s;
    v[3 + 6 * f + 4] = c;
    sincosf(pz * sc, &s, &c);
    v[3 + 6 * f + 2] = s;
    v[3 + 6 * f + 5] = c;
  }
  v[63] = 0.0f;
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    uint4 q;
    q.x = pack2<BF16>(v[8 * ch + 0], v[8 * ch + 1]);
    q.y = pack2<BF16>(v[8 * ch + 2], v[8 * ch + 3]);
    q.z = pack2<BF16>(v[8 * ch + 4], v[8 * ch + 5]);
    q.w = pack2<BF16>(v[8 * ch + 6], v[8 * ch + 7]);
    *reinterpret_cast<uint4*>(dst + ch * kChunkBytes + row * 16) = q;
  }
}

// Epilogue for 64 accumulator columns [col0, col0+64) of this thread's row:
//   f(n, acc) -> value (bias, activation and any side computation happen in `f`);
//   WRITE_A : round the values to 16 bit and store them as 8 chunks of the next A operand
//   ST_TMEM : also store the fp32 values to TMEM columns st_taddr.. (R2L residual stream)
//   d_taddr : TMEM address of (this warp's lane quarter, accumulator column col0)
//   a_dst   : A buffer base + (first chunk index)*kChunkBytes + row*16
template <bool BF16, bool WRITE_A, bool ST_TMEM, class F>
__device__ __forceinline__ void epilogue_cols64(uint32_t d_taddr, uint8_t* a_dst, int col0, uint32_t st_taddr,
                                                F&& f) {
  uint32_t v0[32], v1[32];
  tmem_ld32(d_taddr, v0);
  tmem_ld32(d_taddr + 32, v1);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) v0[i] = __float_as_uint(f(col0 + i, __uint_as_float(v0[i])));
#pragma unroll
  for (int i = 0; i < 32; ++i) v1[i] = __float_as_uint(f(col0 + 32 + i, __uint_as_float(v1[i])));
  if (WRITE_A) {
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      uint4 q;
      q.x = pack2<BF16>(__uint_as_float(v0[8 * ch + 0]), __uint_as_float(v0[8 * ch + 1]));
      q.y = pack2<BF16>(__uint_as_float(v0[8 * ch + 2]), __uint_as_float(v0[8 * ch + 3]));
      q.z = pack2<BF16>(__uint_as_float(v0[8 * ch + 4]), __uint_as_float(v0[8 * ch + 5]));
      q.w = pack2<BF16>(__uint_as_float(v0[8 * ch + 6]), __uint_as_float(v0[8 * ch + 7]));
      *reinterpret_cast<uint4*>(a_dst + ch * kChunkBytes) = q;
    }
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      uint4 q;
      q.x = pack2<BF16>(__uint_as_float(v1[8 * ch + 0]), __uint_as_float(v1[8 * ch + 1]));
      q.y = pack2<BF16>(__uint_as_float(v1[8 * ch + 2]), __uint_as_float(v1[8 * ch + 3]));
      q.z = pack2<BF16>(__uint_as_float(v1[8 * ch + 4]), __uint_as_float(v1[8 * ch + 5]));
      q.w = pack2<BF16>(__uint_as_float(v1[8 * ch + 6]), __uint_as_float(v1[8 * ch + 7]));
      *reinterpret_cast<uint4*>(a_dst + (4 + ch) * kChunkBytes) = q;
    }
  }
  if (ST_TMEM) {
    tmem_st32(st_taddr, v0);
    tmem_st32(st_taddr + 32, v1);
  }
}

// Issue the two K=16 MMAs of one weight stage.
//   a_addr : shared address of the A operand's first chunk for this stage (4 chunks are consumed)
//   b_addr : shared address of the weight stage; lbo_b = N*16
__device__ __forceinline__ void issue_stage(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, uint32_t lbo_b,
                                            uint32_t idesc, bool first_of_layer) {
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const uint64_t ad = make_smem_desc(a_addr + j * 2 * kLboA, kLboA, kSbo);
    const uint64_t bd = make_smem_desc(b_addr + j * 2 * lbo_b, lbo_b, kSbo);
    umma_f16_ss(d_tmem, ad, bd, idesc, (first_of_layer && j == 0) ? 0u : 1u);
  }
}

}  // namespace r2l
