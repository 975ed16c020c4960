// R2L network (NeRF_v3_2 with ResMLP body, W256 x D88), "ping-pong" schedule on a CTA pair with HALF-HEIGHT tiles.
// Same math as mlp_r2l.cu (model/nerf_raybased.py:443-465, 480-544); what changes is how the epilogue latency is hidden.
//
// mlp_r2l.cu keeps ONE 128-ray tile per CTA: the fp32 residual stream (256 TMEM columns) and the hidden accumulator (256)
// fill the 512 columns, so layer l+1's MMAs can only CHASE layer l's epilogue and the tensor pipe idles ~1070 of every
// 3250 cycles (commit -> warps wake -> tcgen05.ld -> cvt/st -> fences -> arrive -> issuer wakes).  A second tile per
// CTA, as in the NeRF kernel, did not fit.
// It does with tcgen05.mma.cta_group::2 at M = 128: each CTA of the pair then contributes 64 rows, the instruction still
// runs at full rate (64 cycles for N = 256, K = 16: measured, scratch/ubench/pair_m128.cu), and a 64-row x 256-column
// accumulator occupies 128 lanes x 128 columns (lanes 0-63 hold output columns 0-127, lanes 64-127 columns 128-255 of
// the same 64 rows: measured with the same probe).  So residual + hidden of one 64-row tile take 256 columns, and a
// CTA holds TWO tiles T0, T1 (2 x 64 rays; the pair works on 256 rays like before).  Each tile has its own issuer
// thread; while one tile's accumulator is converted by the epilogue warps, the other tile's MMAs keep the pipe busy.
// Both tiles consume the SAME weight stages (released when both issuers have committed them), so the L2 -> shared
// weight traffic per ray is unchanged.  Per tile the layer-to-layer chase (four 64-column operand groups, one
// mbarrier each) is kept, so a tile's next layer starts after its first operand group, not after its whole epilogue.
//
// MEASURED (B200, 4 frames per launch): correct (equal to mlp_r2l.cu's pair kernel to a few ulp of the fp32 tail sum), tensor pipe 72 % busy instead
// of 67 % (in-kernel counters: per layer and tile the issuer thread needs ~1575 cycles for 17 MMAs + 11 barrier
// operations — longer than the 1088 cycles the MMAs execute — and that issue time sits on each tile's chain), but 2 %
// SLOWER end to end (102.1 vs 104.5 Mrays/s): both kernels run into the 1 kW power cap, and an M = 128 MMA re-reads the
// 4 KiB weight operand for half as many rows (96 instead of 64 B/cycle of shared-memory operand reads), i.e. more
// energy per FLOP.  Kept as a selectable alternative (R2L_PP=1), not the default.
//
// TMEM columns: T0: x [0,128), h [128,256); T1: x [256,384), h [384,512).  Shared memory: A[t] 32 KiB each (64 rows x 256,
// 16-bit, k-chunk major with 64-row chunks of 1 KiB), the 8-slot weight ring and the bias ring of mlp_r2l.cu.
// Warps: 0-7 epilogue + head encoders (TMEM lane quarter = warp % 4: quarters 0,1 = rows 0-63 / output columns 0-127,
// quarters 2,3 = the same rows / columns 128-255; warps 0-3 take the TMEM column pieces [0,32) and [64,96), warps 4-7
// the other two), 8 weight producer, 9 / 10 the issuers of T0 / T1 (leader CTA).
#include <type_traits>

#include "common.cuh"
#include "mlp_params.cuh"
#include "mlp_tc.cuh"

namespace r2l {

constexpr int kRpThreads = 352;
constexpr int kRpProducerWarp = 8;
constexpr int kRpMmaWarp0 = 9;
constexpr int kRpRows = 64;                          // rows of a tile held by ONE CTA
constexpr int kRpChunkBytes = kRpRows * 16;          // 1 KiB: one 8-column K-chunk of a 64-row A operand
constexpr int kRpGroupBytes = 8 * kRpChunkBytes;     // 8 KiB: a 64-column group
constexpr int kRpABytes = 4 * kRpGroupBytes;         // 32 KiB
constexpr int kRpRing = 8, kRpBiasRing = 4;
constexpr uint32_t kRpStageB = kStageBytes / 2;      // this CTA's N-half of a K = 64 stage (16 KiB)
constexpr uint32_t kRpBiasB = kBiasStageBytes / 2;   // 4 KiB
constexpr uint32_t kRpLboB = 128 * 16;
constexpr uint32_t kRpLboA = kRpChunkBytes;
constexpr int kRpOffA = 0;
constexpr int kRpOffOnes = kRpOffA + 2 * kRpABytes;               // 2 chunks of 1 KiB
constexpr int kRpOffRing = kRpOffOnes + 2 * kRpChunkBytes;
constexpr int kRpOffBiasRing = kRpOffRing + kRpRing * static_cast<int>(kRpStageB);
constexpr int kRpOffWt = kRpOffBiasRing + kRpBiasRing * static_cast<int>(kRpBiasB);   // 3*256 floats
constexpr int kRpOffPart = kRpOffWt + 768 * 4;                    // [2 tiles][64 rows][4 slots] float4
constexpr int kRpOffBars = kRpOffPart + 2 * 64 * 4 * 16;
constexpr int kRpNumBars = 2 * kRpRing + 2 * kRpBiasRing + 8 + 2 + 8 + 2;
constexpr int kRpOffTmem = kRpOffBars + kRpNumBars * 8;
constexpr int kRpSmemBytes = kRpOffTmem + 16;
static_assert(kRpSmemBytes <= 227 * 1024, "R2L ping-pong kernel shared memory exceeds 227 KiB");
static_assert(kRpOffRing % 1024 == 0, "weight ring must stay 1 KiB aligned");

// Write 32 accumulator columns of this thread's row as 4 chunks of a 64-row A operand (chunk stride 1 KiB).
template <bool BF16, bool RELU>
__device__ __forceinline__ void store_sub_rows64(const uint32_t (&v)[32], uint8_t* dst) {
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    uint4 q;
    q.x = cvt2<BF16, RELU>(__uint_as_float(v[8 * ch + 0]), __uint_as_float(v[8 * ch + 1]));
    q.y = cvt2<BF16, RELU>(__uint_as_float(v[8 * ch + 2]), __uint_as_float(v[8 * ch + 3]));
    q.z = cvt2<BF16, RELU>(__uint_as_float(v[8 * ch + 4]), __uint_as_float(v[8 * ch + 5]));
    q.w = cvt2<BF16, RELU>(__uint_as_float(v[8 * ch + 6]), __uint_as_float(v[8 * ch + 7]));
    *reinterpret_cast<uint4*>(dst + ch * kRpChunkBytes) = q;
  }
}

// K = 16 MMAs of one weight stage on a 64-row-per-CTA A operand (cta_group::2, M = 128)
template <int N_MMA>
__device__ __forceinline__ void rp_issue_stage(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, bool fresh) {
  const uint32_t a_lo = desc_lo(a_addr, kRpLboA);
  const uint32_t b_lo = desc_lo(b_addr, kRpLboB);
#pragma unroll
  for (int j = 0; j < N_MMA; ++j) {
    const uint64_t ad = desc_pack(a_lo + j * ((2 * kRpLboA) >> 4));
    const uint64_t bd = desc_pack(b_lo + j * ((2 * kRpLboB) >> 4));
    umma_f16_ss_pair(d_tmem, ad, bd, idesc, (fresh && j == 0) ? 0u : 1u);
  }
}

// PROF: the instantiation behind r2l_resmlp_profile: per-CTA cycle counters prof[blockIdx.x][8]: [0] issuer T0 total,
// [1] T0 waiting for operand groups (a_ready), [2] T0 waiting for weight / bias stages, [3] issuer T1 total, [4] T1
// a_ready, [5] T1 weights, [6] epilogue warp 0 waiting for accumulators, [7] epilogue warp 0 total.
template <bool BF16, bool PROF>
__global__ void __launch_bounds__(kRpThreads, 1) r2l_mlp_pp_kernel(const R2lParams p, const __grid_constant__ R2lPairMaps maps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* const sA = smem + kRpOffA;
  uint8_t* const sOnes = smem + kRpOffOnes;
  uint8_t* const sRing = smem + kRpOffRing;
  uint8_t* const sBiasRing = smem + kRpOffBiasRing;
  float* const sWt = reinterpret_cast<float*>(smem + kRpOffWt);
  float4* const sPart = reinterpret_cast<float4*>(smem + kRpOffPart);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + kRpOffBars);
  uint64_t* const w_full = bars;
  uint64_t* const w_empty = w_full + kRpRing;        // both issuers have committed the stage (2 arrivals, both CTAs)
  uint64_t* const b_full = w_empty + kRpRing;
  uint64_t* const b_empty = b_full + kRpBiasRing;
  uint64_t* const a_ready = b_empty + kRpBiasRing;   // [tile][group]: leader; 4 warps of each CTA have written the group
  uint64_t* const d_full = a_ready + 8;              // [tile]: accumulator of the tile's current layer complete
  uint64_t* const a_free = d_full + 2;               // [tile][block]: head MMAs have consumed block j of A[tile]
  uint64_t* const drained = a_free + 8;              // [tile]: leader; all epilogue warps have read the tile's last x
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + kRpOffTmem);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_chunks = p.n_points / 4;
  const int nb = p.n_blocks;
  const uint32_t rank = cluster_ctarank();
  // a unit = 256 rays: tile t of CTA `rank` holds rays 256 u + 128 t + 64 rank + [0, 64)
  const int n_units = static_cast<int>((p.n_rays + 255) / 256);
  const int unit0 = static_cast<int>(blockIdx.x >> 1);
  const int unit_step = static_cast<int>(gridDim.x >> 1);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kRpRing; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 2);
    }
    for (int i = 0; i < kRpBiasRing; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 2);
    }
    for (int i = 0; i < 8; ++i) {
      mbar_init(&a_ready[i], 8);
      mbar_init(&a_free[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&d_full[i], 1);
      mbar_init(&drained[i], 16);
    }
    mbar_fence_init();
  }
  {   // constant A operand of the bias step: columns 0 and 1 = 1.0 (64 rows, 2 chunks)
    const uint32_t one2 = pack2<BF16>(1.0f, 1.0f);
    for (int r = threadIdx.x; r < kRpRows; r += kRpThreads) {
      *reinterpret_cast<uint4*>(sOnes + r * 16) = make_uint4(one2, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(sOnes + kRpChunkBytes + r * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  for (int i = threadIdx.x; i < 768; i += kRpThreads) sWt[i] = p.w_tail[i];
  fence_proxy_async_smem();
  if (warp == kRpMmaWarp0) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kRpProducerWarp) {
    // ===================== weight producer: this CTA's N-half of every stage, once per unit (both tiles) ==========
    if (lane == 0) {
      uint32_t g = 0, gb = 0;
      const uint8_t* src = nullptr;
      auto push = [&]() {
        const uint32_t slot = g % kRpRing;
        mbar_wait(&w_empty[slot], ((g / kRpRing) & 1) ^ 1, p.dbg, 100 + slot, 8);
        if (rank != 0) {
          tma2d_g2s_pair_bar(sRing + slot * kRpStageB, &maps.m16, 0, static_cast<int>((src + kRpStageB - p.wstream) >> 9),
                             mapa_u32(&w_full[slot], 0));
        } else {
          mbar_expect_tx(&w_full[slot], 2 * kRpStageB);
          bulk_g2s(sRing + slot * kRpStageB, src, kRpStageB, &w_full[slot]);
        }
        src += kStageBytes;
        ++g;
      };
      auto push_bias = [&]() {
        const uint32_t slot = gb % kRpBiasRing;
        mbar_wait(&b_empty[slot], ((gb / kRpBiasRing) & 1) ^ 1, p.dbg, 120 + slot, 8);
        if (rank != 0) {
          tma2d_g2s_pair_bar(sBiasRing + slot * kRpBiasB, &maps.m4, 0, static_cast<int>((src + kRpBiasB - p.wstream) >> 9),
                             mapa_u32(&b_full[slot], 0));
        } else {
          mbar_expect_tx(&b_full[slot], 2 * kRpBiasB);
          bulk_g2s(sBiasRing + slot * kRpBiasB, src, kRpBiasB, &b_full[slot]);
        }
        src += kBiasStageBytes;
        ++gb;
      };
      for (int unit = unit0; unit < n_units; unit += unit_step) {
        src = p.wstream;
        push_bias();
        for (int i = 0; i < n_chunks * 4; ++i) push();
        for (int l = 0; l < 2 * nb; ++l) {
          push_bias();
          for (int i = 0; i < 4; ++i) push();
        }
      }
    }
  } else if (warp == kRpMmaWarp0 || warp == kRpMmaWarp0 + 1) {
    // ===================== MMA issuer of tile t (leader CTA; one thread per tile, ONE copy of the code) =============
    if (rank == 0 && lane == 0) {
      const int t = warp - kRpMmaWarp0;
      const uint32_t idesc = make_idesc_f16(BF16, 2 * kRpRows, 256);
      const uint32_t aA = smem_u32(sA) + t * kRpABytes;
      const uint32_t aOnes = smem_u32(sOnes);
      const uint32_t aRing = smem_u32(sRing);
      const uint32_t aBiasRing = smem_u32(sBiasRing);
      const uint32_t dx = tmem_base + 256u * t, dh = dx + 128u;
      uint64_t* const ar = a_ready + 4 * t;
      uint64_t* const af = a_free + 4 * t;
      uint32_t g = 0, gb = 0, par_a = 0;
      long long t_a = 0, t_w = 0;
      const long long t_start = PROF ? clock64() : 0;
      auto bias_step = [&](uint32_t d_tmem, bool fresh) {
        const uint32_t slot = gb % kRpBiasRing;
        const long long c0 = PROF ? clock64() : 0;
        mbar_wait(&b_full[slot], (gb / kRpBiasRing) & 1, p.dbg, 240 + slot + 30 * t);
        if (PROF) t_w += clock64() - c0;
        tc_fence_after_sync();
        rp_issue_stage<1>(d_tmem, aOnes, aBiasRing + slot * kRpBiasB, idesc, fresh);
        umma_commit_pair(&b_empty[slot]);
        ++gb;
      };
      auto run4 = [&](uint32_t d_tmem, bool free_blocks) {
        for (int st = 0; st < 4; ++st) {
          const uint32_t slot = g % kRpRing;
          if (PROF) {   // separate waits so that the late one is known
            const long long c0 = clock64();
            mbar_wait(&w_full[slot], (g / kRpRing) & 1, p.dbg, 220 + st + 30 * t);
            const long long c1 = clock64();
            mbar_wait(&ar[st], par_a, p.dbg, 210 + st + 30 * t);
            t_w += c1 - c0;
            t_a += clock64() - c1;
          } else {
            mbar_wait2<true>(&ar[st], par_a, &w_full[slot], (g / kRpRing) & 1, p.dbg, 210 + st + 30 * t);
          }
          tc_fence_after_sync();
          rp_issue_stage<4>(d_tmem, aA + st * kRpGroupBytes, aRing + slot * kRpStageB, idesc, false);
          umma_commit_pair(&w_empty[slot]);
          if (free_blocks) umma_commit_pair(&af[st]);
          ++g;
        }
        par_a ^= 1u;
      };
      uint32_t it = 0;
      for (int unit = unit0; unit < n_units; unit += unit_step, ++it) {
        if (it > 0) {   // the previous unit's last epilogue has read x of this tile
          mbar_wait(&drained[t], (it - 1) & 1u, p.dbg, 250 + t);
          tc_fence_after_sync();
        }
        bias_step(dx, true);
        for (int c = 0; c < n_chunks; ++c) run4(dx, c + 1 < n_chunks);
        umma_commit_pair(&d_full[t]);
        for (int b = 0; b < nb; ++b) {
          bias_step(dh, true);
          run4(dh, false);
          umma_commit_pair(&d_full[t]);
          bias_step(dx, false);   // accumulate onto the fp32 residual stream
          run4(dx, false);
          umma_commit_pair(&d_full[t]);
        }
      }
      if (PROF && p.prof != nullptr) {
        long long* o = p.prof + blockIdx.x * 8 + 3 * t;
        o[0] = clock64() - t_start;
        o[1] = t_a;
        o[2] = t_w;
      }
    }
  } else if (warp < 8) {
    // ===================== epilogue / encoder warps =====================
    const int wg = warp >> 2;                         // TMEM column pieces [32 wg, +32) and [64 + 32 wg, +32)
    const int q = warp & 3;                           // TMEM lane quarter
    const int row = (q & 1) * 32 + lane;              // row of the tile held by this CTA
    const int cb = (q >> 1) * 128;                    // output columns cb + TMEM column
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int enc_row = threadIdx.x & 63, enc_j = threadIdx.x >> 6;   // head: one (row, block) per thread and chunk
    uint32_t par_d = 0;      // bit t: parity of the next d_full[t] phase
    uint32_t par_free = 0;   // bit (4 t + j)
    long long t_d = 0;
    const long long t_start = PROF ? clock64() : 0;
    auto wait_d = [&](int t, uint32_t id) {
      const long long c0 = PROF ? clock64() : 0;
      mbar_wait(&d_full[t], (par_d >> t) & 1u, p.dbg, id);
      if (PROF) t_d += clock64() - c0;
      par_d ^= 1u << t;
      tc_fence_after_sync();
    };
    // the two 32-column pieces of accumulator `dcol` (TMEM column base of x or h of a tile), software pipelined;
    // f(piece, out_col0, v); done(piece)
    auto for_pieces = [&](uint32_t dcol, auto&& f, auto&& done) {
      uint32_t va[32], vb[32];
      tmem_ld32(lane_taddr + dcol + 32 * wg, va);
      tmem_ld_wait();
      tmem_ld32(lane_taddr + dcol + 64 + 32 * wg, vb);
      f(0, cb + 32 * wg, va);
      done(0);
      tmem_ld_wait();
      f(1, cb + 64 + 32 * wg, vb);
      done(1);
    };
    for (int unit = unit0; unit < n_units; unit += unit_step) {
      // ---- head: chunks of 4 points (K = 256) per tile; thread = (row enc_row, block enc_j)
      for (int c = 0; c < n_chunks; ++c) {
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
          const long long ray = 256LL * unit + 128 * t + 64 * rank + enc_row;
          const long long ray_c = ray < p.n_rays ? ray : p.n_rays - 1;
          const int pt = c * 4 + enc_j;
          float px = 0.f, py = 0.f, pz = 0.f;
          if (p.cam != nullptr) {
            // fused ray generation: this row's ray from its pixel index and pose (point_sample_kernel's arithmetic)
            const long long gi = p.cam_ray0 + ray_c;
            const long long hw = static_cast<long long>(p.cam_H) * p.cam_W;
            const float* cm = p.cam + 12 * (gi / hw);
            const int pix = static_cast<int>(gi % hw);
            const int h = pix / p.cam_W, w = pix % p.cam_W;
            const float dx_ = __fdiv_rn(__fsub_rn(static_cast<float>(w), static_cast<float>(p.cam_W * 0.5)), p.cam_focal);
            const float dy_ = -__fdiv_rn(__fsub_rn(static_cast<float>(h), static_cast<float>(p.cam_H * 0.5)), p.cam_focal);
            const float rdx = __fadd_rn(__fadd_rn(__fadd_rn(0.0f, __fmul_rn(dx_, __ldg(cm + 0))), __fmul_rn(dy_, __ldg(cm + 1))),
                                        __fmul_rn(-1.0f, __ldg(cm + 2)));
            const float rdy = __fadd_rn(__fadd_rn(__fadd_rn(0.0f, __fmul_rn(dx_, __ldg(cm + 4))), __fmul_rn(dy_, __ldg(cm + 5))),
                                        __fmul_rn(-1.0f, __ldg(cm + 6)));
            const float rdz = __fadd_rn(__fadd_rn(__fadd_rn(0.0f, __fmul_rn(dx_, __ldg(cm + 8))), __fmul_rn(dy_, __ldg(cm + 9))),
                                        __fmul_rn(-1.0f, __ldg(cm + 10)));
            const float zz = __ldg(p.cam_z + pt);
            px = __fadd_rn(__ldg(cm + 3), __fmul_rn(rdx, zz));
            py = __fadd_rn(__ldg(cm + 7), __fmul_rn(rdy, zz));
            pz = __fadd_rn(__ldg(cm + 11), __fmul_rn(rdz, zz));
          } else if (p.embedded == nullptr) {
            const float* prow = p.pts + ray_c * p.pts_stride;
            px = __ldg(prow + 3 * pt);
            py = __ldg(prow + 3 * pt + 1);
            pz = __ldg(prow + 3 * pt + 2);
          }
          uint4 qv[8];
          if (p.embedded == nullptr) {
            encode_point_packed<BF16>(px, py, pz, qv);
          } else {
            // API path: gather the caller's embedding (reference order (3s+c)*21 + f') into block order
            const float* e = p.embedded + ray_c * p.emb_stride + static_cast<long long>(pt) * 63;
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int k = ch * 8 + i;
                int ref = -1;
                if (k < 3) {
                  ref = k * 21 + 20;
                } else if (k < 63) {
                  const int f = (k - 3) / 6, rem = (k - 3) % 6;
                  ref = (rem < 3) ? (rem * 21 + f) : ((rem - 3) * 21 + 10 + f);
                }
                v[i] = (ref >= 0) ? __ldg(e + ref) : 0.0f;
              }
              uint4 qq;
              qq.x = pack2<BF16>(v[0], v[1]);
              qq.y = pack2<BF16>(v[2], v[3]);
              qq.z = pack2<BF16>(v[4], v[5]);
              qq.w = pack2<BF16>(v[6], v[7]);
              // (stored after the a_free wait below: keep it in the array)
              qv[ch] = qq;
            }
          }
          if (c > 0) {
            const int bit = 4 * t + enc_j;
            mbar_wait(&a_free[bit], (par_free >> bit) & 1u, p.dbg, 400 + bit);
            par_free ^= 1u << bit;
          }
          uint8_t* blk = sA + t * kRpABytes + enc_j * kRpGroupBytes + enc_row * 16;
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) *reinterpret_cast<uint4*>(blk + ch * kRpChunkBytes) = qv[ch];
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) lane_arrive<true>(&a_ready[4 * t + enc_j], 2);   // a block comes from 2 warps per CTA
        }
      }
      // ---- head epilogue (x0 = relu(acc) kept in TMEM as the residual stream + A + tail partials), then the body
      float4* const part_row = sPart + row * 4 + (q >> 1) * 2 + wg;   // + t * 256
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        uint8_t* const a_row = sA + t * kRpABytes + row * 16;
        const uint32_t xcol = 256u * t;
        uint64_t* const ar = a_ready + 4 * t;
        float t0 = 0.f, t1 = 0.f, t2 = 0.f;
        wait_d(t, 300 + t);
        for_pieces(
            xcol,
            [&](int, int col0, uint32_t (&v)[32]) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float x = relu_nan(__uint_as_float(v[i]));
                v[i] = __float_as_uint(x);
                if (p.outer_skip) {
                  t0 = fmaf(sWt[col0 + i], x, t0);
                  t1 = fmaf(sWt[256 + col0 + i], x, t1);
                  t2 = fmaf(sWt[512 + col0 + i], x, t2);
                }
              }
              tmem_st32(lane_taddr + xcol + (col0 - cb), v);
              store_sub_rows64<BF16, false>(v, a_row + (col0 >> 3) * kRpChunkBytes);
            },
            [&](int piece) {
              tmem_st_wait();
              warp_arrive<true>(&ar[(q >> 1) * 2 + piece], lane);
            });
        part_row[t * 256] = make_float4(t0, t1, t2, 0.f);
      }
      for (int b = 0; b < nb; ++b) {
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {   // W1: h = relu(acc_h) -> A[t]
          uint8_t* const a_row = sA + t * kRpABytes + row * 16;
          uint64_t* const ar = a_ready + 4 * t;
          wait_d(t, 310 + t);
          for_pieces(
              256u * t + 128u,
              [&](int, int col0, uint32_t (&v)[32]) { store_sub_rows64<BF16, true>(v, a_row + (col0 >> 3) * kRpChunkBytes); },
              [&](int piece) { warp_arrive<true>(&ar[(q >> 1) * 2 + piece], lane); });
        }
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {   // W2: x = acc_x -> A[t]   (last block: tail partials instead)
          uint8_t* const a_row = sA + t * kRpABytes + row * 16;
          uint64_t* const ar = a_ready + 4 * t;
          wait_d(t, 320 + t);
          if (b + 1 < nb) {
            for_pieces(
                256u * t,
                [&](int, int col0, uint32_t (&v)[32]) { store_sub_rows64<BF16, false>(v, a_row + (col0 >> 3) * kRpChunkBytes); },
                [&](int piece) { warp_arrive<true>(&ar[(q >> 1) * 2 + piece], lane); });
          } else {
            float t0 = 0.f, t1 = 0.f, t2 = 0.f;
            for_pieces(
                256u * t,
                [&](int, int col0, uint32_t (&v)[32]) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) {
                    const float x = __uint_as_float(v[i]);
                    t0 = fmaf(sWt[col0 + i], x, t0);
                    t1 = fmaf(sWt[256 + col0 + i], x, t1);
                    t2 = fmaf(sWt[512 + col0 + i], x, t2);
                  }
                },
                [&](int) {});
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) lane_arrive<true>(&drained[t]);
            float4 acc = part_row[t * 256];
            acc.x += t0, acc.y += t1, acc.z += t2;
            part_row[t * 256] = acc;
          }
        }
      }
      // ---- tail: the four partial sums of a row (2 column halves x 2 piece sets), bias, sigmoid, store
      named_bar_sync(1, 256);
      if (warp < 4) {   // 128 threads = 2 tiles x 64 rows
        const int t = warp >> 1, r = (warp & 1) * 32 + lane;
        const float4* pr = sPart + t * 256 + r * 4;
        const float4 s0 = pr[0], s1 = pr[1], s2 = pr[2], s3 = pr[3];
        float o0 = ((s0.x + s1.x) + (s2.x + s3.x)) + p.b_tail[0];
        float o1 = ((s0.y + s1.y) + (s2.y + s3.y)) + p.b_tail[1];
        float o2 = ((s0.z + s1.z) + (s2.z + s3.z)) + p.b_tail[2];
        const long long ray = 256LL * unit + 128 * t + 64 * rank + r;
        const bool valid = ray < p.n_rays;
        if (valid) note_nonfinite(p.dbg, o0 + o1 + o2, ray);   // before the sigmoid hides an inf
        if (p.sigmoid_out) {
          o0 = 1.0f / (1.0f + expf(-o0));
          o1 = 1.0f / (1.0f + expf(-o1));
          o2 = 1.0f / (1.0f + expf(-o2));
        }
        if (valid) {
          if (p.n_peer == 0) {
            p.rgb[3 * ray + 0] = o0;
            p.rgb[3 * ray + 1] = o1;
            p.rgb[3 * ray + 2] = o2;
          } else {
            // fused gather: this rank's rows go straight into every GPU's frame buffer (see mlp_r2l.cu)
#pragma unroll 1
            for (int g = 0; g < p.n_peer; ++g) {
              float* o = p.rgb_peer[g] + 3 * (p.peer_row0 + ray);
              o[0] = o0;
              o[1] = o1;
              o[2] = o2;
            }
            __threadfence_system();
          }
        }
      }
      named_bar_sync(2, 256);   // sPart may be rewritten by the next unit's head epilogue
    }
    if (PROF && p.prof != nullptr && threadIdx.x == 0) {
      p.prof[blockIdx.x * 8 + 6] = t_d;
      p.prof[blockIdx.x * 8 + 7] = clock64() - t_start;
    }
  }

  // ---- teardown ----
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == kRpMmaWarp0) {
    tc_fence_after_sync();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

template <bool BF16, bool PROF>
static int launch_r2l_pp(const R2lParams& p, const R2lPairMaps& maps, int grid, cudaStream_t st) {
  R2L_CUDA(cudaFuncSetAttribute(r2l_mlp_pp_kernel<BF16, PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRpSmemBytes));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(kRpThreads);
  cfg.dynamicSmemBytes = kRpSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  R2L_CUDA(cudaLaunchKernelEx(&cfg, r2l_mlp_pp_kernel<BF16, PROF>, p, maps));
  count_launch();
  return R2L_OK;
}

// CTA-pair ping-pong R2L kernel: weights in the pair layout (the same stream as the pair kernel of mlp_r2l.cu), grid even
int r2l_mlp_pp_launch(bool bf16, const R2lParams& p, const R2lPairMaps& maps, int grid, cudaStream_t st) {
  if (p.prof != nullptr)
    return bf16 ? launch_r2l_pp<true, true>(p, maps, grid, st) : launch_r2l_pp<false, true>(p, maps, grid, st);
  return bf16 ? launch_r2l_pp<true, false>(p, maps, grid, st) : launch_r2l_pp<false, false>(p, maps, grid, st);
}

}  // namespace r2l
