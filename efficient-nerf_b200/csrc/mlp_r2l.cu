// Persistent fused R2L network (NeRF_v3_2 with ResMLP body, W256 x D88) on tcgen05/TMEM.
// Reference math: NeRF_v3_2.forward model/nerf_raybased.py:539-544, ResMLP.forward :461-465,
// PositionalEmbedder.__call__ :198-208 on PointSampler points :94-126.
//
// One tile = 128 RAYS.  Per tile (see mlp_tc.cuh for the execution model):
//   head : x0 = relu(W_h enc(pts) + b_h); K = 64 per 3-D point (63 features + pad, weights permuted at
//          pack time), encoded by the epilogue warps straight into the shared-memory A buffer in chunks
//          of 4 points (K = 256); block j of the next chunk is written as soon as the MMAs have consumed
//          block j of the current one.  The head accumulates in TMEM columns [256,512) ("D2").
//   body : 43 x [ h = relu(W1 x + b1)  (accumulator D1 = columns [0,256)) ;
//                 x = x + res_scale*(W2 h + b2)  (ACCUMULATED onto D2) ]
//          The residual stream lives in D2 in fp32 for the whole body: x0 is stored back there once
//          (tcgen05.st) and every W2 layer accumulates onto it (tcgen05.mma accumulate flag), so the
//          residual add is free and exact fp32; b1/b2 ride along as the folded bias step.
//   tail : rgb = sigmoid(W_t (x_43 + x0) + b_t)   (outer skip, use_residual) on CUDA cores in fp32:
//          W_t x0 is accumulated at the head epilogue, W_t x_43 at the last epilogue.
// Inputs of the head, one of three (R2lParams): the sampled points `pts` (PointSampler output), the caller's embedding
// `embedded` (API path), or NOTHING but camera and pose (`cam`: r2l_resmlp_render) — then each row generates its ray
// from pixel index and c2w with point_sample_kernel's arithmetic (rays.cu) and the frame is one kernel from pose to rgb.
// Output: `rgb`, or — fused tile gather of a ray-sharded frame — the same rows of every GPU's frame buffer
// (`rgb_peer`, peer-to-peer stores; the caller's cross-GPU barrier publishes them).
// Weight stream per tile: per layer one 8 KiB bias stage + K/64 stages of 32 KiB (P = 16: 12.4 MB, L2 resident).
#include "common.cuh"
#include "mlp_params.cuh"
#include "mlp_tc.cuh"

namespace r2l {

constexpr int kR2lThreads = 320;
constexpr int kR2lProducerWarp = 8;
constexpr int kR2lMmaWarp = 9;
// The weight ring holds 128 KiB: 4 stages of 32 KiB, or (CTA pair: each CTA keeps its N-half) 8 half stages of 16 KiB.
constexpr int kR2lRingBytes = 4 * kStageBytes;
constexpr int kR2lBiasRingBytes = 2 * kBiasStageBytes;   // 2 bias stages of 8 KiB, or 4 halves of 4 KiB
constexpr int kR2lMaxRing = 8, kR2lMaxBiasRing = 4;
constexpr int kR2lOffA = 0;
constexpr int kR2lOffOnes = kR2lOffA + kABufBytes;
constexpr int kR2lOffRing = kR2lOffOnes + kOnesBytes;
constexpr int kR2lOffBiasRing = kR2lOffRing + kR2lRingBytes;
constexpr int kR2lOffWt = kR2lOffBiasRing + kR2lBiasRingBytes;   // 3*256 floats
constexpr int kR2lOffPart = kR2lOffWt + 768 * 4;                 // 128*4 floats
constexpr int kR2lOffBars = kR2lOffPart + 128 * 4 * 4;
constexpr int kR2lNumBars = 2 * kR2lMaxRing + 2 * kR2lMaxBiasRing + 4 + 2 + 4 + 1;
constexpr int kR2lOffTmem = kR2lOffBars + kR2lNumBars * 8;
constexpr int kR2lSmemBytes = kR2lOffTmem + 16;
static_assert(kR2lSmemBytes <= 227 * 1024, "R2L kernel shared memory exceeds 227 KiB");
static_assert(kR2lOffRing % 1024 == 0, "weight ring must stay 1 KiB aligned");

// PAIR = true: two CTAs of one cluster (a TPC's SM pair) work on two 128-ray tiles with tcgen05.mma.cta_group::2
// (M = 256): each CTA streams only ITS N-half of every weight stage (half the L2 -> shared-memory traffic and half
// the B-operand shared-memory reads per SM), the leader CTA's MMA thread issues for both, and all "operand ready"
// barriers live in the leader and collect the warps of both CTAs; MMA completion is multicast to both.
// DBG = true: the instantiation behind r2l_resmlp_debug_head / r2l_resmlp_profile (tests, profiling) that also dumps
// the head layer's accumulators and keeps the in-kernel cycle counters.
// Inside the unrolled head epilogue that hook was 1500 SASS instructions (24 KB of a 130 KB kernel) — and these
// kernels are sensitive to code size (see mlp_nerf_pp.cu) — so the production instantiation does not contain it.
template <bool BF16, bool PAIR, bool DBG>
__global__ void __launch_bounds__(kR2lThreads, 1) r2l_mlp_kernel(const R2lParams p, const __grid_constant__ R2lPairMaps maps) {
  constexpr int kRing = PAIR ? 8 : 4;
  constexpr int kBiasRing = PAIR ? 4 : 2;
  constexpr uint32_t kStageB = PAIR ? kStageBytes / 2 : kStageBytes;          // bytes of a stage held by THIS CTA
  constexpr uint32_t kBiasB = PAIR ? kBiasStageBytes / 2 : kBiasStageBytes;
  constexpr uint32_t kLboB = (PAIR ? 128 : 256) * 16;
  constexpr int kCtas = PAIR ? 2 : 1;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* const sA = smem + kR2lOffA;
  uint8_t* const sOnes = smem + kR2lOffOnes;
  uint8_t* const sRing = smem + kR2lOffRing;
  uint8_t* const sBiasRing = smem + kR2lOffBiasRing;
  float* const sWt = reinterpret_cast<float*>(smem + kR2lOffWt);
  float* const sPart = reinterpret_cast<float*>(smem + kR2lOffPart);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + kR2lOffBars);
  uint64_t* const w_full = bars;
  uint64_t* const w_empty = bars + kR2lMaxRing;
  uint64_t* const b_full = bars + 2 * kR2lMaxRing;
  uint64_t* const b_empty = b_full + kR2lMaxBiasRing;
  uint64_t* const a_ready = b_empty + kR2lMaxBiasRing;   // [64-column group 0..3]
  uint64_t* const d_full = a_ready + 4;            // [0] = D1 (cols 0..255), [1] = D2 (cols 256..511)
  uint64_t* const a_free = d_full + 2;             // [block 0..3]: head MMAs finished reading block j of A
  uint64_t* const drained = a_free + 4;            // all 8 epilogue warps have read the tile's final accumulator (D2)
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + kR2lOffTmem);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_chunks = p.n_points / 4;
  const int nb = p.n_blocks;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;          // 0 = leader (issues the MMAs)
  // work units: one tile per CTA; a pair takes tiles 2u and 2u+1 of unit u
  const int n_units = PAIR ? (p.n_tiles + 1) / 2 : p.n_tiles;
  const int unit0 = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int unit_step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  // ---- one-time setup ----
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRing; ++i) {
      mbar_init(&w_full[i], 1);   // PAIR: the peer's half stage completes its bytes on the leader's barrier
      mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < kBiasRing; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&a_ready[i], 8 * kCtas);
    mbar_init(&d_full[0], 1);
    mbar_init(&d_full[1], 1);
    for (int i = 0; i < 4; ++i) mbar_init(&a_free[i], 1);
    mbar_init(drained, 8 * kCtas);
    mbar_fence_init();
  }
  write_ones_block<BF16>(sOnes, threadIdx.x, kR2lThreads);
  for (int i = threadIdx.x; i < 768; i += kR2lThreads) sWt[i] = p.w_tail[i];
  fence_proxy_async_smem();
  if (warp == kR2lMmaWarp) {
    if (PAIR)
      tmem_alloc_pair(tmem_slot, 512);
    else
      tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // the peer's barriers are initialised before anybody signals them
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kR2lProducerWarp) {
    // ===================== weight producer (every CTA streams its own half in PAIR mode) =====================
    if (lane == 0) {
      uint32_t g = 0, gb = 0;
      const uint8_t* src = nullptr;
      auto push = [&]() {
        const uint32_t slot = g % kRing;
        mbar_wait_x<PAIR>(&w_empty[slot], ((g / kRing) & 1) ^ 1, p.dbg, 100 + slot);
        if (PAIR && rank != 0) {
          // cp.async.bulk.tensor...cta_group::2 may signal the LEADER's barrier: no relay hop (see mlp_nerf_pp.cu)
          tma2d_g2s_pair_bar(sRing + slot * kStageB, &maps.m16, 0, static_cast<int>((src + kStageB - p.wstream) >> 9),
                             mapa_u32(&w_full[slot], 0));
        } else {
          mbar_expect_tx(&w_full[slot], PAIR ? 2 * kStageB : kStageB);
          bulk_g2s(sRing + slot * kStageB, src, kStageB, &w_full[slot]);
        }
        src += kStageBytes;
        ++g;
      };
      auto push_bias = [&]() {
        const uint32_t slot = gb % kBiasRing;
        mbar_wait_x<PAIR>(&b_empty[slot], ((gb / kBiasRing) & 1) ^ 1, p.dbg, 120 + slot);
        if (PAIR && rank != 0) {
          tma2d_g2s_pair_bar(sBiasRing + slot * kBiasB, &maps.m4, 0, static_cast<int>((src + kBiasB - p.wstream) >> 9),
                             mapa_u32(&b_full[slot], 0));
        } else {
          mbar_expect_tx(&b_full[slot], PAIR ? 2 * kBiasB : kBiasB);
          bulk_g2s(sBiasRing + slot * kBiasB, src, kBiasB, &b_full[slot]);
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
  } else if (warp == kR2lMmaWarp) {
    if (PAIR && rank != 0) {
      // peer CTA: nothing to do here (its half stages signal the leader's barriers directly)
    } else
    // ===================== MMA issuer (leader CTA) =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(BF16, kCtas * kTileM, 256);
      const uint32_t aA = smem_u32(sA);
      const uint32_t aOnes = smem_u32(sOnes);
      const uint32_t aRing = smem_u32(sRing);
      const uint32_t d1 = tmem_base, d2 = tmem_base + 256;
      const uint32_t aBiasRing = smem_u32(sBiasRing);
      uint32_t g = 0, gb = 0;
      uint32_t par_a = 0;   // parity of the a_ready phase the next layer / chunk waits for (all 4 groups in step)
      const bool prof = DBG && p.prof != nullptr;
      long long t_a = 0, t_w = 0, t_a0 = 0;   // t_a0: the part of t_a spent on a layer's FIRST stage
      const long long t_start = prof ? clock64() : 0;
      // the K=16 bias step of a layer (needs no activations)
      auto bias_step = [&](uint32_t d_tmem, bool fresh) {
        const uint32_t slot = gb % kBiasRing;
        const long long c0 = prof ? clock64() : 0;
        mbar_wait_x<PAIR>(&b_full[slot], (gb / kBiasRing) & 1, p.dbg, 240 + slot);
        if (prof) t_w += clock64() - c0;
        tc_fence_after_sync();
        issue_bias_stage<PAIR>(d_tmem, aOnes, aBiasRing + slot * kBiasB, kLboB, idesc, fresh);
        umma_commit_x<PAIR>(&b_empty[slot]);
        ++gb;
      };
      // 4 stages (K = 256) of A accumulated into d_tmem; `free_blocks`: release A blocks to the head encoders
      auto run4 = [&](uint32_t d_tmem, bool free_blocks) {
        for (int st = 0; st < 4; ++st) {
          const uint32_t slot = g % kRing;
          const long long c0 = prof ? clock64() : 0;
          mbar_wait2<PAIR>(&a_ready[st], par_a, &w_full[slot], (g / kRing) & 1, p.dbg, 210 + st);
          if (prof) {
            const long long dt = clock64() - c0;
            t_a += dt;
            if (st == 0) t_a0 += dt;
          }
          tc_fence_after_sync();
          issue_stage<4, PAIR>(d_tmem, aA + st * kGroupBytes, aRing + slot * kStageB, kLboB, idesc, false);
          umma_commit_x<PAIR>(&w_empty[slot]);
          if (free_blocks) umma_commit_x<PAIR>(&a_free[st]);
          ++g;
        }
        par_a ^= 1u;
      };
      uint32_t it = 0;
      for (int unit = unit0; unit < n_units; unit += unit_step, ++it) {
        // head accumulates in D2; W1 of block 0 then writes D1 (consecutive layers never share a TMEM buffer).
        // The previous tile's last epilogue reads D2 without handing anything back through a_ready, so the
        // first (overwriting) MMA of this tile waits until all epilogue warps have drained it.
        if (it > 0) {
          mbar_wait_x<PAIR>(drained, (it - 1) & 1u, p.dbg, 250);
          tc_fence_after_sync();
        }
        bias_step(d2, true);
        for (int c = 0; c < n_chunks; ++c) run4(d2, c + 1 < n_chunks);
        umma_commit_x<PAIR>(&d_full[1]);
        for (int b = 0; b < nb; ++b) {
          bias_step(d1, true);
          run4(d1, false);
          umma_commit_x<PAIR>(&d_full[0]);
          bias_step(d2, false);   // accumulate onto the fp32 residual stream
          run4(d2, false);
          umma_commit_x<PAIR>(&d_full[1]);
        }
      }
      if (prof) {
        long long* o = p.prof + blockIdx.x * 8;
        o[0] = clock64() - t_start;   // MMA thread: total
        o[1] = t_a;                   // waiting for A groups + weight stages (joint wait)
        o[2] = t_w;                   // waiting for bias stages
        if (PAIR) o[8] = t_a0;        // (the peer CTA's unused slot 0) the first-stage share of o[1]
      }
    }
  } else {
    // ===================== epilogue / encoder warpgroups =====================
    const int wg = warp >> 2;                    // owns the 32-column pieces wg, wg+2, wg+4, wg+6
    const int row = (warp & 3) * 32 + lane;      // tile row == TMEM lane
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const bool prof = DBG && p.prof != nullptr && (threadIdx.x & 127) == 0;
    long long t_d = 0, t_enc = 0;
    const long long t_start = prof ? clock64() : 0;
    uint32_t par_d = 0;      // bit dbuf: parity of the next d_full phase
    uint32_t par_free = 0;   // bit j: parity of the next a_free[j] phase
    uint8_t* const a_row = sA + row * 16;
    auto wait_d = [&](int db, uint32_t id) {
      const long long cd = prof ? clock64() : 0;
      mbar_wait_x<PAIR>(&d_full[db], (par_d >> db) & 1u, p.dbg, id);
      if (prof) t_d += clock64() - cd;
      par_d ^= 1u << db;
      tc_fence_after_sync();
    };
    // One 256-wide layer's epilogue: this warp converts the 32-column pieces q = wg, wg+2, wg+4, wg+6 (columns
    // 32q) of its 32 rows, software pipelined (the TMEM load of the next piece is in flight while the current one
    // is processed).  The two warpgroups interleave, so the 64-column K-groups complete in consumption order
    // (group q/2 after ONE piece time each) and the first MMA of the next layer waits for a single piece.
    // f(col0, v) consumes 32 fp32 values starting at column col0; done(group) signals the piece's group.
    auto for_pieces = [&](uint32_t d_col0, auto&& f, auto&& done) {
      uint32_t va[32], vb[32];
      const uint32_t c0 = 32 * wg;
      tmem_ld32(lane_taddr + d_col0 + c0, va);
      tmem_ld_wait();
      tmem_ld32(lane_taddr + d_col0 + c0 + 64, vb);
      f(c0, va);
      done(0);
      tmem_ld_wait();
      tmem_ld32(lane_taddr + d_col0 + c0 + 128, va);
      f(c0 + 64, vb);
      done(1);
      tmem_ld_wait();
      tmem_ld32(lane_taddr + d_col0 + c0 + 192, vb);
      f(c0 + 128, va);
      done(2);
      tmem_ld_wait();
      f(c0 + 192, vb);
      done(3);
    };
    for (int unit = unit0; unit < n_units; unit += unit_step) {
      const int tile = PAIR ? (2 * unit + static_cast<int>(rank)) : unit;   // may be == n_tiles for the peer: rows clamp
      const long long ray = static_cast<long long>(tile) * kTileM + row;
      const bool valid = ray < p.n_rays;
      const long long ray_c = valid ? ray : (p.n_rays - 1);
      const float* prow = (p.pts != nullptr) ? p.pts + ray_c * p.pts_stride : nullptr;
      const float* erow = (p.embedded != nullptr) ? p.embedded + ray_c * p.emb_stride : nullptr;
      // fused ray generation: this row's ray from its pixel index and pose (point_sample_kernel's arithmetic)
      float rox = 0.f, roy = 0.f, roz = 0.f, rdx = 0.f, rdy = 0.f, rdz = 0.f;
      if (p.cam != nullptr) {
        const long long g = p.cam_ray0 + ray_c;
        const long long hw = static_cast<long long>(p.cam_H) * p.cam_W;
        const float* c = p.cam + 12 * (g / hw);
        const int pix = static_cast<int>(g % hw);
        const int h = pix / p.cam_W, w = pix % p.cam_W;
        const float dx = __fdiv_rn(__fsub_rn(static_cast<float>(w), static_cast<float>(p.cam_W * 0.5)), p.cam_focal);
        const float dy = -__fdiv_rn(__fsub_rn(static_cast<float>(h), static_cast<float>(p.cam_H * 0.5)), p.cam_focal);
        rdx = __fadd_rn(__fadd_rn(__fadd_rn(0.0f, __fmul_rn(dx, __ldg(c + 0))), __fmul_rn(dy, __ldg(c + 1))),
                        __fmul_rn(-1.0f, __ldg(c + 2)));
        rdy = __fadd_rn(__fadd_rn(__fadd_rn(0.0f, __fmul_rn(dx, __ldg(c + 4))), __fmul_rn(dy, __ldg(c + 5))),
                        __fmul_rn(-1.0f, __ldg(c + 6)));
        rdz = __fadd_rn(__fadd_rn(__fadd_rn(0.0f, __fmul_rn(dx, __ldg(c + 8))), __fmul_rn(dy, __ldg(c + 9))),
                        __fmul_rn(-1.0f, __ldg(c + 10)));
        rox = __ldg(c + 3), roy = __ldg(c + 7), roz = __ldg(c + 11);
      }
      // ---- head: chunks of 4 points (K = 256); this WG encodes blocks j = wg and wg+2 of every chunk
      const long long ce = prof ? clock64() : 0;
      for (int c = 0; c < n_chunks; ++c) {
#pragma unroll 1
        for (int bi = 0; bi < 2; ++bi) {
          const int j = wg + 2 * bi;
          const int pt = c * 4 + j;
          float px = 0.f, py = 0.f, pz = 0.f;
          if (p.cam != nullptr) {
            const float zz = __ldg(p.cam_z + pt);
            px = __fadd_rn(rox, __fmul_rn(rdx, zz));
            py = __fadd_rn(roy, __fmul_rn(rdy, zz));
            pz = __fadd_rn(roz, __fmul_rn(rdz, zz));
          } else if (erow == nullptr) {
            px = __ldg(prow + 3 * pt);
            py = __ldg(prow + 3 * pt + 1);
            pz = __ldg(prow + 3 * pt + 2);
          }
          if (c > 0) {
            mbar_wait_x<PAIR>(&a_free[j], (par_free >> j) & 1u, p.dbg, 400 + j);
            par_free ^= 1u << j;
          }
          uint8_t* blk = sA + j * 8 * kChunkBytes;
          if (erow == nullptr) {
            encode_point_block<BF16>(blk, row, px, py, pz);
          } else {
            // API path: gather the caller's embedding (reference order (3s+c)*21 + f') into block order
            const float* e = erow + static_cast<long long>(pt) * 63;
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
              uint4 q;
              q.x = pack2<BF16>(v[0], v[1]);
              q.y = pack2<BF16>(v[2], v[3]);
              q.z = pack2<BF16>(v[4], v[5]);
              q.w = pack2<BF16>(v[6], v[7]);
              *reinterpret_cast<uint4*>(blk + ch * kChunkBytes + row * 16) = q;
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) lane_arrive<PAIR>(&a_ready[j], 2);   // 8 arrivals per CTA and phase; a block comes from 4 warps
        }
      }
      if (prof) t_enc += clock64() - ce;

      float t0 = 0.f, t1 = 0.f, t2 = 0.f;
      // ---- head epilogue: x0 = relu(D2) -> stored back in place (fp32 residual stream), A, tail partials
      wait_d(1, 300);
      for_pieces(
          256,
          [&](uint32_t col0, uint32_t (&v)[32]) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float x = relu_nan(__uint_as_float(v[i]));
              if (DBG && p.dbg_head_acc != nullptr) {   // debug hook: raw accumulator (bias included) and x0
                const long long o = (static_cast<long long>(tile) * kTileM + row) * 256 + col0 + i;
                p.dbg_head_acc[o] = __uint_as_float(v[i]);
                p.dbg_head_x0[o] = x;
              }
              v[i] = __float_as_uint(x);
              if (p.outer_skip) {
                t0 = fmaf(sWt[col0 + i], x, t0);
                t1 = fmaf(sWt[256 + col0 + i], x, t1);
                t2 = fmaf(sWt[512 + col0 + i], x, t2);
              }
            }
            tmem_st32(lane_taddr + 256 + col0, v);
            store_sub<BF16, false>(v, a_row + (col0 >> 5) * kSubBytes);
          },
          [&](int g) {
            tmem_st_wait();
            warp_arrive<PAIR>(&a_ready[g], lane);
          });
      // ---- body
      for (int b = 0; b < nb; ++b) {
        // W1: h = relu(D1) -> A
        wait_d(0, 310);
        for_pieces(
            0, [&](uint32_t col0, uint32_t (&v)[32]) { store_sub<BF16, true>(v, a_row + (col0 >> 5) * kSubBytes); },
            [&](int g) { warp_arrive<PAIR>(&a_ready[g], lane); });
        // W2: x = D2 -> A   (last block: tail partials instead)
        wait_d(1, 320);
        if (b + 1 < nb) {
          for_pieces(
              256, [&](uint32_t col0, uint32_t (&v)[32]) { store_sub<BF16, false>(v, a_row + (col0 >> 5) * kSubBytes); },
              [&](int g) { warp_arrive<PAIR>(&a_ready[g], lane); });
        } else {
          for_pieces(
              256,
              [&](uint32_t col0, uint32_t (&v)[32]) {
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
          if (lane == 0) lane_arrive<PAIR>(drained);
        }
      }
      // ---- tail: combine the two column sets, bias, sigmoid
      if (wg == 1) {
        sPart[row * 4 + 0] = t0;
        sPart[row * 4 + 1] = t1;
        sPart[row * 4 + 2] = t2;
        named_bar_arrive(1, 256);
        named_bar_sync(2, 256);   // WG0 has consumed sPart
      } else {
        named_bar_sync(1, 256);
        float o0 = t0 + sPart[row * 4 + 0] + p.b_tail[0];
        float o1 = t1 + sPart[row * 4 + 1] + p.b_tail[1];
        float o2 = t2 + sPart[row * 4 + 2] + p.b_tail[2];
        named_bar_arrive(2, 256);
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
            // fused gather: this rank's rows go straight into every GPU's frame buffer (P2P stores over NVLink; a
            // tile is 1.5 KB per peer).  Kernel completion + the symmetric-memory barrier that follows on the stream
            // publish them; the fence orders them before this CTA's exit at system scope.
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
    }
    if (prof) {
      long long* o = p.prof + blockIdx.x * 8 + 3 + wg * 2;
      o[0] = t_d;                                  // WG: waiting for accumulators
      o[1] = (clock64() - t_start) - t_d - t_enc;  // WG: epilogue work (everything else)
      if (wg == 0) p.prof[blockIdx.x * 8 + 7] = t_enc;
    }
  }

  // ---- teardown ----
  tc_fence_before_sync();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // both CTAs are done with their TMEM and with each other's barriers
  if (warp == kR2lMmaWarp) {
    tc_fence_after_sync();
    if (PAIR)
      tmem_dealloc_pair(tmem_base, 512);
    else
      tmem_dealloc(tmem_base, 512);
  }
}

template <bool BF16, bool PAIR, bool DBG>
int launch_r2l_(const R2lParams& p, const R2lPairMaps& maps, int grid, cudaStream_t st) {
  R2L_CUDA(cudaFuncSetAttribute(r2l_mlp_kernel<BF16, PAIR, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                kR2lSmemBytes));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(kR2lThreads);
  cfg.dynamicSmemBytes = kR2lSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  R2L_CUDA(cudaLaunchKernelEx(&cfg, r2l_mlp_kernel<BF16, PAIR, DBG>, p, maps));
  count_launch();
  return R2L_OK;
}
template <bool BF16, bool PAIR>
int launch_r2l(const R2lParams& p, const R2lPairMaps& maps, int grid, cudaStream_t st) {
  return (p.dbg_head_acc != nullptr || p.prof != nullptr) ? launch_r2l_<BF16, PAIR, true>(p, maps, grid, st)
                                   : launch_r2l_<BF16, PAIR, false>(p, maps, grid, st);
}

// pair: CTA-pair mode (the weights must have been packed in the pair layout); grid is then a multiple of 2
int r2l_mlp_launch(bool bf16, bool pair, const R2lParams& p, const R2lPairMaps* maps, int grid, cudaStream_t st) {
  static const R2lPairMaps none = {};
  if (pair) {
    R2L_CHECK_ARG(maps != nullptr, "r2l_mlp_launch: the CTA-pair kernel needs tensor maps");
    return bf16 ? launch_r2l<true, true>(p, *maps, grid, st) : launch_r2l<false, true>(p, *maps, grid, st);
  }
  return bf16 ? launch_r2l<true, false>(p, none, grid, st) : launch_r2l<false, false>(p, none, grid, st);
}

}  // namespace r2l
