// R2L network (NeRF_v3_2 with ResMLP body, W256 x D88), "v3" schedule: CTA pair + double-buffered activations +
// N-half split.  Same math as mlp_r2l.cu (model/nerf_raybased.py:443-544, :198-208); see that file for the
// per-tile data flow.  What changes is the schedule that hides the epilogue latency:
//
//   * Two CTAs of one cluster (the SM pair of a TPC) run tcgen05.mma.cta_group::2 (M = 256: one 128-ray tile per
//     CTA).  Each CTA streams only its quarter-blocks of every weight stage from L2 (half the traffic), which frees
//     enough shared memory for a SECOND activation buffer: layer l reads A[q] and its epilogue writes A[q^1].
//   * Every 256-wide layer is computed as two N-halves H0 = columns [0,128), H1 = [128,256) (MMA N = 128):
//         bias(H0,H1) | K0: H0,H1 | K1: H0,H1 | K2: H0 | K3: H0 -> commit H0 | K2: H1 | K3: H1 -> commit H1
//     H0's epilogue (TMEM -> 16-bit -> A[q^1] columns 0..127) runs while the tensor pipe still works on H1, and the
//     next layer can do its bias step and K-stages 0,1 (which only need columns 0..127) for BOTH halves — 1152
//     tensor cycles — before it needs anything from H1's epilogue.  The pipe idles only if H0's epilogue latency
//     exceeds the 512 cycles of H1's K2/K3 plus the next bias step.
//   * Weight stage layout per CTA r: for each N-half h the 64 rows n = 128h + 64r + i, k-chunk major (LBO 1024).
//   * TWO issuer threads, one per N-half.  Measured (scratch/ubench/mmaissue.cu): a thread issues a tcgen05.mma in
//     ~50 cycles and a tcgen05.commit costs it ~190 cycles, so ONE thread cannot feed N = 128 MMAs (64 tensor
//     cycles each) plus the commits that release weight slots; two threads with independent accumulators can.
//     Issuer H1 holds back its K2/K3 MMAs until issuer H0 has issued its own (mbarrier h0_ahead), which keeps the
//     H0-first order on the tensor pipe.
#include "common.cuh"
#include "mlp_params.cuh"
#include "mlp_tc.cuh"

namespace r2l {

constexpr int kV3Threads = 352;
constexpr int kV3ProducerWarp = 8;
constexpr int kV3MmaWarp = 9;      // issues the MMAs of N-half H0 (leader CTA) / relays weight arrivals (peer CTA)
constexpr int kV3MmaWarp1 = 10;    // issues the MMAs of N-half H1 (leader CTA)
constexpr int kV3Ring = 4;                          // 16 KiB: this CTA's part of a K=64 stage (2 halves x 64 rows)
constexpr int kV3BiasRing = 2;                      // 4 KiB: this CTA's part of a bias stage
constexpr uint32_t kV3StageB = kStageBytes / 2;     // 16 KiB
constexpr uint32_t kV3HalfB = kV3StageB / 2;        // 8 KiB: one N-half block (64 rows x 64 K)
constexpr uint32_t kV3BiasB = kBiasStageBytes / 2;  // 4 KiB
constexpr uint32_t kV3BiasHalfB = kV3BiasB / 2;     // 2 KiB
constexpr uint32_t kV3LboB = 64 * 16;               // 64 B-rows per CTA and MMA
constexpr int kV3OffA = 0;                          // two activation buffers
constexpr int kV3OffOnes = kV3OffA + 2 * kABufBytes;
constexpr int kV3OffRing = kV3OffOnes + kOnesBytes;
constexpr int kV3OffBiasRing = kV3OffRing + kV3Ring * kV3StageB;
constexpr int kV3OffWt = kV3OffBiasRing + kV3BiasRing * kV3BiasB;   // 3*256 floats
constexpr int kV3OffPart = kV3OffWt + 768 * 4;                      // 128*4 floats
constexpr int kV3OffBars = kV3OffPart + 128 * 4 * 4;
constexpr int kV3NumBars = 2 * kV3Ring + 2 * kV3BiasRing + 8 + 4 + 8 + 1 + 1;
constexpr int kV3OffTmem = kV3OffBars + kV3NumBars * 8;
constexpr int kV3SmemBytes = kV3OffTmem + 16;
static_assert(kV3SmemBytes <= 227 * 1024, "R2L v3 kernel shared memory exceeds 227 KiB");
static_assert(kV3OffRing % 1024 == 0, "weight ring must stay 1 KiB aligned");

template <bool BF16>
__global__ void __launch_bounds__(kV3Threads, 1) r2l_mlp_v3_kernel(const R2lParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* const sA = smem + kV3OffA;
  uint8_t* const sOnes = smem + kV3OffOnes;
  uint8_t* const sRing = smem + kV3OffRing;
  uint8_t* const sBiasRing = smem + kV3OffBiasRing;
  float* const sWt = reinterpret_cast<float*>(smem + kV3OffWt);
  float* const sPart = reinterpret_cast<float*>(smem + kV3OffPart);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + kV3OffBars);
  uint64_t* const w_full = bars;                          // leader: own producer + peer relay
  uint64_t* const w_empty = w_full + kV3Ring;
  uint64_t* const b_full = w_empty + kV3Ring;
  uint64_t* const b_empty = b_full + kV3BiasRing;
  // [abuf*4 + 64-column group], in the leader, 16 warp arrivals.  One set per activation buffer: the head encoders
  // run up to two chunks (= both buffers) ahead of the MMAs, and an mbarrier must not complete two phases before its
  // waiter has seen the first.
  uint64_t* const a_ready = b_empty + kV3BiasRing;
  uint64_t* const d_full = a_ready + 8;                   // [dbuf*2 + half]
  uint64_t* const a_free = d_full + 4;                    // [abuf*4 + block]: head MMAs finished reading that block
  uint64_t* const drained = a_free + 8;                   // in the leader, 16 warp arrivals
  uint64_t* const h0_ahead = drained + 1;                 // leader: issuer H0 has issued K2,K3 of the current layer
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + kV3OffTmem);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_chunks = p.n_points / 4;
  const int nb = p.n_blocks;
  const uint32_t rank = cluster_ctarank();                // 0 = leader (issues the MMAs)
  const int n_units = (p.n_tiles + 1) / 2;                // a pair takes tiles 2u and 2u+1 of unit u
  const int unit0 = static_cast<int>(blockIdx.x >> 1);
  const int unit_step = static_cast<int>(gridDim.x >> 1);
  const int passes_per_tile = n_chunks + 2 * nb;          // K-passes (each flips the activation buffer)

  // ---- one-time setup ----
  if (threadIdx.x == 0) {
    for (int i = 0; i < kV3Ring; ++i) {
      mbar_init(&w_full[i], rank == 0 ? 2 : 1);
      mbar_init(&w_empty[i], 2);   // one tcgen05.commit from each of the two issuer threads
    }
    for (int i = 0; i < kV3BiasRing; ++i) {
      mbar_init(&b_full[i], rank == 0 ? 2 : 1);
      mbar_init(&b_empty[i], 2);
    }
    for (int i = 0; i < 8; ++i) mbar_init(&a_ready[i], 16);
    for (int i = 0; i < 4; ++i) mbar_init(&d_full[i], 1);
    for (int i = 0; i < 8; ++i) mbar_init(&a_free[i], 2);
    mbar_init(drained, 16);
    mbar_init(h0_ahead, 1);
    mbar_fence_init();
  }
  write_ones_block<BF16>(sOnes, threadIdx.x, kV3Threads);
  for (int i = threadIdx.x; i < 768; i += kV3Threads) sWt[i] = p.w_tail[i];
  fence_proxy_async_smem();
  if (warp == kV3MmaWarp) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers are initialised before anybody signals them
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  // ===================== MMA issuer for N-half h (leader CTA; one thread per half) =====================
  auto mma_issuer = [&](const int h) {
    const uint32_t idesc = make_idesc_f16(BF16, 2 * kTileM, 128);
    const uint32_t aA = smem_u32(sA);
    const uint32_t aOnes = smem_u32(sOnes);
    const uint32_t aRing = smem_u32(sRing) + h * kV3HalfB;
    const uint32_t aBiasRing = smem_u32(sBiasRing) + h * kV3BiasHalfB;
    const uint32_t dh = 128u * h;
    uint32_t g = 0, gb = 0;
    uint32_t par_a = 0;   // bit q: parity of the a_ready phase the next K-pass over buffer q waits for
    uint32_t q = 0;       // activation buffer the next K-pass reads
    uint32_t par_h0 = 0;  // parity of the next h0_ahead phase (issuer H1 only)
    const bool prof = p.prof != nullptr && h == 0;
    long long t_a = 0, t_w = 0;
    const long long t_start = prof ? clock64() : 0;
    // the K=16 bias step of this half (needs no activations)
    auto bias_step = [&](uint32_t d, bool fresh) {
      const uint32_t slot = gb % kV3BiasRing;
      const long long c0 = prof ? clock64() : 0;
      mbar_wait(&b_full[slot], (gb / kV3BiasRing) & 1, p.dbg, 240 + slot);
      if (prof) t_w += clock64() - c0;
      tc_fence_after_sync();
      issue_bias_stage<true>(d + dh, aOnes, aBiasRing + slot * kV3BiasB, kV3LboB, idesc, fresh);
      umma_commit_pair(&b_empty[slot]);
      ++gb;
    };
    // wait for K-group st of A[q] and for the weight stage, then issue this half's four K=16 MMAs
    auto stage = [&](uint32_t d, int st) -> uint32_t {
      const uint32_t slot = (g + st) % kV3Ring;
      const long long c0 = prof ? clock64() : 0;
      mbar_wait2(&a_ready[q * 4 + st], (par_a >> q) & 1u, &w_full[slot], ((g + st) / kV3Ring) & 1, p.dbg,
                 210 + st + 20 * h);
      if (prof) t_a += clock64() - c0;
      tc_fence_after_sync();
      issue_stage<4, true>(d + dh, aA + q * kABufBytes + st * kGroupBytes, aRing + slot * kV3StageB, kV3LboB, idesc,
                           false);
      return slot;
    };
    auto end_pass = [&]() {
      g += 4;
      par_a ^= 1u << q;
      q ^= 1u;
    };
    // one K = 256 pass over A[q] into accumulator d; H0's K2/K3 go to the tensor pipe before H1's
    auto layer = [&](uint32_t d, int db) {
      umma_commit_pair(&w_empty[stage(d, 0)]);
      umma_commit_pair(&w_empty[stage(d, 1)]);
      if (h == 1) {
        mbar_wait(h0_ahead, par_h0, p.dbg, 260);
        par_h0 ^= 1u;
      }
      const uint32_t s2 = stage(d, 2);
      const uint32_t s3 = stage(d, 3);
      if (h == 0) mbar_arrive(h0_ahead);
      umma_commit_pair(&d_full[db * 2 + h]);
      umma_commit_pair(&w_empty[s2]);
      umma_commit_pair(&w_empty[s3]);
      end_pass();
    };
    const uint32_t d1 = tmem_base, d2 = tmem_base + 256;
    uint32_t it = 0;
    for (int unit = unit0; unit < n_units; unit += unit_step, ++it) {
      // The previous tile's last epilogue reads D2 without handing anything back through a_ready: the first
      // (overwriting) MMA of this tile waits until all 16 epilogue warps of the pair have drained it.
      if (it > 0) {
        mbar_wait(drained, (it - 1) & 1u, p.dbg, 250);
        tc_fence_after_sync();
      }
      // ---- head: K = 64 per point, chunks of 4 points alternate between the two A buffers; accumulates in D2
      bias_step(d2, true);
      for (int c = 0; c < n_chunks; ++c) {
        for (int st = 0; st < 4; ++st) {
          const uint32_t s = stage(d2, st);
          umma_commit_pair(&w_empty[s]);
          if (c + 2 < n_chunks) umma_commit_pair(&a_free[q * 4 + st]);   // encoders run two chunks ahead
        }
        end_pass();
      }
      umma_commit_pair(&d_full[2 + h]);
      // ---- body
      for (int b = 0; b < nb; ++b) {
        bias_step(d1, true);
        layer(d1, 0);
        bias_step(d2, false);   // accumulate onto the fp32 residual stream
        layer(d2, 1);
      }
    }
    if (prof) {
      long long* o = p.prof + blockIdx.x * 8;
      o[0] = clock64() - t_start;   // issuer H0: total
      o[1] = t_a;                   // waiting for A groups + weight stages (joint wait)
      o[2] = t_w;                   // waiting for bias stages
    }
  };

  if (warp == kV3ProducerWarp) {
    // ===================== weight producer: this CTA's part of every stage =====================
    if (lane == 0) {
      uint32_t g = 0, gb = 0;
      const uint8_t* src = nullptr;
      auto push = [&]() {
        const uint32_t slot = g % kV3Ring;
        mbar_wait(&w_empty[slot], ((g / kV3Ring) & 1) ^ 1, p.dbg, 100 + slot, 8);
        mbar_expect_tx(&w_full[slot], kV3StageB);
        bulk_g2s(sRing + slot * kV3StageB, src + rank * kV3StageB, kV3StageB, &w_full[slot]);
        src += kStageBytes;
        ++g;
      };
      auto push_bias = [&]() {
        const uint32_t slot = gb % kV3BiasRing;
        mbar_wait(&b_empty[slot], ((gb / kV3BiasRing) & 1) ^ 1, p.dbg, 120 + slot, 8);
        mbar_expect_tx(&b_full[slot], kV3BiasB);
        bulk_g2s(sBiasRing + slot * kV3BiasB, src + rank * kV3BiasB, kV3BiasB, &b_full[slot]);
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
  } else if (warp == kV3MmaWarp) {
    if (rank != 0) {
      // ===================== peer CTA: relay "my part of the stage has landed" to the leader =====================
      if (lane == 0) {
        uint32_t g = 0, gb = 0;
        auto relay = [&]() {
          const uint32_t slot = g % kV3Ring;
          mbar_wait(&w_full[slot], (g / kV3Ring) & 1, p.dbg, 150 + slot, 8);
          mbar_arrive_cluster(mapa_u32(&w_full[slot], 0));
          ++g;
        };
        auto relay_bias = [&]() {
          const uint32_t slot = gb % kV3BiasRing;
          mbar_wait(&b_full[slot], (gb / kV3BiasRing) & 1, p.dbg, 170 + slot, 8);
          mbar_arrive_cluster(mapa_u32(&b_full[slot], 0));
          ++gb;
        };
        for (int unit = unit0; unit < n_units; unit += unit_step) {
          relay_bias();
          for (int i = 0; i < n_chunks * 4; ++i) relay();
          for (int l = 0; l < 2 * nb; ++l) {
            relay_bias();
            for (int i = 0; i < 4; ++i) relay();
          }
        }
      }
    } else if (lane == 0) {
      mma_issuer(0);
    }
  } else if (warp == kV3MmaWarp1) {
    if (rank == 0 && lane == 0) mma_issuer(1);
  } else {
    // ===================== epilogue / encoder warpgroups (both CTAs) =====================
    const int wg = warp >> 2;                    // owns the 32-column pieces wg, wg+2, wg+4, wg+6
    const int row = (warp & 3) * 32 + lane;      // tile row == TMEM lane
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const bool prof = p.prof != nullptr && (threadIdx.x & 127) == 0;
    long long t_d = 0, t_enc = 0;
    const long long t_start = prof ? clock64() : 0;
    uint32_t par_d = 0;      // bit (dbuf*2+half): parity of the next d_full phase
    uint32_t par_free = 0;   // bit (abuf*4+block): parity of the next a_free phase
    uint32_t q = 0;          // activation buffer the next K-pass reads (same sequence as the MMA thread)
    auto wait_d = [&](int idx, uint32_t id) {
      const long long cd = prof ? clock64() : 0;
      mbar_wait(&d_full[idx], (par_d >> idx) & 1u, p.dbg, id, 4);
      if (prof) t_d += clock64() - cd;
      par_d ^= 1u << idx;
      tc_fence_after_sync();
    };
    // One layer's epilogue: this warp converts the 32-column pieces at columns 32*wg + 64*i (i = 0..3) of its 32
    // rows; pieces 0,1 belong to N-half H0 (ready first), pieces 2,3 to H1.  The TMEM load of the second piece of a
    // half is in flight while the first is processed.  f(col0, v); done(group i) signals the piece's K-group.
    auto for_pieces = [&](int db, auto&& f, auto&& done) {
      uint32_t va[32], vb[32];
      const uint32_t c0 = 32 * wg;
      const uint32_t dcol = db * 256;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        wait_d(db * 2 + h, 300 + db * 2 + h);
        tmem_ld32(lane_taddr + dcol + c0 + 128 * h, va);
        tmem_ld_wait();
        tmem_ld32(lane_taddr + dcol + c0 + 128 * h + 64, vb);
        f(c0 + 128 * h, va);
        done(2 * h);
        tmem_ld_wait();
        f(c0 + 128 * h + 64, vb);
        done(2 * h + 1);
      }
    };
    for (int unit = unit0; unit < n_units; unit += unit_step) {
      const int tile = 2 * unit + static_cast<int>(rank);   // may be == n_tiles for the peer: rows clamp, no stores
      const long long ray = static_cast<long long>(tile) * kTileM + row;
      const bool valid = ray < p.n_rays;
      const long long ray_c = valid ? ray : (p.n_rays - 1);
      const float* prow = (p.pts != nullptr) ? p.pts + ray_c * p.pts_stride : nullptr;
      const float* erow = (p.embedded != nullptr) ? p.embedded + ray_c * p.emb_stride : nullptr;
      // ---- head: chunk c goes to buffer q^(c&1); this WG encodes blocks j = wg and wg+2 of every chunk
      const long long ce = prof ? clock64() : 0;
      for (int c = 0; c < n_chunks; ++c) {
        const uint32_t qb = q ^ (c & 1);
#pragma unroll 1
        for (int bi = 0; bi < 2; ++bi) {
          const int j = wg + 2 * bi;
          const int pt = c * 4 + j;
          float px = 0.f, py = 0.f, pz = 0.f;
          if (erow == nullptr) {
            px = __ldg(prow + 3 * pt);
            py = __ldg(prow + 3 * pt + 1);
            pz = __ldg(prow + 3 * pt + 2);
          }
          if (c >= 2) {
            const int fi = qb * 4 + j;
            mbar_wait(&a_free[fi], (par_free >> fi) & 1u, p.dbg, 400 + fi, 4);
            par_free ^= 1u << fi;
          }
          uint8_t* blk = sA + qb * kABufBytes + j * kGroupBytes;
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
              uint4 qq;
              qq.x = pack2<BF16>(v[0], v[1]);
              qq.y = pack2<BF16>(v[2], v[3]);
              qq.z = pack2<BF16>(v[4], v[5]);
              qq.w = pack2<BF16>(v[6], v[7]);
              *reinterpret_cast<uint4*>(blk + ch * kChunkBytes + row * 16) = qq;
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) lane_arrive<true>(&a_ready[qb * 4 + j], 2);   // 16 arrivals per phase; 4 warps per CTA encode a block
        }
      }
      q ^= static_cast<uint32_t>(n_chunks & 1);
      if (prof) t_enc += clock64() - ce;

      float t0 = 0.f, t1 = 0.f, t2 = 0.f;
      // ---- head epilogue: x0 = relu(D2) -> stored back in place (fp32 residual stream), A[q], tail partials
      {
        uint8_t* const a_row = sA + q * kABufBytes + row * 16;
        for_pieces(
            1,
            [&](uint32_t col0, uint32_t (&v)[32]) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float x = fmaxf(__uint_as_float(v[i]), 0.0f);
                if (p.dbg_head_acc != nullptr && valid) {   // debug hook: raw accumulator (bias included) and x0
                  const long long o = ray * 256 + col0 + i;
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
              warp_arrive<true>(&a_ready[q * 4 + g], lane);
            });
      }
      // ---- body: every layer reads A[q] and writes A[q^1]
      for (int b = 0; b < nb; ++b) {
        {
          // W1: h = relu(D1)
          uint8_t* const a_row = sA + (q ^ 1u) * kABufBytes + row * 16;
          for_pieces(
              0, [&](uint32_t col0, uint32_t (&v)[32]) { store_sub<BF16, true>(v, a_row + (col0 >> 5) * kSubBytes); },
              [&](int g) { warp_arrive<true>(&a_ready[(q ^ 1u) * 4 + g], lane); });
          q ^= 1u;
        }
        if (b + 1 < nb) {
          // W2: x = D2
          uint8_t* const a_row = sA + (q ^ 1u) * kABufBytes + row * 16;
          for_pieces(
              1, [&](uint32_t col0, uint32_t (&v)[32]) { store_sub<BF16, false>(v, a_row + (col0 >> 5) * kSubBytes); },
              [&](int g) { warp_arrive<true>(&a_ready[(q ^ 1u) * 4 + g], lane); });
          q ^= 1u;
        } else {
          // last W2: tail partials from the fp32 residual stream
          for_pieces(
              1,
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
          q ^= 1u;
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) lane_arrive<true>(drained);
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
        if (p.sigmoid_out) {
          o0 = 1.0f / (1.0f + expf(-o0));
          o1 = 1.0f / (1.0f + expf(-o1));
          o2 = 1.0f / (1.0f + expf(-o2));
        }
        if (valid) {
          p.rgb[3 * ray + 0] = o0;
          p.rgb[3 * ray + 1] = o1;
          p.rgb[3 * ray + 2] = o2;
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
  cluster_sync_all();   // both CTAs are done with their TMEM and with each other's barriers
  if (warp == kV3MmaWarp) {
    tc_fence_after_sync();
    tmem_dealloc_pair(tmem_base, 512);
  }
  (void)passes_per_tile;
}

template <bool BF16>
int launch_r2l_v3(const R2lParams& p, int grid, cudaStream_t st) {
  R2L_CUDA(cudaFuncSetAttribute(r2l_mlp_v3_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kV3SmemBytes));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(kV3Threads);
  cfg.dynamicSmemBytes = kV3SmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  R2L_CUDA(cudaLaunchKernelEx(&cfg, r2l_mlp_v3_kernel<BF16>, p));
  return R2L_OK;
}

// grid must be even (CTA pairs); the weights must have been packed in the v3 layout (mlp_api.cu)
int r2l_mlp_v3_launch(bool bf16, const R2lParams& p, int grid, cudaStream_t st) {
  return bf16 ? launch_r2l_v3<true>(p, grid, st) : launch_r2l_v3<false>(p, grid, st);
}

}  // namespace r2l
