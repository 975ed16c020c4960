// Persistent fused R2L network (NeRF_v3_2 with ResMLP body, W256 x D88) on tcgen05/TMEM.
// Reference math: NeRF_v3_2.forward model/nerf_raybased.py:539-544, ResMLP.forward :461-465,
// PositionalEmbedder.__call__ :198-208 on PointSampler points :94-126.
//
// One tile = 128 RAYS.  Per tile:
//   head : x0 = relu(W_h enc(pts) + b_h); K = 64 per 3-D point (63 features + pad, weights
//          permuted at pack time), encoded by the epilogue warps straight into the shared-memory
//          A buffers in chunks of 4 points (K = 256) that ping-pong between the two buffers.
//   body : 43 x [ h = relu(W1 x + b1) ;  x = x + res_scale*(W2 h + b2) ]
//          The residual stream lives in TMEM columns [256,512) in fp32 for the whole body:
//          the head accumulates there, x0 = relu(.) is stored back in place (tcgen05.st) and every W2
//          layer ACCUMULATES onto it
//          (tcgen05.mma with the accumulate flag), so the residual add costs nothing and is
//          exact fp32.  The biases b2 are folded into a per-block cumulative bias cb_b that is
//          added when x is read back (x_{b+1} = D2 + cb_b); W1 layers use TMEM columns [0,256).
//   tail : rgb = sigmoid(W_t (x_43 + x0) + b_t)   (outer skip, use_residual) on CUDA cores in
//          fp32: W_t x0 is accumulated at the head epilogue, W_t x_43 at the last epilogue.
// Weight stream per tile: (P/4)*8 + 43*16 stages of 16 KiB (P = 16: 720 stages, 11.8 MB, L2 resident).
#include "common.cuh"
#include "mlp_params.cuh"
#include "mlp_tc.cuh"

namespace r2l {

constexpr int kR2lRing = 5;
constexpr int kR2lOffA0 = 0;
constexpr int kR2lOffA1 = kR2lOffA0 + kABufBytes;
constexpr int kR2lOffRing = kR2lOffA1 + kABufBytes;
constexpr int kR2lOffPart = kR2lOffRing + kR2lRing * kStageBytes;   // 128*3 floats
constexpr int kR2lOffBars = kR2lOffPart + 128 * 4 * 4;
constexpr int kR2lNumBars = 2 * kR2lRing + 4 + 2 + 2;
constexpr int kR2lOffTmem = kR2lOffBars + kR2lNumBars * 8;
constexpr int kR2lSmemBytes = kR2lOffTmem + 16;
static_assert(kR2lSmemBytes <= 227 * 1024, "R2L kernel shared memory exceeds 227 KiB");

template <bool BF16>
__global__ void __launch_bounds__(kThreads, 1) r2l_mlp_kernel(const R2lParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* const sA0 = smem + kR2lOffA0;   // A buffers are addressed as sA0 + buf*kABufBytes (no local arrays:
                                            // dynamically indexed stack arrays were mis-overlapped by nvcc 12.9)
  uint8_t* sRing = smem + kR2lOffRing;
  float* sPart = reinterpret_cast<float*>(smem + kR2lOffPart);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kR2lOffBars);
  uint64_t* w_full = bars;
  uint64_t* w_empty = bars + kR2lRing;
  uint64_t* a_ready = bars + 2 * kR2lRing;   // [buf*2 + half]
  uint64_t* d_full = a_ready + 4;            // [0] = D1 (cols 0..255), [1] = D2 (cols 256..511)
  uint64_t* a_free = d_full + 2;             // [buf]: head-chunk MMAs finished reading A[buf]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kR2lOffTmem);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_chunks = p.n_points / 4;
  const int nb = p.n_blocks;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kR2lRing; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&a_ready[i], 128);
    mbar_init(&d_full[0], 1);
    mbar_init(&d_full[1], 1);
    mbar_init(&a_free[0], 1);
    mbar_init(&a_free[1], 1);
    mbar_fence_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kProducerWarp) {
    // ===================== weight producer =====================
    if (lane == 0) {
      const int stages_per_tile = n_chunks * 8 + nb * 16;
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const uint8_t* src = p.wstream;
        for (int st = 0; st < stages_per_tile; ++st) {
          const uint32_t slot = g % kR2lRing;
          mbar_wait(&w_empty[slot], ((g / kR2lRing) & 1) ^ 1, p.dbg, 100 + slot);
          mbar_expect_tx(&w_full[slot], kStageBytes);
          bulk_g2s(sRing + slot * kStageBytes, src, kStageBytes, &w_full[slot]);
          src += kStageBytes;
          ++g;
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(BF16, kTileM, 256);
      const uint32_t aA0 = smem_u32(sA0);
      const uint32_t aRing = smem_u32(sRing);
      const uint32_t d1 = tmem_base, d2 = tmem_base + 256;
      uint32_t g = 0;
      uint32_t par_a = 0;   // bit (buf*2+half): parity of the next a_ready phase to wait for
      const bool prof = p.prof != nullptr;
      long long t_a = 0, t_w = 0, t_start = prof ? clock64() : 0;
      // 8 stages (K = 256) from A[buf] into d_tmem
      auto run_k256 = [&](int buf, uint32_t d_tmem, bool fresh) {
        for (int st = 0; st < 8; ++st) {
          if (st == 0 || st == 4) {
            const int bi = buf * 2 + (st >> 2);
            const long long c0 = prof ? clock64() : 0;
            mbar_wait(&a_ready[bi], (par_a >> bi) & 1u, p.dbg, 210 + bi);
            if (prof) t_a += clock64() - c0;
            par_a ^= 1u << bi;
          }
          const uint32_t slot = g % kR2lRing;
          const long long c1 = prof ? clock64() : 0;
          mbar_wait(&w_full[slot], (g / kR2lRing) & 1, p.dbg, 220 + slot);
          if (prof) t_w += clock64() - c1;
          tc_fence_after_sync();
          issue_stage(d_tmem, aA0 + buf * kABufBytes + st * 4 * kChunkBytes, aRing + slot * kStageBytes, 256 * 16, idesc,
                      fresh && st == 0);
          umma_commit(&w_empty[slot]);
          ++g;
        }
      };
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        // head accumulates in D2: the next layer (W1 of block 0) writes D1 and may start on K-half 0 while
        // WG1 is still reading the head accumulators (consecutive layers must never share a TMEM buffer)
        for (int c = 0; c < n_chunks; ++c) {
          run_k256(c & 1, d2, c == 0);
          if (c + 2 < n_chunks) umma_commit(&a_free[c & 1]);
        }
        umma_commit(&d_full[1]);
        for (int b = 0; b < nb; ++b) {
          run_k256(0, d1, true);
          umma_commit(&d_full[0]);
          run_k256(1, d2, false);   // accumulate onto the fp32 residual stream
          umma_commit(&d_full[1]);
        }
      }
      if (prof) {
        long long* o = p.prof + blockIdx.x * 8;
        o[0] = clock64() - t_start;   // MMA thread: total
        o[1] = t_a;                   // waiting for A (epilogues / encoders)
        o[2] = t_w;                   // waiting for weight stages
      }
    }
  } else {
    // ===================== epilogue / encoder warpgroups =====================
    const int wg = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const int c0 = wg * 128;
    const bool prof = p.prof != nullptr && (threadIdx.x & 127) == 0;
    long long t_d = 0, t_enc = 0, t_start = prof ? clock64() : 0;
    uint32_t par_d = 0;      // bit dbuf: parity of the next d_full phase
    uint32_t par_free = 0;   // bit buf: parity of the next a_free phase
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const long long ray = static_cast<long long>(tile) * kTileM + row;
      const bool valid = ray < p.n_rays;
      const long long ray_c = valid ? ray : (p.n_rays - 1);
      const float* prow = (p.pts != nullptr) ? p.pts + ray_c * p.pts_stride : nullptr;
      const float* erow = (p.embedded != nullptr) ? p.embedded + ray_c * p.emb_stride : nullptr;
      // ---- head: encode chunks of 4 points (K = 256); this WG owns points 2*wg, 2*wg+1 of each chunk
      const long long ce = prof ? clock64() : 0;
      for (int c = 0; c < n_chunks; ++c) {
        const int buf = c & 1;
        if (c >= 2) {
          mbar_wait(&a_free[buf], (par_free >> buf) & 1u, p.dbg, 400 + buf);
          par_free ^= 1u << buf;
        }
#pragma unroll 1
        for (int bl = 0; bl < 2; ++bl) {
          const int pt = c * 4 + wg * 2 + bl;
          uint8_t* blk = sA0 + buf * kABufBytes + (wg * 2 + bl) * 8 * kChunkBytes;
          if (erow == nullptr) {
            const float px = __ldg(prow + 3 * pt), py = __ldg(prow + 3 * pt + 1), pz = __ldg(prow + 3 * pt + 2);
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
        }
        fence_proxy_async_smem();
        mbar_arrive(&a_ready[buf * 2 + wg]);
      }
      if (prof) t_enc += clock64() - ce;
      float t0 = 0.f, t1 = 0.f, t2 = 0.f;
      const float* wt = p.w_tail;
      // ---- head epilogue: x0 = relu(D2 + b_h) -> residual stream (written back in place to D2), A[0], tail partials
      {
        const long long cd = prof ? clock64() : 0;
        mbar_wait(&d_full[1], (par_d >> 1) & 1u, p.dbg, 300);
        if (prof) t_d += clock64() - cd;
      }
      par_d ^= 2u;
      tc_fence_after_sync();
      {
        uint8_t* a_dst = sA0 + row * 16;
        const float* bias = p.b_head;
        const bool skip = p.outer_skip != 0;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          const int col0 = c0 + h * 64;
          epilogue_cols64<BF16, true, true>(lane_taddr + 256 + col0, a_dst + (col0 >> 3) * kChunkBytes, col0,
                                            lane_taddr + 256 + col0, [&](int n, float acc) {
                                              const float v = fmaxf(acc + __ldg(bias + n), 0.0f);
                                              if (p.dbg_head_acc != nullptr) {
                                                const long long o = (static_cast<long long>(tile) * kTileM + row) * 256 + n;
                                                p.dbg_head_acc[o] = acc;
                                                p.dbg_head_x0[o] = v;
                                              }
                                              if (skip) {
                                                t0 = fmaf(__ldg(wt + n), v, t0);
                                                t1 = fmaf(__ldg(wt + 256 + n), v, t1);
                                                t2 = fmaf(__ldg(wt + 512 + n), v, t2);
                                              }
                                              return v;
                                            });
        }
        tmem_st_wait();
        fence_proxy_async_smem();
        tc_fence_before_sync();
        mbar_arrive(&a_ready[0 * 2 + wg]);
      }
      // ---- body
      for (int b = 0; b < nb; ++b) {
        // W1: h = relu(D1 + b1) -> A[1]
        {
          const long long cd = prof ? clock64() : 0;
          mbar_wait(&d_full[0], par_d & 1u, p.dbg, 310);
          if (prof) t_d += clock64() - cd;
        }
        par_d ^= 1u;
        tc_fence_after_sync();
        {
          uint8_t* a_dst = sA0 + kABufBytes + row * 16;
          const float* bias = p.b1 + b * 256;
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            const int col0 = c0 + h * 64;
            epilogue_cols64<BF16, true, false>(lane_taddr + col0, a_dst + (col0 >> 3) * kChunkBytes, col0, 0u,
                                               [&](int n, float acc) { return fmaxf(acc + __ldg(bias + n), 0.0f); });
          }
          fence_proxy_async_smem();
          tc_fence_before_sync();
          mbar_arrive(&a_ready[1 * 2 + wg]);
        }
        // W2: x = D2 + cb_b -> A[0]   (last block: tail partials instead)
        {
          const long long cd = prof ? clock64() : 0;
          mbar_wait(&d_full[1], (par_d >> 1) & 1u, p.dbg, 320);
          if (prof) t_d += clock64() - cd;
        }
        par_d ^= 2u;
        tc_fence_after_sync();
        {
          uint8_t* a_dst = sA0 + row * 16;
          const float* bias = p.cb + b * 256;
          if (b + 1 < nb) {
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
              const int col0 = c0 + h * 64;
              epilogue_cols64<BF16, true, false>(lane_taddr + 256 + col0, a_dst + (col0 >> 3) * kChunkBytes, col0,
                                                 0u, [&](int n, float acc) { return acc + __ldg(bias + n); });
            }
            fence_proxy_async_smem();
            tc_fence_before_sync();
            mbar_arrive(&a_ready[0 * 2 + wg]);
          } else {
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
              const int col0 = c0 + h * 64;
              epilogue_cols64<BF16, false, false>(lane_taddr + 256 + col0, nullptr, col0, 0u,
                                                  [&](int n, float acc) {
                                                    const float v = acc + __ldg(bias + n);
                                                    t0 = fmaf(__ldg(wt + n), v, t0);
                                                    t1 = fmaf(__ldg(wt + 256 + n), v, t1);
                                                    t2 = fmaf(__ldg(wt + 512 + n), v, t2);
                                                    return v;
                                                  });
            }
            tc_fence_before_sync();
          }
        }
      }
      // ---- tail: combine the two column halves, bias, sigmoid
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

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

template <bool BF16>
int launch_r2l(const R2lParams& p, int grid, cudaStream_t st) {
  R2L_CUDA(cudaFuncSetAttribute(r2l_mlp_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kR2lSmemBytes));
  r2l_mlp_kernel<BF16><<<grid, kThreads, kR2lSmemBytes, st>>>(p);
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

int r2l_mlp_launch(bool bf16, const R2lParams& p, int grid, cudaStream_t st) {
  return bf16 ? launch_r2l<true>(p, grid, st) : launch_r2l<false>(p, grid, st);
}

}  // namespace r2l
