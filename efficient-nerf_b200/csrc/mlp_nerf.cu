// Persistent fused NeRF MLP (W256 x D8, skip at layer 4, sigma/rgb heads with view
// directions) on tcgen05/TMEM.  Reference math: NeRF.forward model/nerf_raybased.py:377-401,
// fed by run_network main.py:65-87 (positional encoding of pts and viewdirs, helpers:24-74)
// and pts = o + d*z main.py:701,733.
//
// Per 128-sample tile (row g = ray*S + s):
//   P   = encode(o + d*z)                63 features + pad, written straight to shared memory
//   h   = relu(W0 P)  ... relu(W4 h)     K = 64 | 256
//   h   = relu(W5 [h, P])                K = 320  (skip: reference order cat[pts, h], weights permuted)
//   h   = relu(W6 h), relu(W7 h);        sigma = w_a . h + b_a   (fp32 CUDA cores, from the fp32 h)
//   f   = Wf h + bf                      no activation
//   v   = relu(Wv[:, :256] f + vb[ray])  N = 128; vb = bv + Wv[:, 256:] embed(viewdir) precomputed
//                                        per RAY in fp32 (the view branch is constant along a ray)
//   rgb = Wr v + br                      fp32 CUDA cores
//   raw[g] = (rgb, sigma)
// 10 tensor-core layers, 76 weight stages of 32 K-columns (68 x 16 KiB + 8 x 8 KiB) per tile.
#include "common.cuh"
#include "mlp_params.cuh"
#include "mlp_tc.cuh"

namespace r2l {

constexpr int kNerfRing = 4;
constexpr int kNerfSteps = 10;
// shared memory map
constexpr int kNerfOffA0 = 0;
constexpr int kNerfOffA1 = kNerfOffA0 + kABufBytes;
constexpr int kNerfOffP = kNerfOffA1 + kABufBytes;
constexpr int kNerfOffRing = kNerfOffP + kPBlockBytes;
constexpr int kNerfOffBias = kNerfOffRing + kNerfRing * kStageBytes;  // 9*256 floats
constexpr int kNerfOffAlphaW = kNerfOffBias + 9 * 256 * 4;            // 256 floats
constexpr int kNerfOffRgbW = kNerfOffAlphaW + 256 * 4;                // 3*128 floats
constexpr int kNerfOffSigma = kNerfOffRgbW + 384 * 4;                 // 128 floats
constexpr int kNerfOffBars = kNerfOffSigma + 128 * 4;                 // mbarriers
constexpr int kNerfNumBars = 2 * kNerfRing + 4 + 2 + 1;
constexpr int kNerfOffTmem = kNerfOffBars + kNerfNumBars * 8;
constexpr int kNerfSmemBytes = kNerfOffTmem + 16;
static_assert(kNerfSmemBytes <= 227 * 1024, "NeRF kernel shared memory exceeds 227 KiB");

__device__ __forceinline__ int nerf_stages_in_step(int step) { return step == 0 ? 2 : (step == 5 ? 10 : 8); }

template <bool BF16>
__global__ void __launch_bounds__(kThreads, 1) nerf_mlp_kernel(const NerfParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* const sA0 = smem + kNerfOffA0;   // A buffers: sA0 + buf*kABufBytes (no dynamically indexed local arrays:
                                             // nvcc 12.9 overlapped two such stack arrays in the R2L kernel)
  uint8_t* sP = smem + kNerfOffP;
  uint8_t* sRing = smem + kNerfOffRing;
  float* sBias = reinterpret_cast<float*>(smem + kNerfOffBias);
  float* sAlphaW = reinterpret_cast<float*>(smem + kNerfOffAlphaW);
  float* sRgbW = reinterpret_cast<float*>(smem + kNerfOffRgbW);
  float* sSigma = reinterpret_cast<float*>(smem + kNerfOffSigma);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kNerfOffBars);
  uint64_t* w_full = bars;
  uint64_t* w_empty = bars + kNerfRing;
  uint64_t* a_ready = bars + 2 * kNerfRing;      // [buf*2 + half]
  uint64_t* d_full = a_ready + 4;                // [dbuf]
  uint64_t* p_ready = d_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kNerfOffTmem);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- one-time setup ----
  for (int i = threadIdx.x; i < 9 * 256; i += kThreads) sBias[i] = p.bias[i];
  for (int i = threadIdx.x; i < 256; i += kThreads) sAlphaW[i] = p.alpha_w[i];
  for (int i = threadIdx.x; i < 384; i += kThreads) sRgbW[i] = p.rgb_w[i];
  if (threadIdx.x == 0) {
    for (int i = 0; i < kNerfRing; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&a_ready[i], 128);
    mbar_init(&d_full[0], 1);
    mbar_init(&d_full[1], 1);
    mbar_init(p_ready, 128);
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
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const uint8_t* src = p.wstream;
        for (int st = 0; st < 76; ++st) {
          const uint32_t bytes = (st < 68) ? kStageBytes : (kStageBytes / 2);
          const uint32_t slot = g % kNerfRing;
          mbar_wait(&w_empty[slot], ((g / kNerfRing) & 1) ^ 1, p.dbg, 100 + slot);
          mbar_expect_tx(&w_full[slot], bytes);
          bulk_g2s(sRing + slot * kStageBytes, src, bytes, &w_full[slot]);
          src += bytes;
          ++g;
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc256 = make_idesc_f16(BF16, kTileM, 256);
      const uint32_t idesc128 = make_idesc_f16(BF16, kTileM, 128);
      const uint32_t aA0 = smem_u32(sA0);
      const uint32_t aP = smem_u32(sP);
      const uint32_t aRing = smem_u32(sRing);
      uint32_t g = 0, cnt_p = 0;
      uint32_t par_a = 0;   // bit (buf*2+half): parity of the next a_ready phase to wait for
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int step = 0; step < kNerfSteps; ++step) {
          const int nst = nerf_stages_in_step(step);
          const uint32_t d_tmem = tmem_base + (step & 1) * 256;
          const uint32_t idesc = (step == 9) ? idesc128 : idesc256;
          const uint32_t lbo_b = (step == 9) ? 128 * 16 : 256 * 16;
          const int abuf = (step - 1) & 1;
          for (int st = 0; st < nst; ++st) {
            uint32_t a_addr;
            if (step == 0) {
              if (st == 0) {
                mbar_wait(p_ready, cnt_p & 1, p.dbg, 200);
                ++cnt_p;
              }
              a_addr = aP + st * 4 * kChunkBytes;
            } else if (step == 5 && st >= 8) {
              a_addr = aP + (st - 8) * 4 * kChunkBytes;
            } else {
              if (st == 0 || st == 4) {
                const int bi = abuf * 2 + (st >> 2);
                mbar_wait(&a_ready[bi], (par_a >> bi) & 1u, p.dbg, 210 + bi);
                par_a ^= 1u << bi;
              }
              a_addr = aA0 + abuf * kABufBytes + st * 4 * kChunkBytes;
            }
            const uint32_t slot = g % kNerfRing;
            mbar_wait(&w_full[slot], (g / kNerfRing) & 1, p.dbg, 220 + slot);
            tc_fence_after_sync();
            issue_stage(d_tmem, a_addr, aRing + slot * kStageBytes, lbo_b, idesc, st == 0);
            umma_commit(&w_empty[slot]);
            ++g;
          }
          umma_commit(&d_full[step & 1]);
        }
      }
    }
  } else {
    // ===================== epilogue / encoder warpgroups =====================
    const int wg = warp >> 2;                       // 0 | 1 : output column half
    const int row = (warp & 3) * 32 + lane;         // tile row == TMEM lane
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t par_d = 0;   // bit dbuf: parity of the next d_full phase
    bool first = true;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const long long g_row = static_cast<long long>(tile) * kTileM + row;
      const bool valid = g_row < p.n_rows;
      const long long g_clamped = valid ? g_row : (p.n_rows - 1);
      auto encode_tile = [&](int t) {
        long long gr = static_cast<long long>(t) * kTileM + row;
        if (gr >= p.n_rows) gr = p.n_rows - 1;
        if (p.embedded != nullptr) {
          // API path: the caller already embedded the points (63 features per row)
          const float* x = p.embedded + gr * p.emb_stride;
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (ch * 8 + i < 63) ? __ldg(x + ch * 8 + i) : 0.0f;
            uint4 q;
            q.x = pack2<BF16>(v[0], v[1]);
            q.y = pack2<BF16>(v[2], v[3]);
            q.z = pack2<BF16>(v[4], v[5]);
            q.w = pack2<BF16>(v[6], v[7]);
            *reinterpret_cast<uint4*>(sP + ch * kChunkBytes + row * 16) = q;
          }
        } else {
          const long long ray = gr / p.S;
          const float z = __ldg(p.z_vals + gr);
          const float* o = p.rays_o + ray * p.o_stride;
          const float* d = p.rays_d + ray * p.d_stride;
          const float px = __fadd_rn(__ldg(o + 0), __fmul_rn(__ldg(d + 0), z));
          const float py = __fadd_rn(__ldg(o + 1), __fmul_rn(__ldg(d + 1), z));
          const float pz = __fadd_rn(__ldg(o + 2), __fmul_rn(__ldg(d + 2), z));
          encode_point_block<BF16>(sP, row, px, py, pz);
        }
        fence_proxy_async_smem();
        mbar_arrive(p_ready);
      };
      if (wg == 1 && first) encode_tile(tile);
      first = false;
      float sigma_part = 0.0f;
      for (int step = 0; step < kNerfSteps; ++step) {
        const int db = step & 1;
        mbar_wait(&d_full[db], (par_d >> db) & 1u, p.dbg, 300 + step);
        par_d ^= 1u << db;
        tc_fence_after_sync();
        const uint32_t d_taddr = lane_taddr + db * 256;
        if (step <= 8) {
          uint8_t* a_dst = sA0 + (step & 1) * kABufBytes + row * 16;
          const float* bias = sBias + step * 256;
          const int c0 = wg * 128;
          if (step == 7) {
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
              const int col0 = c0 + h * 64;
              epilogue_cols64<BF16, true, false>(d_taddr + col0, a_dst + (col0 >> 3) * kChunkBytes, col0, 0u,
                                          [&](int n, float acc) {
                                            const float v = fmaxf(acc + bias[n], 0.0f);
                                            sigma_part = fmaf(sAlphaW[n], v, sigma_part);
                                            return v;
                                          });
            }
            if (wg == 1) sSigma[row] = sigma_part;
          } else if (step == 8) {
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
              const int col0 = c0 + h * 64;
              epilogue_cols64<BF16, true, false>(d_taddr + col0, a_dst + (col0 >> 3) * kChunkBytes, col0, 0u,
                                          [&](int n, float acc) { return acc + bias[n]; });
            }
          } else {
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
              const int col0 = c0 + h * 64;
              epilogue_cols64<BF16, true, false>(d_taddr + col0, a_dst + (col0 >> 3) * kChunkBytes, col0, 0u,
                                          [&](int n, float acc) { return fmaxf(acc + bias[n], 0.0f); });
            }
          }
          fence_proxy_async_smem();
          tc_fence_before_sync();
          mbar_arrive(&a_ready[(step & 1) * 2 + wg]);
          if (step == 7 && wg == 1) named_bar_arrive(1, 256);
        } else {
          // step 9: view branch (N = 128) -> rgb; WG0 only.  WG1 encodes the next tile meanwhile.
          if (wg == 0) {
            const long long ray = g_clamped / p.S;
            const float* vb = p.vb + ray * 128;
            float r = 0.f, gch = 0.f, b = 0.f;
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
              const int col0 = h * 64;
              epilogue_cols64<BF16, false, false>(d_taddr + col0, nullptr, col0, 0u, [&](int n, float acc) {
                const float v = fmaxf(acc + __ldg(vb + n), 0.0f);
                r = fmaf(sRgbW[n], v, r);
                gch = fmaf(sRgbW[128 + n], v, gch);
                b = fmaf(sRgbW[256 + n], v, b);
                return v;
              });
            }
            named_bar_sync(1, 256);
            const float sigma = sigma_part + sSigma[row] + p.alpha_b;
            if (valid) {
              float4 o;
              o.x = r + p.rgb_b[0];
              o.y = gch + p.rgb_b[1];
              o.z = b + p.rgb_b[2];
              o.w = sigma;
              reinterpret_cast<float4*>(p.raw)[g_row] = o;
            }
            tc_fence_before_sync();
          } else {
            const int next = tile + gridDim.x;
            if (next < p.n_tiles) encode_tile(next);
          }
        }
      }
    }
  }

  // ---- teardown ----
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// Per-ray view-branch bias: vb[ray][n] = bv[n] + sum_j Wvd[n][j] * embed4(viewdir)[j]   (fp32)
// embed4 = NeRF Embedder with L = 4 on the unit view direction (27 features).
__global__ void __launch_bounds__(128)
nerf_view_bias_kernel(long long n_rays, const float* __restrict__ viewdirs, long long v_stride, int pre_embedded,
                      const float* __restrict__ wvd /*[128][27]*/, const float* __restrict__ bv, float* __restrict__ vb) {
  __shared__ float s_w[128 * 27];
  __shared__ float s_e[8][28];
  for (int i = threadIdx.x; i < 128 * 27; i += 128) s_w[i] = wvd[i];
  const int n = threadIdx.x;
  const float b = bv[n];
  for (long long r0 = static_cast<long long>(blockIdx.x) * 8; r0 < n_rays; r0 += static_cast<long long>(gridDim.x) * 8) {
    __syncthreads();
    if (pre_embedded) {
      // rows already hold the 27 embedded view features
      for (int it = threadIdx.x; it < 8 * 27; it += 128) {
        const int rr = it / 27, j = it % 27;
        const long long ray = r0 + rr;
        if (ray < n_rays) s_e[rr][j] = viewdirs[ray * v_stride + j];
      }
    } else if (threadIdx.x < 24) {
      // 8 rays x 3 coords: identity + 4 sincos each
      const int rr = threadIdx.x / 3, c = threadIdx.x % 3;
      const long long ray = r0 + rr;
      if (ray < n_rays) {
        const float v = viewdirs[ray * v_stride + c];
        s_e[rr][c] = v;
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          float s, co;
          sincosf(v * static_cast<float>(1 << f), &s, &co);
          s_e[rr][3 + 6 * f + c] = s;
          s_e[rr][3 + 6 * f + 3 + c] = co;
        }
      }
    }
    __syncthreads();
    for (int rr = 0; rr < 8; ++rr) {
      const long long ray = r0 + rr;
      if (ray >= n_rays) break;
      float acc = b;
#pragma unroll
      for (int j = 0; j < 27; ++j) acc = fmaf(s_w[n * 27 + j], s_e[rr][j], acc);
      vb[ray * 128 + n] = acc;
    }
  }
}

template <bool BF16>
int launch_nerf(const NerfParams& p, int grid, cudaStream_t st) {
  R2L_CUDA(cudaFuncSetAttribute(nerf_mlp_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kNerfSmemBytes));
  nerf_mlp_kernel<BF16><<<grid, kThreads, kNerfSmemBytes, st>>>(p);
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

int nerf_mlp_launch(bool bf16, const NerfParams& p, int grid, cudaStream_t st) {
  return bf16 ? launch_nerf<true>(p, grid, st) : launch_nerf<false>(p, grid, st);
}

int nerf_view_bias_launch(long long n_rays, const float* viewdirs, long long v_stride, int pre_embedded,
                          const float* wvd, const float* bv, float* vb, cudaStream_t st) {
  long long blocks = (n_rays + 7) / 8;
  const long long cap = static_cast<long long>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  nerf_view_bias_kernel<<<static_cast<int>(blocks), 128, 0, st>>>(n_rays, viewdirs, v_stride, pre_embedded, wvd, bv, vb);
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

}  // namespace r2l
