// Persistent fused NeRF MLP (W256 x D8, skip at layer 4, sigma/rgb heads with view
// directions) on tcgen05/TMEM.  Reference math: NeRF.forward model/nerf_raybased.py:377-401,
// fed by run_network main.py:65-87 (positional encoding of pts and viewdirs, helpers:24-74)
// and pts = o + d*z main.py:701,733.
//
// Per 128-sample tile (row g = ray*S + s), 10 tensor-core steps (see mlp_tc.cuh for the execution model):
//   P   = encode(o + d*z)                63 features + pad  -> shared memory (encoder warps, one tile ahead)
//   V   = [encode4(viewdir), 1, 1, 0..]  27 features + the two "ones" columns that carry the view bias
//   0   : h = relu(W0 P + b0)                         K = 16 (bias) + 64
//   1-4 : h = relu(Wl h + bl)                         K = 16 + 256
//   5   : h = relu(W5 [P, h] + b5)                    K = 16 + 64 + 256   (skip; reference order cat[pts, h])
//   6,7 : h = relu(Wl h + bl);  sigma = w_a . h7 + b_a   (fp32 CUDA cores, from the fp32 accumulators)
//   8   : f = Wf h + bf                               no activation
//   9   : v = relu(Wv [V, f])                         N = 128, K = 32 + 256  (bias rides in V's ones columns)
//         rgb = Wr v + br                             fp32 CUDA cores
//   raw[g] = (rgb, sigma)
// Weight stream per tile: 9 bias stages + the view stage (8 KiB each, small ring), 34 K=64 stages of 32 KiB and
// 4 of 16 KiB for the N = 128 layer (main ring).
#include "common.cuh"
#include "mlp_params.cuh"
#include "mlp_tc.cuh"

namespace r2l {

constexpr int kNerfThreads = 448;
constexpr int kNerfProducerWarp = 12;
constexpr int kNerfMmaWarp = 13;
constexpr int kNerfRing = 3;       // 32 KiB weight stages (K = 64)
constexpr int kNerfBiasRing = 1;   // 8 KiB bias / view stages
// shared memory map
constexpr int kNerfOffA = 0;
constexpr int kNerfOffP = kNerfOffA + kABufBytes;                     // 2 point blocks (double buffered)
constexpr int kNerfOffV = kNerfOffP + 2 * kPBlockBytes;               // 2 view blocks
constexpr int kNerfOffOnes = kNerfOffV + 2 * kVBlockBytes;
constexpr int kNerfOffRing = kNerfOffOnes + kOnesBytes;
constexpr int kNerfOffBiasRing = kNerfOffRing + kNerfRing * kStageBytes;
constexpr int kNerfOffAlphaW = kNerfOffBiasRing + kNerfBiasRing * kBiasStageBytes;   // 256 floats
constexpr int kNerfOffRgbW = kNerfOffAlphaW + 256 * 4;                // 3*128 floats
constexpr int kNerfOffPart = kNerfOffRgbW + 384 * 4;                  // 128 x float4
constexpr int kNerfOffAbs = kNerfOffPart + 128 * 16;                  // 128 floats
constexpr int kNerfOffBars = kNerfOffAbs + 128 * 4;
constexpr int kNerfNumBars = 2 * kNerfRing + 2 * kNerfBiasRing + 4 + 2 + 2 + 2 + 1;
constexpr int kNerfOffTmem = kNerfOffBars + kNerfNumBars * 8;
constexpr int kNerfSmemBytes = kNerfOffTmem + 16;
static_assert(kNerfSmemBytes <= 227 * 1024, "NeRF kernel shared memory exceeds 227 KiB");
static_assert(kNerfOffRing % 1024 == 0, "weight ring must stay 1 KiB aligned");

template <bool BF16>
__global__ void __launch_bounds__(kNerfThreads, 1) nerf_mlp_kernel(const NerfParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* const sA = smem + kNerfOffA;
  uint8_t* const sP = smem + kNerfOffP;
  uint8_t* const sV = smem + kNerfOffV;
  uint8_t* const sOnes = smem + kNerfOffOnes;
  uint8_t* const sRing = smem + kNerfOffRing;
  uint8_t* const sBiasRing = smem + kNerfOffBiasRing;
  float* const sAlphaW = reinterpret_cast<float*>(smem + kNerfOffAlphaW);
  float* const sRgbW = reinterpret_cast<float*>(smem + kNerfOffRgbW);
  float4* const sPart = reinterpret_cast<float4*>(smem + kNerfOffPart);
  float* const sAbs = reinterpret_cast<float*>(smem + kNerfOffAbs);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + kNerfOffBars);
  uint64_t* const w_full = bars;
  uint64_t* const w_empty = bars + kNerfRing;
  uint64_t* const b_full = bars + 2 * kNerfRing;
  uint64_t* const b_empty = b_full + kNerfBiasRing;
  uint64_t* const a_ready = b_empty + kNerfBiasRing;   // [64-column group 0..3]
  uint64_t* const d_full = a_ready + 4;             // [dbuf]
  uint64_t* const p_ready = d_full + 2;             // [buf]: P/V blocks of a tile are encoded
  uint64_t* const p_free = p_ready + 2;             // [buf]: all MMAs of the tile that used them have completed
  uint64_t* const drained = p_free + 2;             // all 8 epilogue warps have read the view-branch accumulator (D1)
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + kNerfOffTmem);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- one-time setup ----
  for (int i = threadIdx.x; i < 256; i += kNerfThreads) sAlphaW[i] = p.alpha_w[i];
  for (int i = threadIdx.x; i < 384; i += kNerfThreads) sRgbW[i] = p.rgb_w[i];
  write_ones_block<BF16>(sOnes, threadIdx.x, kNerfThreads);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kNerfRing; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < kNerfBiasRing; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&a_ready[i], 8);
    mbar_init(&d_full[0], 1);
    mbar_init(&d_full[1], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&p_ready[i], 4);
      mbar_init(&p_free[i], 1);
    }
    mbar_init(drained, 8);
    mbar_fence_init();
  }
  fence_proxy_async_smem();
  if (warp == kNerfMmaWarp) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kNerfProducerWarp) {
    // ===================== weight producer =====================
    if (lane == 0) {
      uint32_t g = 0, gb = 0;
      const uint8_t* src = nullptr;
      auto push = [&](uint32_t bytes) {
        const uint32_t slot = g % kNerfRing;
        mbar_wait(&w_empty[slot], ((g / kNerfRing) & 1) ^ 1, p.dbg, 100 + slot);
        mbar_expect_tx(&w_full[slot], bytes);
        bulk_g2s(sRing + slot * kStageBytes, src, bytes, &w_full[slot]);
        src += bytes;
        ++g;
      };
      auto push_bias = [&]() {
        const uint32_t slot = gb % kNerfBiasRing;
        mbar_wait(&b_empty[slot], ((gb / kNerfBiasRing) & 1) ^ 1, p.dbg, 120 + slot);
        mbar_expect_tx(&b_full[slot], kBiasStageBytes);
        bulk_g2s(sBiasRing + slot * kBiasStageBytes, src, kBiasStageBytes, &b_full[slot]);
        src += kBiasStageBytes;
        ++gb;
      };
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        src = p.wstream;
        for (int step = 0; step < 9; ++step) {
          push_bias();
          const int n = (step == 0) ? 1 : (step == 5 ? 5 : 4);
          for (int i = 0; i < n; ++i) push(kStageBytes);
        }
        push_bias();                                          // view stage (N = 128, K = 32)
        for (int i = 0; i < 4; ++i) push(kStageBytes / 2);    // N = 128 layer: 4 stages of 16 KiB
      }
    }
  } else if (warp == kNerfMmaWarp) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc256 = make_idesc_f16(BF16, kTileM, 256);
      const uint32_t idesc128 = make_idesc_f16(BF16, kTileM, 128);
      const uint32_t aA = smem_u32(sA);
      const uint32_t aOnes = smem_u32(sOnes);
      const uint32_t aRing = smem_u32(sRing);
      const uint32_t aBiasRing = smem_u32(sBiasRing);
      uint32_t g = 0, gb = 0;
      uint32_t par_a = 0;   // parity of the a_ready phase the next layer waits for (all 4 groups in step)
      const bool prof = p.prof != nullptr;
      long long t_a = 0, t_w = 0, t_p = 0;
      const long long t_start = prof ? clock64() : 0;
      // wait for the next 8 KiB stage of the small ring (bias / view stage); returns its shared address
      auto next_b = [&]() -> uint32_t {
        const uint32_t slot = gb % kNerfBiasRing;
        const long long c0 = prof ? clock64() : 0;
        mbar_wait(&b_full[slot], (gb / kNerfBiasRing) & 1, p.dbg, 240 + slot);
        if (prof) t_w += clock64() - c0;
        tc_fence_after_sync();
        return aBiasRing + slot * kBiasStageBytes;
      };
      auto release_b = [&]() {
        umma_commit(&b_empty[gb % kNerfBiasRing]);
        ++gb;
      };
      auto bias_step = [&](uint32_t d_tmem) {
        const uint32_t b = next_b();
        issue_bias_stage(d_tmem, aOnes, b, 256 * 16, idesc256, true);
        release_b();
      };
      // one K=64 stage fed from the point block
      auto run_p = [&](uint32_t d_tmem, uint32_t aP) {
        const uint32_t slot = g % kNerfRing;
        const long long c0 = prof ? clock64() : 0;
        mbar_wait(&w_full[slot], (g / kNerfRing) & 1, p.dbg, 220 + slot);
        if (prof) t_w += clock64() - c0;
        tc_fence_after_sync();
        issue_stage<4>(d_tmem, aP, aRing + slot * kStageBytes, 256 * 16, idesc256, false);
        umma_commit(&w_empty[slot]);
        ++g;
      };
      // four K=64 stages fed from the activation buffer, chasing the previous layer's epilogue
      auto run4 = [&](uint32_t d_tmem, uint32_t idesc, uint32_t lbo_b) {
        for (int st = 0; st < 4; ++st) {
          const uint32_t slot = g % kNerfRing;
          const long long c0 = prof ? clock64() : 0;
          mbar_wait2(&a_ready[st], par_a, &w_full[slot], (g / kNerfRing) & 1, p.dbg, 210 + st);
          if (prof) t_a += clock64() - c0;
          tc_fence_after_sync();
          issue_stage<4>(d_tmem, aA + st * kGroupBytes, aRing + slot * kStageBytes, lbo_b, idesc, false);
          umma_commit(&w_empty[slot]);
          ++g;
        }
        par_a ^= 1u;
      };
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        const uint32_t buf = it & 1u;
        const uint32_t aP = smem_u32(sP) + buf * kPBlockBytes;
        const uint32_t aV = smem_u32(sV) + buf * kVBlockBytes;
        {
          const long long c0 = prof ? clock64() : 0;
          mbar_wait(&p_ready[buf], (it >> 1) & 1u, p.dbg, 200 + buf);
          if (prof) t_p += clock64() - c0;
          tc_fence_after_sync();
        }
        const uint32_t d0 = tmem_base, d1 = tmem_base + 256;
        // step 0
        bias_step(d0);
        run_p(d0, aP);
        umma_commit(&d_full[0]);
        // steps 1..8
        for (int step = 1; step <= 8; ++step) {
          const uint32_t d = (step & 1) ? d1 : d0;
          if (step == 1 && it > 0) {
            // D1 was last read by the previous tile's view-branch epilogue, which signals nothing through
            // a_ready: wait until all epilogue warps have drained it before the overwriting bias step
            mbar_wait(drained, (it - 1) & 1u, p.dbg, 250);
            tc_fence_after_sync();
          }
          bias_step(d);
          if (step == 5) run_p(d, aP);
          run4(d, idesc256, 256 * 16);
          umma_commit(&d_full[step & 1]);
        }
        // step 9: view branch, N = 128
        {
          const uint32_t b = next_b();
          issue_stage<2>(d1, aV, b, 128 * 16, idesc128, true);
          release_b();
          run4(d1, idesc128, 128 * 16);
          umma_commit(&d_full[1]);
          umma_commit(&p_free[buf]);
        }
      }
      if (prof) {
        long long* o = p.prof + blockIdx.x * 8;
        o[0] = clock64() - t_start;   // MMA thread: total
        o[1] = t_a;                   // waiting for A groups + weight stages (joint wait)
        o[2] = t_w;                   // waiting for bias / point-block weight stages
        o[7] = t_p;                   // waiting for the encoder
      }
    }
  } else if (warp >= 8) {
    // ===================== encoder warpgroup (one tile ahead of the MMAs) =====================
    const int row = (warp & 3) * 32 + lane;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1u;
      long long gr = static_cast<long long>(tile) * kTileM + row;
      if (gr >= p.n_rows) gr = p.n_rows - 1;
      float px, py, pz, vx, vy, vz;
      const float* x = nullptr;
      if (p.embedded != nullptr) {
        x = p.embedded + gr * p.emb_stride;
        px = py = pz = vx = vy = vz = 0.f;
      } else {
        const long long ray = gr / p.S;
        const float z = __ldg(p.z_vals + gr);
        const float* o = p.rays_o + ray * p.o_stride;
        const float* d = p.rays_d + ray * p.d_stride;
        const float* vd = p.viewdirs + ray * p.v_stride;
        px = __fadd_rn(__ldg(o + 0), __fmul_rn(__ldg(d + 0), z));
        py = __fadd_rn(__ldg(o + 1), __fmul_rn(__ldg(d + 1), z));
        pz = __fadd_rn(__ldg(o + 2), __fmul_rn(__ldg(d + 2), z));
        vx = __ldg(vd + 0);
        vy = __ldg(vd + 1);
        vz = __ldg(vd + 2);
      }
      if (it >= 2) mbar_wait(&p_free[buf], ((it >> 1) - 1) & 1u, p.dbg, 500 + buf);
      uint8_t* const dP = sP + buf * kPBlockBytes;
      uint8_t* const dV = sV + buf * kVBlockBytes;
      float v[32];
      if (x != nullptr) {
        // API path: the caller already embedded the points (63 features) and view directions (27 features)
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) e[i] = (ch * 8 + i < 63) ? __ldg(x + ch * 8 + i) : 0.0f;
          uint4 q;
          q.x = pack2<BF16>(e[0], e[1]);
          q.y = pack2<BF16>(e[2], e[3]);
          q.z = pack2<BF16>(e[4], e[5]);
          q.w = pack2<BF16>(e[6], e[7]);
          *reinterpret_cast<uint4*>(dP + ch * kChunkBytes + row * 16) = q;
        }
#pragma unroll
        for (int i = 0; i < 27; ++i) v[i] = __ldg(x + 63 + i);
      } else {
        encode_point_block<BF16>(dP, row, px, py, pz);
        v[0] = vx;
        v[1] = vy;
        v[2] = vz;
        float s[4], c[4];
        sincos_octaves<4>(vx, s, c);
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          v[3 + 6 * f + 0] = s[f];
          v[3 + 6 * f + 3] = c[f];
        }
        sincos_octaves<4>(vy, s, c);
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          v[3 + 6 * f + 1] = s[f];
          v[3 + 6 * f + 4] = c[f];
        }
        sincos_octaves<4>(vz, s, c);
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          v[3 + 6 * f + 2] = s[f];
          v[3 + 6 * f + 5] = c[f];
        }
      }
      v[27] = 1.0f;   // x bias_hi
      v[28] = 1.0f;   // x bias_lo
      v[29] = v[30] = v[31] = 0.0f;
      store_block32<BF16>(dV, row, v);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[buf]);
    }
  } else {
    // ===================== epilogue warpgroups =====================
    const int wg = warp >> 2;                       // owns the 32-column pieces wg, wg+2, wg+4, wg+6
    const int row = (warp & 3) * 32 + lane;         // tile row == TMEM lane
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint8_t* const a_row = sA + row * 16;
    const bool prof = p.prof != nullptr && (threadIdx.x & 127) == 0;
    long long t_d = 0;
    const long long t_start = prof ? clock64() : 0;
    uint32_t par_d = 0;   // bit dbuf: parity of the next d_full phase
    auto wait_d = [&](int db, uint32_t id) {
      const long long cd = prof ? clock64() : 0;
      mbar_wait(&d_full[db], (par_d >> db) & 1u, p.dbg, id);
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
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const long long g_row = static_cast<long long>(tile) * kTileM + row;
      const bool valid = g_row < p.n_rows;
      float sigma_part = 0.0f;
      float abs_part = 0.0f;   // sum |w_a| relu(h7): scale of the far-sample guard band (nerf_far.cu)
      for (int step = 0; step <= 8; ++step) {
        const int db = step & 1;
        wait_d(db, 300 + step);
        auto signal = [&](int g) { warp_arrive(&a_ready[g], lane); };
        if (step == 7) {
          for_pieces(
              db * 256,
              [&](uint32_t col0, uint32_t (&v)[32]) {
                store_sub<BF16, true>(v, a_row + (col0 >> 5) * kSubBytes);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  const float w = sAlphaW[col0 + i], h = relu_nan(__uint_as_float(v[i]));
                  sigma_part = fmaf(w, h, sigma_part);
                  if ((i & 3) == 0) abs_part = fmaf(fabsf(w), h, abs_part);   // every 4th term: a scale estimate
                }
              },
              signal);
        } else if (step == 8) {
          for_pieces(
              db * 256,
              [&](uint32_t col0, uint32_t (&v)[32]) { store_sub<BF16, false>(v, a_row + (col0 >> 5) * kSubBytes); }, signal);
        } else {
          for_pieces(
              db * 256,
              [&](uint32_t col0, uint32_t (&v)[32]) { store_sub<BF16, true>(v, a_row + (col0 >> 5) * kSubBytes); }, signal);
        }
      }
      // step 9: view branch (N = 128 -> D1 columns [256,384)); this WG owns the 64 columns [64*wg, 64*wg+64)
      wait_d(1, 309);
      float r = 0.f, gch = 0.f, b = 0.f;
      {
        uint32_t va[32], vb[32];
        tmem_ld32(lane_taddr + 256 + 64 * wg, va);
        tmem_ld32(lane_taddr + 256 + 64 * wg + 32, vb);
        tmem_ld_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(drained);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int n = 64 * wg + i;
          const float x = relu_nan(__uint_as_float(va[i]));
          r = fmaf(sRgbW[n], x, r);
          gch = fmaf(sRgbW[128 + n], x, gch);
          b = fmaf(sRgbW[256 + n], x, b);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int n = 64 * wg + 32 + i;
          const float x = relu_nan(__uint_as_float(vb[i]));
          r = fmaf(sRgbW[n], x, r);
          gch = fmaf(sRgbW[128 + n], x, gch);
          b = fmaf(sRgbW[256 + n], x, b);
        }
      }
      if (wg == 1) {
        sPart[row] = make_float4(r, gch, b, sigma_part);
        sAbs[row] = abs_part;
        named_bar_arrive(1, 256);
        named_bar_sync(2, 256);   // WG0 has consumed sPart
      } else {
        named_bar_sync(1, 256);
        const float4 o1 = sPart[row];
        const float a1 = sAbs[row];
        named_bar_arrive(2, 256);
        if (valid) {
          float4 o;
          o.x = r + o1.x + p.rgb_b[0];
          o.y = gch + o1.y + p.rgb_b[1];
          o.z = b + o1.z + p.rgb_b[2];
          o.w = sigma_part + o1.w + p.alpha_b;
          reinterpret_cast<float4*>(p.raw)[g_row] = o;
          note_nonfinite(p.dbg, o.x + o.y + o.z + o.w, g_row);
          nerf_far_flag(p, g_row, g_row / p.S, o.w, 4.0f * (abs_part + a1));
        }
      }
    }
    if (prof) {
      long long* o = p.prof + blockIdx.x * 8 + 3 + wg * 2;
      o[0] = t_d;                          // WG: waiting for accumulators
      o[1] = (clock64() - t_start) - t_d;  // WG: epilogue work (everything else)
    }
  }

  // ---- teardown ----
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kNerfMmaWarp) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

template <bool BF16>
int launch_nerf(const NerfParams& p, int grid, cudaStream_t st) {
  R2L_CUDA(cudaFuncSetAttribute(nerf_mlp_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kNerfSmemBytes));
  nerf_mlp_kernel<BF16><<<grid, kNerfThreads, kNerfSmemBytes, st>>>(p);
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

int nerf_mlp_launch(bool bf16, const NerfParams& p, int grid, cudaStream_t st) {
  return bf16 ? launch_nerf<true>(p, grid, st) : launch_nerf<false>(p, grid, st);
}

}  // namespace r2l
