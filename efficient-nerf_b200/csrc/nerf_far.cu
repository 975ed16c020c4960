// Far-sample sigma fix-up of the fused NeRF path (main.py:578-581, 598).
//
// raw2outputs gives the LAST sample of a ray the interval 1e10 (main.py:579-581), so its alpha = 1 - exp(-relu(sigma) *
// 1e10 * |d|) is a step function of sign(sigma_far): 0 for sigma <= 0, 1 for any positive sigma.  The tensor-core
// kernels compute sigma from 16-bit-operand layers (|error| ~ 3e-5 on random-init nets), which flips that sign on a
// handful of rays per frame — each flip moves rgb_map by up to ~0.4 (white background), far outside the 2e-3 gate.
//
// So the fused kernels FLAG every ray whose far-sample sigma is inside a guard band,
//     |sigma| < max(far_abs, far_rel * sum_i |alpha_w_i| relu(h7_i))
// (the second term scales the band with the magnitude of what was summed: the rounding error of sigma grows with it),
// append the ray to a compacted list (atomicAdd; a few hundred rays per 160 000-ray frame), and this kernel re-evaluates
// ONLY those points through the network's eight point layers + alpha_linear in fp32 and patches raw[ray, S-1, 3] before
// compositing.  The arithmetic is the `precision="fp32"` path's, operation for operation — point = o + d*z with
// separately rounded mul/add (rays.cu), sincosf(x * 2^f) (embed.cu), acc = fmaf(x_k, w_k, acc) for k ascending in the
// reference's K order then + bias (linear_fp32.cu) — so the patched sigma is BIT-IDENTICAL to that path's (tested),
// which is itself within 1e-5 of the torch-CPU reference.
//
// Work per flagged point: 0.98 MFLOP fp32; weights (transposed fp32 copy, 1.96 MB per net) stream from L2 through a
// cp.async ring.  One block = 256 threads works on 32 flagged rays at a time (4 neurons x 8 rays per thread), grid =
// one block per SM looping over the list.
#include "common.cuh"
#include "mlp_params.cuh"

namespace r2l {

constexpr int kFarG = 8;   // flagged rays per block iteration

// Wt layout (floats): for layer l, rows k = 0..K_l-1 of 256 outputs each, layers concatenated:
//   K = 63, 256, 256, 256, 256, 319 (reference order: cat[pts(63), h(256)]), 256, 256  -> 1918 rows
// then biases [8][256], then alpha_w [256].
__host__ __device__ constexpr int far_layer_k(int l) { return l == 0 ? 63 : (l == 5 ? 319 : 256); }
__host__ __device__ constexpr int far_layer_row0(int l) {
  int r = 0;
  for (int i = 0; i < l; ++i) r += far_layer_k(i);
  return r;
}
constexpr int kFarRows = 63 + 4 * 256 + 319 + 2 * 256;   // 1918
constexpr int kFarBiasOff = kFarRows * 256;
constexpr int kFarAlphaOff = kFarBiasOff + 8 * 256;
constexpr int kFarFloats = kFarAlphaOff + 256;

// W [256, K] row-major -> Wt rows [K][256]
__global__ void far_transpose_kernel(const float* __restrict__ W, int K, float* __restrict__ Wt) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 256 * K) return;
  const int k = idx / 256, n = idx % 256;
  Wt[idx] = W[n * K + k];
}

// The whole net is one sequence of 1918 weight rows of 1 KiB, streamed through a 4-deep cp.async ring of 16-row chunks
// that runs ahead ACROSS layer boundaries (with a handful of flagged rays per launch nothing else hides the L2 latency:
// a first version with loads issued from the FMA loop cost 0.4 ms for 30 rays).  A block works on 32 flagged rays at
// a time: thread (s, j) = (tid / 64, tid % 64) owns output neurons 4j..4j+3 of the 8 rays of sub-batch s — per weight
// row one LDS.128 of weights and two broadcast LDS.128 of activations feed 32 FMAs, and the ring is shared by all 32
// rays.  (The 8-rays-per-block version was LDS-bound at 3 loads per 8 FMAs: 300 us for the ~2600 rays the coarse
// pass of a frame flags; ncu r2d.)  Every output still accumulates k = 0, 1, 2, ... in order: bit-identical to the
// fp32 path.
constexpr int kFarChunk = 16;                       // rows per chunk (16 KiB)
constexpr int kFarStages = 4;
constexpr int kFarSub = 4;                          // sub-batches of kFarG rays per block
constexpr int kFarBatch = kFarSub * kFarG;          // 32 rays per block iteration
constexpr int kFarRingFloats = kFarStages * kFarChunk * 256;
constexpr int kFarEmbFloats = kFarSub * 64 * kFarG;
constexpr int kFarHFloats = 2 * kFarSub * 256 * kFarG;
constexpr int kFarSmemBytes = (kFarRingFloats + kFarEmbFloats + kFarHFloats) * 4;

__device__ __forceinline__ void far_cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst))), "l"(src)
               : "memory");
}

__global__ void __launch_bounds__(256, 1)
nerf_far_fixup_kernel(const float* __restrict__ Wt, float alpha_b, const int* __restrict__ list,
                      const int* __restrict__ count, int cap, int* __restrict__ stats, const float* __restrict__ rays_o,
                      long long o_stride, const float* __restrict__ rays_d, long long d_stride,
                      const float* __restrict__ z_vals, int S, float* __restrict__ out, int compact) {
  extern __shared__ __align__(16) float s_far[];
  float* const s_ring = s_far;                              // [kFarStages][kFarChunk][256]
  float* const s_emb = s_ring + kFarRingFloats;             // [sub][64][kFarG]: embedded point, reference order
  float* const s_h = s_emb + kFarEmbFloats;                 // [2][sub][256][kFarG]: hidden activations, ping-pong
  __shared__ float s_pt[kFarBatch][3];
  int n = *count;
  if (n > cap) n = cap;
  if (blockIdx.x == 0 && threadIdx.x == 0) stats[0] = *count;   // flagged by the last forward (may exceed cap)
  const int tid = threadIdx.x;
  const int sub = tid >> 6, j4 = (tid & 63) * 4;              // sub-batch, first of this thread's 4 neurons
  constexpr int n_chunks = (kFarRows + kFarChunk - 1) / kFarChunk;   // 120 (the last one holds 14 rows)
  auto issue = [&](int c) {   // chunk c -> ring slot c % kFarStages (rows past the end read the bias block: harmless)
    if (c < n_chunks) {
      float* dst = s_ring + (c % kFarStages) * kFarChunk * 256;
      const float* src = Wt + static_cast<size_t>(c) * kFarChunk * 256;
#pragma unroll
      for (int i = 0; i < 4; ++i) far_cp16(dst + (tid + 256 * i) * 4, src + (tid + 256 * i) * 4);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int g0 = blockIdx.x * kFarBatch; g0 < n; g0 += gridDim.x * kFarBatch) {
    const int ng = min(kFarBatch, n - g0);
    __syncthreads();
    for (int c = 0; c < kFarStages - 1; ++c) issue(c);
    if (tid < kFarBatch * 3) {
      const int g = tid / 3, c = tid % 3;
      float v = 0.0f;
      if (g < ng) {
        const long long ray = list[g0 + g];
        const float z = z_vals[ray * S + (S - 1)];
        v = __fadd_rn(rays_o[ray * o_stride + c], __fmul_rn(rays_d[ray * d_stride + c], z));   // main.py:701
      }
      s_pt[g][c] = v;
      s_emb[((g / kFarG) * 64 + c) * kFarG + (g % kFarG)] = v;
    }
    __syncthreads();
    for (int it = tid; it < kFarBatch * 30; it += 256) {   // (ray, coord, freq): Embedder order 3 + 6 f + {c, 3 + c}
      const int g = it / 30, r = it % 30, c = r / 10, f = r % 10;
      float sn, co;
      sincosf(s_pt[g][c] * exp2f(static_cast<float>(f)), &sn, &co);
      float* e = s_emb + (g / kFarG) * 64 * kFarG + (g % kFarG);
      e[(3 + 6 * f + c) * kFarG] = sn;
      e[(3 + 6 * f + 3 + c) * kFarG] = co;
    }
    // rows are consumed in order; `l` / `k` = layer and input index of the current row
    int l = 0, k = 0, cur = 0;
    float acc[4][kFarG];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int g = 0; g < kFarG; ++g) acc[q][g] = 0.0f;
    const float* const emb_s = s_emb + sub * 64 * kFarG;
    for (int c = 0; c < n_chunks; ++c) {
      asm volatile("cp.async.wait_group %0;" ::"n"(kFarStages - 2) : "memory");   // chunk c has landed (this thread's part)
      __syncthreads();            // ... everybody's part, and everybody is done with chunk c - 1 (and with s_emb / s_h writes)
      issue(c + kFarStages - 1);  // into the slot chunk c - 1 occupied
      const float* wrow = s_ring + (c % kFarStages) * kFarChunk * 256 + j4;
      const int rows = min(kFarChunk, kFarRows - c * kFarChunk);
      for (int r = 0; r < rows;) {
        // input of this row: layer 0 and the first 63 rows of the skip layer read the embedding (cat[input_pts, h])
        const bool from_emb = (l == 0) || (l == 5 && k < 63);
        const float* x = from_emb ? emb_s + k * kFarG : s_h + ((cur * kFarSub + sub) * 256 + (l == 5 ? k - 63 : k)) * kFarG;
        const int k_end = (l == 5 && k < 63) ? 63 : far_layer_k(l);   // rows left before the input source changes
        auto fma_row = [&](const float4& w4, const float4& xa, const float4& xb) {
          const float wq[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            acc[q][0] = fmaf(xa.x, wq[q], acc[q][0]);
            acc[q][1] = fmaf(xa.y, wq[q], acc[q][1]);
            acc[q][2] = fmaf(xa.z, wq[q], acc[q][2]);
            acc[q][3] = fmaf(xa.w, wq[q], acc[q][3]);
            acc[q][4] = fmaf(xb.x, wq[q], acc[q][4]);
            acc[q][5] = fmaf(xb.y, wq[q], acc[q][5]);
            acc[q][6] = fmaf(xb.z, wq[q], acc[q][6]);
            acc[q][7] = fmaf(xb.w, wq[q], acc[q][7]);
          }
        };
        if (sub * kFarG >= ng) {
          // this sub-batch holds no ray of the (last, partial) batch: skip the arithmetic, keep the control flow
          const int adv = min(min(4, rows - r), k_end - k);
          r += adv, k += adv;
          if (k != far_layer_k(l)) continue;
          --k;
        } else if (r + 4 <= rows && k + 4 <= k_end) {
          // four rows at a time: all twelve shared-memory loads are issued before the first FMA (same FMA order per
          // accumulator), otherwise every row exposes an LDS round trip with two warps per scheduler
          float4 w4[4], xa[4], xb[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            w4[j] = *reinterpret_cast<const float4*>(wrow + (r + j) * 256);
            xa[j] = *reinterpret_cast<const float4*>(x + j * kFarG);
            xb[j] = *reinterpret_cast<const float4*>(x + j * kFarG + 4);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) fma_row(w4[j], xa[j], xb[j]);
          r += 4, k += 4;
          if (k != far_layer_k(l)) continue;
          --k;   // layer complete: fall through to the hand-over below (which increments k again)
        } else {
          fma_row(*reinterpret_cast<const float4*>(wrow + r * 256), *reinterpret_cast<const float4*>(x),
                  *reinterpret_cast<const float4*>(x + 4));
          ++r;
        }
        if (++k == far_layer_k(l)) {   // layer complete: bias, relu, hand the activations to the next layer
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(Wt + kFarBiasOff + l * 256 + j4));
          const float bq[4] = {b4.x, b4.y, b4.z, b4.w};
          const int nxt = (l == 0) ? 0 : (cur ^ 1);
          __syncthreads();             // every thread has finished reading s_h[cur] ... (uniform branch)
          float* out = s_h + ((nxt * kFarSub + sub) * 256 + j4) * kFarG;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 o0, o1;
            o0.x = fmaxf(acc[q][0] + bq[q], 0.0f), o0.y = fmaxf(acc[q][1] + bq[q], 0.0f);
            o0.z = fmaxf(acc[q][2] + bq[q], 0.0f), o0.w = fmaxf(acc[q][3] + bq[q], 0.0f);
            o1.x = fmaxf(acc[q][4] + bq[q], 0.0f), o1.y = fmaxf(acc[q][5] + bq[q], 0.0f);
            o1.z = fmaxf(acc[q][6] + bq[q], 0.0f), o1.w = fmaxf(acc[q][7] + bq[q], 0.0f);
            *reinterpret_cast<float4*>(out + q * kFarG) = o0;
            *reinterpret_cast<float4*>(out + q * kFarG + 4) = o1;
#pragma unroll
            for (int g = 0; g < kFarG; ++g) acc[q][g] = 0.0f;
          }
          cur = nxt;
          ++l, k = 0;
          __syncthreads();             // ... and sees the new layer's input
        }
      }
    }
    if (tid < ng) {   // alpha_linear: sequential k, like the fp32 path's K loop
      float a = 0.0f;
      const float* aw = Wt + kFarAlphaOff;
      const float* h = s_h + ((cur * kFarSub + tid / kFarG) * 256) * kFarG + (tid % kFarG);
      for (int kk = 0; kk < 256; ++kk) a = fmaf(h[kk * kFarG], __ldg(aw + kk), a);
      // row of the patched sigma: the ray's own row of raw, or (fused compositing) row `list position` of the compact copy
      const long long orow = compact ? static_cast<long long>(g0 + tid) : static_cast<long long>(list[g0 + tid]);
      out[(orow * S + (S - 1)) * 4 + 3] = a + alpha_b;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
}

int nerf_far_pack(const float* const* pts_w, const float* const* pts_b, const float* alpha_w, float* Wt,
                  cudaStream_t st) {
  for (int l = 0; l < 8; ++l) {
    const int K = far_layer_k(l);
    far_transpose_kernel<<<(256 * K + 255) / 256, 256, 0, st>>>(pts_w[l], K, Wt + static_cast<size_t>(far_layer_row0(l)) * 256);
    R2L_LAUNCH_CHECK();
    R2L_CUDA(cudaMemcpyAsync(Wt + kFarBiasOff + l * 256, pts_b[l], 256 * 4, cudaMemcpyDeviceToDevice, st));
  }
  R2L_CUDA(cudaMemcpyAsync(Wt + kFarAlphaOff, alpha_w, 256 * 4, cudaMemcpyDeviceToDevice, st));
  return R2L_OK;
}

size_t nerf_far_weight_bytes() { return sizeof(float) * kFarFloats; }

int nerf_far_fixup_launch(const float* Wt, float alpha_b, const int* list, const int* count, int cap, int* stats,
                          const NerfParams& p, float* out, int compact, cudaStream_t st) {
  const int grid = sm_count();
  static bool attr_set = false;
  if (!attr_set) {
    R2L_CUDA(cudaFuncSetAttribute(nerf_far_fixup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFarSmemBytes));
    attr_set = true;
  }
  nerf_far_fixup_kernel<<<grid, 256, kFarSmemBytes, st>>>(Wt, alpha_b, list, count, cap, stats, p.rays_o, p.o_stride, p.rays_d,
                                             p.d_stride, p.z_vals, p.S, out, compact);
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

}  // namespace r2l
