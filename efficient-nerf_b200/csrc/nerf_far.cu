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
// Work per flagged point: 0.98 MFLOP fp32; weights (transposed fp32 copy, 1.96 MB per net) stream from L2.  One block =
// 256 threads = the 256 output neurons of a layer, 8 flagged rays at a time (8 accumulators per thread, the group's
// activations broadcast from shared memory), grid = 2 blocks per SM looping over the list.
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

__global__ void __launch_bounds__(256)
nerf_far_fixup_kernel(const float* __restrict__ Wt, float alpha_b, const int* __restrict__ list,
                      const int* __restrict__ count, int cap, int* __restrict__ stats, const float* __restrict__ rays_o,
                      long long o_stride, const float* __restrict__ rays_d, long long d_stride,
                      const float* __restrict__ z_vals, int S, float* __restrict__ raw) {
  __shared__ __align__(16) float s_emb[64][kFarG];    // embedded point, reference order (63 used)
  __shared__ __align__(16) float s_h[2][256][kFarG];  // hidden activations, ping-pong
  __shared__ float s_pt[kFarG][3];
  int n = *count;
  if (n > cap) n = cap;
  if (blockIdx.x == 0 && threadIdx.x == 0) stats[0] = *count;   // flagged by the last forward (may exceed cap)
  const int tid = threadIdx.x;
  for (int g0 = blockIdx.x * kFarG; g0 < n; g0 += gridDim.x * kFarG) {
    const int ng = min(kFarG, n - g0);
    __syncthreads();
    if (tid < kFarG * 3) {
      const int g = tid / 3, c = tid % 3;
      float v = 0.0f;
      if (g < ng) {
        const long long ray = list[g0 + g];
        const float z = z_vals[ray * S + (S - 1)];
        v = __fadd_rn(rays_o[ray * o_stride + c], __fmul_rn(rays_d[ray * d_stride + c], z));   // main.py:701
      }
      s_pt[g][c] = v;
      s_emb[c][g] = v;
    }
    __syncthreads();
    if (tid < kFarG * 30) {   // (ray, coord, freq): Embedder order 3 + 6 f + {c, 3 + c}  (helpers:24-56)
      const int g = tid / 30, r = tid % 30, c = r / 10, f = r % 10;
      float s, co;
      sincosf(s_pt[g][c] * exp2f(static_cast<float>(f)), &s, &co);
      s_emb[3 + 6 * f + c][g] = s;
      s_emb[3 + 6 * f + 3 + c][g] = co;
    }
    __syncthreads();
    int cur = 0;
    for (int l = 0; l < 8; ++l) {
      float acc[kFarG];
#pragma unroll
      for (int g = 0; g < kFarG; ++g) acc[g] = 0.0f;
      const float* w = Wt + static_cast<size_t>(far_layer_row0(l)) * 256 + tid;
      auto run = [&](const float (*x)[kFarG], int K) {
        int k = 0;
        for (; k + 4 <= K; k += 4) {   // 4 weight loads in flight
          const float w0 = __ldg(w + (k + 0) * 256), w1 = __ldg(w + (k + 1) * 256);
          const float w2 = __ldg(w + (k + 2) * 256), w3 = __ldg(w + (k + 3) * 256);
          const float wk[4] = {w0, w1, w2, w3};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 xa = *reinterpret_cast<const float4*>(&x[k + j][0]);
            const float4 xb = *reinterpret_cast<const float4*>(&x[k + j][4]);
            acc[0] = fmaf(xa.x, wk[j], acc[0]);
            acc[1] = fmaf(xa.y, wk[j], acc[1]);
            acc[2] = fmaf(xa.z, wk[j], acc[2]);
            acc[3] = fmaf(xa.w, wk[j], acc[3]);
            acc[4] = fmaf(xb.x, wk[j], acc[4]);
            acc[5] = fmaf(xb.y, wk[j], acc[5]);
            acc[6] = fmaf(xb.z, wk[j], acc[6]);
            acc[7] = fmaf(xb.w, wk[j], acc[7]);
          }
        }
        for (; k < K; ++k) {
          const float wv = __ldg(w + k * 256);
#pragma unroll
          for (int g = 0; g < kFarG; ++g) acc[g] = fmaf(x[k][g], wv, acc[g]);
        }
        w += static_cast<size_t>(K) * 256;
      };
      if (l == 0) {
        run(s_emb, 63);
      } else if (l == 5) {   // skip layer: cat[input_pts, h]  (model/nerf_raybased.py:381-385)
        run(s_emb, 63);
        run(s_h[cur], 256);
      } else {
        run(s_h[cur], 256);
      }
      const float b = __ldg(Wt + kFarBiasOff + l * 256 + tid);
      float* out = &s_h[l == 0 ? 0 : (cur ^ 1)][tid][0];
#pragma unroll
      for (int g = 0; g < kFarG; ++g) out[g] = fmaxf(acc[g] + b, 0.0f);
      if (l > 0) cur ^= 1;
      __syncthreads();
    }
    if (tid < ng) {   // alpha_linear: sequential k, like the fp32 path's K loop
      float acc = 0.0f;
      const float* aw = Wt + kFarAlphaOff;
      for (int k = 0; k < 256; ++k) acc = fmaf(s_h[cur][k][tid], __ldg(aw + k), acc);
      const long long ray = list[g0 + tid];
      raw[(ray * S + (S - 1)) * 4 + 3] = acc + alpha_b;
    }
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
                          const NerfParams& p, cudaStream_t st) {
  const int grid = 2 * sm_count();
  nerf_far_fixup_kernel<<<grid, 256, 0, st>>>(Wt, alpha_b, list, count, cap, stats, p.rays_o, p.o_stride, p.rays_d,
                                             p.d_stride, p.z_vals, p.S, p.raw);
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

}  // namespace r2l
