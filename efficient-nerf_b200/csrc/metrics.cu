// Image metrics of the render_path caller, on the device: once a frame renders in
// milliseconds, the reference's per-frame PSNR / SSIM stage (main.py:330-335, 384-391) is what is left.
//
//   r2l_image_error : abs error map |a-b| (main.py:331) + per-image sum of squared differences in double
//                     (img2mse = mean((x-y)^2), helpers:19; mse2psnr helpers:20) in ONE pass over the two images.
//   r2l_ssim        : utils/ssim_torch.py:27-54 as called through main.py:46 — per channel, zero-padded 11x11 Gaussian
//                     (sigma 1.5) filtering of x, y, x^2, y^2, xy, the SSIM map, and its sum per image.
//                     The Gaussian is separable: a block stages a (32+10)x(32+10) tile of both images in shared
//                     memory, filters horizontally into five [42][32] maps, then vertically per output pixel.
//                     (The reference convolves with the fp32 outer product of the same 11 taps; the separable form
//                     differs by fp32 summation order only: ~1e-7 on the SSIM value.)
// Images are the renderer's layout: [n_img][H][W][3] fp32 (HWC), contiguous.  Both kernels are HBM-bound:
// 24 B/pixel in (+12 B out for the error map).
#include "common.cuh"

namespace r2l {

constexpr int kErrThreads = 256;

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, o);
    hi = __shfl_xor_sync(0xffffffffu, hi, o);
    v += __hiloint2double(hi, lo);
  }
  return v;
}

// grid.y = image; grid.x strides over the image's floats
__global__ void __launch_bounds__(kErrThreads)
image_error_kernel(long long n_per_img, const float* __restrict__ a, const float* __restrict__ b,
                   float* __restrict__ abs_err, double* __restrict__ sum_sq) {
  const long long base = static_cast<long long>(blockIdx.y) * n_per_img;
  double acc = 0.0;
  for (long long i = static_cast<long long>(blockIdx.x) * kErrThreads + threadIdx.x; i < n_per_img;
       i += static_cast<long long>(gridDim.x) * kErrThreads) {
    const float d = __fsub_rn(__ldg(a + base + i), __ldg(b + base + i));
    if (abs_err != nullptr) abs_err[base + i] = fabsf(d);
    acc += static_cast<double>(__fmul_rn(d, d));
  }
  __shared__ double s_part[kErrThreads / 32];
  acc = warp_sum_f64(acc);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < kErrThreads / 32) ? s_part[threadIdx.x] : 0.0;
    v = warp_sum_f64(v);
    if (threadIdx.x == 0) atomicAdd(sum_sq + blockIdx.y, v);
  }
}

constexpr int kSsimTile = 32;
constexpr int kSsimR = 5;                       // 11 taps
constexpr int kSsimIn = kSsimTile + 2 * kSsimR;  // 42

struct SsimTaps {
  float w[11];
};

// grid = (tiles_x, tiles_y, n_img * 3); block = 32 x 8
__global__ void __launch_bounds__(256)
ssim_kernel(int H, int W, const float* __restrict__ a, const float* __restrict__ b, long long img_stride,
            const SsimTaps taps, double* __restrict__ sum_out) {
  __shared__ float s_a[kSsimIn][kSsimIn + 1];
  __shared__ float s_b[kSsimIn][kSsimIn + 1];
  __shared__ float s_h[5][kSsimIn][kSsimTile + 1];
  __shared__ double s_part[8];
  const int img = blockIdx.z / 3, ch = blockIdx.z % 3;
  const float* ia = a + static_cast<long long>(img) * img_stride + ch;
  const float* ib = b + static_cast<long long>(img) * img_stride + ch;
  const int x0 = blockIdx.x * kSsimTile - kSsimR, y0 = blockIdx.y * kSsimTile - kSsimR;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  for (int i = tid; i < kSsimIn * kSsimIn; i += 256) {
    const int r = i / kSsimIn, c = i % kSsimIn;
    const int y = y0 + r, x = x0 + c;
    float va = 0.0f, vb = 0.0f;   // zero padding (F.conv2d padding=5)
    if (y >= 0 && y < H && x >= 0 && x < W) {
      const long long off = (static_cast<long long>(y) * W + x) * 3;
      va = __ldg(ia + off);
      vb = __ldg(ib + off);
    }
    s_a[r][c] = va;
    s_b[r][c] = vb;
  }
  __syncthreads();
  // horizontal pass: 42 rows x 32 columns, five quantities
  for (int i = tid; i < kSsimIn * kSsimTile; i += 256) {
    const int r = i / kSsimTile, c = i % kSsimTile;
    float h1 = 0.f, h2 = 0.f, h11 = 0.f, h22 = 0.f, h12 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float w = taps.w[k], pa = s_a[r][c + k], pb = s_b[r][c + k];
      h1 = fmaf(w, pa, h1);
      h2 = fmaf(w, pb, h2);
      h11 = fmaf(w, pa * pa, h11);
      h22 = fmaf(w, pb * pb, h22);
      h12 = fmaf(w, pa * pb, h12);
    }
    s_h[0][r][c] = h1;
    s_h[1][r][c] = h2;
    s_h[2][r][c] = h11;
    s_h[3][r][c] = h22;
    s_h[4][r][c] = h12;
  }
  __syncthreads();
  // vertical pass + SSIM map: thread (x, y) handles rows y, y+8, y+16, y+24 of column x
  double acc = 0.0;
  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
  for (int r = threadIdx.y; r < kSsimTile; r += 8) {
    const int c = threadIdx.x;
    const int y = blockIdx.y * kSsimTile + r, x = blockIdx.x * kSsimTile + c;
    if (y >= H || x >= W) continue;
    float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float w = taps.w[k];
      m1 = fmaf(w, s_h[0][r + k][c], m1);
      m2 = fmaf(w, s_h[1][r + k][c], m2);
      e11 = fmaf(w, s_h[2][r + k][c], e11);
      e22 = fmaf(w, s_h[3][r + k][c], e22);
      e12 = fmaf(w, s_h[4][r + k][c], e12);
    }
    const float m1s = m1 * m1, m2s = m2 * m2, m12 = m1 * m2;
    const float v1 = e11 - m1s, v2 = e22 - m2s, v12 = e12 - m12;
    const float num = (2.0f * m12 + C1) * (2.0f * v12 + C2);
    const float den = (m1s + m2s + C1) * (v1 + v2 + C2);
    acc += static_cast<double>(__fdiv_rn(num, den));
  }
  acc = warp_sum_f64(acc);
  if (threadIdx.x == 0) s_part[threadIdx.y] = acc;
  __syncthreads();
  if (tid == 0) {
    double v = 0.0;
    for (int i = 0; i < 8; ++i) v += s_part[i];
    atomicAdd(sum_out + img, v);
  }
}

}  // namespace r2l

using namespace r2l;

extern "C" {

// abs_err (optional) [n_img][n_per_img] = |a - b|;  sum_sq [n_img] (double, ZEROED by this call) = sum (a-b)^2.
// MSE of image i = sum_sq[i] / n_per_img (helpers:19); PSNR = -10 log10(MSE) (helpers:20).
int r2l_image_error(int n_img, long long n_per_img, const float* a, const float* b, float* abs_err, double* sum_sq,
                    void* stream) {
  R2L_CHECK_ARG(n_img >= 0 && n_per_img >= 0, "r2l_image_error: bad sizes");
  if (n_img == 0) return R2L_OK;
  R2L_CHECK_ARG(n_img <= 65535, "r2l_image_error: more than 65535 images per call");
  R2L_CHECK_ARG(a && b && sum_sq, "r2l_image_error: null pointer");
  auto st = static_cast<cudaStream_t>(stream);
  R2L_CUDA(cudaMemsetAsync(sum_sq, 0, sizeof(double) * n_img, st));
  if (n_per_img == 0) return R2L_OK;
  long long bx = (n_per_img + kErrThreads - 1) / kErrThreads;
  const long long cap = (static_cast<long long>(sm_count()) * 8 + n_img - 1) / n_img;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  image_error_kernel<<<dim3(static_cast<unsigned>(bx), static_cast<unsigned>(n_img)), kErrThreads, 0, st>>>(
      n_per_img, a, b, abs_err, sum_sq);
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

// a, b: [n_img][H][W][3] fp32 (image i at a + i*img_stride floats); taps: the 11 normalised Gaussian taps
// (utils/ssim_torch.py:10-16, computed by the caller exactly as the reference does); sum_out [n_img] (double, ZEROED
// by this call) = sum of the SSIM map over the 3 channels; SSIM of image i = sum_out[i] / (3 H W).
int r2l_ssim(int n_img, int H, int W, const float* a, const float* b, long long img_stride, const float* taps,
             double* sum_out, void* stream) {
  R2L_CHECK_ARG(n_img >= 0 && H > 0 && W > 0, "r2l_ssim: bad sizes");
  if (n_img == 0) return R2L_OK;
  R2L_CHECK_ARG(n_img * 3 <= 65535, "r2l_ssim: more than 21845 images per call");
  R2L_CHECK_ARG(a && b && taps && sum_out, "r2l_ssim: null pointer");
  R2L_CHECK_ARG(img_stride >= static_cast<long long>(H) * W * 3, "r2l_ssim: bad image stride");
  auto st = static_cast<cudaStream_t>(stream);
  SsimTaps t;
  for (int i = 0; i < 11; ++i) t.w[i] = taps[i];   // host pointer: 11 floats
  R2L_CUDA(cudaMemsetAsync(sum_out, 0, sizeof(double) * n_img, st));
  const dim3 grid((W + kSsimTile - 1) / kSsimTile, (H + kSsimTile - 1) / kSsimTile, n_img * 3);
  ssim_kernel<<<grid, dim3(32, 8), 0, st>>>(H, W, a, b, img_stride, t, sum_out);
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

}  // extern "C"
