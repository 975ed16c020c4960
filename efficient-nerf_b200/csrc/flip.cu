// FLIP image-difference metric of the render_path caller on the device (utils/flip_loss.py:69-140,
// called from main.py:370-379 on the [-1,1]-rescaled frame stacks).  LDR-FLIP (Andersson et al. 2020):
//
//   colour pipeline   sRGB -> linear RGB -> XYZ -> YCxCz; contrast-sensitivity filtering of the three opponent planes
//                     (Gaussian kernels A / RG / BY, radius 10 px at 67 px/degree, replicate padding); back to linear
//                     RGB, clamp to [0,1]; -> L*a*b*, Hunt adjustment, HyAB distance ^ 0.7, remapped to [0,1];
//   feature pipeline  first- and second-derivative-of-Gaussian filters (radius 9) of the normalised luminance plane in
//                     x and y; | |edges_ref| - |edges_test| | and the same for points, max, (./sqrt 2) ^ 0.5;
//   FLIP              = colour_error ^ (1 - feature_error), per pixel; main.py reports the mean over the stack.
//
// Two launches for a whole stack of frames: (1) a point-wise pass turns both images into planar YCxCz (the only
// transcendental per pixel there: the sRGB decode), (2) one kernel per 16x16 tile stages the six planes with their
// halo in shared memory and does all five 2-D filters of both images plus the colour math per output pixel — the
// reference's ~60 eager torch kernels and five cuDNN convolutions per call.  Filter taps are computed by the caller
// exactly as the reference computes them (numpy double -> fp32) and passed in; summation order inside a filter is
// row-major over the taps (cuDNN's is unspecified): agreement ~1e-6 on the FLIP map.
// ~5.5 k FMA per pixel and image pair, 24 B/pixel in: compute-bound on CUDA cores, ~30 us per 400x400 frame.
#include "common.cuh"

namespace r2l {

struct FlipConsts {
  float A[9];        // linear RGB -> XYZ (fp32, utils/flip_loss.py:293-303)
  float Ainv[9];     // torch.inverse(A) in fp32
  float illum[3];    // A * (1,1,1): D65 reference illuminant
  float cmax;        // HyAB(hunt(lab(green)), hunt(lab(blue))) ^ qc
  float qc, qf, pc, pt;
  int r_csf, r_feat;  // filter radii
};

constexpr int kFlipTile = 16;

__device__ __forceinline__ float srgb_decode(float v) {
  v = fminf(fmaxf(v, 0.0f), 1.0f);
  return v > 0.04045f ? powf((v + 0.055f) / 1.055f, 2.4f) : v / 12.92f;
}
__device__ __forceinline__ void mat3(const float* M, float a, float b, float c, float& x, float& y, float& z) {
  x = M[0] * a + M[1] * b + M[2] * c;
  y = M[3] * a + M[4] * b + M[5] * c;
  z = M[6] * a + M[7] * b + M[8] * c;
}

// (1) planes[img][which][3][H][W]: YCxCz of the test (which = 0) and reference (1) image; v -> v * scale + offset first
__global__ void __launch_bounds__(256)
flip_prep_kernel(int H, int W, const float* __restrict__ test, const float* __restrict__ ref, long long img_stride,
                 float sc_t, float of_t, float sc_r, float of_r, FlipConsts c, float* __restrict__ planes) {
  const long long hw = static_cast<long long>(H) * W;
  const int img = blockIdx.y;
  for (long long p = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; p < hw; p += static_cast<long long>(gridDim.x) * 256) {
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const float* src = (which == 0 ? test : ref) + img * img_stride + 3 * p;
      const float sc = which == 0 ? sc_t : sc_r, of = which == 0 ? of_t : of_r;
      const float r = srgb_decode(src[0] * sc + of), g = srgb_decode(src[1] * sc + of), b = srgb_decode(src[2] * sc + of);
      float X, Y, Z;
      mat3(c.A, r, g, b, X, Y, Z);
      X /= c.illum[0], Y /= c.illum[1], Z /= c.illum[2];
      float* dst = planes + (static_cast<long long>(img) * 2 + which) * 3 * hw + p;
      dst[0] = 116.0f * Y - 16.0f;
      dst[hw] = 500.0f * (X - Y);
      dst[2 * hw] = 200.0f * (Y - Z);
    }
  }
}

__device__ __forceinline__ float lab_f(float t) {
  const float delta = 6.0f / 29.0f;
  return t > 0.00885f ? cbrtf(t) : t / (3.0f * delta * delta) + 4.0f / 29.0f;
}

// filtered opponent colour -> Hunt-adjusted L*a*b*
__device__ __forceinline__ void opponent_to_hunt_lab(const FlipConsts& c, float y, float cx, float cz, float& L, float& a,
                                                     float& b) {
  const float fy = (y + 16.0f) / 116.0f;
  float X = (fy + cx / 500.0f) * c.illum[0], Y = fy * c.illum[1], Z = (fy - cz / 200.0f) * c.illum[2];
  float r, g, bl;
  mat3(c.Ainv, X, Y, Z, r, g, bl);
  r = fminf(fmaxf(r, 0.0f), 1.0f), g = fminf(fmaxf(g, 0.0f), 1.0f), bl = fminf(fmaxf(bl, 0.0f), 1.0f);
  mat3(c.A, r, g, bl, X, Y, Z);
  const float fx = lab_f(X / c.illum[0]), fyy = lab_f(Y / c.illum[1]), fz = lab_f(Z / c.illum[2]);
  L = 116.0f * fyy - 16.0f;
  a = (0.01f * L) * (500.0f * (fx - fyy));
  b = (0.01f * L) * (200.0f * (fyy - fz));
}

// taps: [3][(2 r_csf + 1)^2] CSF filters A, RG, BY, then [2][(2 r_feat + 1)^2] edge / point x-filters (y = transpose)
__global__ void __launch_bounds__(kFlipTile * kFlipTile)
flip_main_kernel(int H, int W, const float* __restrict__ planes, const float* __restrict__ taps, FlipConsts c,
                 float* __restrict__ flip_map, double* __restrict__ sum_out) {
  extern __shared__ float smem[];
  const int R = c.r_csf > c.r_feat ? c.r_csf : c.r_feat;
  const int TW = kFlipTile + 2 * R;
  const int n_csf = (2 * c.r_csf + 1) * (2 * c.r_csf + 1), n_feat = (2 * c.r_feat + 1) * (2 * c.r_feat + 1);
  float* s_taps = smem;                              // 3 n_csf + 2 n_feat
  float* s_tile = smem + 3 * n_csf + 2 * n_feat;     // [2 images][3 planes][TW][TW]
  const int tid = threadIdx.y * kFlipTile + threadIdx.x;
  const int img = blockIdx.z;
  const long long hw = static_cast<long long>(H) * W;
  for (int i = tid; i < 3 * n_csf + 2 * n_feat; i += kFlipTile * kFlipTile) s_taps[i] = taps[i];
  const int x0 = blockIdx.x * kFlipTile - R, y0 = blockIdx.y * kFlipTile - R;
  for (int i = tid; i < 6 * TW * TW; i += kFlipTile * kFlipTile) {
    const int pl = i / (TW * TW), rem = i % (TW * TW), ty = rem / TW, tx = rem % TW;
    const int gx = min(max(x0 + tx, 0), W - 1), gy = min(max(y0 + ty, 0), H - 1);   // replicate padding
    s_tile[i] = planes[(static_cast<long long>(img) * 6 + pl) * hw + static_cast<long long>(gy) * W + gx];
  }
  __syncthreads();
  const int px = blockIdx.x * kFlipTile + threadIdx.x, py = blockIdx.y * kFlipTile + threadIdx.y;
  float value = 0.0f;
  if (px < W && py < H) {
    float lab[2][3];
    float feat[2][2];   // [image][edge, point] gradient magnitudes
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const float* t0 = s_tile + which * 3 * TW * TW;
      float f[3] = {0.f, 0.f, 0.f};
      const int o = R - c.r_csf;
      for (int dy = 0; dy <= 2 * c.r_csf; ++dy) {
        const float* row = t0 + (threadIdx.y + o + dy) * TW + threadIdx.x + o;
        const float* wr = s_taps + dy * (2 * c.r_csf + 1);
        for (int dx = 0; dx <= 2 * c.r_csf; ++dx) {
          f[0] = fmaf(row[dx], wr[dx], f[0]);
          f[1] = fmaf(row[TW * TW + dx], wr[n_csf + dx], f[1]);
          f[2] = fmaf(row[2 * TW * TW + dx], wr[2 * n_csf + dx], f[2]);
        }
      }
      opponent_to_hunt_lab(c, f[0], f[1], f[2], lab[which][0], lab[which][1], lab[which][2]);
      // feature filters on the normalised luminance (Y + 16) / 116 of the UNFILTERED plane
      float ex = 0.f, ey = 0.f, qx = 0.f, qy = 0.f;
      const int of = R - c.r_feat, nf = 2 * c.r_feat + 1;
      const float* we = s_taps + 3 * n_csf;
      const float* wp = we + n_feat;
      for (int dy = 0; dy < nf; ++dy) {
        const float* row = t0 + (threadIdx.y + of + dy) * TW + threadIdx.x + of;
        for (int dx = 0; dx < nf; ++dx) {
          const float v = (row[dx] + 16.0f) / 116.0f;
          ex = fmaf(v, we[dy * nf + dx], ex);
          ey = fmaf(v, we[dx * nf + dy], ey);   // transposed filter
          qx = fmaf(v, wp[dy * nf + dx], qx);
          qy = fmaf(v, wp[dx * nf + dy], qy);
        }
      }
      feat[which][0] = sqrtf(ex * ex + ey * ey);
      feat[which][1] = sqrtf(qx * qx + qy * qy);
    }
    const float da = lab[1][1] - lab[0][1], db = lab[1][2] - lab[0][2];
    const float hyab = fabsf(lab[1][0] - lab[0][0]) + sqrtf(da * da + db * db);
    const float pe = powf(hyab, c.qc);
    const float pccmax = c.pc * c.cmax;
    const float dEc = pe < pccmax ? (c.pt / pccmax) * pe : c.pt + ((pe - pccmax) / (c.cmax - pccmax)) * (1.0f - c.pt);
    float dEf = fmaxf(fabsf(feat[1][0] - feat[0][0]), fabsf(feat[0][1] - feat[1][1]));
    dEf = fminf(fmaxf(powf(0.70710678118654752f * dEf, c.qf), 0.0f), 1.0f);
    value = powf(dEc, 1.0f - dEf);
    if (flip_map != nullptr) flip_map[img * hw + static_cast<long long>(py) * W + px] = value;
  }
  // per-image sum (double) for the mean
  double v = static_cast<double>(value);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, o);
    hi = __shfl_xor_sync(0xffffffffu, hi, o);
    v += __hiloint2double(hi, lo);
  }
  __shared__ double s_part[kFlipTile * kFlipTile / 32];
  if ((tid & 31) == 0) s_part[tid >> 5] = v;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int i = 0; i < kFlipTile * kFlipTile / 32; ++i) s += s_part[i];
    atomicAdd(sum_out + img, s);
  }
}

}  // namespace r2l

using namespace r2l;

extern "C" {

// test, ref: [n_img][H][W][3] fp32 (image i at + i*img_stride floats), each value mapped v -> v*scale + offset on load
// (main.py:366-368 rescales both stacks to [-1,1] before LPIPS / FLIP).  consts: HOST pointer to 26 floats
// (A[9], Ainv[9], illuminant[3], cmax, qc, qf, pc, pt) + r_csf, r_feat passed separately; taps: DEVICE pointer,
// [3][(2 r_csf+1)^2] + [2][(2 r_feat+1)^2] floats.  workspace: DEVICE [n_img][2][3][H][W] floats.
// flip_map (optional) [n_img][H][W]; sum_out [n_img] double (zeroed here): FLIP of image i = sum_out[i] / (H W).
int r2l_flip(int n_img, int H, int W, const float* test, const float* ref, long long img_stride, double scale_test,
             double offset_test, double scale_ref, double offset_ref, const float* consts, int r_csf, int r_feat,
             const float* taps, float* workspace, float* flip_map, double* sum_out, void* stream) {
  R2L_CHECK_ARG(n_img >= 0 && H > 0 && W > 0, "r2l_flip: bad sizes");
  if (n_img == 0) return R2L_OK;
  R2L_CHECK_ARG(n_img <= 65535, "r2l_flip: more than 65535 images per call");
  R2L_CHECK_ARG(test && ref && consts && taps && workspace && sum_out, "r2l_flip: null pointer");
  R2L_CHECK_ARG(r_csf >= 0 && r_csf <= 24 && r_feat >= 0 && r_feat <= 24, "r2l_flip: filter radius outside [0, 24]");
  R2L_CHECK_ARG(img_stride >= static_cast<long long>(H) * W * 3, "r2l_flip: bad image stride");
  auto st = static_cast<cudaStream_t>(stream);
  FlipConsts c;
  for (int i = 0; i < 9; ++i) c.A[i] = consts[i], c.Ainv[i] = consts[9 + i];
  for (int i = 0; i < 3; ++i) c.illum[i] = consts[18 + i];
  c.cmax = consts[21], c.qc = consts[22], c.qf = consts[23], c.pc = consts[24], c.pt = consts[25];
  c.r_csf = r_csf, c.r_feat = r_feat;
  R2L_CUDA(cudaMemsetAsync(sum_out, 0, sizeof(double) * n_img, st));
  const long long hw = static_cast<long long>(H) * W;
  long long bx = (hw + 255) / 256;
  const long long cap = (static_cast<long long>(sm_count()) * 8 + n_img - 1) / n_img;
  if (bx > cap) bx = cap;
  flip_prep_kernel<<<dim3(static_cast<unsigned>(bx), static_cast<unsigned>(n_img)), 256, 0, st>>>(
      H, W, test, ref, img_stride, static_cast<float>(scale_test), static_cast<float>(offset_test),
      static_cast<float>(scale_ref), static_cast<float>(offset_ref), c, workspace);
  R2L_LAUNCH_CHECK();
  const int R = r_csf > r_feat ? r_csf : r_feat, TW = kFlipTile + 2 * R;
  const int n_taps = 3 * (2 * r_csf + 1) * (2 * r_csf + 1) + 2 * (2 * r_feat + 1) * (2 * r_feat + 1);
  const size_t smem = sizeof(float) * (static_cast<size_t>(n_taps) + 6ull * TW * TW);
  R2L_CHECK_ARG(smem <= 200 * 1024, "r2l_flip: filters too large for shared memory");
  if (smem > 48 * 1024)
    R2L_CUDA(cudaFuncSetAttribute(flip_main_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const dim3 grid((W + kFlipTile - 1) / kFlipTile, (H + kFlipTile - 1) / kFlipTile, n_img);
  flip_main_kernel<<<grid, dim3(kFlipTile, kFlipTile), smem, st>>>(H, W, workspace, taps, c, flip_map, sum_out);
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

}  // extern "C"
