// Ray generation, NDC warp, view-direction normalisation, stratified depths and the
// R2L point sampler.  All arithmetic uses explicitly rounded single operations
// (__fmul_rn/__fadd_rn/__fdiv_rn, never contracted into FMA) so that results are
// bit-identical to the reference's eager fp32 tensor ops:
//   get_rays       utils/run_nerf_raybased_helpers.py:231-257
//   ndc_rays       utils/run_nerf_raybased_helpers.py:260-279
//   viewdirs       main.py:148-157
//   z_vals/perturb main.py:676-699
//   PointSampler   model/nerf_raybased.py:76-126
#include "common.cuh"

namespace r2l {

__device__ __forceinline__ void pixel_dir(int w, int h, float half_w, float half_h, float focal, float& dx,
                                          float& dy, float& dz) {
  // dirs = [(i - W*.5)/focal, -(j - H*.5)/focal, -1]        (helpers:238-240)
  dx = __fdiv_rn(__fsub_rn(static_cast<float>(w), half_w), focal);
  dy = -__fdiv_rn(__fsub_rn(static_cast<float>(h), half_h), focal);
  dz = -1.0f;
}

__device__ __forceinline__ float rot_row(float dx, float dy, float dz, const float* __restrict__ c) {
  // torch.sum(dirs[..., None, :] * c2w[:3,:3], -1): products rounded separately, summed left to right
  float acc = __fadd_rn(0.0f, __fmul_rn(dx, c[0]));
  acc = __fadd_rn(acc, __fmul_rn(dy, c[1]));
  acc = __fadd_rn(acc, __fmul_rn(dz, c[2]));
  return acc;
}

__global__ void get_rays_kernel(int H, int W, float focal, const float* __restrict__ c2w, float* __restrict__ ro,
                                float* __restrict__ rd) {
  const long long n = static_cast<long long>(H) * W;
  const float half_w = static_cast<float>(W * 0.5);
  const float half_h = static_cast<float>(H * 0.5);
  float c[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) c[i] = __ldg(c2w + i);
  for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < n;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int h = static_cast<int>(p / W), w = static_cast<int>(p % W);
    float dx, dy, dz;
    pixel_dir(w, h, half_w, half_h, focal, dx, dy, dz);
    rd[3 * p + 0] = rot_row(dx, dy, dz, c + 0);
    rd[3 * p + 1] = rot_row(dx, dy, dz, c + 4);
    rd[3 * p + 2] = rot_row(dx, dy, dz, c + 8);
    ro[3 * p + 0] = c[3];
    ro[3 * p + 1] = c[7];
    ro[3 * p + 2] = c[11];
  }
}

__global__ void ndc_rays_kernel(long long n, float sw, float sh, float near, float two_near, float neg_two_near,
                                const float* __restrict__ ro, const float* __restrict__ rd, float* __restrict__ oo,
                                float* __restrict__ od) {
  for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < n;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    float ox = ro[3 * p], oy = ro[3 * p + 1], oz = ro[3 * p + 2];
    const float dx = rd[3 * p], dy = rd[3 * p + 1], dz = rd[3 * p + 2];
    // t = -(near + o_z) / d_z ; o = o + t*d                      (helpers:262-263)
    const float t = __fdiv_rn(-__fadd_rn(near, oz), dz);
    ox = __fadd_rn(ox, __fmul_rn(t, dx));
    oy = __fadd_rn(oy, __fmul_rn(t, dy));
    oz = __fadd_rn(oz, __fmul_rn(t, dz));
    // projection                                                 (helpers:266-274)
    const float o0 = __fdiv_rn(__fmul_rn(sw, ox), oz);
    const float o1 = __fdiv_rn(__fmul_rn(sh, oy), oz);
    const float o2 = __fadd_rn(1.0f, __fdiv_rn(two_near, oz));
    const float d0 = __fmul_rn(sw, __fsub_rn(__fdiv_rn(dx, dz), __fdiv_rn(ox, oz)));
    const float d1 = __fmul_rn(sh, __fsub_rn(__fdiv_rn(dy, dz), __fdiv_rn(oy, oz)));
    const float d2 = __fdiv_rn(neg_two_near, oz);
    oo[3 * p] = o0;
    oo[3 * p + 1] = o1;
    oo[3 * p + 2] = o2;
    od[3 * p] = d0;
    od[3 * p + 1] = d1;
    od[3 * p + 2] = d2;
  }
}

__global__ void normalize_dirs_kernel(long long n, const float* __restrict__ d, long long d_stride,
                                      float* __restrict__ out) {
  for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < n;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float x = d[p * d_stride], y = d[p * d_stride + 1], z = d[p * d_stride + 2];
    const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    out[3 * p] = __fdiv_rn(x, nrm);
    out[3 * p + 1] = __fdiv_rn(y, nrm);
    out[3 * p + 2] = __fdiv_rn(z, nrm);
  }
}

// Pluecker ray coordinates [d, o x d] (model/nerf_raybased.py:170-190).  ATen's CPU cross kernel evaluates each
// component a1*b2 - a2*b1 as fma(a1, b2, -round(a2*b1)) (pinned against the reference's outputs in
// tests/golden/sampler_extra.npz: bit-exact in that form, 73/630 elements off by 1 ulp with two rounded products);
// the same form here.  o_stride == 0 broadcasts one origin (sample_test_plucker).
__global__ void plucker_kernel(long long n, const float* __restrict__ o, long long o_stride,
                               const float* __restrict__ d, long long d_stride, float* __restrict__ out) {
  for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < n;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float ox = o[p * o_stride], oy = o[p * o_stride + 1], oz = o[p * o_stride + 2];
    const float dx = d[p * d_stride], dy = d[p * d_stride + 1], dz = d[p * d_stride + 2];
    float* r = out + 6 * p;
    r[0] = dx;
    r[1] = dy;
    r[2] = dz;
    r[3] = __fmaf_rn(oy, dz, -__fmul_rn(oz, dy));
    r[4] = __fmaf_rn(oz, dx, -__fmul_rn(ox, dz));
    r[5] = __fmaf_rn(ox, dy, -__fmul_rn(oy, dx));
  }
}

// z_vals = near*(1-t) + far*t  (or lindisp), optional stratified perturbation (main.py:676-699)
__global__ void z_vals_kernel(long long n, int S, const float* __restrict__ near, const float* __restrict__ far,
                              long long nf_stride, const float* __restrict__ t_vals, int lindisp,
                              const float* __restrict__ t_rand, float* __restrict__ z_out) {
  const long long total = n * S;
  for (long long g = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; g < total;
       g += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = g / S;
    const int s = static_cast<int>(g % S);
    const float nr = near[r * nf_stride], fr = far[r * nf_stride];
    auto zval = [&](int si) -> float {
      const float t = __ldg(t_vals + si);
      const float omt = __fsub_rn(1.0f, t);
      if (!lindisp) return __fadd_rn(__fmul_rn(nr, omt), __fmul_rn(fr, t));
      const float a = __fmul_rn(__fdiv_rn(1.0f, nr), omt);
      const float b = __fmul_rn(__fdiv_rn(1.0f, fr), t);
      return __fdiv_rn(1.0f, __fadd_rn(a, b));
    };
    float z = zval(s);
    if (t_rand != nullptr) {
      // mids = .5*(z[1:]+z[:-1]); upper=[mids, z[-1]]; lower=[z[0], mids]; z = lower + (upper-lower)*t_rand
      const float zl = (s > 0) ? zval(s - 1) : z;
      const float zu = (s < S - 1) ? zval(s + 1) : z;
      const float lower = (s > 0) ? __fmul_rn(0.5f, __fadd_rn(z, zl)) : z;
      const float upper = (s < S - 1) ? __fmul_rn(0.5f, __fadd_rn(zu, z)) : z;
      z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t_rand[g]));
    }
    z_out[g] = z;
  }
}

// pts[r, s, :] = o + d * z   (main.py:701, model/nerf_raybased.py:100,124)
__global__ void points_from_rays_kernel(long long n, int S, const float* __restrict__ ro, long long o_stride,
                                        const float* __restrict__ rd, long long d_stride,
                                        const float* __restrict__ z, long long z_stride, float* __restrict__ pts) {
  const long long total = n * S;
  for (long long g = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; g < total;
       g += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = g / S;
    const int s = static_cast<int>(g % S);
    const float zz = z[r * z_stride + s];
#pragma unroll
    for (int k = 0; k < 3; ++k)
      pts[3 * g + k] = __fadd_rn(ro[r * o_stride + k], __fmul_rn(rd[r * d_stride + k], zz));
  }
}

// PointSampler.sample_test: rays from the pixel grid + c2w, then pts on shared z_vals
// (model/nerf_raybased.py:94-102).  One thread per (ray, sample); blockIdx.y = pose of a batch (c2w [n][3][4],
// pts [n][H*W][S*3]).
__global__ void point_sample_kernel(int H, int W, float focal, const float* __restrict__ c2w,
                                    const float* __restrict__ z_vals, int S, float* __restrict__ pts) {
  const long long total = static_cast<long long>(H) * W * S;
  c2w += 12 * blockIdx.y;
  pts += 3 * total * blockIdx.y;
  const float half_w = static_cast<float>(W * 0.5);
  const float half_h = static_cast<float>(H * 0.5);
  float c[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) c[i] = __ldg(c2w + i);
  for (long long g = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; g < total;
       g += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = g / S;
    const int s = static_cast<int>(g % S);
    const int h = static_cast<int>(p / W), w = static_cast<int>(p % W);
    float dx, dy, dz;
    pixel_dir(w, h, half_w, half_h, focal, dx, dy, dz);
    const float zz = __ldg(z_vals + s);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float d = rot_row(dx, dy, dz, c + 4 * k);
      pts[3 * g + k] = __fadd_rn(c[4 * k + 3], __fmul_rn(d, zz));
    }
  }
}

static inline int grid_for(long long n, int block) {
  long long g = (n + block - 1) / block;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace r2l

using namespace r2l;

extern "C" {

int r2l_get_rays(int H, int W, double focal, const float* c2w, float* rays_o, float* rays_d, void* stream) {
  R2L_CHECK_ARG(H > 0 && W > 0 && focal != 0.0, "r2l_get_rays: bad H/W/focal");
  R2L_CHECK_ARG(c2w && rays_o && rays_d, "r2l_get_rays: null pointer");
  const long long n = static_cast<long long>(H) * W;
  get_rays_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      H, W, static_cast<float>(focal), c2w, rays_o, rays_d);
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

int r2l_ndc_rays(long long n, int H, int W, double focal, double near, const float* rays_o, const float* rays_d,
                 float* out_o, float* out_d, void* stream) {
  R2L_CHECK_ARG(n >= 0 && H > 0 && W > 0 && focal != 0.0, "r2l_ndc_rays: bad sizes");
  if (n == 0) return R2L_OK;
  R2L_CHECK_ARG(rays_o && rays_d && out_o && out_d, "r2l_ndc_rays: null pointer");
  // python-side scalars are evaluated in double, then rounded once when they meet an fp32 tensor
  const float sw = static_cast<float>(-1.0 / (W / (2.0 * focal)));
  const float sh = static_cast<float>(-1.0 / (H / (2.0 * focal)));
  ndc_rays_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      n, sw, sh, static_cast<float>(near), static_cast<float>(2.0 * near), static_cast<float>(-2.0 * near), rays_o,
      rays_d, out_o, out_d);
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

int r2l_normalize_dirs(long long n, const float* dirs, long long stride, float* out, void* stream) {
  R2L_CHECK_ARG(n >= 0 && stride >= 3, "r2l_normalize_dirs: bad sizes");
  if (n == 0) return R2L_OK;
  R2L_CHECK_ARG(dirs && out, "r2l_normalize_dirs: null pointer");
  normalize_dirs_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(n, dirs, stride, out);
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

int r2l_plucker(long long n, const float* rays_o, long long o_stride, const float* rays_d, long long d_stride,
                float* out, void* stream) {
  R2L_CHECK_ARG(n >= 0 && (o_stride == 0 || o_stride >= 3) && d_stride >= 3, "r2l_plucker: bad sizes");
  if (n == 0) return R2L_OK;
  R2L_CHECK_ARG(rays_o && rays_d && out, "r2l_plucker: null pointer");
  plucker_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(n, rays_o, o_stride, rays_d, d_stride,
                                                                                 out);
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

int r2l_z_vals(long long n, int S, const float* near, const float* far, long long nf_stride, const float* t_vals,
               int lindisp, const float* t_rand, float* z_out, void* stream) {
  R2L_CHECK_ARG(n >= 0 && S > 0, "r2l_z_vals: bad sizes");
  if (n == 0) return R2L_OK;
  R2L_CHECK_ARG(near && far && t_vals && z_out, "r2l_z_vals: null pointer");
  z_vals_kernel<<<grid_for(n * S, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(n, S, near, far, nf_stride,
                                                                                    t_vals, lindisp, t_rand, z_out);
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

int r2l_points_from_rays(long long n, int S, const float* rays_o, long long o_stride, const float* rays_d,
                         long long d_stride, const float* z, long long z_stride, float* pts, void* stream) {
  R2L_CHECK_ARG(n >= 0 && S > 0, "r2l_points_from_rays: bad sizes");
  if (n == 0) return R2L_OK;
  R2L_CHECK_ARG(rays_o && rays_d && z && pts, "r2l_points_from_rays: null pointer");
  points_from_rays_kernel<<<grid_for(n * S, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      n, S, rays_o, o_stride, rays_d, d_stride, z, z_stride, pts);
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

int r2l_point_sample(int H, int W, double focal, const float* c2w, const float* z_vals, int S, float* pts,
                     void* stream) {
  R2L_CHECK_ARG(H > 0 && W > 0 && S > 0 && focal != 0.0, "r2l_point_sample: bad sizes");
  R2L_CHECK_ARG(c2w && z_vals && pts, "r2l_point_sample: null pointer");
  const long long total = static_cast<long long>(H) * W * S;
  point_sample_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      H, W, static_cast<float>(focal), c2w, z_vals, S, pts);
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

int r2l_point_sample_batch(int n_poses, int H, int W, double focal, const float* c2w, const float* z_vals, int S,
                           float* pts, void* stream) {
  R2L_CHECK_ARG(n_poses >= 0 && n_poses <= 65535 && H > 0 && W > 0 && S > 0 && focal != 0.0,
                "r2l_point_sample_batch: bad sizes");
  if (n_poses == 0) return R2L_OK;
  R2L_CHECK_ARG(c2w && z_vals && pts, "r2l_point_sample_batch: null pointer");
  const long long total = static_cast<long long>(H) * W * S;
  int gx = grid_for(total, 256) / n_poses;
  if (gx < 1) gx = 1;
  point_sample_kernel<<<dim3(static_cast<unsigned>(gx), static_cast<unsigned>(n_poses)), 256, 0,
                        static_cast<cudaStream_t>(stream)>>>(H, W, static_cast<float>(focal), c2w, z_vals, S, pts);
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

}  // extern "C"
