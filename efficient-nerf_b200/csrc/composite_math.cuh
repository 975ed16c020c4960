// Arithmetic of the alpha compositing (raw2outputs, main.py:556-621) shared by the stand-alone kernels (composite.cu)
// and the compositor inside the NeRF ping-pong kernel (mlp_nerf_pp.cu): ONE definition, so the fused frame and the
// raw2outputs call produce the same bits.
#pragma once
#include <cuda_runtime.h>

namespace r2l {

__device__ __forceinline__ double shfl_up_f64(double v, int delta) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_up_sync(0xffffffffu, lo, delta);
  hi = __shfl_up_sync(0xffffffffu, hi, delta);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_idx_f64(double v, int src) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(0xffffffffu, lo, src);
  hi = __shfl_sync(0xffffffffu, hi, src);
  return __hiloint2double(hi, lo);
}

// 1 / (1 + 2^(-x*log2 e)): FMUL + MUFU.EX2 + FADD + MUFU.RCP (flush-to-zero is harmless: 1 + tiny = 1, rcp(inf) = 0)
__device__ __forceinline__ float fast_sigmoid(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__fmul_rn(x, -1.4426950408889634f)));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fadd_rn(1.0f, e)));
  return r;
}

// alpha = 1 - exp(-relu(sigma) * dist); relu / exp propagate NaN in the reference (main.py:579-600)
__device__ __forceinline__ float comp_alpha(float sigma, float dist) {
  const float rl = fmaxf(sigma, 0.0f);
  float a = __fsub_rn(1.0f, expf(__fmul_rn(-rl, dist)));
  if (sigma != sigma) a = sigma;
  return a;
}
// distance to the next sample (the last one: 1e10), scaled by |rays_d| (main.py:579-581)
__device__ __forceinline__ float comp_dist(float z, float z_next, bool last, float dnorm) {
  const float d = last ? 1e10f : __fsub_rn(z_next, z);
  return __fmul_rn(d, dnorm);
}
__device__ __forceinline__ float comp_dnorm(float dx, float dy, float dz) {
  return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
}
// 1 / max(1e-10, depth / acc) with torch.max's NaN propagation (0/0 when all weights are zero; main.py:614-615)
__device__ __forceinline__ float comp_disp(float depth, float acc) {
  const float q = __fdiv_rn(depth, acc);
  const float m = (q != q) ? q : fmaxf(1e-10f, q);
  return __fdiv_rn(1.0f, m);
}

// The "blocked" composite: lane `sl` (0 .. LPR-1) of a ray owns its K CONSECUTIVE samples [sl*K, sl*K+K); the
// transmittance is a sequential fp64 product inside the lane and ONE segmented fp64 scan over the ray's LPR lanes.
// On return w[j] = weight of the lane's j-th sample and (ar, ag, ab, adepth, aacc) hold the RAY's sums on every lane
// of the ray.  FULL: the ray has exactly K*LPR samples (no per-sample predicate).
template <int K, bool FULL, int LPR>
__device__ __forceinline__ void blocked_composite(const float4 (&rv)[K], const float (&zr)[K], float z_next_lane, int sl,
                                                  int S, bool ray_ok, float dnorm, const float* noise_row,
                                                  float (&w)[K], float& ar, float& ag, float& ab, float& adepth,
                                                  float& aacc) {
  const int base = sl * K;      // index of this lane's first sample within its ray
  float alpha[K];
  double excl_in[K];
  double run = 1.0;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int i = base + j;
    const bool valid = ray_ok && (FULL || i < S);
    const float zn = (j + 1 < K) ? zr[j + 1] : z_next_lane;
    const float dist = comp_dist(zr[j], zn, i == S - 1, dnorm);
    float sigma = rv[j].w;
    if (noise_row != nullptr && valid) sigma = __fadd_rn(sigma, __ldg(noise_row + i));
    const float a = comp_alpha(sigma, dist);
    alpha[j] = a;
    excl_in[j] = run;
    if (valid) run *= static_cast<double>(__fadd_rn(__fsub_rn(1.0f, a), 1e-10f));
  }
  double p = run;   // inclusive scan of the lane totals, segmented by ray
#pragma unroll
  for (int o = 1; o < LPR; o <<= 1) {
    const double q = shfl_up_f64(p, o);
    if (sl >= o) p *= q;
  }
  double excl = shfl_up_f64(p, 1);
  if (sl == 0) excl = 1.0;
  ar = 0.f, ag = 0.f, ab = 0.f, adepth = 0.f, aacc = 0.f;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const float T = static_cast<float>(excl * excl_in[j]);
    const bool valid = ray_ok && (FULL || base + j < S);
    const float wj = valid ? __fmul_rn(alpha[j], T) : 0.0f;
    w[j] = wj;
    if (valid) {
      // sigmoid with ex2.approx / rcp.approx: |error| <= s(1-s)*(2+1.16|x|) ulp + 1 ulp < 1.2e-7 absolute, inside
      // the 2e-6 gate of the fp32 path; it removes three IEEE-division slow-path stubs per sample.
      const float sr = fast_sigmoid(rv[j].x), sg = fast_sigmoid(rv[j].y), sb = fast_sigmoid(rv[j].z);
      ar = __fadd_rn(ar, __fmul_rn(wj, sr));
      ag = __fadd_rn(ag, __fmul_rn(wj, sg));
      ab = __fadd_rn(ab, __fmul_rn(wj, sb));
      adepth = __fadd_rn(adepth, __fmul_rn(wj, zr[j]));
      aacc = __fadd_rn(aacc, wj);
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) {
    ar += __shfl_xor_sync(0xffffffffu, ar, o);
    ag += __shfl_xor_sync(0xffffffffu, ag, o);
    ab += __shfl_xor_sync(0xffffffffu, ab, o);
    adepth += __shfl_xor_sync(0xffffffffu, adepth, o);
    aacc += __shfl_xor_sync(0xffffffffu, aacc, o);
  }
}

}  // namespace r2l
