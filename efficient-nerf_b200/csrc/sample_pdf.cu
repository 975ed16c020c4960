// Hierarchical inverse-CDF sampling (sample_pdf) and the sorted merge of coarse and
// fine depths.  Reference: utils/run_nerf_raybased_helpers.py:283-330, called from
// main.py:722-732 (the reference runs this on the HOST with ATen CPU kernels and
// round-trips through PCIe every chunk; here it stays on the device).
//
// Bit-exactness contract (restated and pinned against torch CPU in oracle/sample_pdf_np.py):
//   w      = weights + 1e-5f                      one fp32 add
//   total  = sum(w) in ATen `vectorized_inner_sum` order: 8 vector lanes x 4 interleaved
//            accumulators, folded 0+=1,2,3; scalar tail summed first, then the 8 lanes
//            added left to right (rows shorter than 8: the same 4-accumulator scheme on scalars).
//            ASSUMPTION: this is ATen's AVX2 kernel (8 fp32 lanes).  torch 2.11 dispatches it for `sum` on AVX-512
//            hosts as well (measured here and on the GPU boxes, SURVEY 8a-8: torch.backends.cpu.get_cpu_capability()
//            says AVX512 but the fp32 row sums match the 8-lane order on 200 k rows); a build whose `sum` used 16
//            lanes would round `total` differently, and the bit-exact claim would then hold against goldens made
//            with ATEN_CPU_CAPABILITY=avx2 (tests/golden/sample_pdf.npz records the capability it was made with).
//   pdf    = w / total                            correctly rounded fp32 divide
//   cdf    = [0, cumsum(pdf)] accumulated in DOUBLE, each entry rounded to fp32.  All partial
//            sums of these fp32 values are exact in double (values >= ~1e-7, sum <= ~1, < 53
//            bits of span), so a warp prefix scan gives the same bits as the sequential loop.
//   inds   = searchsorted(cdf, u, right=True) = #{cdf_j <= u}     branch-free bisection
//   t      = (u - cdf[below]) / denom ; sample = bins[below] + t*(bins[above]-bins[below])
//            with separately rounded fp32 sub/div/mul/add (no FMA contraction).
//
// One warp per ray: coalesced loads of bins/weights into a per-warp shared-memory row,
// prefix sum by warp shuffles, every lane then inverts the CDF for Ni/32 samples.
// Algorithmic HBM bytes per ray: 4*(nb + nb-1) in, 4*Ni out (+ 4*Ni when u is per-ray).
#include "common.cuh"

namespace r2l {

constexpr int kPdfWarps = 8;

__device__ __forceinline__ double shfl_up_f64_(double v, int delta) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_up_sync(0xffffffffu, lo, delta);
  hi = __shfl_up_sync(0xffffffffu, hi, delta);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_idx_f64_(double v, int src) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(0xffffffffu, lo, src);
  hi = __shfl_sync(0xffffffffu, hi, src);
  return __hiloint2double(hi, lo);
}

// Correctly rounded num / denom that keeps a ZERO dividend off the IEEE-divide slow path (one lane with num == 0
// otherwise drags the whole warp through ~50 extra instructions): det=True has u_0 = cdf_0 = 0 in every ray and
// u_last = 1 = cdf_last in most.  +-0 / denom = +-0 exactly for a finite denom > 0; everything else divides normally.
__device__ __forceinline__ float div_zero_fast(float num, float denom) {
  const bool zero = (num == 0.0f) && (denom > 0.0f) && (denom < __int_as_float(0x7f800000));
  const float t = __fdiv_rn(zero ? 1.0f : num, denom);
  return zero ? num : t;
}

// Sum of w[0..n) in ATen-CPU order.  Executed by the whole warp; result valid on every lane.
__device__ __forceinline__ float aten_row_sum(const float* w, int n, int lane) {
  float total = 0.0f;
  if (n < 8) {
    if (lane == 0) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      const int m = n / 4;
      for (int i = 0; i < m; ++i) {
        a0 = __fadd_rn(a0, w[4 * i]);
        a1 = __fadd_rn(a1, w[4 * i + 1]);
        a2 = __fadd_rn(a2, w[4 * i + 2]);
        a3 = __fadd_rn(a3, w[4 * i + 3]);
      }
      for (int i = 4 * m; i < n; ++i) a0 = __fadd_rn(a0, w[i]);
      a0 = __fadd_rn(a0, a1);
      a0 = __fadd_rn(a0, a2);
      a0 = __fadd_rn(a0, a3);
      total = a0;
    }
    return __shfl_sync(0xffffffffu, total, 0);
  }
  const int vec_size = n >> 3;
  const int size_ilp = vec_size >> 2;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (lane < 8) {
    for (int i = 0; i < size_ilp; ++i) {
      const float* p = w + (i * 4) * 8 + lane;
      a0 = __fadd_rn(a0, p[0]);
      a1 = __fadd_rn(a1, p[8]);
      a2 = __fadd_rn(a2, p[16]);
      a3 = __fadd_rn(a3, p[24]);
    }
    for (int i = size_ilp * 4; i < vec_size; ++i) a0 = __fadd_rn(a0, w[i * 8 + lane]);
    a0 = __fadd_rn(a0, a1);
    a0 = __fadd_rn(a0, a2);
    a0 = __fadd_rn(a0, a3);
  }
  if (lane == 0) {
    for (int k = vec_size * 8; k < n; ++k) total = __fadd_rn(total, w[k]);
  }
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const float part = __shfl_sync(0xffffffffu, a0, m);
    if (lane == 0) total = __fadd_rn(total, part);
  }
  return __shfl_sync(0xffffffffu, total, 0);
}

__global__ void __launch_bounds__(kPdfWarps * 32)
sample_pdf_kernel(long long n_rays, int nb, int Ni, const float* __restrict__ bins, long long bins_stride,
                  const float* __restrict__ weights, long long w_stride, const float* __restrict__ u,
                  int u_per_ray, float* __restrict__ samples, long long* __restrict__ inds_out, int row_pad) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  float* s_w = sm + static_cast<size_t>(wib) * 3 * row_pad;
  float* s_cdf = s_w + row_pad;
  float* s_bins = s_cdf + row_pad;
  const int nw = nb - 1;
  // largest power of two <= nb, first bisection step
  int p2 = 1;
  while ((p2 << 1) <= nb) p2 <<= 1;
  const long long warp0 = static_cast<long long>(blockIdx.x) * kPdfWarps + wib;
  const long long nwarps = static_cast<long long>(gridDim.x) * kPdfWarps;
  for (long long ray = warp0; ray < n_rays; ray += nwarps) {
    const float* rb = bins + ray * bins_stride;
    const float* rw = weights + ray * w_stride;
    for (int j = lane; j < nb; j += 32) s_bins[j] = __ldg(rb + j);
    for (int j = lane; j < nw; j += 32) s_w[j] = __fadd_rn(__ldg(rw + j), 1e-5f);
    __syncwarp();
    const float total = aten_row_sum(s_w, nw, lane);
    // cdf = [0, cumsum(pdf)] in double
    double carry = 0.0;
    if (lane == 0) s_cdf[0] = 0.0f;
    for (int c0 = 0; c0 < nw; c0 += 32) {
      const int j = c0 + lane;
      double p = (j < nw) ? static_cast<double>(__fdiv_rn(s_w[j], total)) : 0.0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double q = shfl_up_f64_(p, o);
        if (lane >= o) p += q;
      }
      if (j < nw) s_cdf[j + 1] = static_cast<float>(carry + p);
      carry += shfl_idx_f64_(p, 31);
    }
    __syncwarp();
    for (int k = lane; k < Ni; k += 32) {
      const float uk = u_per_ray ? __ldg(u + ray * Ni + k) : __ldg(u + k);
      int pos = 0;
      for (int s = p2; s > 0; s >>= 1) {
        const int np = pos + s;
        if (np <= nb && s_cdf[np - 1] <= uk) pos = np;
      }
      const int below = max(pos - 1, 0);
      const int above = min(pos, nb - 1);
      const float cb = s_cdf[below], ca = s_cdf[above];
      const float bb = s_bins[below], ba = s_bins[above];
      float denom = __fsub_rn(ca, cb);
      if (denom < 1e-5f) denom = 1.0f;
      const float t = __fdiv_rn(__fsub_rn(uk, cb), denom);
      samples[ray * Ni + k] = __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
      if (inds_out != nullptr) inds_out[ray * Ni + k] = pos;
    }
    __syncwarp();
  }
}

// ---- specialised fast path: nb = NB (compile time) <= 64, Ni = 32*NIT -------------------------------------------
// The generic kernel above runs ~1100 warp instructions per ray (dynamic-trip loops, two chunk scans, a bounds-
// checked bisection, four scalar shared loads per sample): 12 % of the HBM roofline (ncu, r1b).  With the sizes of
// every BASELINE config known at compile time (64 coarse samples -> nb = 63; 128 or 64 fine samples):
//   * the ATen-order row sum is straight-line code executed redundantly by all lanes (no divergence),
//   * lane l owns pdf entries 2l, 2l+1: one fp64 warp scan per ray (partial sums are exact in double, see header),
//   * cdf and bins are interleaved as float2 so `below`/`above` cost one 64-bit shared load each, the cdf row is
//     padded with +inf so the unrolled bisection needs no bounds check, and a shared `u` table lives in registers.
template <int N>
__device__ __forceinline__ float aten_row_sum_ct(const float* w, int lane) {
  static_assert(N >= 8, "vectorised ATen path only");
  constexpr int vec_size = N >> 3, size_ilp = vec_size >> 2;
  const float* p = w + (lane & 7);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
  for (int i = 0; i < size_ilp; ++i) {
    a0 = __fadd_rn(a0, p[(i * 4) * 8]);
    a1 = __fadd_rn(a1, p[(i * 4 + 1) * 8]);
    a2 = __fadd_rn(a2, p[(i * 4 + 2) * 8]);
    a3 = __fadd_rn(a3, p[(i * 4 + 3) * 8]);
  }
#pragma unroll
  for (int i = size_ilp * 4; i < vec_size; ++i) a0 = __fadd_rn(a0, p[i * 8]);
  a0 = __fadd_rn(a0, a1);
  a0 = __fadd_rn(a0, a2);
  a0 = __fadd_rn(a0, a3);
  float total = 0.0f;
#pragma unroll
  for (int k = vec_size * 8; k < N; ++k) total = __fadd_rn(total, w[k]);
#pragma unroll
  for (int m = 0; m < 8; ++m) total = __fadd_rn(total, __shfl_sync(0xffffffffu, a0, m));
  return total;
}

template <int NB, int NIT>
__global__ void __launch_bounds__(kPdfWarps * 32)
sample_pdf_fast_kernel(long long n_rays, const float* __restrict__ bins, long long bins_stride,
                       const float* __restrict__ weights, long long w_stride, const float* __restrict__ u,
                       int u_per_ray, float* __restrict__ samples, long long* __restrict__ inds_out) {
  static_assert(NB >= 9 && NB <= 64, "one 64-entry row per warp");
  constexpr int NW = NB - 1, Ni = 32 * NIT;
  constexpr int P2 = (NB >= 64) ? 64 : (NB >= 32) ? 32 : (NB >= 16) ? 16 : 8;   // first bisection step
  __shared__ __align__(16) float s_w_all[kPdfWarps][64];
  __shared__ __align__(16) float2 s_cb_all[kPdfWarps][128];   // (cdf_j, bins_j); cdf padded with +inf up to 2*P2
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  float* const s_w = s_w_all[wib];
  float2* const s_cb = s_cb_all[wib];
  float ureg[NIT];
  if (!u_per_ray) {
#pragma unroll
    for (int it = 0; it < NIT; ++it) ureg[it] = __ldg(u + it * 32 + lane);
  }
  for (int j = NB + lane; j < 2 * P2; j += 32) s_cb[j] = make_float2(__int_as_float(0x7f800000), 0.f);
  const long long warp0 = static_cast<long long>(blockIdx.x) * kPdfWarps + wib;
  const long long nwarps = static_cast<long long>(gridDim.x) * kPdfWarps;
  for (long long ray = warp0; ray < n_rays; ray += nwarps) {
    const float* rb = bins + ray * bins_stride;
    const float* rw = weights + ray * w_stride;
    const float b0 = __ldg(rb + lane);
    const float b1 = (lane + 32 < NB) ? __ldg(rb + lane + 32) : 0.f;
    s_w[lane] = __fadd_rn(__ldg(rw + lane), 1e-5f);
    if (lane + 32 < NW) s_w[lane + 32] = __fadd_rn(__ldg(rw + lane + 32), 1e-5f);
    if (u_per_ray) {
#pragma unroll
      for (int it = 0; it < NIT; ++it) ureg[it] = __ldg(u + ray * Ni + it * 32 + lane);
    }
    __syncwarp();
    const float total = aten_row_sum_ct<NW>(s_w, lane);
    const float2 wp = *reinterpret_cast<const float2*>(s_w + 2 * lane);
    const double p0 = (2 * lane < NW) ? static_cast<double>(__fdiv_rn(wp.x, total)) : 0.0;
    const double p1 = (2 * lane + 1 < NW) ? static_cast<double>(__fdiv_rn(wp.y, total)) : 0.0;
    double p = p0 + p1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double q = shfl_up_f64_(p, o);
      if (lane >= o) p += q;
    }
    double excl = shfl_up_f64_(p, 1);
    if (lane == 0) excl = 0.0;
    __syncwarp();   // previous ray's readers of s_cb are done (loop-carried), this ray's s_w reads are done
    // cdf[0] = 0; cdf[j+1] = float(sum_{i<=j} pdf_i)
    s_cb[lane].y = b0;
    if (lane + 32 < NB) s_cb[lane + 32].y = b1;
    if (lane == 0) s_cb[0].x = 0.0f;
    if (2 * lane < NW) s_cb[2 * lane + 1].x = static_cast<float>(excl + p0);
    if (2 * lane + 1 < NW) s_cb[2 * lane + 2].x = static_cast<float>(excl + (p0 + p1));
    __syncwarp();
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const float uk = ureg[it];
      int pos = 0;
#pragma unroll
      for (int st = P2; st > 0; st >>= 1) {
        if (s_cb[pos + st - 1].x <= uk) pos += st;
      }
      const int below = max(pos - 1, 0);
      const int above = min(pos, NB - 1);
      const float2 lo = s_cb[below], hi = s_cb[above];
      float denom = __fsub_rn(hi.x, lo.x);
      if (denom < 1e-5f) denom = 1.0f;
      const float t = div_zero_fast(__fsub_rn(uk, lo.x), denom);
      samples[ray * Ni + it * 32 + lane] = __fadd_rn(lo.y, __fmul_rn(t, __fsub_rn(hi.y, lo.y)));
      if (inds_out != nullptr) inds_out[ray * Ni + it * 32 + lane] = pos;
    }
  }
}

// ---- fused hierarchical sampling: mids + sample_pdf + sorted merge + z_std in ONE kernel ----------------------------
// main.py:720-733, 750 for the configuration every BASELINE render uses (64 coarse samples, a shared `u` table):
//   z_vals_mid = .5 * (z[1:] + z[:-1]);  z_samples = sample_pdf(z_vals_mid, weights[1:-1], Ni, det);
//   z_all = sort(cat[z, z_samples]);     z_std = std(z_samples, unbiased=False)
// The unfused path runs two elementwise torch kernels, sample_pdf and merge_sort and moves 2.5 KB per ray through
// HBM; here a ray costs 512 B in and 4*(64+Ni) B out, and the merge is almost free: a sample drawn from bin
// [mid_b, mid_{b+1}] has exactly the coarse depths z_0..z_b (and possibly z_{b+1}) below it, so its merged position
// is k + b + 1 + (z_{b+1} <= s) (verified / corrected by a local scan, so any ascending z is handled exactly); the
// coarse depths find theirs by a 7-step search over the (ascending) samples in shared memory.  With a stochastic
// per-ray u (create_data) the samples come out unordered: they alone are bitonic-sorted in shared memory, then the
// same rank merge runs with searches on both sides.  Rows whose z is not ascending take a full bitonic sort.
// Same arithmetic, in the same order, as sample_pdf_fast_kernel and merge_sort_kernel: results are bit-identical.
template <int NIT>
__global__ void __launch_bounds__(kPdfWarps * 32)
hier_sample_kernel(long long n_rays, const float* __restrict__ z_vals, const float* __restrict__ weights,
                   const float* __restrict__ u, int u_per_ray, float* __restrict__ z_out, float* __restrict__ z_std,
                   float* __restrict__ samples_out, long long* __restrict__ inds_out) {
  constexpr int NS = 64, NB = 63, NW = 62, Ni = 32 * NIT, NO = NS + Ni, P2 = 32;
  constexpr int PSORT = (NO <= 128) ? 128 : 256;
  __shared__ __align__(16) float s_w_all[kPdfWarps][64];
  __shared__ __align__(16) float2 s_cb_all[kPdfWarps][64];   // (cdf_j, bins_j), j < 63; [63] = (+inf, 0)
  __shared__ __align__(16) float s_z_all[kPdfWarps][64];
  __shared__ __align__(16) float s_o_all[kPdfWarps][PSORT];  // samples (fast path) / sort buffer (fallback)
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  float* const s_w = s_w_all[wib];
  float2* const s_cb = s_cb_all[wib];
  float* const s_z = s_z_all[wib];
  float* const s_o = s_o_all[wib];
  float ureg[NIT];
  if (!u_per_ray) {
#pragma unroll
    for (int it = 0; it < NIT; ++it) ureg[it] = __ldg(u + it * 32 + lane);
  }
  if (lane == 0) s_cb[63] = make_float2(__int_as_float(0x7f800000), 0.f);
  const long long warp0 = static_cast<long long>(blockIdx.x) * kPdfWarps + wib;
  const long long nwarps = static_cast<long long>(gridDim.x) * kPdfWarps;
  for (long long ray = warp0; ray < n_rays; ray += nwarps) {
    const float* rz = z_vals + ray * NS;
    const float* rw = weights + ray * NS + 1;   // weights[..., 1:-1]
    const float z0 = __ldg(rz + lane), z1 = __ldg(rz + lane + 32);
    s_z[lane] = z0;
    s_z[lane + 32] = z1;
    s_w[lane] = __fadd_rn(__ldg(rw + lane), 1e-5f);
    if (lane + 32 < NW) s_w[lane + 32] = __fadd_rn(__ldg(rw + lane + 32), 1e-5f);
    if (u_per_ray) {
#pragma unroll
      for (int it = 0; it < NIT; ++it) ureg[it] = __ldg(u + ray * Ni + it * 32 + lane);
    }
    __syncwarp();
    // bins_j = .5 * (z_{j+1} + z_j): one rounded add, an exact halving
    const float b0 = __fmul_rn(0.5f, __fadd_rn(s_z[lane + 1], z0));
    const float b1 = (lane + 32 < NB) ? __fmul_rn(0.5f, __fadd_rn(s_z[lane + 33], z1)) : 0.f;
    const bool z_asc = __all_sync(0xffffffffu, (z0 <= s_z[lane + 1]) && (lane + 33 > 63 || z1 <= s_z[lane + 33]));
    bool asc = true;   // samples ascending and free of NaNs
    const float total = aten_row_sum_ct<NW>(s_w, lane);
    const float2 wp = *reinterpret_cast<const float2*>(s_w + 2 * lane);
    const double p0 = (2 * lane < NW) ? static_cast<double>(__fdiv_rn(wp.x, total)) : 0.0;
    const double p1 = (2 * lane + 1 < NW) ? static_cast<double>(__fdiv_rn(wp.y, total)) : 0.0;
    double p = p0 + p1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double q = shfl_up_f64_(p, o);
      if (lane >= o) p += q;
    }
    double excl = shfl_up_f64_(p, 1);
    if (lane == 0) excl = 0.0;
    __syncwarp();
    s_cb[lane].y = b0;
    if (lane + 32 < NB) s_cb[lane + 32].y = b1;
    if (lane == 0) s_cb[0].x = 0.0f;
    if (2 * lane < NW) s_cb[2 * lane + 1].x = static_cast<float>(excl + p0);
    if (2 * lane + 1 < NW) s_cb[2 * lane + 2].x = static_cast<float>(excl + (p0 + p1));
    __syncwarp();
    float sv[NIT];
    int cle[NIT];      // #{coarse z <= sample}
    double sum = 0.0;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const float uk = ureg[it];
      int pos = 0;
#pragma unroll
      for (int st = P2; st > 0; st >>= 1) {
        if (s_cb[pos + st - 1].x <= uk) pos += st;
      }
      const int below = max(pos - 1, 0);
      const int above = min(pos, NB - 1);
      const float2 lo = s_cb[below], hi = s_cb[above];
      float denom = __fsub_rn(hi.x, lo.x);
      if (denom < 1e-5f) denom = 1.0f;
      const float t = div_zero_fast(__fsub_rn(uk, lo.x), denom);
      const float sm_ = __fadd_rn(lo.y, __fmul_rn(t, __fsub_rn(hi.y, lo.y)));
      sv[it] = sm_;
      sum += static_cast<double>(sm_);
      s_o[it * 32 + lane] = sm_;
      if (samples_out != nullptr) samples_out[ray * Ni + it * 32 + lane] = sm_;
      if (inds_out != nullptr) inds_out[ray * Ni + it * 32 + lane] = pos;
      // coarse depths not above the sample: z_0..z_below lie below mid_below <= sample; then a local scan
      int c = below + 1;
      while (c < NS && s_z[c] <= sm_) ++c;
      while (c > 0 && s_z[c - 1] > sm_) --c;
      cle[it] = c;
    }
    __syncwarp();
    // samples ascending?  (sample k lives at s_o[k], k = it*32 + lane)
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int k = it * 32 + lane;
      if (k + 1 < Ni) asc = asc && (sv[it] <= s_o[k + 1]);
      asc = asc && (sv[it] == sv[it]);   // a NaN sample fails both this and the no_nan test: general path
    }
    if (z_std != nullptr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        int lo_ = __double2loint(sum), hi_ = __double2hiint(sum);
        lo_ = __shfl_xor_sync(0xffffffffu, lo_, o);
        hi_ = __shfl_xor_sync(0xffffffffu, hi_, o);
        sum += __hiloint2double(hi_, lo_);
      }
      const double mean = sum / Ni;
      double ss = 0.0;
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const double d = static_cast<double>(sv[it]) - mean;
        ss += d * d;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        int lo_ = __double2loint(ss), hi_ = __double2hiint(ss);
        lo_ = __shfl_xor_sync(0xffffffffu, lo_, o);
        hi_ = __shfl_xor_sync(0xffffffffu, hi_, o);
        ss += __hiloint2double(hi_, lo_);
      }
      if (lane == 0) z_std[ray] = static_cast<float>(sqrt(ss / Ni));
    }
    float* const out = z_out + ray * NO;
    const bool s_asc = __all_sync(0xffffffffu, asc);
    bool s_nan = false;
#pragma unroll
    for (int it = 0; it < NIT; ++it) s_nan = s_nan || (sv[it] != sv[it]);
    const bool no_nan = !__any_sync(0xffffffffu, s_nan);
    if (z_asc && s_asc) {
      // stable rank merge, coarse depths first on ties: rank(z_i) = i + #{s < z_i}, rank(s_k) = k + #{z <= s_k}
#pragma unroll
      for (int it = 0; it < NIT; ++it) out[it * 32 + lane + cle[it]] = sv[it];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = lane + 32 * h;
        const float v = h ? z1 : z0;
        int lo_ = 0, hi_ = Ni;   // first sample >= v
        while (lo_ < hi_) {
          const int mid = (lo_ + hi_) >> 1;
          if (s_o[mid] < v) lo_ = mid + 1; else hi_ = mid;
        }
        out[i + lo_] = v;
      }
    } else if (z_asc && no_nan) {
      // stochastic u (create_data, training-style renders): the samples come out unordered.  Sort the Ni samples
      // alone (bitonic, Ni is a power of two), then the same rank merge with searches on both sides.
      __syncwarp();
      for (int k = 2; k <= Ni; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
          for (int it = 0; it < NIT; ++it) {
            const int i = it * 32 + lane;
            const int ixj = i ^ j;
            if (ixj > i) {
              const float a = s_o[i], b = s_o[ixj];
              const bool up = ((i & k) == 0);
              if ((a > b) == up) {
                s_o[i] = b;
                s_o[ixj] = a;
              }
            }
          }
          __syncwarp();
        }
      }
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int k = it * 32 + lane;
        const float v = s_o[k];
        int lo_ = 0, hi_ = NS;   // first coarse depth > v
        while (lo_ < hi_) {
          const int mid = (lo_ + hi_) >> 1;
          if (s_z[mid] <= v) lo_ = mid + 1; else hi_ = mid;
        }
        out[k + lo_] = v;
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = lane + 32 * h;
        const float v = h ? z1 : z0;
        int lo_ = 0, hi_ = Ni;   // first sample >= v
        while (lo_ < hi_) {
          const int mid = (lo_ + hi_) >> 1;
          if (s_o[mid] < v) lo_ = mid + 1; else hi_ = mid;
        }
        out[i + lo_] = v;
      }
    } else {
      // general path: bitonic sort of [samples, z, +inf padding] in shared memory (values only)
      __syncwarp();
      s_o[Ni + lane] = z0;
      s_o[Ni + 32 + lane] = z1;
      for (int j = NO + lane; j < PSORT; j += 32) s_o[j] = __int_as_float(0x7f800000);
      __syncwarp();
      for (int k = 2; k <= PSORT; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int i = lane; i < PSORT; i += 32) {
            const int ixj = i ^ j;
            if (ixj > i) {
              const float a = s_o[i], b = s_o[ixj];
              const bool up = ((i & k) == 0);
              if ((a > b) == up) {
                s_o[i] = b;
                s_o[ixj] = a;
              }
            }
          }
          __syncwarp();
        }
      }
      for (int j = lane; j < NO; j += 32) out[j] = s_o[j];
    }
    __syncwarp();
  }
}

// ---- sample_pdf for an ASCENDING shared u table (det=True), 63 bins: the search-free scheme of
// hier_sample_det_kernel below (see its header) without the merge: inds from marks + a max-scan, NIT consecutive
// samples per lane, 16-byte stores.  Same arithmetic as sample_pdf_fast_kernel: samples and inds are bit-identical.
template <int NIT>
__global__ void __launch_bounds__(kPdfWarps * 32)
sample_pdf_det_kernel(long long n_rays, const float* __restrict__ bins, long long bins_stride,
                      const float* __restrict__ weights, long long w_stride, const float* __restrict__ u,
                      float* __restrict__ samples, long long* __restrict__ inds_out) {
  constexpr int NB = 63, NW = 62, Ni = 32 * NIT, P2 = 32;
  constexpr unsigned FULL = 0xffffffffu;
  const float kInf = __int_as_float(0x7f800000);
  __shared__ __align__(16) float s_up[Ni + 4];               // [0] = -inf, [k + 1] = u_k, [Ni + 1..] = +inf
  __shared__ __align__(16) float s_w_all[kPdfWarps][64];
  __shared__ __align__(16) float2 s_cb_all[kPdfWarps][64];   // (cdf_j, bins_j), j < 63; [63] = (+inf, 0)
  __shared__ __align__(16) int s_i_all[kPdfWarps][Ni + 4];   // marks for inds (index r in [0, Ni])
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  float* const s_w = s_w_all[wib];
  float2* const s_cb = s_cb_all[wib];
  int* const s_i = s_i_all[wib];
  for (int j = threadIdx.x; j < Ni; j += kPdfWarps * 32) s_up[j + 1] = __ldg(u + j);
  if (threadIdx.x == 0) s_up[0] = -kInf;
  if (threadIdx.x < 3) s_up[Ni + 1 + threadIdx.x] = kInf;
  for (int j = lane; j < Ni + 4; j += 32) s_i[j] = 0;
  if (lane == 0) s_cb[63] = make_float2(kInf, 0.f);
  __syncthreads();
  float ureg[NIT];
#pragma unroll
  for (int i = 0; i < NIT; ++i) ureg[i] = s_up[NIT * lane + i + 1];
  const long long warp0 = static_cast<long long>(blockIdx.x) * kPdfWarps + wib;
  const long long nwarps = static_cast<long long>(gridDim.x) * kPdfWarps;
  // the next ray's row is fetched while this one is processed (one ray per warp in flight otherwise)
  float nb0 = 0.f, nb1 = 0.f, nw0 = 0.f, nw1 = 0.f;
  if (warp0 < n_rays) {
    nb0 = __ldg(bins + warp0 * bins_stride + lane);
    if (lane + 32 < NB) nb1 = __ldg(bins + warp0 * bins_stride + lane + 32);
    nw0 = __ldg(weights + warp0 * w_stride + lane);
    if (lane + 32 < NW) nw1 = __ldg(weights + warp0 * w_stride + lane + 32);
  }
  for (long long ray = warp0; ray < n_rays; ray += nwarps) {
    const float b0 = nb0, b1 = nb1;
    s_w[lane] = __fadd_rn(nw0, 1e-5f);
    if (lane + 32 < NW) s_w[lane + 32] = __fadd_rn(nw1, 1e-5f);
    if (ray + nwarps < n_rays) {
      const float* rb = bins + (ray + nwarps) * bins_stride;
      const float* rw = weights + (ray + nwarps) * w_stride;
      nb0 = __ldg(rb + lane);
      if (lane + 32 < NB) nb1 = __ldg(rb + lane + 32);
      nw0 = __ldg(rw + lane);
      if (lane + 32 < NW) nw1 = __ldg(rw + lane + 32);
    }
    __syncwarp();
    const float total = aten_row_sum_ct<NW>(s_w, lane);
    const float2 wp = *reinterpret_cast<const float2*>(s_w + 2 * lane);
    const double p0 = (2 * lane < NW) ? static_cast<double>(__fdiv_rn(wp.x, total)) : 0.0;
    const double p1 = (2 * lane + 1 < NW) ? static_cast<double>(__fdiv_rn(wp.y, total)) : 0.0;
    double p = p0 + p1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double q = shfl_up_f64_(p, o);
      if (lane >= o) p += q;
    }
    double excl = shfl_up_f64_(p, 1);
    if (lane == 0) excl = 0.0;
    const float c1 = static_cast<float>(excl + p0);          // cdf[2 lane + 1]
    const float c2 = static_cast<float>(excl + (p0 + p1));   // cdf[2 lane + 2]
    const bool mono = __all_sync(FULL, (p0 >= 0.0) && (p1 >= 0.0) && (c2 < kInf));
    __syncwarp();
    s_cb[lane].y = b0;
    if (lane + 32 < NB) s_cb[lane + 32].y = b1;
    if (lane == 0) s_cb[0].x = 0.0f;
    if (2 * lane < NW) s_cb[2 * lane + 1].x = c1;
    if (2 * lane + 1 < NW) s_cb[2 * lane + 2].x = c2;
    int pos[NIT];
    if (mono) {
      int gA = min(max(static_cast<int>(ceilf(c1 * static_cast<float>(Ni - 1))), 0), Ni);
      int gB = min(max(static_cast<int>(ceilf(c2 * static_cast<float>(Ni - 1))), 0), Ni);
      gA += (s_up[gA + 1] < c1 ? 1 : 0) - (s_up[gA] >= c1 ? 1 : 0);
      gB += (s_up[gB + 1] < c2 ? 1 : 0) - (s_up[gB] >= c2 ? 1 : 0);
      const bool off = !(s_up[gA] < c1) || !(s_up[gA + 1] >= c1) || !(s_up[gB] < c2) || !(s_up[gB + 1] >= c2);
      if (__any_sync(FULL, off)) {   // not a linspace (ties, clusters): walk
        while (gA > 0 && s_up[gA] >= c1) --gA;
        while (gA < Ni && s_up[gA + 1] < c1) ++gA;
        while (gB > 0 && s_up[gB] >= c2) --gB;
        while (gB < Ni && s_up[gB + 1] < c2) ++gB;
        __syncwarp();
      }
      const int rA = (lane < 31) ? gA : Ni, rB = (lane < 31) ? gB : Ni;
      const int rN = __shfl_down_sync(FULL, rA, 1);          // lane 30 sees lane 31's Ni: "last entry"
      if (lane == 0 && rA != 0) s_i[0] = 1;                  // cdf_0 = 0: r_0 = 0
      if (rA != rB) s_i[rA] = 2 * lane + 2;
      if (lane < 31 && rB != rN) s_i[rB] = 2 * lane + 3;
      __syncwarp();
      int m[NIT];
#pragma unroll
      for (int i = 0; i < NIT; ++i) {
        m[i] = s_i[NIT * lane + i];
        if (i > 0) m[i] = max(m[i], m[i - 1]);
      }
#pragma unroll
      for (int i = 0; i < NIT; ++i) s_i[NIT * lane + i] = 0;
      int incl = m[NIT - 1];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int q = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl = max(incl, q);
      }
      int ex = __shfl_up_sync(FULL, incl, 1);
      if (lane == 0) ex = 0;
#pragma unroll
      for (int i = 0; i < NIT; ++i) pos[i] = max(m[i], ex);
    } else {
      __syncwarp();
#pragma unroll
      for (int i = 0; i < NIT; ++i) {
        int q = 0;
#pragma unroll
        for (int st = P2; st > 0; st >>= 1) {
          if (s_cb[q + st - 1].x <= ureg[i]) q += st;
        }
        pos[i] = q;
      }
    }
    float sv[NIT];
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      const int below = max(pos[i] - 1, 0);
      const int above = min(pos[i], NB - 1);
      const float2 lo = s_cb[below], hi = s_cb[above];
      float denom = __fsub_rn(hi.x, lo.x);
      if (denom < 1e-5f) denom = 1.0f;
      const float t = div_zero_fast(__fsub_rn(ureg[i], lo.x), denom);
      sv[i] = __fadd_rn(lo.y, __fmul_rn(t, __fsub_rn(hi.y, lo.y)));
    }
    float* so = samples + ray * Ni + NIT * lane;
    if constexpr (NIT == 4) *reinterpret_cast<float4*>(so) = make_float4(sv[0], sv[1], sv[2], sv[3]);
    else *reinterpret_cast<float2*>(so) = make_float2(sv[0], sv[1]);
    if (inds_out != nullptr) {
      long long* io = inds_out + ray * Ni + NIT * lane;
#pragma unroll
      for (int i = 0; i < NIT; i += 2) *reinterpret_cast<longlong2*>(io + i) = make_longlong2(pos[i], pos[i + 1]);
    }
    __syncwarp();
  }
}

// ---- the same fused step for an ASCENDING shared u table (det=True: linspace(0,1,Ni) — every render) ---------------
// hier_sample_kernel<4> executes 975 warp instructions per ray and is issue-bound (ncu r1g: issue slots 76 % busy at
// 13 % of the DRAM roofline): 4 x 6-step bisections, divergent local scans for the merge positions and two 7-step
// searches of the coarse depths.  With u ascending none of the searches is needed:
//   * inds_k = #{j : cdf_j <= u_k} = #{j : r_j <= k} with r_j = #{k : u_k < cdf_j}.  The lane that owns cdf_j gets r_j
//     from ceil(cdf_j (Ni-1)) corrected against the real table (exact for ANY ascending table, O(1) for a linspace),
//     the last j of every run of equal r marks s_i[r_j] = j+1, and an inclusive max-scan over k turns the marks into
//     inds (cdf ascending => r ascending => one writer per address, no atomics).
//   * a lane owns NIT CONSECUTIVE samples: neighbours compare in registers, samples / inds leave as 16-byte stores.
//   * merged position of sample k = k + cle_k, cle_k = #{z <= s_k} = below+1+(z[below+1] <= s_k), verified with two
//     more compares (any ascending z is still handled exactly: a failed check falls back to the scan);
//   * merged position of z_i = i + #{s_k < z_i} = i + #{k : cle_k <= i}: the same mark + max-scan trick over cle.
// Arithmetic and results are those of hier_sample_kernel (samples, inds and merged depths bit-identical; z_std sums the
// same doubles in another order: equal to the last ulp of fp32 except when a rounding boundary is hit).  Rows with a
// non-monotone or non-finite cdf (negative / NaN weights) take the bisection; rows whose z or samples are not
// ascending take the bitonic sort.
template <int NIT>
__global__ void __launch_bounds__(kPdfWarps * 32, 5)
hier_sample_det_kernel(long long n_rays, const float* __restrict__ z_vals, const float* __restrict__ weights,
                       const float* __restrict__ u, float* __restrict__ z_out, float* __restrict__ z_std,
                       float* __restrict__ samples_out, long long* __restrict__ inds_out) {
  constexpr int NS = 64, NB = 63, NW = 62, Ni = 32 * NIT, NO = NS + Ni, P2 = 32;
  constexpr int PSORT = (NO <= 128) ? 128 : 256;
  constexpr unsigned FULL = 0xffffffffu;
  const float kInf = __int_as_float(0x7f800000);
  __shared__ __align__(16) float s_up[Ni + 4];               // [0] = -inf, [k + 1] = u_k, [Ni + 1..] = +inf
  __shared__ __align__(16) float s_w_all[kPdfWarps][64];
  __shared__ __align__(16) float2 s_cb_all[kPdfWarps][64];   // (cdf_j, bins_j), j < 63; [63] = (+inf, 0)
  __shared__ __align__(16) float s_z_all[kPdfWarps][68];     // [64..67] = +inf
  __shared__ __align__(16) int s_i_all[kPdfWarps][Ni + 4];   // marks for inds (index r in [0, Ni])
  __shared__ __align__(16) int s_c_all[kPdfWarps][68];       // marks for the coarse ranks (index cle in [0, 64])
  __shared__ __align__(16) float s_o_all[kPdfWarps][PSORT];  // bitonic fallback only
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  float* const s_w = s_w_all[wib];
  float2* const s_cb = s_cb_all[wib];
  float* const s_z = s_z_all[wib];
  int* const s_i = s_i_all[wib];
  int* const s_c = s_c_all[wib];
  float* const s_o = s_o_all[wib];
  for (int j = threadIdx.x; j < Ni; j += kPdfWarps * 32) s_up[j + 1] = __ldg(u + j);
  if (threadIdx.x == 0) s_up[0] = -kInf;
  if (threadIdx.x < 3) s_up[Ni + 1 + threadIdx.x] = kInf;
  for (int j = lane; j < Ni + 4; j += 32) s_i[j] = 0;
  for (int j = lane; j < 68; j += 32) s_c[j] = 0;
  if (lane < 4) s_z[64 + lane] = kInf;
  if (lane == 0) s_cb[63] = make_float2(kInf, 0.f);
  __syncthreads();
  float ureg[NIT];
#pragma unroll
  for (int i = 0; i < NIT; ++i) ureg[i] = s_up[NIT * lane + i + 1];
  const long long warp0 = static_cast<long long>(blockIdx.x) * kPdfWarps + wib;
  const long long nwarps = static_cast<long long>(gridDim.x) * kPdfWarps;
  // the next ray's row is fetched while this one is processed (one ray per warp in flight otherwise)
  float nz0 = 0.f, nz1 = 0.f, nw0 = 0.f, nw1 = 0.f;
  if (warp0 < n_rays) {
    nz0 = __ldg(z_vals + warp0 * NS + lane), nz1 = __ldg(z_vals + warp0 * NS + lane + 32);
    nw0 = __ldg(weights + warp0 * NS + 1 + lane);
    if (lane + 32 < NW) nw1 = __ldg(weights + warp0 * NS + 33 + lane);
  }
  for (long long ray = warp0; ray < n_rays; ray += nwarps) {
    const float z0 = nz0, z1 = nz1;
    s_z[lane] = z0;
    s_z[lane + 32] = z1;
    s_w[lane] = __fadd_rn(nw0, 1e-5f);
    if (lane + 32 < NW) s_w[lane + 32] = __fadd_rn(nw1, 1e-5f);
    if (ray + nwarps < n_rays) {
      const float* rz = z_vals + (ray + nwarps) * NS;
      nz0 = __ldg(rz + lane), nz1 = __ldg(rz + lane + 32);
      nw0 = __ldg(weights + (ray + nwarps) * NS + 1 + lane);   // weights[..., 1:-1]
      if (lane + 32 < NW) nw1 = __ldg(weights + (ray + nwarps) * NS + 33 + lane);
    }
    __syncwarp();
    const float zn0 = s_z[lane + 1], zn1 = s_z[lane + 33];   // [64] = +inf
    const float b0 = __fmul_rn(0.5f, __fadd_rn(zn0, z0));
    const float b1 = (lane + 32 < NB) ? __fmul_rn(0.5f, __fadd_rn(zn1, z1)) : 0.f;
    const bool z_asc = __all_sync(FULL, (z0 <= zn0) && (lane == 31 || z1 <= zn1));
    const float total = aten_row_sum_ct<NW>(s_w, lane);
    const float2 wp = *reinterpret_cast<const float2*>(s_w + 2 * lane);
    const double p0 = (2 * lane < NW) ? static_cast<double>(__fdiv_rn(wp.x, total)) : 0.0;
    const double p1 = (2 * lane + 1 < NW) ? static_cast<double>(__fdiv_rn(wp.y, total)) : 0.0;
    double p = p0 + p1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double q = shfl_up_f64_(p, o);
      if (lane >= o) p += q;
    }
    double excl = shfl_up_f64_(p, 1);
    if (lane == 0) excl = 0.0;
    const float c1 = static_cast<float>(excl + p0);          // cdf[2 lane + 1]
    const float c2 = static_cast<float>(excl + (p0 + p1));   // cdf[2 lane + 2]
    const bool mono = __all_sync(FULL, (p0 >= 0.0) && (p1 >= 0.0) && (c2 < kInf));
    __syncwarp();
    s_cb[lane].y = b0;
    if (lane + 32 < NB) s_cb[lane + 32].y = b1;
    if (lane == 0) s_cb[0].x = 0.0f;
    if (2 * lane < NW) s_cb[2 * lane + 1].x = c1;
    if (2 * lane + 1 < NW) s_cb[2 * lane + 2].x = c2;
    int pos[NIT];
    if (mono) {
      // r = #{k : u_k < c} for this lane's two cdf entries: guess from the linspace, one branch-free correction step,
      // then verified (s_up[k + 1] = u_k, s_up[0] = -inf, s_up[Ni + 1..] = +inf: no bounds checks).  Lane 31 owns no
      // entry: it runs the same code on its (valid) copies of the last cdf value and is overridden below.
      int gA = min(max(static_cast<int>(ceilf(c1 * static_cast<float>(Ni - 1))), 0), Ni);
      int gB = min(max(static_cast<int>(ceilf(c2 * static_cast<float>(Ni - 1))), 0), Ni);
      gA += (s_up[gA + 1] < c1 ? 1 : 0) - (s_up[gA] >= c1 ? 1 : 0);
      gB += (s_up[gB + 1] < c2 ? 1 : 0) - (s_up[gB] >= c2 ? 1 : 0);
      const bool off = !(s_up[gA] < c1) || !(s_up[gA + 1] >= c1) || !(s_up[gB] < c2) || !(s_up[gB + 1] >= c2);
      if (__any_sync(FULL, off)) {   // not a linspace (ties, clusters): walk
        while (gA > 0 && s_up[gA] >= c1) --gA;
        while (gA < Ni && s_up[gA + 1] < c1) ++gA;
        while (gB > 0 && s_up[gB] >= c2) --gB;
        while (gB < Ni && s_up[gB + 1] < c2) ++gB;
        __syncwarp();
      }
      const int rA = (lane < 31) ? gA : Ni, rB = (lane < 31) ? gB : Ni;
      const int rN = __shfl_down_sync(FULL, rA, 1);          // lane 30 sees lane 31's Ni: "last entry"
      if (lane == 0 && rA != 0) s_i[0] = 1;                  // cdf_0 = 0: r_0 = 0
      if (rA != rB) s_i[rA] = 2 * lane + 2;
      if (lane < 31 && rB != rN) s_i[rB] = 2 * lane + 3;
      __syncwarp();
      int m[NIT];
#pragma unroll
      for (int i = 0; i < NIT; ++i) {
        m[i] = s_i[NIT * lane + i];
        if (i > 0) m[i] = max(m[i], m[i - 1]);
      }
#pragma unroll
      for (int i = 0; i < NIT; ++i) s_i[NIT * lane + i] = 0;
      if (lane == 0) s_i[Ni] = 0;
      int incl = m[NIT - 1];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int q = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl = max(incl, q);
      }
      int ex = __shfl_up_sync(FULL, incl, 1);
      if (lane == 0) ex = 0;
#pragma unroll
      for (int i = 0; i < NIT; ++i) pos[i] = max(m[i], ex);
    } else {
      __syncwarp();
#pragma unroll
      for (int i = 0; i < NIT; ++i) {
        int q = 0;
#pragma unroll
        for (int st = P2; st > 0; st >>= 1) {
          if (s_cb[q + st - 1].x <= ureg[i]) q += st;
        }
        pos[i] = q;
      }
    }
    float sv[NIT];
    int cle[NIT];      // #{coarse z <= sample}
    double sum = 0.0;
    bool bad = false;
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      const float uk = ureg[i];
      const int below = max(pos[i] - 1, 0);
      const int above = min(pos[i], NB - 1);
      const float2 lo = s_cb[below], hi = s_cb[above];
      float denom = __fsub_rn(hi.x, lo.x);
      if (denom < 1e-5f) denom = 1.0f;
      const float t = div_zero_fast(__fsub_rn(uk, lo.x), denom);
      const float sm_ = __fadd_rn(lo.y, __fmul_rn(t, __fsub_rn(hi.y, lo.y)));
      sv[i] = sm_;
      sum += static_cast<double>(sm_);
      const bool a = s_z[below] <= sm_, b = s_z[below + 1] <= sm_, d = s_z[below + 2] <= sm_;
      cle[i] = below + 1 + (b ? 1 : 0);
      bad = bad || !a || d;
    }
    if (__any_sync(FULL, bad)) {
#pragma unroll
      for (int i = 0; i < NIT; ++i) {
        int c = cle[i];
        while (c < NS && s_z[c] <= sv[i]) ++c;
        while (c > 0 && s_z[c - 1] > sv[i]) --c;
        cle[i] = c;
      }
    }
    if (samples_out != nullptr) {
      float* so = samples_out + ray * Ni + NIT * lane;
      if constexpr (NIT == 4) *reinterpret_cast<float4*>(so) = make_float4(sv[0], sv[1], sv[2], sv[3]);
      else *reinterpret_cast<float2*>(so) = make_float2(sv[0], sv[1]);
    }
    if (inds_out != nullptr) {
      long long* io = inds_out + ray * Ni + NIT * lane;
#pragma unroll
      for (int i = 0; i < NIT; i += 2) *reinterpret_cast<longlong2*>(io + i) = make_longlong2(pos[i], pos[i + 1]);
    }
    // samples ascending and free of NaNs?
    const float s_next = __shfl_down_sync(FULL, sv[0], 1);
    bool asc = (lane == 31) || (sv[NIT - 1] <= s_next);
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      if (i + 1 < NIT) asc = asc && (sv[i] <= sv[i + 1]);
      asc = asc && (sv[i] == sv[i]);
    }
    if (z_std != nullptr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        int lo_ = __double2loint(sum), hi_ = __double2hiint(sum);
        lo_ = __shfl_xor_sync(FULL, lo_, o);
        hi_ = __shfl_xor_sync(FULL, hi_, o);
        sum += __hiloint2double(hi_, lo_);
      }
      const double mean = sum / Ni;
      double ss = 0.0;
#pragma unroll
      for (int i = 0; i < NIT; ++i) {
        const double dd = static_cast<double>(sv[i]) - mean;
        ss += dd * dd;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        int lo_ = __double2loint(ss), hi_ = __double2hiint(ss);
        lo_ = __shfl_xor_sync(FULL, lo_, o);
        hi_ = __shfl_xor_sync(FULL, hi_, o);
        ss += __hiloint2double(hi_, lo_);
      }
      if (lane == 0) z_std[ray] = static_cast<float>(sqrt(ss / Ni));
    }
    float* const out = z_out + ray * NO;
    if (z_asc && __all_sync(FULL, asc)) {
      // stable rank merge, coarse depths first on ties: rank(s_k) = k + #{z <= s_k}, rank(z_i) = i + #{s < z_i}
#pragma unroll
      for (int i = 0; i < NIT; ++i) out[NIT * lane + i + cle[i]] = sv[i];
      const int c_next = __shfl_down_sync(FULL, cle[0], 1);
#pragma unroll
      for (int i = 0; i < NIT; ++i) {
        const int nxt = (i + 1 < NIT) ? cle[(i + 1) % NIT] : c_next;
        if ((i == NIT - 1 && lane == 31) || cle[i] != nxt) s_c[cle[i]] = NIT * lane + i + 1;
      }
      __syncwarp();
      const int2 mk = *reinterpret_cast<const int2*>(s_c + 2 * lane);
      *reinterpret_cast<int2*>(s_c + 2 * lane) = make_int2(0, 0);
      const int m1 = max(mk.x, mk.y);
      int incl = m1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int q = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl = max(incl, q);
      }
      int ex = __shfl_up_sync(FULL, incl, 1);
      if (lane == 0) ex = 0;
      const float2 zz = *reinterpret_cast<const float2*>(s_z + 2 * lane);
      out[2 * lane + max(ex, mk.x)] = zz.x;
      out[2 * lane + 1 + max(ex, m1)] = zz.y;
    } else {
      // general path: bitonic sort of [samples, z, +inf padding] in shared memory (values only)
#pragma unroll
      for (int i = 0; i < NIT; ++i) s_o[NIT * lane + i] = sv[i];
      s_o[Ni + lane] = z0;
      s_o[Ni + 32 + lane] = z1;
      for (int j = NO + lane; j < PSORT; j += 32) s_o[j] = kInf;
      __syncwarp();
      for (int k = 2; k <= PSORT; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int i = lane; i < PSORT; i += 32) {
            const int ixj = i ^ j;
            if (ixj > i) {
              const float a = s_o[i], b = s_o[ixj];
              const bool up = ((i & k) == 0);
              if ((a > b) == up) {
                s_o[i] = b;
                s_o[ixj] = a;
              }
            }
          }
          __syncwarp();
        }
      }
      for (int j = lane; j < NO; j += 32) out[j] = s_o[j];
    }
    __syncwarp();
  }
}

// z_out = sort(cat[z_a, z_b]) per ray (values only, ascending; main.py:730-732) and
// z_std = std(z_b, unbiased=False) (main.py:750).  Warp-level bitonic sort in shared memory.
__global__ void __launch_bounds__(kPdfWarps * 32)
merge_sort_kernel(long long n_rays, int na, int nbv, const float* __restrict__ za, const float* __restrict__ zb,
                  float* __restrict__ z_out, float* __restrict__ z_std, int P) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  float* s = sm + static_cast<size_t>(wib) * P;
  const int n = na + nbv;
  const long long warp0 = static_cast<long long>(blockIdx.x) * kPdfWarps + wib;
  const long long nwarps = static_cast<long long>(gridDim.x) * kPdfWarps;
  for (long long ray = warp0; ray < n_rays; ray += nwarps) {
    for (int j = lane; j < na; j += 32) s[j] = __ldg(za + ray * na + j);
    double sum = 0.0;
    for (int j = lane; j < nbv; j += 32) {
      const float v = __ldg(zb + ray * nbv + j);
      s[na + j] = v;
      sum += static_cast<double>(v);
    }
    for (int j = n + lane; j < P; j += 32) s[j] = __int_as_float(0x7f800000);  // +inf padding
    __syncwarp();
    if (z_std != nullptr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        int lo = __double2loint(sum), hi = __double2hiint(sum);
        lo = __shfl_xor_sync(0xffffffffu, lo, o);
        hi = __shfl_xor_sync(0xffffffffu, hi, o);
        sum += __hiloint2double(hi, lo);
      }
      const double mean = sum / nbv;
      double ss = 0.0;
      for (int j = lane; j < nbv; j += 32) {
        const double d = static_cast<double>(s[na + j]) - mean;
        ss += d * d;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        int lo = __double2loint(ss), hi = __double2hiint(ss);
        lo = __shfl_xor_sync(0xffffffffu, lo, o);
        hi = __shfl_xor_sync(0xffffffffu, hi, o);
        ss += __hiloint2double(hi, lo);
      }
      if (lane == 0) z_std[ray] = static_cast<float>(sqrt(ss / nbv));
    }
    // Fast path (every deterministic render, and stratified z_a): both inputs already ascending -> stable merge
    // by rank: rank(a_i) = i + #{b < a_i}, rank(b_j) = j + #{a <= b_j} (binary searches in shared memory).
    // The result of a values-only sort does not depend on the algorithm, so this is bit-identical to the
    // general path below.
    bool sorted = true;
    for (int j = lane; j + 1 < na; j += 32) sorted = sorted && (s[j] <= s[j + 1]);
    for (int j = lane; j + 1 < nbv; j += 32) sorted = sorted && (s[na + j] <= s[na + j + 1]);
    if (__all_sync(0xffffffffu, sorted)) {
      const float* a = s;
      const float* b = s + na;
      float* out = z_out + ray * n;
      for (int i = lane; i < na; i += 32) {
        const float v = a[i];
        int lo = 0, hi = nbv;   // first index with b[idx] >= v
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (b[mid] < v) lo = mid + 1; else hi = mid;
        }
        out[i + lo] = v;
      }
      for (int j = lane; j < nbv; j += 32) {
        const float v = b[j];
        int lo = 0, hi = na;    // first index with a[idx] > v
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (a[mid] <= v) lo = mid + 1; else hi = mid;
        }
        out[j + lo] = v;
      }
      __syncwarp();
      continue;
    }
    // general path: bitonic sort of P elements
    for (int k = 2; k <= P; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = lane; i < P; i += 32) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const float a = s[i], b = s[ixj];
            const bool up = ((i & k) == 0);
            if ((a > b) == up) {
              s[i] = b;
              s[ixj] = a;
            }
          }
        }
        __syncwarp();
      }
    }
    for (int j = lane; j < n; j += 32) z_out[ray * n + j] = s[j];
    __syncwarp();
  }
}

}  // namespace r2l

using namespace r2l;

extern "C" {

int r2l_sample_pdf(long long n_rays, int nb, int Ni, const float* bins, long long bins_stride,
                   const float* weights, long long w_stride, const float* u, int u_per_ray, float* samples,
                   long long* inds_out, void* stream) {
  R2L_CHECK_ARG(n_rays >= 0 && nb >= 2 && Ni > 0, "r2l_sample_pdf: bad sizes");
  R2L_CHECK_ARG(nb - 1 < 512, "r2l_sample_pdf: rows of >= 512 weights are not supported (ATen cascade sum)");
  if (n_rays == 0) return R2L_OK;
  R2L_CHECK_ARG(bins && weights && u && samples, "r2l_sample_pdf: null pointer");
  R2L_CHECK_ARG(bins_stride >= nb && w_stride >= nb - 1, "r2l_sample_pdf: bad strides");
  const int row_pad = ((nb + 31) / 32) * 32;
  const size_t smem = static_cast<size_t>(kPdfWarps) * 3 * row_pad * sizeof(float);
  long long blocks = (n_rays + kPdfWarps - 1) / kPdfWarps;
  const long long cap = static_cast<long long>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  R2L_CHECK_ARG(u_per_ray >= 0 && u_per_ray <= 2, "r2l_sample_pdf: u_per_ray must be 0 (shared table), 1 (per ray) or 2 "
                "(shared ASCENDING table)");
  if (nb == 63 && (Ni == 128 || Ni == 64) && u_per_ray == 2 && (reinterpret_cast<uintptr_t>(samples) & 15) == 0 &&
      (inds_out == nullptr || (reinterpret_cast<uintptr_t>(inds_out) & 15) == 0)) {   // det=True: search-free kernel
    auto st = static_cast<cudaStream_t>(stream);
    if (Ni == 128)
      sample_pdf_det_kernel<4><<<static_cast<int>(blocks), kPdfWarps * 32, 0, st>>>(n_rays, bins, bins_stride, weights,
                                                                                   w_stride, u, samples, inds_out);
    else
      sample_pdf_det_kernel<2><<<static_cast<int>(blocks), kPdfWarps * 32, 0, st>>>(n_rays, bins, bins_stride, weights,
                                                                                   w_stride, u, samples, inds_out);
    R2L_LAUNCH_CHECK();
    return R2L_OK;
  }
  if (u_per_ray == 2) u_per_ray = 0;   // other sizes: an ascending table is just a shared table
  if (nb == 63 && (Ni == 128 || Ni == 64)) {   // 64 coarse samples: every BASELINE config
    auto st = static_cast<cudaStream_t>(stream);
    if (Ni == 128)
      sample_pdf_fast_kernel<63, 4><<<static_cast<int>(blocks), kPdfWarps * 32, 0, st>>>(
          n_rays, bins, bins_stride, weights, w_stride, u, u_per_ray, samples, inds_out);
    else
      sample_pdf_fast_kernel<63, 2><<<static_cast<int>(blocks), kPdfWarps * 32, 0, st>>>(
          n_rays, bins, bins_stride, weights, w_stride, u, u_per_ray, samples, inds_out);
    R2L_LAUNCH_CHECK();
    return R2L_OK;
  }
  if (smem > 48 * 1024)
    R2L_CUDA(cudaFuncSetAttribute(sample_pdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sample_pdf_kernel<<<static_cast<int>(blocks), kPdfWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
      n_rays, nb, Ni, bins, bins_stride, weights, w_stride, u, u_per_ray, samples, inds_out, row_pad);
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

int r2l_hier_sample(long long n_rays, int n_coarse, int Ni, const float* z_vals, const float* weights, const float* u,
                    int u_per_ray, float* z_out, float* z_std, float* samples, long long* inds_out, void* stream) {
  R2L_CHECK_ARG(n_rays >= 0, "r2l_hier_sample: bad sizes");
  R2L_CHECK_ARG(n_coarse == 64 && (Ni == 128 || Ni == 64),
                "r2l_hier_sample: fused path needs 64 coarse samples and 64 or 128 fine samples (got %d, %d); use "
                "r2l_sample_pdf + r2l_merge_sorted", n_coarse, Ni);
  if (n_rays == 0) return R2L_OK;
  R2L_CHECK_ARG(z_vals && weights && u && z_out, "r2l_hier_sample: null pointer");
  R2L_CHECK_ARG(u_per_ray >= 0 && u_per_ray <= 2, "r2l_hier_sample: u_per_ray must be 0 (shared table), 1 (per ray) or "
                "2 (shared ASCENDING table)");
  long long blocks = (n_rays + kPdfWarps - 1) / kPdfWarps;
  const long long cap = static_cast<long long>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  auto st = static_cast<cudaStream_t>(stream);
  const bool aligned = ((reinterpret_cast<uintptr_t>(samples) | reinterpret_cast<uintptr_t>(inds_out)) & 15) == 0;
  if (u_per_ray == 2 && !aligned) u_per_ray = 0;   // the search-free kernel stores 16 bytes per lane
  if (u_per_ray == 2) {   // ONE ascending table shared by all rays (det=True): the search-free kernel
    if (Ni == 128)
      hier_sample_det_kernel<4><<<static_cast<int>(blocks), kPdfWarps * 32, 0, st>>>(n_rays, z_vals, weights, u, z_out,
                                                                                    z_std, samples, inds_out);
    else
      hier_sample_det_kernel<2><<<static_cast<int>(blocks), kPdfWarps * 32, 0, st>>>(n_rays, z_vals, weights, u, z_out,
                                                                                    z_std, samples, inds_out);
    R2L_LAUNCH_CHECK();
    return R2L_OK;
  }
  if (Ni == 128)
    hier_sample_kernel<4><<<static_cast<int>(blocks), kPdfWarps * 32, 0, st>>>(n_rays, z_vals, weights, u, u_per_ray,
                                                                              z_out, z_std, samples, inds_out);
  else
    hier_sample_kernel<2><<<static_cast<int>(blocks), kPdfWarps * 32, 0, st>>>(n_rays, z_vals, weights, u, u_per_ray,
                                                                              z_out, z_std, samples, inds_out);
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

int r2l_merge_sorted(long long n_rays, int na, int nbv, const float* za, const float* zb, float* z_out,
                     float* z_std, void* stream) {
  R2L_CHECK_ARG(n_rays >= 0 && na >= 0 && nbv > 0, "r2l_merge_sorted: bad sizes");
  R2L_CHECK_ARG(na + nbv <= 4096, "r2l_merge_sorted: more than 4096 depths per ray");
  if (n_rays == 0) return R2L_OK;
  R2L_CHECK_ARG((za || na == 0) && zb && z_out, "r2l_merge_sorted: null pointer");
  int P = 32;
  while (P < na + nbv) P <<= 1;
  const size_t smem = static_cast<size_t>(kPdfWarps) * P * sizeof(float);
  long long blocks = (n_rays + kPdfWarps - 1) / kPdfWarps;
  const long long cap = static_cast<long long>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (smem > 48 * 1024)
    R2L_CUDA(cudaFuncSetAttribute(merge_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  merge_sort_kernel<<<static_cast<int>(blocks), kPdfWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
      n_rays, na, nbv, za, zb, z_out, z_std, P);
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

}  // extern "C"
