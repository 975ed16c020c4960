// raw2outputs: alpha compositing along each ray as a warp-level scan.
// Reference: main.py:556-621 (twins: utils/create_data.py:335-402,
// model/nerf_raybased.py:226-295, utils/run_nerf_raybased_helpers.py:77-144).
//
//   dists_i = z_{i+1}-z_i (last = 1e10), times |rays_d|
//   alpha_i = 1 - exp(-relu(sigma_i + noise_i) * dists_i)
//   T_i     = prod_{j<i} (1 - alpha_j + 1e-10)        (reference: ATen CPU cumprod accumulates in double)
//   w_i     = alpha_i * T_i
//   rgb_map = sum w_i * sigmoid(rgb_i);  depth = sum w_i z_i;  acc = sum w_i
//   disp    = 1 / max(1e-10, depth/acc)  with torch.max NaN propagation;  white_bkgd: rgb += 1-acc
//
// One warp per ray; lane l owns samples l, l+32, ... so every load/store is a fully
// coalesced 128/512-byte transaction (raw is read as float4).  The transmittance is an
// exclusive product scan: 5 shuffle steps per 32-sample chunk in double precision plus a
// running carry — double so that the result rounds to the same fp32 value as the
// reference's sequential double-accumulated cumprod.
// Algorithmic HBM bytes per ray: 16S (raw) + 4S (z) + 12 (d) in, 4S (weights) + 24 out.
#include "common.cuh"
#include "composite_math.cuh"

namespace r2l {

constexpr int kCompWarps = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// CH > 0: S <= 32*CH and ALL of a ray's loads (raw as float4, z, noise) are issued before the first scan step,
// so a warp keeps CH*640 bytes in flight instead of one 32-sample chunk; CH == 0: any S, chunk-by-chunk loads.
template <int CH>
__global__ void __launch_bounds__(kCompWarps * 32)
raw2outputs_kernel(long long n_rays, int S, const float4* __restrict__ raw, const float* __restrict__ z_vals,
                   const float* __restrict__ rays_d, long long d_stride, const float* __restrict__ noise,
                   int white_bkgd, float* __restrict__ rgb_map, float* __restrict__ disp_map,
                   float* __restrict__ acc_map, float* __restrict__ weights, float* __restrict__ depth_map) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = static_cast<long long>(blockIdx.x) * kCompWarps + (threadIdx.x >> 5);
  const long long nwarps = static_cast<long long>(gridDim.x) * kCompWarps;
  const int chunks = (CH > 0) ? CH : (S + 31) / 32;
  constexpr int NR = (CH > 0) ? CH : 1;
  for (long long ray = warp0; ray < n_rays; ray += nwarps) {
    const float4* rraw = raw + ray * S;
    const float* rz = z_vals + ray * S;
    float4 rv_[NR];
    float zi_[NR], nz_[NR];
    if (CH > 0) {
#pragma unroll
      for (int c = 0; c < NR; ++c) {
        const int i = c * 32 + lane;
        rv_[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        zi_[c] = 0.f;
        nz_[c] = 0.f;
        if (i < S) {
          rv_[c] = __ldg(rraw + i);
          zi_[c] = __ldg(rz + i);
          if (noise != nullptr) nz_[c] = __ldg(noise + ray * S + i);
        }
      }
    }
    const float dx = rays_d[ray * d_stride], dy = rays_d[ray * d_stride + 1], dz = rays_d[ray * d_stride + 2];
    const float dnorm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    double carry = 1.0;
    float ar = 0.f, ag = 0.f, ab = 0.f, adepth = 0.f, aacc = 0.f;
#pragma unroll
    for (int c = 0; c < chunks; ++c) {
      const int i = c * 32 + lane;
      const bool valid = i < S;
      float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
      float zi = 0.f, zn = 0.f, nz = 0.f;
      if (CH > 0) {
        rv = rv_[c % NR];
        zi = zi_[c % NR];
        nz = nz_[c % NR];
        // z of the next sample: the next lane's value; lane 31 takes lane 0 of the next chunk
        zn = __shfl_down_sync(0xffffffffu, zi, 1);
        const float z_first_next = __shfl_sync(0xffffffffu, (c + 1 < NR) ? zi_[(c + 1) % NR] : 0.f, 0);
        if (lane == 31) zn = z_first_next;
      } else if (valid) {
        rv = __ldg(rraw + i);
        zi = __ldg(rz + i);
        if (i + 1 < S) zn = __ldg(rz + i + 1);
        if (noise != nullptr) nz = __ldg(noise + ray * S + i);
      }
      float dist = (i == S - 1) ? 1e10f : __fsub_rn(zn, zi);
      dist = __fmul_rn(dist, dnorm);
      const float sigma = (noise != nullptr) ? __fadd_rn(rv.w, nz) : rv.w;
      const float rl = fmaxf(sigma, 0.0f);
      float alpha = __fsub_rn(1.0f, expf(__fmul_rn(-rl, dist)));
      if (sigma != sigma) alpha = sigma;  // relu/exp propagate NaN in the reference
      const float tt = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);
      double p = valid ? static_cast<double>(tt) : 1.0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double q = shfl_up_f64(p, o);
        if (lane >= o) p *= q;
      }
      double excl = shfl_up_f64(p, 1);
      if (lane == 0) excl = 1.0;
      const float T = static_cast<float>(carry * excl);
      carry *= shfl_idx_f64(p, 31);
      const float w = valid ? __fmul_rn(alpha, T) : 0.0f;
      if (valid) {
        if (weights != nullptr) weights[ray * S + i] = w;
        const float sr = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-rv.x)));
        const float sg = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-rv.y)));
        const float sb = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-rv.z)));
        ar = __fadd_rn(ar, __fmul_rn(w, sr));
        ag = __fadd_rn(ag, __fmul_rn(w, sg));
        ab = __fadd_rn(ab, __fmul_rn(w, sb));
        adepth = __fadd_rn(adepth, __fmul_rn(w, zi));
        aacc = __fadd_rn(aacc, w);
      }
    }
    ar = warp_sum(ar);
    ag = warp_sum(ag);
    ab = warp_sum(ab);
    adepth = warp_sum(adepth);
    aacc = warp_sum(aacc);
    if (lane == 0) {
      if (white_bkgd) {
        const float bg = __fsub_rn(1.0f, aacc);
        ar = __fadd_rn(ar, bg);
        ag = __fadd_rn(ag, bg);
        ab = __fadd_rn(ab, bg);
      }
      if (rgb_map != nullptr) {
        rgb_map[3 * ray] = ar;
        rgb_map[3 * ray + 1] = ag;
        rgb_map[3 * ray + 2] = ab;
      }
      if (depth_map != nullptr) depth_map[ray] = adepth;
      if (acc_map != nullptr) acc_map[ray] = aacc;
      if (disp_map != nullptr) {
        const float q = __fdiv_rn(adepth, aacc);
        // torch.max(1e-10, q) propagates NaN (0/0 when all weights are zero); fmaxf would not
        const float m = (q != q) ? q : fmaxf(1e-10f, q);
        disp_map[ray] = __fdiv_rn(1.0f, m);
      }
    }
  }
}

// ---- blocked layout (S <= 32*K): the fast path --------------------------------------------------------------------
// The lane-strided kernel above spends ~280 warp instructions per 32 samples, two thirds of them in the five-step
// fp64 shuffle scan it runs for EVERY 32-sample chunk: it is issue-bound at 42 % of the HBM roofline (ncu, r1b).
// Here lane l owns the K CONSECUTIVE samples [l*K, l*K+K): the running product inside a lane is a plain sequential
// fp64 multiply and ONE warp scan per ray (not per chunk) combines the 32 lane totals.  Global traffic stays fully
// coalesced: a ray's row goes global -> shared with 16-byte cp.async (LDGSTS, no register staging) into a per-warp,
// double-buffered row whose per-lane stride is padded (K+1 float4 / K|1 floats) so that the blocked reads are
// bank-conflict free; the next ray's row is in flight while the current one is composited.  Weights go back through
// shared memory so the global store is coalesced too.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kBlkWarps = 4;
template <int K> struct BlkLayout {
  static constexpr int KR = K + 1;   // float4 per lane in the raw row (padded)
  static constexpr int KZ = K + 1;   // floats per lane in the z / weight rows (same padded stride: one offset table)
  static constexpr int kRawBytes = 32 * KR * 16;
  static constexpr int kZBytes = 32 * KZ * 4;
  static constexpr int kStageBytes = kRawBytes + kZBytes;
  static constexpr int kWarpBytes = 2 * kStageBytes + kZBytes;   // two input stages + the weight row
};

// RPW rays per warp: a warp stages RPW CONSECUTIVE rays (one contiguous chunk of RPW*S samples) and lane l owns the K
// consecutive samples [l*K, l*K+K) of that chunk, i.e. 32/RPW lanes per ray.  With S = 64 and one ray per warp a lane
// holds only 2 samples and the per-ray warp-level work (fp64 scan, five reductions, output stores) dominates: the
// kernel was issue-bound at 52 % of the HBM roofline (ncu r1g).  Four rays per warp (8 lanes x 8 samples per ray) run
// a 3-step segmented scan and 3-step reductions once for FOUR rays.
// FULL: RPW*S == 32*K exactly (64 / 128 / 192 / 256 samples: every BASELINE config) -> no per-sample predicates
// (whole rays beyond n_rays in the last warp are still masked).
template <int K, bool FULL, int RPW>
__global__ void __launch_bounds__(kBlkWarps * 32)
raw2outputs_blocked_kernel(long long n_rays, int S, const float4* __restrict__ raw, const float* __restrict__ z_vals,
                           const float* __restrict__ rays_d, long long d_stride, const float* __restrict__ noise,
                           int white_bkgd, float* __restrict__ rgb_map, float* __restrict__ disp_map,
                           float* __restrict__ acc_map, float* __restrict__ weights, float* __restrict__ depth_map) {
  using L = BlkLayout<K>;
  constexpr int LPR = 32 / RPW;   // lanes per ray
  extern __shared__ __align__(16) uint8_t smem_blk[];
  const int lane = threadIdx.x & 31;
  const int sl = lane % LPR;      // lane within its ray
  uint8_t* const wbase = smem_blk + static_cast<size_t>(threadIdx.x >> 5) * L::kWarpBytes;
  float* const s_wt = reinterpret_cast<float*>(wbase + 2 * L::kStageBytes);
  const long long n_groups = (n_rays + RPW - 1) / RPW;
  const long long warp0 = static_cast<long long>(blockIdx.x) * kBlkWarps + (threadIdx.x >> 5);
  const long long nwarps = static_cast<long long>(gridDim.x) * kBlkWarps;
  const long long total = n_rays * S;   // samples in the whole call

  // shared-memory slot (in elements) of chunk element i = c*32 + lane: owner lane i/K, position i%K; loop invariant
  int slot[K];
#pragma unroll
  for (int c = 0; c < K; ++c) {
    const int i = c * 32 + lane;
    slot[c] = (i / K) * L::KR + (i % K);
  }
  // elements of group g's chunk that exist: RPW*S, less for the last group / for S < 32K/RPW (generic path, RPW = 1)
  auto chunk_len = [&](long long g) -> int {
    const long long rest = total - g * RPW * S;
    return static_cast<int>(rest < static_cast<long long>(RPW) * S ? rest : static_cast<long long>(RPW) * S);
  };
  auto issue = [&](long long g, int stage) {
    float4* s_raw = reinterpret_cast<float4*>(wbase + stage * L::kStageBytes);
    float* s_z = reinterpret_cast<float*>(wbase + stage * L::kStageBytes + L::kRawBytes);
    const float4* rraw = raw + g * RPW * S;
    const float* rz = z_vals + g * RPW * S;
    const int n = chunk_len(g);
#pragma unroll
    for (int c = 0; c < K; ++c) {
      const int i = c * 32 + lane;
      if (i < n) {
        cp_async16(s_raw + slot[c], rraw + i);
        cp_async4(s_z + slot[c], rz + i);
      }
    }
  };

  int stage = 0;
  if (warp0 < n_groups) issue(warp0, 0);
  cp_async_commit();
  for (long long g = warp0; g < n_groups; g += nwarps, stage ^= 1) {
    if (g + nwarps < n_groups) issue(g + nwarps, stage ^ 1);
    cp_async_commit();
    const long long ray = g * RPW + lane / LPR;      // this lane's ray
    const bool ray_ok = ray < n_rays;
    const long long rd_ray = ray_ok ? ray : n_rays - 1;
    const float dx = __ldg(rays_d + rd_ray * d_stride), dy = __ldg(rays_d + rd_ray * d_stride + 1),
                dz = __ldg(rays_d + rd_ray * d_stride + 2);
    const float dnorm = comp_dnorm(dx, dy, dz);
    cp_async_wait<1>();
    __syncwarp();
    const float4* s_raw = reinterpret_cast<const float4*>(wbase + stage * L::kStageBytes) + lane * L::KR;
    const float* s_z = reinterpret_cast<const float*>(wbase + stage * L::kStageBytes + L::kRawBytes) + lane * L::KZ;
    float4 rv[K];
    float zr[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      rv[j] = s_raw[j];
      zr[j] = s_z[j];
    }
    const float z_next_lane = __shfl_down_sync(0xffffffffu, zr[0], 1);
    float w[K];
    float ar, ag, ab, adepth, aacc;
    blocked_composite<K, FULL, LPR>(rv, zr, z_next_lane, sl, S, ray_ok, dnorm,
                                    noise != nullptr ? noise + ray * S : nullptr, w, ar, ag, ab, adepth, aacc);
#pragma unroll
    for (int j = 0; j < K; ++j) s_wt[lane * L::KZ + j] = w[j];
    __syncwarp();
    if (weights != nullptr) {
      const int n = chunk_len(g);
      float* wrow = weights + g * RPW * S;
#pragma unroll
      for (int c = 0; c < K; ++c) {
        const int i = c * 32 + lane;
        if (i < n) wrow[i] = s_wt[slot[c]];
      }
    }
    if (sl == 0 && ray_ok) {
      if (white_bkgd) {
        const float bg = __fsub_rn(1.0f, aacc);
        ar = __fadd_rn(ar, bg);
        ag = __fadd_rn(ag, bg);
        ab = __fadd_rn(ab, bg);
      }
      if (rgb_map != nullptr) {
        rgb_map[3 * ray] = ar;
        rgb_map[3 * ray + 1] = ag;
        rgb_map[3 * ray + 2] = ab;
      }
      if (depth_map != nullptr) depth_map[ray] = adepth;
      if (acc_map != nullptr) acc_map[ray] = aacc;
      if (disp_map != nullptr) disp_map[ray] = comp_disp(adepth, aacc);
    }
    __syncwarp();   // every lane is done with this stage and with s_wt before they are overwritten
  }
  cp_async_wait<0>();
}

template <int K, bool FULL, int RPW>
static int launch_blocked(long long n_rays, int S, const float4* raw4, const float* z_vals, const float* rays_d,
                          long long d_stride, const float* noise, int white_bkgd, float* rgb_map, float* disp_map,
                          float* acc_map, float* weights, float* depth_map, cudaStream_t st) {
  constexpr int smem = kBlkWarps * BlkLayout<K>::kWarpBytes;
  static int ctas_per_sm = 0;   // occupancy is a property of the kernel; query once
  if (ctas_per_sm == 0) {
    if (smem > 48 * 1024)
      R2L_CUDA(cudaFuncSetAttribute(raw2outputs_blocked_kernel<K, FULL, RPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int n = 0;
    R2L_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, raw2outputs_blocked_kernel<K, FULL, RPW>, kBlkWarps * 32, smem));
    ctas_per_sm = n > 0 ? n : 1;
  }
  const long long n_groups = (n_rays + RPW - 1) / RPW;
  long long blocks = (n_groups + kBlkWarps - 1) / kBlkWarps;
  const long long cap = static_cast<long long>(sm_count()) * ctas_per_sm;   // one resident wave, grid-stride inside
  if (blocks > cap) blocks = cap;
  raw2outputs_blocked_kernel<K, FULL, RPW><<<static_cast<int>(blocks), kBlkWarps * 32, smem, st>>>(
      n_rays, S, raw4, z_vals, rays_d, d_stride, noise, white_bkgd, rgb_map, disp_map, acc_map, weights, depth_map);
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

// ---- the flagged rays of a fused NeRF frame (mlp_nerf_pp.cu compositor + nerf_far.cu) ----------------------------
// Entry e < min(*count, cap) of `list` names a ray whose staged rows were copied to far_raw[e][S] and whose far-sample
// sigma the fp32 fix-up has just patched there: composite it again (same blocked arithmetic, same K / lanes per ray
// as r2l_raw2outputs uses for this S) and overwrite the ray's outputs.
template <int K, int RPW>
__global__ void __launch_bounds__(kBlkWarps * 32)
raw2outputs_list_kernel(const int* __restrict__ list, const int* __restrict__ count, int cap, int S,
                        const float4* __restrict__ far_raw, const float* __restrict__ z_vals,
                        const float* __restrict__ rays_d, long long d_stride, int white_bkgd, float* __restrict__ rgb_map,
                        float* __restrict__ disp_map, float* __restrict__ acc_map, float* __restrict__ weights,
                        float* __restrict__ depth_map) {
  constexpr int LPR = 32 / RPW;
  const int lane = threadIdx.x & 31;
  const int sl = lane % LPR;
  int n = *count;
  if (n > cap) n = cap;
  const int n_groups = (n + RPW - 1) / RPW;
  for (int g = blockIdx.x * kBlkWarps + (threadIdx.x >> 5); g < n_groups; g += gridDim.x * kBlkWarps) {
    const int e = g * RPW + lane / LPR;
    const bool ok = e < n;
    const int ec = ok ? e : n - 1;
    const long long ray = list[ec];
    const float dnorm = comp_dnorm(__ldg(rays_d + ray * d_stride), __ldg(rays_d + ray * d_stride + 1),
                                   __ldg(rays_d + ray * d_stride + 2));
    const int base = sl * K;
    float4 rv[K];
    float zr[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      rv[j] = far_raw[static_cast<long long>(ec) * S + base + j];
      zr[j] = __ldg(z_vals + ray * S + base + j);
    }
    const float z_next_lane = __shfl_down_sync(0xffffffffu, zr[0], 1);
    float w[K];
    float ar, ag, ab, adepth, aacc;
    blocked_composite<K, true, LPR>(rv, zr, z_next_lane, sl, S, ok, dnorm, nullptr, w, ar, ag, ab, adepth, aacc);
    if (ok && weights != nullptr) {
#pragma unroll
      for (int j = 0; j < K; ++j) weights[ray * S + base + j] = w[j];
    }
    if (sl == 0 && ok) {
      if (white_bkgd) {
        const float bg = __fsub_rn(1.0f, aacc);
        ar = __fadd_rn(ar, bg);
        ag = __fadd_rn(ag, bg);
        ab = __fadd_rn(ab, bg);
      }
      if (rgb_map != nullptr) {
        rgb_map[3 * ray] = ar;
        rgb_map[3 * ray + 1] = ag;
        rgb_map[3 * ray + 2] = ab;
      }
      if (depth_map != nullptr) depth_map[ray] = adepth;
      if (acc_map != nullptr) acc_map[ray] = aacc;
      if (disp_map != nullptr) disp_map[ray] = comp_disp(adepth, aacc);
    }
  }
}

bool fused_composite_shape(int S, int* K, int* RPW) {
  int k = 0, r = 0;
  if (S == 64) k = 8, r = 4;
  else if (S == 128) k = 8, r = 2;
  else if (S == 192) k = 6, r = 1;
  else if (S == 256) k = 8, r = 1;
  if (K) *K = k;
  if (RPW) *RPW = r;
  return k != 0;
}

int raw2outputs_list_launch(const int* list, const int* count, int cap, int S, const float4* far_raw,
                            const float* z_vals, const float* rays_d, long long d_stride, int white_bkgd,
                            float* rgb_map, float* disp_map, float* acc_map, float* weights, float* depth_map,
                            cudaStream_t st) {
  const int grid = sm_count();
#define R2L_LIST(K, RPW)                                                                                             \
  raw2outputs_list_kernel<K, RPW><<<grid, kBlkWarps * 32, 0, st>>>(list, count, cap, S, far_raw, z_vals, rays_d,      \
                                                                    d_stride, white_bkgd, rgb_map, disp_map, acc_map, \
                                                                    weights, depth_map)
  if (S == 64) R2L_LIST(8, 4);
  else if (S == 128) R2L_LIST(8, 2);
  else if (S == 192) R2L_LIST(6, 1);
  else if (S == 256) R2L_LIST(8, 1);
  else return fail(R2L_ERR_UNSUPPORTED, "raw2outputs_list: S = %d has no fused compositing shape", S);
#undef R2L_LIST
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

}  // namespace r2l

using namespace r2l;

extern "C" {

int r2l_raw2outputs(long long n_rays, int S, const float* raw, const float* z_vals, const float* rays_d,
                    long long d_stride, const float* noise, int white_bkgd, float* rgb_map, float* disp_map,
                    float* acc_map, float* weights, float* depth_map, void* stream) {
  R2L_CHECK_ARG(n_rays >= 0 && S >= 2 && d_stride >= 3,
                "r2l_raw2outputs: need S >= 2 samples per ray (the reference degenerates to empty weights at S = 1)");
  if (n_rays == 0) return R2L_OK;
  R2L_CHECK_ARG(raw && z_vals && rays_d, "r2l_raw2outputs: null input pointer");
  R2L_CHECK_ARG((reinterpret_cast<uintptr_t>(raw) & 15) == 0, "r2l_raw2outputs: raw must be 16-byte aligned");
  long long blocks = (n_rays + kCompWarps - 1) / kCompWarps;
  const long long cap = static_cast<long long>(sm_count()) * 8;  // 8 resident 256-thread CTAs per SM
  if (blocks > cap) blocks = cap;
  auto st = static_cast<cudaStream_t>(stream);
  const float4* raw4 = reinterpret_cast<const float4*>(raw);
#define R2L_BLOCKED(K)                                                                                              \
  return (S == 32 * K)                                                                                              \
             ? launch_blocked<K, true, 1>(n_rays, S, raw4, z_vals, rays_d, d_stride, noise, white_bkgd, rgb_map,    \
                                          disp_map, acc_map, weights, depth_map, st)                                 \
             : launch_blocked<K, false, 1>(n_rays, S, raw4, z_vals, rays_d, d_stride, noise, white_bkgd, rgb_map,   \
                                           disp_map, acc_map, weights, depth_map, st)
  // the coarse pass (64 samples): four rays per warp; 128 samples (LLFF fine pass): two
  if (S == 64)
    return launch_blocked<8, true, 4>(n_rays, S, raw4, z_vals, rays_d, d_stride, noise, white_bkgd, rgb_map, disp_map,
                                      acc_map, weights, depth_map, st);
  if (S == 128)
    return launch_blocked<8, true, 2>(n_rays, S, raw4, z_vals, rays_d, d_stride, noise, white_bkgd, rgb_map, disp_map,
                                      acc_map, weights, depth_map, st);
  if (S <= 64) R2L_BLOCKED(2);
  if (S <= 128) R2L_BLOCKED(4);
  if (S <= 192) R2L_BLOCKED(6);
  if (S <= 256) R2L_BLOCKED(8);
#undef R2L_BLOCKED
  raw2outputs_kernel<0><<<static_cast<int>(blocks), kCompWarps * 32, 0, st>>>(
      n_rays, S, raw4, z_vals, rays_d, d_stride, noise, white_bkgd, rgb_map, disp_map, acc_map, weights, depth_map);
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

}  // extern "C"
