// Kernel parameter blocks and launcher prototypes shared by mlp_api.cu, mlp_nerf.cu, mlp_r2l.cu.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc_common.cuh"

namespace r2l {

struct NerfParams {
  const uint8_t* wstream;   // packed weight stages (16-bit), consumption order
  const float* alpha_w;     // [256]
  const float* rgb_w;       // [3][128]
  float alpha_b;
  float rgb_b[3];
  const float* rays_o;
  const float* rays_d;
  const float* viewdirs;    // [n_rays][v_stride] unit view directions
  long long o_stride, d_stride, v_stride;
  const float* z_vals;      // [n_rays*S]
  int S;
  long long n_rows;         // n_rays*S
  long long n_rays;         // ray-sample calls (embedded == nullptr)
  float* raw;               // [n_rows][4]
  int n_tiles;
  DebugBuf* dbg;
  const float* embedded;    // optional [n_rows][emb_stride]: 63 embedded-point + 27 embedded-view features
  long long emb_stride;
  long long* prof;          // optional [gridDim.x][8] cycle counters (see r2l_nerf_profile)
  int prof_mode;            // ping-pong kernel: 1 = report ring turnaround instead of the wait split
  const float* vb;          // ping-pong kernel only: per-ray view-branch bias [n_rays][128] (nerf_view_bias_kernel)
  // far-sample guard band (nerf_far.cu): rays whose LAST sample has |sigma| < max(far_abs, far_rel * sum|w_a| relu(h7))
  // are appended to far_list (atomicAdd on far_count; entries beyond far_cap are dropped, the count keeps growing)
  int* far_list;            // nullptr: no flagging
  int* far_count;
  int far_cap;
  float far_abs, far_rel;
  // ping-pong kernel: CTA c owns the CONTIGUOUS tiles [c*T, (c+1)*T), T = tiles_per_cta (even; with fused compositing
  // a multiple of the tiles a compositing group spans, so that no ray straddles two CTAs)
  int tiles_per_cta;
  // fused compositing (ping-pong kernel, UNI rows, S in {64,128,192,256}; raw2outputs main.py:556-621 inside the MLP
  // kernel): the last epilogue stages (rgb, sigma) rows in a per-CTA ring in GLOBAL memory (512 rows = 8 KB per CTA:
  // it never leaves the L2) instead of writing raw [N,S,4]; the otherwise idle encoder warps composite every ray
  // whose last sample has arrived, with the arithmetic of raw2outputs_blocked_kernel (composite.cu: comp_K consecutive
  // samples per lane, 32/comp_RPW lanes per ray).  Rays whose far sample is flagged (far_list) are composited too,
  // and their staged rows are ALSO copied to comp_far_raw[slot] (slot = position in far_list) for the fix-up pass.
  float4* comp_ring;        // nullptr: write raw (no compositing)
  int* comp_aux;            // [grid][512]: far slot of a ray's LAST sample (-1 = not flagged), same indexing as the ring
  float4* comp_far_raw;     // [comp_far_cap][S]
  int comp_far_cap;
  int comp_K, comp_RPW;
  int white_bkgd;
  float* o_rgb;             // [n_rays][3]
  float* o_disp;            // [n_rays] or nullptr
  float* o_acc;             // [n_rays] or nullptr
  float* o_depth;           // [n_rays] or nullptr
  float* o_weights;         // [n_rays][S] or nullptr
};

// far-sample sigma fix-up (nerf_far.cu)
size_t nerf_far_weight_bytes();
int nerf_far_pack(const float* const* pts_w, const float* const* pts_b, const float* alpha_w, float* Wt, cudaStream_t st);
// out: raw [n_rays][S][4] (compact == 0: sigma of ray r's last sample) or the compact copy [cap][S][4] of the flagged
// rays' rows (compact == 1: list position i -> row i)
int nerf_far_fixup_launch(const float* Wt, float alpha_b, const int* list, const int* count, int cap, int* stats,
                          const NerfParams& p, float* out, int compact, cudaStream_t st);

// fused compositing (composite.cu): the (samples per lane, rays per warp) r2l_raw2outputs uses for S samples — the
// compositor inside the ping-pong kernel follows the same partition; false: S has no fused shape
bool fused_composite_shape(int S, int* K, int* RPW);
int raw2outputs_list_launch(const int* list, const int* count, int cap, int S, const float4* far_raw,
                            const float* z_vals, const float* rays_d, long long d_stride, int white_bkgd,
                            float* rgb_map, float* disp_map, float* acc_map, float* weights, float* depth_map,
                            cudaStream_t st);

// The epilogue's test: is row g_row the far sample of its ray, and is its sigma inside the guard band?
// Returns the ray's position in far_list (-1: not flagged / not a far sample / dropped).
__device__ __forceinline__ int nerf_far_flag(const NerfParams& p, long long g_row, long long ray, float sigma,
                                             float abs_sum) {
  if (p.far_list != nullptr && g_row - ray * p.S == p.S - 1 && fabsf(sigma) < fmaxf(p.far_abs, p.far_rel * abs_sum)) {
    const int idx = atomicAdd(p.far_count, 1);
    if (idx < p.far_cap) {
      p.far_list[idx] = static_cast<int>(ray);
      return idx;
    }
  }
  return -1;
}

// fp32 weights of the two heads that run on CUDA cores (alpha_linear [256], rgb_linear [3][128]), passed BY VALUE as a
// kernel parameter: with compile-time indices every FFMA takes its weight straight from the constant bank
// (c[0x0][imm]) — no load instruction, no register, no shared-memory traffic under the MMAs' operand reads.  With the
// weights in shared memory the rgb dot product made the last epilogue 2800 cycles long (in-kernel trace: 1850 of them
// in 48 LDS.128 + 192 FFMA per thread, latency-bound because the 128-register budget leaves nothing to hoist loads).
struct NerfHeadW {
  float alpha_w[256];
  float rgb_w[384];
};

// Tensor maps over the ping-pong kernel's stage stream seen as rows of 256 x 16-bit (512 bytes): boxes of 32 / 16 / 8
// rows = one CTA's half of a K=64 stage (N 256), of a K=64 stage (N 128), of a bias stage
struct NerfPpMaps {
  CUtensorMap m16, m8, m4;
};

constexpr int kMaxPeers = 8;   // GPUs of one box

struct R2lParams {
  const uint8_t* wstream;
  const float* w_tail;      // [3][256]
  float b_tail[3];
  int n_blocks;
  int n_points;             // points per ray (multiple of 4)
  const float* pts;         // [n_rays][pts_stride] (n_points*3 used)
  long long pts_stride;
  long long n_rays;
  float* rgb;               // [n_rays][3]
  int sigmoid_out;          // 1: tail has Sigmoid
  int outer_skip;           // 1: body(x) + x  (args.use_residual)
  int n_tiles;
  DebugBuf* dbg;
  const float* embedded;    // optional [n_rays][emb_stride]: reference-layout PositionalEmbedder output
  long long emb_stride;
  long long* prof;          // optional [gridDim.x][8] cycle counters (see r2l_resmlp_profile)
  float* dbg_head_acc;      // optional debug dump [n_tiles*128][256]: head accumulators (bias included)
  float* dbg_head_x0;       // optional debug dump [n_tiles*128][256]: x0 = relu(acc)
  // fused tile gather (ray-sharded frames): when n_peer > 0 the tail stores row `ray` of this launch to
  // rgb_peer[g] + 3 * (peer_row0 + ray) for every g — the same frame buffer on every GPU of the box, reached by
  // peer-to-peer stores over NVLink (symmetric memory) — instead of p.rgb: the all-gather needs no kernel of its own
  float* rgb_peer[kMaxPeers];
  int n_peer;
  long long peer_row0;
  // fused ray generation + point sampling (PointSampler.sample_test, model/nerf_raybased.py:94-102): when cam != nullptr
  // the head computes its points from the pixel index instead of reading p.pts — row r of this launch is ray
  // cam_ray0 + r of the pose-major range [n_poses][H*W]; cam = c2w [n_poses][3][4], cam_z = the sampler's z_vals
  // [n_points]; the same rounded operations, in the same order, as point_sample_kernel (rays.cu)
  const float* cam;
  const float* cam_z;
  int cam_H, cam_W;
  float cam_focal;
  long long cam_ray0;
};

int nerf_mlp_launch(bool bf16, const NerfParams& p, int grid, cudaStream_t st);
// Tensor maps over the R2L pair-layout stage stream (rows of 512 bytes): boxes of 32 / 8 rows = one CTA's half of a
// K=64 weight stage / of a bias stage
struct R2lPairMaps {
  CUtensorMap m16, m4;
};
int r2l_mlp_launch(bool bf16, bool pair, const R2lParams& p, const R2lPairMaps* maps, int grid, cudaStream_t st);
// CTA-pair ping-pong kernel with two 64-row tiles per CTA (mlp_r2l_pp.cu): pair-layout weights, grid even
int r2l_mlp_pp_launch(bool bf16, const R2lParams& p, const R2lPairMaps& maps, int grid, cudaStream_t st);
// CTA-pair "ping-pong" NeRF kernel (mlp_nerf_pp.cu): grid even, pair-layout stream without the view stage
int nerf_mlp_pp_launch(bool bf16, const NerfParams& p, const NerfPpMaps& maps, const NerfHeadW& hw, int grid,
                       cudaStream_t st);
int nerf_view_bias_launch(long long n_rays, const float* viewdirs, long long v_stride, int pre_embedded,
                          const float* wvd, const float* bv, float* vb, cudaStream_t st);

}  // namespace r2l
