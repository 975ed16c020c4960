// Kernel parameter blocks and launcher prototypes shared by mlp_api.cu, mlp_nerf.cu, mlp_r2l.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc_common.cuh"

namespace r2l {

struct NerfParams {
  const uint8_t* wstream;   // packed weight stages (16-bit), consumption order
  const float* bias;        // [9][256]: pts_linears 0..7, feature_linear
  const float* alpha_w;     // [256]
  const float* rgb_w;       // [3][128]
  float alpha_b;
  float rgb_b[3];
  const float* vb;          // [n_rays][128] per-ray view-branch bias (fp32)
  const float* rays_o;
  const float* rays_d;
  long long o_stride, d_stride;
  const float* z_vals;      // [n_rays*S]
  int S;
  long long n_rows;         // n_rays*S
  float* raw;               // [n_rows][4]
  int n_tiles;
  DebugBuf* dbg;
  const float* embedded;    // optional [n_rows][emb_stride]: pre-embedded points (NeRF.forward API path)
  long long emb_stride;
};

struct R2lParams {
  const uint8_t* wstream;
  const float* b_head;      // [256]
  const float* b1;          // [n_blocks][256]
  const float* cb;          // [n_blocks][256] cumulative res_scale*b2
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
  float* dbg_head_acc;      // optional debug dump [n_tiles*128][256]: raw head accumulators
  float* dbg_head_x0;       // optional debug dump [n_tiles*128][256]: x0 = relu(acc + b_head)
};

int nerf_mlp_launch(bool bf16, const NerfParams& p, int grid, cudaStream_t st);
int nerf_view_bias_launch(long long n_rays, const float* viewdirs, long long v_stride, int pre_embedded,
                          const float* wvd, const float* bv, float* vb, cudaStream_t st);
int r2l_mlp_launch(bool bf16, const R2lParams& p, int grid, cudaStream_t st);

}  // namespace r2l
