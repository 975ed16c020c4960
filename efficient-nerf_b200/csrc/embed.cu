// Stand-alone positional encodings (HBM-bound; the fused MLP kernels produce the
// same features directly into shared memory instead).
//   NeRF layout  Embedder.embed          utils/run_nerf_raybased_helpers.py:24-56
//       [x(D), sin(2^0 x)(D), cos(2^0 x)(D), ..., sin(2^{L-1} x)(D), cos(2^{L-1} x)(D)]
//   R2L  layout  PositionalEmbedder()    model/nerf_raybased.py:191-208
//       per input coordinate c: [sin(2^0 x_c)..sin(2^{L-1} x_c), cos(2^0 x_c)..cos(2^{L-1} x_c), x_c]
// The frequency multiply is exact (powers of two); sinf/cosf are CUDA's accurate
// single-precision routines (no fast-math), within 2 ulp of the reference's sin/cos.
//
// Each block stages a tile of output rows in shared memory and streams it out with
// 16-byte coalesced stores: algorithmic traffic is 4*D in + 4*D*(1+2L) out per row.
#include "common.cuh"

namespace r2l {

constexpr int kEmbedThreads = 256;

// rows: number of D-vectors; out row length = D*(1+2L) (include_input) or D*2L.
template <bool R2L_LAYOUT>
__global__ void __launch_bounds__(kEmbedThreads)
embed_kernel(long long rows, int D, int L, int include_input, const float* __restrict__ x, float* __restrict__ out,
             int rows_per_tile) {
  extern __shared__ float tile[];
  const int out_dim = D * (2 * L + (include_input ? 1 : 0));
  const long long n_tiles = (rows + rows_per_tile - 1) / rows_per_tile;
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const long long row0 = t * rows_per_tile;
    const int nrows = static_cast<int>(min(static_cast<long long>(rows_per_tile), rows - row0));
    const int n_in = nrows * D;
    // one work item = (row, coord, freq): one sincos pair; plus identity items
    const int n_sc = n_in * L;
    for (int it = threadIdx.x; it < n_sc; it += kEmbedThreads) {
      const int f = it % L;
      const int rc = it / L;  // row*D + c
      const int c = rc % D, r = rc / D;
      const float v = x[(row0 * D) + rc];
      const float arg = v * exp2f(static_cast<float>(f));  // exact: power-of-two scaling
      float s, co;
      sincosf(arg, &s, &co);
      float* orow = tile + static_cast<size_t>(r) * out_dim;
      if (R2L_LAYOUT) {
        float* oc = orow + c * (2 * L + (include_input ? 1 : 0));
        oc[f] = s;
        oc[L + f] = co;
      } else {
        const int base = (include_input ? D : 0) + 2 * D * f;
        orow[base + c] = s;
        orow[base + D + c] = co;
      }
    }
    if (include_input) {
      for (int it = threadIdx.x; it < n_in; it += kEmbedThreads) {
        const int c = it % D, r = it / D;
        const float v = x[(row0 * D) + it];
        float* orow = tile + static_cast<size_t>(r) * out_dim;
        if (R2L_LAYOUT)
          orow[c * (2 * L + 1) + 2 * L] = v;
        else
          orow[c] = v;
      }
    }
    __syncthreads();
    // stream the tile out: contiguous in global memory
    const long long o0 = row0 * out_dim;
    const int n_out = nrows * out_dim;
    float* gout = out + o0;
    if ((reinterpret_cast<uintptr_t>(gout) & 15) == 0) {
      const int n4 = n_out / 4;
      const float4* t4 = reinterpret_cast<const float4*>(tile);
      float4* g4 = reinterpret_cast<float4*>(gout);
      for (int i = threadIdx.x; i < n4; i += kEmbedThreads) g4[i] = t4[i];
      for (int i = n4 * 4 + threadIdx.x; i < n_out; i += kEmbedThreads) gout[i] = tile[i];
    } else {
      for (int i = threadIdx.x; i < n_out; i += kEmbedThreads) gout[i] = tile[i];
    }
    __syncthreads();
  }
}

}  // namespace r2l

using namespace r2l;

extern "C" {

// layout: 0 = NeRF (Embedder), 1 = R2L (PositionalEmbedder).  x is [rows, D] contiguous.
int r2l_embed(long long rows, int D, int L, int include_input, int layout, const float* x, float* out,
              void* stream) {
  R2L_CHECK_ARG(rows >= 0 && D > 0 && L >= 0 && L <= 24, "r2l_embed: bad sizes");
  R2L_CHECK_ARG(layout == 0 || layout == 1, "r2l_embed: layout must be 0 (NeRF) or 1 (R2L)");
  if (rows == 0) return R2L_OK;
  R2L_CHECK_ARG(x && out, "r2l_embed: null pointer");
  const int out_dim = D * (2 * L + (include_input ? 1 : 0));
  R2L_CHECK_ARG(out_dim > 0, "r2l_embed: empty output");
  // ~32 KB tiles; rows_per_tile*out_dim*4 multiple of 16 when rows_per_tile % 4 == 0
  int rows_per_tile = (32 * 1024 / 4) / out_dim;
  rows_per_tile = (rows_per_tile / 4) * 4;
  if (rows_per_tile < 4) rows_per_tile = 4;
  const size_t smem = static_cast<size_t>(rows_per_tile) * out_dim * sizeof(float);
  R2L_CHECK_ARG(smem <= 200 * 1024, "r2l_embed: row too wide (%d floats)", out_dim);
  const long long n_tiles = (rows + rows_per_tile - 1) / rows_per_tile;
  long long grid = n_tiles;
  const long long cap = static_cast<long long>(sm_count()) * 6;
  if (grid > cap) grid = cap;
  auto st = static_cast<cudaStream_t>(stream);
  if (layout == 0) {
    if (smem > 48 * 1024)
      R2L_CUDA(cudaFuncSetAttribute(embed_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    embed_kernel<false><<<static_cast<int>(grid), kEmbedThreads, smem, st>>>(rows, D, L, include_input, x, out,
                                                                            rows_per_tile);
  } else {
    if (smem > 48 * 1024)
      R2L_CUDA(cudaFuncSetAttribute(embed_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    embed_kernel<true><<<static_cast<int>(grid), kEmbedThreads, smem, st>>>(rows, D, L, include_input, x, out,
                                                                           rows_per_tile);
  }
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

}  // extern "C"
