// Probe: tcgen05.mma.cta_group::2 with M = 128 (64 rows per CTA) — where do the accumulator rows / columns land in
// each CTA's TMEM, and how long does such an MMA take (the microarchitecture notes say max(M,128) N / (256 cg) = 64
// cycles for N = 256: full rate)?  Basis of a ping-pong R2L kernel with two 64-row tiles per CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scratch/ubench/pair_m128 scratch/ubench/pair_m128.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "../../efficient-nerf_b200/csrc/mlp_tc.cuh"
using namespace r2l;

constexpr int N = 256, K = 16;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe(int M_total, int lane_off, float* out /*[2 cta][128 lanes][512 cols]*/, long long* tim, int n_time) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 64);
  uint8_t* sA1 = smem + 1024;                 // rows x 16 (row-id pattern)
  uint8_t* sA2 = sA1 + 4096;                  // rows x 16 (ones)
  uint8_t* sB1 = sA2 + 4096;                  // N/2 x 16 (ones)
  uint8_t* sB2 = sB1 + 4096;                  // N/2 x 16 (col-id pattern)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int rows = M_total / 2;               // rows per CTA
  for (int i = threadIdx.x; i < 4 * 4096 / 4; i += 128) reinterpret_cast<uint32_t*>(sA1)[i] = 0u;
  __syncthreads();
  // A[r][0]: global row id + 1 (A1) / 1 (A2); k-chunk major with `rows` rows: offset(r, k) = (k/8)*rows*16 + r*16 + (k%8)*2
  for (int r = threadIdx.x; r < rows; r += 128) {
    *reinterpret_cast<__half*>(sA1 + r * 16) = __float2half(static_cast<float>(rank * rows + r + 1));
    *reinterpret_cast<__half*>(sA2 + r * 16) = __float2half(1.0f);
  }
  for (int n = threadIdx.x; n < N / 2; n += 128) {
    *reinterpret_cast<__half*>(sB1 + n * 16) = __float2half(1.0f);
    *reinterpret_cast<__half*>(sB2 + n * 16) = __float2half(static_cast<float>(rank * (N / 2) + n + 1));
  }
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1); mbar_fence_init(); }
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // sentinel in all 512 columns of this warp's 32 lanes
  {
    uint32_t v[32];
    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(-7777.0f);
    for (int c = 0; c < 512; c += 32) tmem_st32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    tmem_st_wait();
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_f16(false, M_total, N);
    const uint32_t lboA = rows * 16, lboB = (N / 2) * 16;
    const uint32_t d = tmem_base + (static_cast<uint32_t>(lane_off) << 16);
    umma_f16_ss_pair(d, make_smem_desc(smem_u32(sA1), lboA, 128), make_smem_desc(smem_u32(sB1), lboB, 128), idesc, 0u);
    umma_f16_ss_pair(d + 256, make_smem_desc(smem_u32(sA2), lboA, 128), make_smem_desc(smem_u32(sB2), lboB, 128), idesc, 0u);
    umma_commit_pair(&bars[0]);
  }
  __syncwarp();
  { uint32_t spins = 0; while (!mbar_try_wait_cluster(&bars[0], 0)) { if (++spins > (1u << 26)) __trap(); } }
  tc_fence_after_sync();
  for (int c = 0; c < 512; c += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(static_cast<size_t>(rank) * 128 + warp * 32 + lane) * 512 + c + i] = __uint_as_float(v[i]);
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  // timing: n_time back-to-back MMAs
  if (rank == 0 && threadIdx.x == 0 && n_time > 0) {
    const uint32_t idesc = make_idesc_f16(false, M_total, N);
    const uint32_t lboA = rows * 16, lboB = (N / 2) * 16;
    const uint64_t ad = make_smem_desc(smem_u32(sA1), lboA, 128), bd = make_smem_desc(smem_u32(sB1), lboB, 128);
    const long long t0 = clock64();
    for (int i = 0; i < n_time; ++i) umma_f16_ss_pair(tmem_base, ad, bd, idesc, 1u);
    const long long t1 = clock64();
    umma_commit_pair(&bars[1]);
    while (!mbar_try_wait_cluster(&bars[1], 0)) {}
    const long long t2 = clock64();
    tim[0] = t1 - t0;
    tim[1] = t2 - t0;
  } else if (n_time > 0 && threadIdx.x == 0) {
    while (!mbar_try_wait_cluster(&bars[1], 0)) {}
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc_pair(tmem_base, 512); }
}

int main() {
  float* out; long long* tim;
  cudaMalloc(&out, 2 * 128 * 512 * 4); cudaMalloc(&tim, 64);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
  std::vector<float> h(2 * 128 * 512);
  for (int M_total : {256, 128}) for (int lane_off : {0, 64}) {
    if (M_total == 256 && lane_off) continue;
    probe<<<2, 128, 32 * 1024>>>(M_total, lane_off, out, tim, 2048);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("M=%d lane_off=%d: %s\n", M_total, lane_off, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h.data(), out, h.size() * 4, cudaMemcpyDeviceToHost);
    long long t[2]; cudaMemcpy(t, tim, 16, cudaMemcpyDeviceToHost);
    printf("== M=%d (rows per CTA %d), D lane offset %d: %.1f cycles issue / %.1f cycles complete per MMA (N=256, K=16)\n",
           M_total, M_total / 2, lane_off, t[0] / 2048.0, t[1] / 2048.0);
    for (int cta = 0; cta < 2; ++cta) {
      printf(" CTA %d: lane -> (row id from D1[col 0], written columns of D1, col id range from D2)\n", cta);
      for (int lane = 0; lane < 128; lane += 8) {
        const float* row = &h[(cta * 128 + lane) * 512];
        int first = -1, last = -1, cnt = 0;
        for (int c = 0; c < 256; ++c) if (row[c] != -7777.0f) { if (first < 0) first = c; last = c; ++cnt; }
        int f2 = -1, l2 = -1;
        for (int c = 256; c < 512; ++c) if (row[c] != -7777.0f) { if (f2 < 0) f2 = c; l2 = c; }
        printf("   lane %3d: row %5.0f  D1 cols [%d..%d] (%d written)  D2: col ids %.0f..%.0f at cols [%d..%d]\n", lane,
               first >= 0 ? row[first] - 1 : -1.f, first, last, cnt, f2 >= 0 ? row[f2] - 1 : -1.f, l2 >= 0 ? row[l2] - 1 : -1.f, f2, l2);
      }
    }
  }
  return 0;
}
