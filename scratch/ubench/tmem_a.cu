// Probe: tcgen05.mma with the A operand in TENSOR MEMORY (kind::f16).  Which TMEM cell (lane, column, half) is element
// A[m][k]?  Hypothesis: lane = m, 32-bit column c holds (k = 2c, 2c + 1), low half = even k, K = 16 per instruction =
// 8 columns.  Each probe t writes 1.0 into cell position t (column t / 2, half t % 2) of every row and zero elsewhere;
// B[n][k] = (k + 1) + 100 (n + 1), so D[m][n] = (k(t) + 1) + 100 (n + 1) reveals k(t).  Row check: a second pass writes
// (m + 1) into position 0.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scratch/ubench/tmem_a scratch/ubench/tmem_a.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "../../efficient-nerf_b200/csrc/mlp_tc.cuh"
using namespace r2l;

constexpr int N = 64, K = 16;

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(int t_pos, int row_mode, float* out /*[128][N]*/) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 64);
  uint8_t* sB = smem + 1024;   // N x 16 halves, k-chunk major: offset(n, k) = (k/8)*N*16 + n*16 + (k%8)*2
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < N * K; i += 128) {
    const int n = i / K, k = i % K;
    *reinterpret_cast<__half*>(sB + (k / 8) * N * 16 + n * 16 + (k % 8) * 2) = __float2half(static_cast<float>((k + 1) + 100 * (n + 1)));
  }
  if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_fence_init(); }
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  // A region: columns [256, 288) of this thread's lane; D: columns [0, N)
  {
    uint32_t v[32];
    for (int i = 0; i < 32; ++i) v[i] = 0u;
    const float val = row_mode ? static_cast<float>(warp * 32 + lane + 1) : 1.0f;
    const uint32_t h = __half_as_ushort(__float2half(val));
    v[t_pos / 2] = (t_pos & 1) ? (h << 16) : h;
    tmem_st32(lane_addr + 256, v);
    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(-7777.0f);
    tmem_st32(lane_addr, v);
    tmem_st32(lane_addr + 32, v);
    tmem_st_wait();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_f16(false, 128, N);
    umma_f16_ts(tmem_base, tmem_base + 256, make_smem_desc(smem_u32(sB), N * 16, 128), idesc, 0u);
    umma_commit(&bars[0]);
  }
  __syncwarp();
  { uint32_t spins = 0; while (!mbar_try_wait(&bars[0], 0)) { if (++spins > (1u << 26)) __trap(); } }
  tc_fence_after_sync();
  for (int c = 0; c < N; c += 32) {
    uint32_t v[32];
    tmem_ld32(lane_addr + c, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * N + c + i] = __uint_as_float(v[i]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc(tmem_base, 512); }
}

int main() {
  float* out;
  cudaMalloc(&out, 128 * N * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024);
  std::vector<float> h(128 * N);
  for (int t = 0; t < 24; ++t) {
    probe<<<1, 128, 16 * 1024>>>(t, 0, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("t=%d: %s\n", t, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h.data(), out, h.size() * 4, cudaMemcpyDeviceToHost);
    // D[m][n] = (k+1) + 100(n+1) if the cell maps to k; 0 if the cell is not read
    printf("pos %2d (col %2d half %d): D[0][0]=%8.1f D[0][1]=%8.1f D[5][0]=%8.1f D[127][63]=%8.1f  -> k=%d\n", t, t / 2, t & 1,
           h[0], h[1], h[5 * N], h[127 * N + 63], h[0] != 0.f ? (int)(h[0] - 100.f) - 1 : -1);
  }
  probe<<<1, 128, 16 * 1024>>>(0, 1, out);
  cudaDeviceSynchronize();
  cudaMemcpy(h.data(), out, h.size() * 4, cudaMemcpyDeviceToHost);
  printf("row mode (value m+1 at pos 0, times B[n][k0]): ");
  for (int m : {0, 1, 31, 32, 64, 100, 127}) printf(" lane %d: %.1f (expect %.1f)", m, h[m * N], (m + 1) * 101.0f);
  printf("\n");
  return 0;
}
