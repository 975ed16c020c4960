// Microbenchmark 4: tcgen05.mma issue rate from one thread vs two threads (different warps, different accumulators).
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "../../efficient-nerf_b200/csrc/mlp_tc.cuh"
using namespace r2l;

__global__ void __launch_bounds__(128, 1) k(int N, int n_mma, int n_issuers, int commit_every, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 64);
  uint8_t* sA = smem + 1024;            // 128 x 64 K
  uint8_t* sB = smem + 1024 + 16384;    // 256 x 64 K
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(sA)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1); mbar_fence_init(); }
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (lane == 0 && warp < n_issuers) {
    const uint32_t idesc = make_idesc_f16(false, 128, N);
    const uint32_t d = tmem_base + warp * 256;
    const uint32_t lbo_b = N * 16;
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 4) {
      issue_stage<4, false>(d, smem_u32(sA), smem_u32(sB), lbo_b, idesc, i == 0);
      if (commit_every && ((i / 4) % commit_every) == commit_every - 1) umma_commit(&bars[2 + warp]);
    }
    const long long t1 = clock64();
    umma_commit(&bars[warp]);
    mbar_wait(&bars[warp], 0, nullptr, 0);
    const long long t2 = clock64();
    out[warp * 2] = t1 - t0;
    out[warp * 2 + 1] = t2 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc(tmem_base, 512); }
}

int main() {
  long long* out; cudaMalloc(&out, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int n_mma = 2048;
  struct Cfg { int N, issuers, commit; };
  for (Cfg c : {Cfg{256, 1, 0}, Cfg{128, 1, 0}, Cfg{64, 1, 0}, Cfg{256, 1, 1}, Cfg{128, 1, 1}, Cfg{128, 1, 2}, Cfg{256, 2, 0}, Cfg{128, 2, 0}, Cfg{64, 2, 0}, Cfg{128, 2, 1}}) {
    for (int rep = 0; rep < 2; ++rep) { k<<<1, 128, 64 * 1024>>>(c.N, n_mma, c.issuers, c.commit, out); if (cudaDeviceSynchronize() != cudaSuccess) { printf("err\n"); return 1; } }
    long long h[4]; cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
    printf("N=%3d issuers=%d commit_every=%d stages: issue %.1f cyc/MMA, complete %.1f cyc/MMA (per issuer; tensor floor %d)", c.N, c.issuers, c.commit,
           (double)h[0] / n_mma, (double)h[1] / n_mma, c.N / 2);
    if (c.issuers == 2) printf("  | issuer1: %.1f / %.1f", (double)h[2] / n_mma, (double)h[3] / n_mma);
    printf("\n");
  }
  return 0;
}
