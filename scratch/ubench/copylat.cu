// Microbenchmark 3: one thread issues n bulk copies back-to-back to DISTINCT slots / mbarriers, then waits for
// each in order and records the clock.  Shows whether copies overlap in the copy engine.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "../../efficient-nerf_b200/csrc/tc_common.cuh"
using namespace r2l;

__global__ void __launch_bounds__(32, 1) k(const uint8_t* buf, int n, int bytes, int same_bar, long long* out, int rounds) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint8_t* data = smem + 1024;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) mbar_init(&bar[i], 1);
    mbar_fence_init();
    for (int r = 0; r < rounds; ++r) {
      const uint8_t* src = buf + (size_t)((blockIdx.x * 7 + r * 13) % 64) * (192 * 1024);
      const long long t0 = clock64();
      if (same_bar) {
        mbar_expect_tx(&bar[0], n * bytes);
        for (int i = 0; i < n; ++i) bulk_g2s(data + (size_t)i * bytes, src + (size_t)i * bytes, bytes, &bar[0]);
      } else {
        for (int i = 0; i < n; ++i) {
          mbar_expect_tx(&bar[i], bytes);
          bulk_g2s(data + (size_t)i * bytes, src + (size_t)i * bytes, bytes, &bar[i]);
        }
      }
      const long long t1 = clock64();
      long long* o = out + ((size_t)blockIdx.x * rounds + r) * 16;
      o[15] = t1 - t0;
      for (int i = 0; i < (same_bar ? 1 : n); ++i) {
        mbar_wait(&bar[i], r & 1, nullptr, 0);
        o[i] = clock64() - t0;
      }
    }
  }
}

int main() {
  const size_t total = 64ull * 192 * 1024;
  uint8_t* buf; cudaMalloc(&buf, total); cudaMemset(buf, 1, total);
  long long* out; cudaMalloc(&out, 148 * 8 * 16 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  const int rounds = 8;
  struct Cfg { int grid, n, bytes, same; };
  std::vector<Cfg> cfgs = {{1, 1, 16384, 0}, {1, 2, 16384, 0}, {1, 4, 16384, 0}, {1, 8, 16384, 0}, {1, 12, 16384, 0},
                           {1, 8, 16384, 1}, {1, 1, 65536, 0}, {1, 3, 65536, 0}, {1, 1, 196608, 0}, {1, 12, 1024, 0},
                           {148, 8, 16384, 0}, {148, 12, 16384, 0}, {148, 1, 196608, 0}, {148, 3, 65536, 0}};
  for (auto& c : cfgs) {
    k<<<c.grid, 32, 1024 + 12 * 16384 + 1024>>>(buf, c.n, c.bytes, c.same, out, rounds);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<long long> h((size_t)c.grid * rounds * 16);
    cudaMemcpy(h.data(), out, h.size() * 8, cudaMemcpyDeviceToHost);
    // report last round of CTA 0 and mean over CTAs of final completion
    const long long* o = &h[(size_t)(rounds - 1) * 16];
    printf("grid %3d n %2d x %6d B %s: issue %4lld cyc; completions:", c.grid, c.n, c.bytes, c.same ? "same-bar" : "own-bar ", o[15]);
    for (int i = 0; i < (c.same ? 1 : c.n); ++i) printf(" %lld", o[i]);
    double mean = 0;
    const int last = c.same ? 0 : c.n - 1;
    for (int b = 0; b < c.grid; ++b) mean += h[((size_t)b * rounds + rounds - 1) * 16 + last];
    mean /= c.grid;
    printf("   | mean final %.0f cyc -> %.1f B/cyc/SM\n", mean, (double)c.n * c.bytes / mean);
  }
  return 0;
}
