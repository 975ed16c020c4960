// Microbenchmark: how fast can every SM stream the SAME L2-resident weight buffer into shared memory with
// 1-D bulk copies through a ring (the fused MLP kernels' producer pattern)?
//   mode 0: all CTAs read the same addresses in lockstep
//   mode 1: CTA i starts at stage offset (i * skew) % n_stages (same buffer, phase shifted)
//   mode 2: CTA i reads copy (i % n_copies) of the buffer
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "../../efficient-nerf_b200/csrc/tc_common.cuh"
using namespace r2l;

__global__ void __launch_bounds__(64, 1) stream_kernel(const uint8_t* buf, size_t copy_bytes, int n_copies, int stage_bytes,
                                                       int ring, int n_stages, int passes, int mode, int skew,
                                                       long long* cycles, int consume_delay, int pieces, int mt) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + 32;
  uint8_t* data = smem + 1024;
  if (threadIdx.x == 0) {
    for (int i = 0; i < ring; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_fence_init();
  }
  __syncthreads();
  const uint8_t* base = buf + (mode == 2 ? (size_t)(blockIdx.x % n_copies) * copy_bytes : 0);
  const int start = (mode == 1) ? (int)(((long long)blockIdx.x * skew) % n_stages) : 0;
  const long long total = (long long)n_stages * passes;
  const long long t0 = clock64();
  if (threadIdx.x < 32 && (threadIdx.x == 0 || (mt && threadIdx.x < pieces))) {
    const int pb = stage_bytes / pieces;
    for (long long g = 0; g < total; ++g) {
      const int slot = g % ring;
      mbar_wait(&empty[slot], ((g / ring) & 1) ^ 1, nullptr, 0);
      const int st = (int)((g + start) % n_stages);
      if (mt) {
        if (threadIdx.x == 0) mbar_expect_tx(&full[slot], stage_bytes);
        __syncwarp((1u << pieces) - 1);
        const int pc = threadIdx.x;
        bulk_g2s(data + (size_t)slot * stage_bytes + pc * pb, base + (size_t)st * stage_bytes + pc * pb, pb, &full[slot]);
      } else {
        mbar_expect_tx(&full[slot], stage_bytes);
        for (int pc = 0; pc < pieces; ++pc)
          bulk_g2s(data + (size_t)slot * stage_bytes + pc * pb, base + (size_t)st * stage_bytes + pc * pb, pb, &full[slot]);
      }
    }
  } else if (threadIdx.x == 32) {
    for (long long g = 0; g < total; ++g) {
      const int slot = g % ring;
      mbar_wait(&full[slot], (g / ring) & 1, nullptr, 0);
      if (consume_delay > 0) { const long long c = clock64(); while (clock64() - c < consume_delay) {} }
      mbar_arrive(&empty[slot]);
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

int main(int argc, char** argv) {
  const size_t copy_bytes = 12ull << 20;
  const int max_copies = 8;
  uint8_t* buf; cudaMalloc(&buf, copy_bytes * max_copies); cudaMemset(buf, 1, copy_bytes * max_copies);
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  struct Cfg { int grid, stage, ring, mode, skew, copies, delay; const char* name; int pieces = 1; int mt = 0; };
  std::vector<Cfg> cfgs = {
    {1, 16384, 9, 0, 0, 1, 0, "1 CTA, 16K x9"},
    {1, 65536, 3, 0, 0, 1, 0, "1 CTA, 64K x3"},
    {1, 4096, 9, 0, 0, 1, 0, "1 CTA, 4K x9"},
    {1, 1024, 9, 0, 0, 1, 0, "1 CTA, 1K x9"},
    {1, 16384, 9, 0, 0, 1, 0, "1 CTA, 16K x9, 4 pieces 1 thread", 4, 0},
    {1, 16384, 9, 0, 0, 1, 0, "1 CTA, 16K x9, 4 pieces 4 threads", 4, 1},
    {1, 16384, 9, 0, 0, 1, 0, "1 CTA, 16K x9, 16 pieces 16 threads", 16, 1},
    {1, 16384, 2, 0, 0, 1, 0, "1 CTA, 16K x2"},
    {1, 16384, 1, 0, 0, 1, 0, "1 CTA, 16K x1"},
    {148, 65536, 3, 0, 0, 1, 0, "148 CTA lockstep, 64K x3"},
    {148, 32768, 6, 0, 0, 1, 0, "148 CTA lockstep, 32K x6"},
    {148, 32768, 6, 1, 97, 1, 0, "148 CTA skew 97, 32K x6"},
    {148, 16384, 9, 0, 0, 1, 0, "148 CTA, 16K x9, 4 pieces 4 threads", 4, 1},
    {148, 8192, 18, 1, 97, 1, 128, "148 CTA skew, 8K x18 consumer 128"},
  };
  for (auto& c : cfgs) {
    const int n_stages = (int)(copy_bytes / c.stage);
    const int passes = 4;
    const int smem = 1024 + c.ring * c.stage;
    for (int rep = 0; rep < 2; ++rep) {
      stream_kernel<<<c.grid, 64, smem>>>(buf, copy_bytes, c.copies, c.stage, c.ring, n_stages, passes, c.mode, c.skew, cyc, c.delay, c.pieces, c.mt);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
    }
    std::vector<long long> h(148);
    cudaMemcpy(h.data(), cyc, c.grid * 8, cudaMemcpyDeviceToHost);
    double mean = 0; long long mx = 0;
    for (int i = 0; i < c.grid; ++i) { mean += h[i]; if (h[i] > mx) mx = h[i]; }
    mean /= c.grid;
    const double bytes = (double)copy_bytes * passes;
    printf("%-55s mean %.1f B/cyc/SM  (slowest CTA %.1f)  chip %.0f B/cyc  cyc/stage %.0f\n", c.name, bytes / mean, bytes / mx,
           bytes / mean * c.grid, mean / ((double)n_stages * passes));
  }
  return 0;
}
