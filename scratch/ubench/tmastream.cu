// Microbenchmark 2: same ring streaming, but with TENSOR TMA (cp.async.bulk.tensor.2d) over a 2-D view of the
// flat weight stream [rows][64 x u16 = 128 B]; box = 64 x R rows.  Compares with the 1-D bulk copy.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include "../../efficient-nerf_b200/csrc/tc_common.cuh"
using namespace r2l;

__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap tm, const uint8_t* buf, int use_tensor,
                                                       int stage_bytes, int ring, int n_stages, int passes,
                                                       long long* cycles, int consume_delay, unsigned* checksum) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + 32;
  uint8_t* data = smem + 1024;
  if (threadIdx.x == 0) {
    for (int i = 0; i < ring; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_fence_init();
  }
  __syncthreads();
  const long long total = (long long)n_stages * passes;
  const int rows_per_stage = stage_bytes / 128;
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    for (long long g = 0; g < total; ++g) {
      const int slot = g % ring;
      mbar_wait(&empty[slot], ((g / ring) & 1) ^ 1, nullptr, 0);
      mbar_expect_tx(&full[slot], stage_bytes);
      const int st = (int)(g % n_stages);
      if (use_tensor) {
        // boxes of at most 256 rows
        for (int r = 0; r < rows_per_stage; r += 256)
          tma_2d(data + (size_t)slot * stage_bytes + (size_t)r * 128, &tm, 0, st * rows_per_stage + r, &full[slot]);
      } else {
        bulk_g2s(data + (size_t)slot * stage_bytes, buf + (size_t)st * stage_bytes, stage_bytes, &full[slot]);
      }
    }
  } else if (threadIdx.x == 32) {
    unsigned acc = 0;
    for (long long g = 0; g < total; ++g) {
      const int slot = g % ring;
      mbar_wait(&full[slot], (g / ring) & 1, nullptr, 0);
      if (g < n_stages) acc += *reinterpret_cast<unsigned*>(data + (size_t)slot * stage_bytes + 4 * (g % 64)) ^
                               *reinterpret_cast<unsigned*>(data + (size_t)slot * stage_bytes + stage_bytes - 4);
      if (consume_delay > 0) { const long long c = clock64(); while (clock64() - c < consume_delay) {} }
      mbar_arrive(&empty[slot]);
    }
    cycles[blockIdx.x] = clock64() - t0;
    if (blockIdx.x == 0) *checksum = acc;
  }
}

int main() {
  const size_t bytes = 12ull << 20;
  uint8_t* buf; cudaMalloc(&buf, bytes);
  std::vector<uint32_t> h(bytes / 4);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint32_t)(i * 2654435761u);
  cudaMemcpy(buf, h.data(), bytes, cudaMemcpyHostToDevice);
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  unsigned* chk; cudaMalloc(&chk, 4);
  PFN_cuTensorMapEncodeTiled enc = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qres);
  if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  struct Cfg { int grid, stage, ring, tensor, box_rows, delay; const char* name; };
  std::vector<Cfg> cfgs = {
    {1, 16384, 9, 0, 0, 0, "1 CTA 1-D bulk 16K x9"},
    {1, 16384, 9, 1, 128, 0, "1 CTA tensor 16K x9 (box 128 rows)"},
    {1, 32768, 6, 1, 256, 0, "1 CTA tensor 32K x6 (box 256 rows)"},
    {1, 8192, 12, 1, 64, 0, "1 CTA tensor 8K x12 (box 64 rows)"},
    {148, 16384, 9, 1, 128, 0, "148 CTA tensor 16K x9"},
    {148, 32768, 6, 1, 256, 0, "148 CTA tensor 32K x6"},
    {148, 8192, 12, 1, 64, 0, "148 CTA tensor 8K x12"},
    {148, 16384, 4, 1, 128, 0, "148 CTA tensor 16K x4"},
    {148, 16384, 9, 1, 128, 256, "148 CTA tensor 16K x9, consumer 256 cyc/stage"},
    {148, 16384, 6, 1, 128, 256, "148 CTA tensor 16K x6, consumer 256 cyc/stage"},
    {148, 65536, 3, 1, 256, 0, "148 CTA tensor 64K x3 (2 boxes)"},
  };
  for (auto& c : cfgs) {
    CUtensorMap tm; memset(&tm, 0, sizeof(tm));
    if (c.tensor) {
      cuuint64_t gdim[2] = {64, bytes / 128};
      cuuint64_t gstr[1] = {128};
      cuuint32_t box[2] = {64, (cuuint32_t)c.box_rows};
      cuuint32_t estr[2] = {1, 1};
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", c.name, (int)r); continue; }
    }
    const int n_stages = (int)(bytes / c.stage), passes = 4, smem = 1024 + c.ring * c.stage;
    for (int rep = 0; rep < 2; ++rep) {
      stream_kernel<<<c.grid, 64, smem>>>(tm, buf, c.tensor, c.stage, c.ring, n_stages, passes, cyc, c.delay, chk);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
    }
    std::vector<long long> hc(148);
    cudaMemcpy(hc.data(), cyc, c.grid * 8, cudaMemcpyDeviceToHost);
    unsigned got; cudaMemcpy(&got, chk, 4, cudaMemcpyDeviceToHost);
    unsigned want = 0;
    for (int g = 0; g < n_stages; ++g) want += h[((size_t)g * c.stage) / 4 + (g % 64)] ^ h[((size_t)(g + 1) * c.stage) / 4 - 1];
    double mean = 0; long long mx = 0;
    for (int i = 0; i < c.grid; ++i) { mean += hc[i]; if (hc[i] > mx) mx = hc[i]; }
    mean /= c.grid;
    const double b = (double)bytes * passes;
    printf("%-50s %6.1f B/cyc/SM (slowest %6.1f) chip %6.0f B/cyc  cyc/stage %5.0f  data %s\n", c.name, b / mean, b / mx,
           b / mean * c.grid, mean / ((double)n_stages * passes), got == want ? "ok" : "MISMATCH");
  }
  return 0;
}
