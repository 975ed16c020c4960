// TMEM -> register read throughput (tcgen05.ld.32x32b.x32): NW warps of one CTA sweep a 128-lane x 512-column
// accumulator region repeatedly.  Reports bytes / cycle / SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../efficient-nerf_b200/csrc/mlp_tc.cuh"
using namespace r2l;

// mma_mode: 0 = no MMAs; 1 = a 17th warp issues M128xN256xK16 MMAs back to back into columns [0,256) while the
// loads sweep columns [256,512); 2 = MMAs and loads on the same columns [0,256)
template <int SHAPE>
__global__ void __launch_bounds__(544, 1) k(int nwarps, int iters, long long* out, uint32_t* sink, int mma_mode,
                                            int do_sts) {
  extern __shared__ __align__(1024) uint8_t dsm[];   // 64 KiB A + 32 KiB B of garbage operands + 64 KiB STS target
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  __shared__ volatile int stop;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); stop = 0; }
  for (int i = threadIdx.x; i < (64 + 32) * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(dsm)[i] = 0;
  fence_proxy_async_smem();
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (mma_mode == 1 ? 256 : 0);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp == 16) {
    if (mma_mode != 0 && (threadIdx.x & 31) == 0) {
      const uint32_t idesc = make_idesc_f16(false, 128, 256);
      const uint32_t aA = smem_u32(dsm), aB = smem_u32(dsm + 64 * 1024);
      uint32_t par = 0;
      long long n = 0;
      while (!stop) {
        for (int j = 0; j < 16; ++j)
          umma_f16_ss(slot, make_smem_desc(aA + (j & 15) * 4096, 2048, 128), make_smem_desc(aB + (j & 3) * 8192, 4096, 128), idesc, 1u);
        umma_commit(&bar);
        mbar_wait(&bar, par, nullptr, 0);
        par ^= 1u;
        n += 16;
      }
      out[148 + blockIdx.x] = n;
    }
  } else if (warp < nwarps) {
    const int groups = nwarps / 4;          // warps sharing a lane quarter split the 512 columns
    const int g = warp >> 2;
    const int cols_per = (mma_mode ? 256 : 512) / groups;
    uint8_t* const sts_dst = dsm + 96 * 1024 + (threadIdx.x & 127) * 16;
    for (int it = 0; it < iters; ++it) {
      for (int c = g * cols_per; c < (g + 1) * cols_per; c += 64) {
        uint32_t va[32], vb[32];
        tmem_ld32(base + c, va);
        tmem_ld32(base + c + 32, vb);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= va[i] + vb[i];
        if (do_sts == 2) {   // the real thing: cvt.rn.relu.f16x2 + 4 x 16-byte stores per 32 columns (mlp_tc.cuh)
          store_sub<false, true>(va, dsm + 96 * 1024 + ((c >> 5) & 7) * kSubBytes + (threadIdx.x & 127) * 16);
          store_sub<false, true>(vb, dsm + 96 * 1024 + (((c >> 5) + 1) & 7) * kSubBytes + (threadIdx.x & 127) * 16);
        } else if (do_sts == 3) {   // + fp32 bias add from shared memory (broadcast 16-byte loads)
          const float4* b4 = reinterpret_cast<const float4*>(dsm + 64 * 1024);
#pragma unroll
          for (int i4 = 0; i4 < 8; ++i4) {
            const float4 ba = b4[((c >> 2) + i4) & 63], bb = b4[((c >> 2) + 8 + i4) & 63];
            va[4 * i4 + 0] = __float_as_uint(__uint_as_float(va[4 * i4 + 0]) + ba.x);
            va[4 * i4 + 1] = __float_as_uint(__uint_as_float(va[4 * i4 + 1]) + ba.y);
            va[4 * i4 + 2] = __float_as_uint(__uint_as_float(va[4 * i4 + 2]) + ba.z);
            va[4 * i4 + 3] = __float_as_uint(__uint_as_float(va[4 * i4 + 3]) + ba.w);
            vb[4 * i4 + 0] = __float_as_uint(__uint_as_float(vb[4 * i4 + 0]) + bb.x);
            vb[4 * i4 + 1] = __float_as_uint(__uint_as_float(vb[4 * i4 + 1]) + bb.y);
            vb[4 * i4 + 2] = __float_as_uint(__uint_as_float(vb[4 * i4 + 2]) + bb.z);
            vb[4 * i4 + 3] = __float_as_uint(__uint_as_float(vb[4 * i4 + 3]) + bb.w);
          }
          store_sub<false, true>(va, dsm + 96 * 1024 + ((c >> 5) & 7) * kSubBytes + (threadIdx.x & 127) * 16);
          store_sub<false, true>(vb, dsm + 96 * 1024 + (((c >> 5) + 1) & 7) * kSubBytes + (threadIdx.x & 127) * 16);
        } else if (do_sts) {   // what the MLP epilogue stores: 8 x 16 bytes per 64 columns, 512 contiguous bytes per warp
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<uint4*>(sts_dst + ((c / 8 + q) & 31) * 2048) = make_uint4(va[4 * q], va[4 * q + 1], vb[4 * q], vb[4 * q + 1]);
        }
      }
    }
    __syncwarp();
    if (threadIdx.x == 0) stop = 1;   // warp 0 done -> stop the MMA warp (all load warps take the same time)
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[threadIdx.x] = acc;
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc(slot, 512); }
}

int main() {
  long long* out; uint32_t* sink;
  cudaMalloc(&out, 8 * 148 * 2); cudaMalloc(&sink, 4 * 544);
  const int iters = 64;
  const int smem = 160 * 1024;
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int sts = 0; sts < 4; ++sts)
  for (int mode = 0; mode < 2; ++mode)
  for (int nw : {8, 16}) {
    k<0><<<148, 544, smem>>>(nw, iters, out, sink, mode, sts);
    cudaDeviceSynchronize();
    k<0><<<148, 544, smem>>>(nw, iters, out, sink, mode, sts);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[296]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0, mm = 0; for (int i = 0; i < 148; ++i) { avg += h[i]; mm += h[148 + i]; } avg /= 148; mm /= 148;
    const double bytes = 128.0 * (mode ? 256 : 512) * 4 * iters;
    printf("sts %d mma_mode %d warps %2d: %.1f B/cycle/SM; one 128x256 fp32 accumulator = %.0f cycles; MMA pipe busy %.0f%%\n",
           sts, mode, nw, bytes / avg, 131072.0 / (bytes / avg), mode ? 100.0 * mm * 128 / avg : 0.0);
  }
  return 0;
}
