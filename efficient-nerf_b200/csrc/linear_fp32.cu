// Full-precision (fp32 FMA, CUDA-core) linear layer:  Y = act(X W^T + b [+ R]).
// This is the `precision="fp32"` path of the NeRF / R2L forward: it exists so that every
// architecture variant of the reference's nn.Linear stacks (model/nerf_raybased.py:337-401,
// 443-544) has an exact-arithmetic CUDA implementation, and so that the tensor-core kernels
// can be validated against fp32 on the device.  The fused tcgen05 kernels (mlp_tc.cu) are
// the fast path.
//
// Classic 128x128x8 register-tiled SGEMM, 256 threads, 8x8 outputs per thread, both
// operands K-contiguous ("TN"), double-buffered through registers.
#include "common.cuh"

namespace r2l {

constexpr int BM = 128, BN = 128, BK = 8;

// act: 0 = identity, 1 = relu, 2 = sigmoid
__global__ void __launch_bounds__(256)
linear_fp32_kernel(int M, int N, int K, const float* __restrict__ X, long long ldx, const float* __restrict__ W,
                   long long ldw, const float* __restrict__ bias, float* __restrict__ Y, long long ldy, int act,
                   const float* __restrict__ R, long long ldr) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int lrow = tid >> 1, lk = (tid & 1) * 4;
  const int ty = tid >> 4, tx = tid & 15;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const bool a_ok = (m0 + lrow) < M, b_ok = (n0 + lrow) < N;
  const float* xa = X + static_cast<long long>(m0 + lrow) * ldx + lk;
  const float* wb = W + static_cast<long long>(n0 + lrow) * ldw + lk;
  auto ldg4 = [](const float* p, bool ok) -> float4 {
    return ok ? __ldg(reinterpret_cast<const float4*>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  const int nk = (K + BK - 1) / BK;
  float4 ra = ldg4(xa, a_ok && lk < K), rb = ldg4(wb, b_ok && lk < K);
  auto sts = [&](int buf) {
    As[buf][lk + 0][lrow] = ra.x;
    As[buf][lk + 1][lrow] = ra.y;
    As[buf][lk + 2][lrow] = ra.z;
    As[buf][lk + 3][lrow] = ra.w;
    Bs[buf][lk + 0][lrow] = rb.x;
    Bs[buf][lk + 1][lrow] = rb.y;
    Bs[buf][lk + 2][lrow] = rb.z;
    Bs[buf][lk + 3][lrow] = rb.w;
  };
  sts(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) {
      const int k = (kt + 1) * BK + lk;
      ra = ldg4(xa + (kt + 1) * BK, a_ok && k < K);
      rb = ldg4(wb + (kt + 1) * BK, b_ok && k < K);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sts(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= N) continue;
      float v = acc[i][j] + (bias != nullptr ? __ldg(bias + n) : 0.f);
      if (R != nullptr) v += R[static_cast<long long>(m) * ldr + n];
      if (act == 1)
        v = fmaxf(v, 0.f);
      else if (act == 2)
        v = 1.0f / (1.0f + expf(-v));
      Y[static_cast<long long>(m) * ldy + n] = v;
    }
  }
}

}  // namespace r2l

using namespace r2l;

extern "C" {

// X [M, K] (row stride ldx), W [N, K] (row stride ldw), Y [M, N] (row stride ldy), optional
// residual R [M, N] (row stride ldr).  K, ldx, ldw must be multiples of 4 and X, W 16-byte
// aligned (rows are read as float4); pad with zeros on the host side.
int r2l_linear_fp32(long long M, int N, int K, const float* X, long long ldx, const float* W, long long ldw,
                    const float* bias, float* Y, long long ldy, int act, const float* R, long long ldr,
                    void* stream) {
  R2L_CHECK_ARG(M >= 0 && N > 0 && K > 0, "r2l_linear_fp32: bad sizes");
  R2L_CHECK_ARG(K % 4 == 0 && ldx % 4 == 0 && ldw % 4 == 0, "r2l_linear_fp32: K/ldx/ldw must be multiples of 4");
  R2L_CHECK_ARG(act >= 0 && act <= 2, "r2l_linear_fp32: bad activation");
  if (M == 0) return R2L_OK;
  R2L_CHECK_ARG(X && W && Y, "r2l_linear_fp32: null pointer");
  R2L_CHECK_ARG(((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(W)) & 15) == 0,
                "r2l_linear_fp32: X and W must be 16-byte aligned");
  R2L_CHECK_ARG(M < (1LL << 31) - BM, "r2l_linear_fp32: M too large");
  dim3 grid((N + BN - 1) / BN, static_cast<unsigned>((M + BM - 1) / BM));
  R2L_CHECK_ARG(grid.y <= 65535, "r2l_linear_fp32: M too large for one launch (chunk it)");
  linear_fp32_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<int>(M), N, K, X, ldx, W,
                                                                          ldw, bias, Y, ldy, act, R, ldr);
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

}  // extern "C"
