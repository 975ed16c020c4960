// C-ABI front end of the fused MLP kernels: weight packing (reference nn.Linear layout ->
// 16-bit K-chunk-major stage stream), model handles, forward entry points and a minimal
// tcgen05 GEMM probe used by the GPU unit tests.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "mlp_params.cuh"
#include "mlp_tc.cuh"

namespace r2l {

// ---------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------
static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}
static std::atomic<long long> g_kernel_launches{0};
void count_launch() { g_kernel_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ---------------------------------------------------------------------------------
// weight packing
// ---------------------------------------------------------------------------------
// W [N, *] fp32 row-major (row stride ldw)  ->  stage stream of 16-bit values.
// Logical operand B[n][k], k in [0, Kpad): source column kmap[k] (or k when kmap == nullptr),
// -1 / out-of-range -> 0.  Stage s holds k in [64s, 64s+64) (kStageK), k-chunk major:
//   element offset = s*(N*64) + ((k%64)/8)*(N*8) + n*8 + (k%8)
// pair != 0: CTA-pair layout — every stage is split into two N-halves (rows [0,N/2) for the leader CTA, then rows
// [N/2,N) for its peer), each half k-chunk major with N/2 rows, so each CTA bulk-copies one contiguous half stage.
// g_pack_max: bit pattern of the largest |value| packed since it was last reset (non-negative floats order like their
// bit patterns; NaN patterns are larger than every finite one, so a NaN weight fails the range check too).
__device__ unsigned int g_pack_max;

__device__ __forceinline__ void track_max(float v) {
  const unsigned int m = __reduce_max_sync(__activemask(), __float_as_uint(fabsf(v)));
  if ((threadIdx.x & 31) == (__ffs(__activemask()) - 1) && m > 0x477fe000u) atomicMax(&g_pack_max, m);   // > 65504
}

__device__ __forceinline__ void pack_layer_elem(const float* __restrict__ W, long long ldw, int N, int K_src, int Kpad,
                                                const int* __restrict__ kmap, float scale, uint16_t* __restrict__ dst,
                                                int bf16, int pair, long long idx) {
  {
    const int n = static_cast<int>(idx / Kpad), k = static_cast<int>(idx % Kpad);
    const int ks = (kmap != nullptr) ? kmap[k] : k;
    float v = 0.0f;
    if (ks >= 0 && ks < K_src) v = W[n * ldw + ks] * scale;
    track_max(v);
    uint16_t bits;
    if (bf16)
      bits = __bfloat16_as_ushort(__float2bfloat16_rn(v));
    else
      bits = __half_as_ushort(__float2half_rn(v));
    const int s = k / kStageK, kk = k % kStageK;
    const int kst = (Kpad - s * kStageK < kStageK) ? (Kpad - s * kStageK) : kStageK;   // K columns in this stage
    if (pair == 2) {
      // v3 layout (CTA pair + N-half split): 64-row blocks ordered (CTA r, half h), n = 128h + 64r + i
      const int h = n / 128, r = (n % 128) / 64, i = n % 64, nh = N / 128;
      dst[static_cast<long long>(s) * N * kStageK + (static_cast<long long>(r) * nh + h) * 64 * kst + (kk >> 3) * 512 +
          i * 8 + (kk & 7)] = bits;
    } else if (pair) {
      const int Nh = N / 2, h = n / Nh, nn = n % Nh;
      dst[static_cast<long long>(s) * N * kStageK + static_cast<long long>(h) * Nh * kst + (kk >> 3) * (Nh * 8) + nn * 8 +
          (kk & 7)] = bits;
    } else {
      dst[static_cast<long long>(s) * N * kStageK + (kk >> 3) * (N * 8) + n * 8 + (kk & 7)] = bits;
    }
  }
}
__global__ void pack_layer_kernel(const float* __restrict__ W, long long ldw, int N, int K_src, int Kpad,
                                  const int* __restrict__ kmap, float scale, uint16_t* __restrict__ dst, int bf16,
                                  int pair) {
  const long long total = static_cast<long long>(N) * Kpad;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x)
    pack_layer_elem(W, ldw, N, K_src, Kpad, kmap, scale, dst, bf16, pair, idx);
}

// 16-bit hi / lo split of a bias value: hi = rn16(b), lo = rn16(b - hi); hi + lo reproduces b to ~2^-22 relative
// (fp16) / 2^-16 (bf16).  Both are multiplied by exact 1.0 operands and accumulated in fp32 by the MMA.
__device__ __forceinline__ uint16_t bias_part(float b, int part, int bf16) {
  if (bf16) {
    const __nv_bfloat16 hi = __float2bfloat16_rn(b);
    return (part == 0) ? __bfloat16_as_ushort(hi) : __bfloat16_as_ushort(__float2bfloat16_rn(b - __bfloat162float(hi)));
  }
  const __half hi = __float2half_rn(b);
  return (part == 0) ? __half_as_ushort(hi) : __half_as_ushort(__float2half_rn(b - __half2float(hi)));
}

// Bias step of a layer: one K=16 stage [N x 16] whose k=0 / k=1 columns hold the hi / lo split of scale*bias
// (the matching A operand is the constant ones block), all other columns zero.
__device__ __forceinline__ void pack_bias_elem(const float* __restrict__ bias, int N, float scale,
                                               uint16_t* __restrict__ dst, int bf16, int pair, int idx) {
  const int n = idx / 16, k = idx % 16;
  track_max(k == 0 ? bias[n] * scale : 0.0f);
  const uint16_t bits = (k < 2) ? bias_part(bias[n] * scale, k, bf16) : static_cast<uint16_t>(0);
  if (pair == 2) {   // v3: 64-row blocks ordered (CTA r, half h), n = 128h + 64r + i
    const int h = n / 128, r = (n % 128) / 64, i = n % 64, nh = N / 128;
    dst[(r * nh + h) * 64 * 16 + (k >> 3) * 512 + i * 8 + (k & 7)] = bits;
  } else if (pair) {   // two N-halves, each k-chunk major with N/2 rows (see pack_layer_kernel)
    const int Nh = N / 2, h = n / Nh, nn = n % Nh;
    dst[h * (Nh * 16) + (k >> 3) * (Nh * 8) + nn * 8 + (k & 7)] = bits;
  } else {
    dst[(k >> 3) * (N * 8) + n * 8 + (k & 7)] = bits;
  }
}
__global__ void pack_bias_kernel(const float* __restrict__ bias, int N, float scale, uint16_t* __restrict__ dst,
                                 int bf16, int pair) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * 16) return;
  pack_bias_elem(bias, N, scale, dst, bf16, pair, idx);
}
// The 2 * n_blocks body layers of an R2L net (all 256 x 256, each [bias stage][weight stages]) in ONE launch:
// blockIdx.y = body layer l (even: W1 / b1 of block l / 2, odd: W2 / b2 scaled by res_scale).  ptrs = device array
// [w1 (n_blocks) | b1 | w2 | b2] of the caller's parameter pointers.
__global__ void pack_r2l_body_kernel(const float* const* __restrict__ ptrs, int n_blocks, float res_scale,
                                     uint16_t* __restrict__ dst, int bf16, int pair) {
  const int l = blockIdx.y, b = l >> 1, second = l & 1;
  const float* W = ptrs[(second ? 2 : 0) * n_blocks + b];
  const float* bias = ptrs[(second ? 3 : 1) * n_blocks + b];
  const float scale = second ? res_scale : 1.0f;
  uint16_t* d = dst + static_cast<long long>(l) * (256 * 16 + 256 * 256);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 256 * 16 + 256 * 256; idx += gridDim.x * blockDim.x) {
    if (idx < 256 * 16)
      pack_bias_elem(bias, 256, scale, d, bf16, pair, idx);
    else
      pack_layer_elem(W, 256, 256, 256, 256, nullptr, scale, d + 256 * 16, bf16, pair, idx - 256 * 16);
  }
}

// View stage of the NeRF view branch: [128 x 32], k < 27 -> views_w[n][256 + k] (embedded view direction),
// k = 27 / 28 -> hi / lo split of views_b[n] (the V block holds 1.0 there), k = 29..31 -> 0.
__global__ void pack_view_stage_kernel(const float* __restrict__ views_w, long long ldw, const float* __restrict__ views_b,
                                       uint16_t* __restrict__ dst, int bf16) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 128 * 32) return;
  const int n = idx / 32, k = idx % 32;
  uint16_t bits = 0;
  if (k < 27) {
    const float v = views_w[n * ldw + 256 + k];
    track_max(v);
    bits = bf16 ? __bfloat16_as_ushort(__float2bfloat16_rn(v)) : __half_as_ushort(__float2half_rn(v));
  } else if (k < 29) {
    bits = bias_part(views_b[n], k - 27, bf16);
  }
  dst[(k >> 3) * (128 * 8) + n * 8 + (k & 7)] = bits;
}

struct Mlp {
  int kind = 0;  // 0 = NeRF, 1 = R2L ResMLP
  bool bf16 = false;
  int pair = 0;   // weight layout / kernel: 0 single CTA, 1 CTA pair (cta_group::2), 2 "v3" (pair + N-half split)
  uint8_t* wstream = nullptr;
  size_t wbytes = 0;
  float* aux = nullptr;
  DebugBuf* dbg_host = nullptr;  // pinned, mapped: readable after a device trap
  DebugBuf* dbg_dev = nullptr;
  // NeRF
  float alpha_b = 0.f, rgb_b[3] = {0.f, 0.f, 0.f};
  int nerf_pp = 0;                 // 1: CTA-pair ping-pong kernel (mlp_nerf_pp.cu)
  uint8_t* wstream_pp = nullptr;   // pair-layout stage stream without the view stage
  float* view_tab = nullptr;       // [128][27] view half of views_linears[0].weight, then [128] its bias (fp32)
  float* vb_ws = nullptr;          // workspace: per-ray view-branch bias [vb_cap rays][128]
  long long vb_cap = 0;
  NerfPpMaps* pp_maps = nullptr;   // tensor maps over wstream_pp
  NerfHeadW* head_w = nullptr;     // host copy of alpha_linear / rgb_linear weights (kernel parameter, constant bank)
  // far-sample sigma fix-up (nerf_far.cu): fp32 transposed point layers, the flagged-ray list and its counters
  float* far_wt = nullptr;
  int* far_ws = nullptr;           // [0] count of this forward, [1] count of the last finished forward, [2..] ray list
  long long far_cap = 0;
  int far_mode = 1;                // 0: off (raw sigma of the 16-bit kernels), 1: flag + fp32 fix-up
  float far_abs = 1e-4f, far_rel = 9.765625e-4f;   // guard band: |sigma| < max(far_abs, far_rel * sum |w_a| relu(h7))
  // fused compositing (r2l_nerf_render): per-CTA staging rings + far slots, the compact copy of the flagged rays' rows,
  // and the raw workspace of the unfused route
  float4* comp_ring = nullptr;
  int* comp_aux = nullptr;
  float4* comp_far_raw = nullptr;
  size_t comp_far_bytes = 0;
  float* raw_ws = nullptr;
  size_t raw_ws_bytes = 0;
  int comp_mode = 1;               // 0: r2l_nerf_render always takes the unfused route (MLP -> raw -> raw2outputs)
  R2lPairMaps* r2l_maps = nullptr; // R2L pair mode: tensor maps over wstream
  // R2L
  int n_points = 0, n_blocks = 0, sigmoid_out = 1, outer_skip = 1;
  int r2l_pp = 0;                  // 1: ping-pong kernel (needs pair == 1)
  float b_tail[3] = {0.f, 0.f, 0.f};
};

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
static int encode_rows_map(CUtensorMap* out, void* base, size_t bytes, int box_rows) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    R2L_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
    if (sym == nullptr || q != cudaDriverEntryPointSuccess)
      return fail(R2L_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    fn = reinterpret_cast<EncodeFn>(sym);
  }
  const cuuint64_t dims[2] = {256, static_cast<cuuint64_t>(bytes / 512)};
  const cuuint64_t strides[1] = {512};
  const cuuint32_t box[2] = {256, static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(R2L_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  return R2L_OK;
}

// NeRF handles use the CTA-pair ping-pong kernel (mlp_nerf_pp.cu: 41.6 vs 44.8 ms per 400x400 frame on B200);
// R2L_NERF_PP=0 selects the single-CTA chasing kernel (mlp_nerf.cu) instead.
static int nerf_pp_default() {
  const char* e = getenv("R2L_NERF_PP");
  return (e != nullptr && e[0] == '0') ? 0 : 1;
}
// R2L handles use the CTA-pair kernel (tcgen05.mma.cta_group::2: half the L2 -> shared-memory weight traffic and half
// the B-operand shared-memory reads per FLOP).  With a relay thread forwarding the peer's stage arrivals it only
// matched the single-CTA kernel (1.677 vs 1.678 ms per frame: higher clocks under the power cap, more cycles);
// with the peer's half stages signalling the leader's barrier directly (cp.async.bulk.tensor...cta_group::2) it is
// 1.9 % faster sustained (99.6 vs 97.8 Mrays/s at 1.68 vs 1.57 GHz).  R2L_PAIR=0 selects the single-CTA kernel.
static int pair_mode_default() {
  const char* e = getenv("R2L_PAIR");
  return (e != nullptr && e[0] == '0') ? 0 : 1;
}
// R2L_PP=1 selects the ping-pong kernel (mlp_r2l_pp.cu: two 64-row tiles per CTA, M = 128 cta_group::2 MMAs, same packed
// weights, results equal to a few ulp of the fp32 tail sum) for pair-mode handles.  Measured on B200 (A/B in one process, 4 frames per launch):
// 102.1 vs 104.5 Mrays/s for the one-tile-per-CTA pair kernel of mlp_r2l.cu, at 1612 vs 1630 MHz under the power cap
// — the tensor pipe is busier (72 % vs 67 % by the in-kernel counters) but every M = 128 MMA re-reads the weight
// operand for half as many rows (96 instead of 64 B/cycle of operand reads), which costs more energy per FLOP than the
// overlap buys.  So the default stays 0; the kernel is kept as a tested alternative.
static int r2l_pp_default() {
  const char* e = getenv("R2L_PP");
  return (e != nullptr && e[0] == '1') ? 1 : 0;
}

// r2l_nerf_render's route.  Fused compositing (R2L_NERF_FUSED=1 or r2l_nerf_render_mode) removes the raw [N,S,4] round
// trip — 1.3 GB of HBM traffic per 400x400 frame — and is bit-identical, but the frame is NOT faster: A/B on one box,
// 160 000 rays x (64 + 192) samples: two-step route 35.9-36.0 ms, fused 36.4-36.5 ms, the same kernel with the
// compositor's body skipped 35.95 ms.  The compositor runs on the encoder warps, which share their schedulers with the
// epilogue warps: its ~1 k instructions per unit land on the epilogues every tile-layer waits for.  The MLP kernel is
// tensor-bound, the stand-alone raw2outputs runs at 80 % of the HBM roofline and costs 0.23 ms; so the default is 0.
static int nerf_fused_default() {
  const char* e = getenv("R2L_NERF_FUSED");
  return (e != nullptr && e[0] == '1') ? 1 : 0;
}

// Range check of everything packed between pack_max_reset and pack_max_check (fp16 operands only: bf16 has fp32's range)
static int pack_max_reset(cudaStream_t st) {
  const unsigned int zero = 0;
  R2L_CUDA(cudaMemcpyToSymbolAsync(g_pack_max, &zero, sizeof(zero), 0, cudaMemcpyHostToDevice, st));
  return R2L_OK;
}
static int pack_max_check(bool bf16, cudaStream_t st, const char* who) {
  unsigned int bits = 0;
  R2L_CUDA(cudaMemcpyFromSymbolAsync(&bits, g_pack_max, sizeof(bits), 0, cudaMemcpyDeviceToHost, st));
  R2L_CUDA(cudaStreamSynchronize(st));
  if (!bf16 && bits > 0x477fe000u) {
    float v;
    memcpy(&v, &bits, sizeof(v));
    return fail(R2L_ERR_RANGE, "%s: a weight / bias of magnitude %g does not fit fp16 operands (max 65504): use "
                "precision 'bf16' or 'fp32' for this model", who, static_cast<double>(v));
  }
  return R2L_OK;
}

static int alloc_debug(Mlp* m) {
  R2L_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&m->dbg_host), sizeof(DebugBuf), cudaHostAllocMapped));
  memset(m->dbg_host, 0, sizeof(DebugBuf));
  R2L_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&m->dbg_dev), m->dbg_host, 0));
  return R2L_OK;
}

static int pack_layer(const float* W, long long ldw, int N, int K_src, int Kpad, const std::vector<int>* kmap,
                      float scale, uint16_t* dst, bool bf16, cudaStream_t st, std::vector<int*>& scratch,
                      int pair = 0) {
  int* d_kmap = nullptr;
  if (kmap != nullptr) {
    R2L_CUDA(cudaMalloc(reinterpret_cast<void**>(&d_kmap), sizeof(int) * Kpad));
    scratch.push_back(d_kmap);
    R2L_CUDA(cudaMemcpyAsync(d_kmap, kmap->data(), sizeof(int) * Kpad, cudaMemcpyHostToDevice, st));
  }
  const long long total = static_cast<long long>(N) * Kpad;
  const int blocks = static_cast<int>((total + 255) / 256);
  pack_layer_kernel<<<blocks, 256, 0, st>>>(W, ldw, N, K_src, Kpad, d_kmap, scale, dst, bf16 ? 1 : 0, pair);
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

static int pack_bias(const float* bias, int N, float scale, uint16_t* dst, bool bf16, cudaStream_t st,
                     int pair = 0) {
  pack_bias_kernel<<<(N * 16 + 255) / 256, 256, 0, st>>>(bias, N, scale, dst, bf16 ? 1 : 0, pair);
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

static void destroy(Mlp* m) {
  if (m == nullptr) return;
  if (m->wstream) cudaFree(m->wstream);
  if (m->wstream_pp) cudaFree(m->wstream_pp);
  if (m->view_tab) cudaFree(m->view_tab);
  if (m->vb_ws) cudaFree(m->vb_ws);
  if (m->far_wt) cudaFree(m->far_wt);
  if (m->far_ws) cudaFree(m->far_ws);
  if (m->comp_ring) cudaFree(m->comp_ring);
  if (m->comp_aux) cudaFree(m->comp_aux);
  if (m->comp_far_raw) cudaFree(m->comp_far_raw);
  if (m->raw_ws) cudaFree(m->raw_ws);
  if (m->aux) cudaFree(m->aux);
  if (m->dbg_host) cudaFreeHost(m->dbg_host);
  delete m->pp_maps;
  delete m->head_w;
  delete m->r2l_maps;
  delete m;
}

// NeRF aux layout (floats): the two heads that run on CUDA cores
constexpr int kNerfAuxAlphaW = 0;                  // 256
constexpr int kNerfAuxRgbW = kNerfAuxAlphaW + 256; // 384
constexpr int kNerfAuxTotal = kNerfAuxRgbW + 384;

// ---------------------------------------------------------------------------------
// tcgen05 GEMM probe: D[128, N] = A[128, K] * B[N, K]^T with 16-bit operands, fp32 accumulate.
// Single CTA, no pipelining; exists to unit-test the descriptor / layout conventions.
// ---------------------------------------------------------------------------------
template <bool BF16>
__global__ void __launch_bounds__(128, 1)
gemm_probe_kernel(const float* __restrict__ A, int K, const uint16_t* __restrict__ Bpacked, int N,
                  float* __restrict__ D, int swap_lbo_sbo, DebugBuf* dbg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                                  // 128 x K 16-bit, chunk major
  uint8_t* sB = smem + kTileM * K * 2;                 // N x K 16-bit, stage stream as packed
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + N * K * 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = threadIdx.x;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // A -> shared (16-bit, k-chunk major)
  for (int ch = 0; ch < K / 8; ++ch) {
    const float* a = A + static_cast<long long>(row) * K + ch * 8;
    uint4 q;
    q.x = pack2<BF16>(a[0], a[1]);
    q.y = pack2<BF16>(a[2], a[3]);
    q.z = pack2<BF16>(a[4], a[5]);
    q.w = pack2<BF16>(a[6], a[7]);
    *reinterpret_cast<uint4*>(sA + ch * kChunkBytes + row * 16) = q;
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t bytes = static_cast<uint32_t>(N) * K * 2;
    mbar_expect_tx(&bars[0], bytes);
    // bulk copies are limited in size per instruction; move stage by stage
    // bulk copies in pieces of 32 K-columns (the packed stream is contiguous)
    const uint32_t piece_bytes = static_cast<uint32_t>(N) * 64;
    for (uint32_t off = 0; off < bytes; off += piece_bytes)
      bulk_g2s(sB + off, reinterpret_cast<const uint8_t*>(Bpacked) + off, piece_bytes, &bars[0]);
    mbar_wait(&bars[0], 0, dbg, 1);
    tc_fence_after_sync();
    const uint32_t idesc = make_idesc_f16(BF16, kTileM, N);
    const uint32_t lbo_b = static_cast<uint32_t>(N) * 16;
    // K-step ks (16 columns) lives in stage ks/4 (a partial last stage holds K%64 columns, still k-chunk major)
    for (int ks = 0; ks < K / 16; ++ks) {
      {
        const int s = ks / 4, j = ks % 4;
        const uint32_t a_addr = smem_u32(sA) + ks * 2 * kChunkBytes;
        const uint32_t b_addr = smem_u32(sB) + s * (static_cast<uint32_t>(N) * kStageK * 2) + j * 2 * lbo_b;
        const uint64_t ad = swap_lbo_sbo ? make_smem_desc(a_addr, kSbo, kLboA) : make_smem_desc(a_addr, kLboA, kSbo);
        const uint64_t bd = swap_lbo_sbo ? make_smem_desc(b_addr, kSbo, lbo_b) : make_smem_desc(b_addr, lbo_b, kSbo);
        umma_f16_ss(tmem_base, ad, bd, idesc, ks != 0 ? 1u : 0u);
      }
    }
    umma_commit(&bars[1]);
  }
  __syncwarp();
  mbar_wait(&bars[1], 0, dbg, 2);
  tc_fence_after_sync();
  const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  for (int c = 0; c < N; c += 32) {
    uint32_t v[32];
    tmem_ld32(lane_taddr + c, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) D[static_cast<long long>(row) * N + c + i] = __uint_as_float(v[i]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 256);
  }
  (void)lane;
}

// CTA-pair probe: D[256, N] = A[256, K] * B[N, K]^T with tcgen05.mma.cta_group::2.  Two CTAs (one cluster): CTA r
// holds A rows [128r, 128r+128) and the N-half r of B; the leader issues the MMAs, both read their own 128 D rows.
template <bool BF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
gemm_probe_pair_kernel(const float* __restrict__ A, int K, const uint16_t* __restrict__ Bpacked, int N,
                       float* __restrict__ D, DebugBuf* dbg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int Nh = N / 2;
  uint8_t* sA = smem;                                  // 128 x K 16-bit, chunk major
  uint8_t* sB = smem + kTileM * K * 2;                 // Nh x K 16-bit: this CTA's half of every stage
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + Nh * K * 2);   // [0] B landed (local), [1] done, [2] peer ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  const int row = threadIdx.x;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc_pair(tmem_slot, 256);
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  for (int ch = 0; ch < K / 8; ++ch) {
    const float* a = A + (static_cast<long long>(rank) * kTileM + row) * K + ch * 8;
    uint4 q;
    q.x = pack2<BF16>(a[0], a[1]);
    q.y = pack2<BF16>(a[2], a[3]);
    q.z = pack2<BF16>(a[4], a[5]);
    q.w = pack2<BF16>(a[6], a[7]);
    *reinterpret_cast<uint4*>(sA + ch * kChunkBytes + row * 16) = q;
  }
  fence_proxy_async_smem();
  __syncthreads();
  const int n_stages = (K + kStageK - 1) / kStageK;
  if (threadIdx.x == 0) {
    // this CTA's half of every stage (a partial last stage holds K % 64 columns)
    mbar_expect_tx(&bars[0], static_cast<uint32_t>(Nh) * K * 2);
    for (int s = 0; s < n_stages; ++s) {
      const int kst = (K - s * kStageK < kStageK) ? (K - s * kStageK) : kStageK;
      const uint8_t* src = reinterpret_cast<const uint8_t*>(Bpacked) + static_cast<size_t>(s) * N * kStageK * 2 +
                           static_cast<size_t>(rank) * Nh * kst * 2;
      bulk_g2s(sB + static_cast<size_t>(s) * Nh * kStageK * 2, src, static_cast<uint32_t>(Nh) * kst * 2, &bars[0]);
    }
    mbar_wait(&bars[0], 0, dbg, 1);
    if (rank == 1) {
      mbar_arrive_cluster(mapa_u32(&bars[2], 0));      // tell the leader: my A and B halves are in place
    } else {
      while (!mbar_try_wait_cluster(&bars[2], 0)) {}
      tc_fence_after_sync();
      const uint32_t idesc = make_idesc_f16(BF16, 2 * kTileM, N);
      const uint32_t lbo_b = static_cast<uint32_t>(Nh) * 16;
      for (int ks = 0; ks < K / 16; ++ks) {
        const int s = ks / 4, j = ks % 4;
        const uint32_t a_addr = smem_u32(sA) + ks * 2 * kChunkBytes;
        const uint32_t b_addr = smem_u32(sB) + s * (static_cast<uint32_t>(Nh) * kStageK * 2) + j * 2 * lbo_b;
        umma_f16_ss_pair(tmem_base, make_smem_desc(a_addr, kLboA, kSbo), make_smem_desc(b_addr, lbo_b, kSbo), idesc,
                         ks != 0 ? 1u : 0u);
      }
      umma_commit_pair(&bars[1]);
    }
  }
  __syncwarp();
  {
    uint32_t spins = 0;
    while (!mbar_try_wait_cluster(&bars[1], 0)) {
      if (++spins > R2L_WATCHDOG_SPINS) __trap();
    }
  }
  tc_fence_after_sync();
  const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  for (int c = 0; c < N; c += 32) {
    uint32_t v[32];
    tmem_ld32(lane_taddr + c, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i)
      D[(static_cast<long long>(rank) * kTileM + row) * N + c + i] = __uint_as_float(v[i]);
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc_pair(tmem_base, 256);
  }
}

}  // namespace r2l

using namespace r2l;

extern "C" {

const char* r2l_last_error(void) { return g_last_error.c_str(); }
int r2l_abi_version(void) { return 7; }

// Number of CUDA kernels this library has launched so far in this process (all entry points, all streams).
long long r2l_kernel_launches(void) { return g_kernel_launches.load(std::memory_order_relaxed); }

// D [128, N] fp32 = A [128, K] fp32 (rounded to 16 bit) x W [N, K]^T fp32 (rounded to 16 bit).
// K multiple of 32, <= 320; N multiple of 32, 32..256.  dtype: 0 = fp16, 1 = bf16.
int r2l_tc_gemm_probe(int dtype, int N, int K, const float* A, const float* W, float* D, int swap_lbo_sbo,
                      void* stream) {
  R2L_CHECK_ARG(dtype == 0 || dtype == 1, "r2l_tc_gemm_probe: dtype must be 0 (fp16) or 1 (bf16)");
  R2L_CHECK_ARG(N >= 32 && N <= 256 && N % 32 == 0, "r2l_tc_gemm_probe: N must be a multiple of 32 in [32,256]");
  R2L_CHECK_ARG(K >= 32 && K <= 320 && K % 32 == 0, "r2l_tc_gemm_probe: bad K");
  R2L_CHECK_ARG(A && W && D, "r2l_tc_gemm_probe: null pointer");
  R2L_CHECK_ARG(kTileM * K * 2 + N * K * 2 + 64 <= 227 * 1024, "r2l_tc_gemm_probe: operands exceed shared memory");
  auto st = static_cast<cudaStream_t>(stream);
  uint16_t* packed = nullptr;
  R2L_CUDA(cudaMalloc(reinterpret_cast<void**>(&packed), static_cast<size_t>(N) * K * 2));
  std::vector<int*> scratch;
  int rc = pack_layer(W, K, N, K, K, nullptr, 1.0f, packed, dtype == 1, st, scratch);
  DebugBuf* dbg_host = nullptr;
  DebugBuf* dbg_dev = nullptr;
  if (rc == R2L_OK) {
    if (cudaHostAlloc(reinterpret_cast<void**>(&dbg_host), sizeof(DebugBuf), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(reinterpret_cast<void**>(&dbg_dev), dbg_host, 0) != cudaSuccess)
      rc = fail(R2L_ERR_CUDA, "r2l_tc_gemm_probe: debug buffer allocation failed");
    else
      memset(dbg_host, 0, sizeof(DebugBuf));
  }
  if (rc == R2L_OK) {
    const int smem = kTileM * K * 2 + N * K * 2 + 64;
    cudaError_t e;
    if (dtype == 1) {
      e = cudaFuncSetAttribute(gemm_probe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e == cudaSuccess) gemm_probe_kernel<true><<<1, 128, smem, st>>>(A, K, packed, N, D, swap_lbo_sbo, dbg_dev);
    } else {
      e = cudaFuncSetAttribute(gemm_probe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e == cudaSuccess) gemm_probe_kernel<false><<<1, 128, smem, st>>>(A, K, packed, N, D, swap_lbo_sbo, dbg_dev);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      rc = fail(R2L_ERR_CUDA, "r2l_tc_gemm_probe: %s (watchdog flag=%u barrier=%u)", cudaGetErrorString(e),
                dbg_host ? dbg_host->flag : 0u, dbg_host ? dbg_host->barrier_id : 0u);
    }
  }
  if (dbg_host) cudaFreeHost(dbg_host);
  cudaFree(packed);
  return rc;
}

// CTA-pair variant of the probe: D [256, N] = A [256, K] x W [N, K]^T on tcgen05.mma.cta_group::2 (one cluster of two
// CTAs).  K multiple of 32, <= 256; N multiple of 64, 64..256.
int r2l_tc_gemm_probe_pair(int dtype, int N, int K, const float* A, const float* W, float* D, void* stream) {
  R2L_CHECK_ARG(dtype == 0 || dtype == 1, "r2l_tc_gemm_probe_pair: dtype must be 0 (fp16) or 1 (bf16)");
  R2L_CHECK_ARG(N >= 64 && N <= 256 && N % 64 == 0, "r2l_tc_gemm_probe_pair: N must be a multiple of 64 in [64,256]");
  R2L_CHECK_ARG(K >= 32 && K <= 256 && K % 32 == 0, "r2l_tc_gemm_probe_pair: bad K");
  R2L_CHECK_ARG(A && W && D, "r2l_tc_gemm_probe_pair: null pointer");
  auto st = static_cast<cudaStream_t>(stream);
  uint16_t* packed = nullptr;
  R2L_CUDA(cudaMalloc(reinterpret_cast<void**>(&packed), static_cast<size_t>(N) * K * 2));
  std::vector<int*> scratch;
  int rc = pack_layer(W, K, N, K, K, nullptr, 1.0f, packed, dtype == 1, st, scratch, /*pair=*/1);
  DebugBuf* dbg_host = nullptr;
  DebugBuf* dbg_dev = nullptr;
  if (rc == R2L_OK) {
    if (cudaHostAlloc(reinterpret_cast<void**>(&dbg_host), sizeof(DebugBuf), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(reinterpret_cast<void**>(&dbg_dev), dbg_host, 0) != cudaSuccess)
      rc = fail(R2L_ERR_CUDA, "r2l_tc_gemm_probe_pair: debug buffer allocation failed");
    else
      memset(dbg_host, 0, sizeof(DebugBuf));
  }
  if (rc == R2L_OK) {
    const int smem = kTileM * K * 2 + (N / 2) * K * 2 + 64;
    cudaError_t e;
    if (dtype == 1) {
      e = cudaFuncSetAttribute(gemm_probe_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e == cudaSuccess) gemm_probe_pair_kernel<true><<<2, 128, smem, st>>>(A, K, packed, N, D, dbg_dev);
    } else {
      e = cudaFuncSetAttribute(gemm_probe_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e == cudaSuccess) gemm_probe_pair_kernel<false><<<2, 128, smem, st>>>(A, K, packed, N, D, dbg_dev);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess)
      rc = fail(R2L_ERR_CUDA, "r2l_tc_gemm_probe_pair: %s (watchdog flag=%u barrier=%u)", cudaGetErrorString(e),
                dbg_host ? dbg_host->flag : 0u, dbg_host ? dbg_host->barrier_id : 0u);
  }
  if (dbg_host) cudaFreeHost(dbg_host);
  cudaFree(packed);
  return rc;
}

// ------------------------------- NeRF ---------------------------------------------
// Weights in the reference's nn.Linear layout ([out, in] row-major fp32, device memory):
//   pts_w[0] [256,63], pts_w[1..4] [256,256], pts_w[5] [256,319], pts_w[6..7] [256,256]   (model:357-360)
//   views_w [128,283], feature_w [256,256], alpha_w [1,256], rgb_w [3,128]                   (model:363-372)
int r2l_nerf_create(void** out_handle, int dtype, const float* const* pts_w, const float* const* pts_b,
                    const float* views_w, const float* views_b, const float* feature_w, const float* feature_b,
                    const float* alpha_w, const float* alpha_b, const float* rgb_w, const float* rgb_b,
                    void* stream) {
  R2L_CHECK_ARG(out_handle != nullptr, "r2l_nerf_create: null out_handle");
  R2L_CHECK_ARG(dtype == 0 || dtype == 1, "r2l_nerf_create: dtype must be 0 (fp16) or 1 (bf16)");
  R2L_CHECK_ARG(pts_w && pts_b && views_w && views_b && feature_w && feature_b && alpha_w && alpha_b && rgb_w && rgb_b,
                "r2l_nerf_create: null weight pointer");
  for (int i = 0; i < 8; ++i) R2L_CHECK_ARG(pts_w[i] && pts_b[i], "r2l_nerf_create: null pts_linears[%d]", i);
  auto st = static_cast<cudaStream_t>(stream);
  Mlp* m = new Mlp();
  m->kind = 0;
  m->bf16 = dtype == 1;
  std::vector<int*> scratch;
  auto cleanup = [&](int rc) {
    cudaStreamSynchronize(st);
    for (int* s : scratch) cudaFree(s);
    if (rc != R2L_OK) destroy(m);
    return rc;
  };
  // stage stream in consumption order; every 256-wide layer is [K=16 bias stage][K/32 weight stages]:
  //   L0 (K 64) | L1..4 (K 256) | L5 (K 64 point part + 256) | L6,7 | feature | views: V stage (N 128, K 32) + K 256
  const size_t elems = 256ull * (64 + 4 * 256 + 320 + 2 * 256 + 256) + 9ull * 256 * 16 + 128ull * (32 + 256);
  m->wbytes = elems * 2;
  if (cudaMalloc(reinterpret_cast<void**>(&m->wstream), m->wbytes) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&m->aux), sizeof(float) * kNerfAuxTotal) != cudaSuccess)
    return cleanup(fail(R2L_ERR_CUDA, "r2l_nerf_create: cudaMalloc failed"));
  int rc = alloc_debug(m);
  if (rc == R2L_OK) rc = pack_max_reset(st);
  if (rc != R2L_OK) return cleanup(rc);
  uint16_t* dst = reinterpret_cast<uint16_t*>(m->wstream);
  std::vector<int> kmap0(64), kmap5(320);
  for (int k = 0; k < 64; ++k) kmap0[k] = (k < 63) ? k : -1;
  // skip layer: reference input is cat[pts(63), h(256)]; the kernel feeds the point block first, then h
  for (int k = 0; k < 320; ++k) kmap5[k] = (k < 63) ? k : ((k == 63) ? -1 : (63 + (k - 64)));
  size_t off = 0;
  for (int l = 0; l < 8 && rc == R2L_OK; ++l) {
    rc = pack_bias(pts_b[l], 256, 1.0f, dst + off, m->bf16, st);
    off += 256ull * 16;
    if (rc != R2L_OK) break;
    if (l == 0) {
      rc = pack_layer(pts_w[0], 63, 256, 63, 64, &kmap0, 1.0f, dst + off, m->bf16, st, scratch);
      off += 256ull * 64;
    } else if (l == 5) {
      rc = pack_layer(pts_w[5], 319, 256, 319, 320, &kmap5, 1.0f, dst + off, m->bf16, st, scratch);
      off += 256ull * 320;
    } else {
      rc = pack_layer(pts_w[l], 256, 256, 256, 256, nullptr, 1.0f, dst + off, m->bf16, st, scratch);
      off += 256ull * 256;
    }
  }
  if (rc == R2L_OK) {
    rc = pack_bias(feature_b, 256, 1.0f, dst + off, m->bf16, st);
    off += 256ull * 16;
  }
  if (rc == R2L_OK) {
    rc = pack_layer(feature_w, 256, 256, 256, 256, nullptr, 1.0f, dst + off, m->bf16, st, scratch);
    off += 256ull * 256;
  }
  if (rc == R2L_OK) {
    pack_view_stage_kernel<<<(128 * 32 + 255) / 256, 256, 0, st>>>(views_w, 283, views_b, dst + off, m->bf16 ? 1 : 0);
    count_launch();
    if (cudaGetLastError() != cudaSuccess) rc = fail(R2L_ERR_CUDA, "r2l_nerf_create: pack_view_stage launch failed");
    off += 128ull * 32;
  }
  if (rc == R2L_OK) {
    rc = pack_layer(views_w, 283, 128, 256, 256, nullptr, 1.0f, dst + off, m->bf16, st, scratch);
    off += 128ull * 256;
  }
  if (rc != R2L_OK) return cleanup(rc);
  if (off != elems) return cleanup(fail(R2L_ERR_INVALID, "r2l_nerf_create: internal stream size mismatch"));
  m->nerf_pp = nerf_pp_default();
  m->comp_mode = nerf_fused_default();
  if (m->nerf_pp) {
    // second stream for the ping-pong kernel: CTA-pair layout (two N-halves per stage), no view stage — the view
    // half of views_linears[0] becomes a per-ray fp32 bias (nerf_view_bias_kernel) from view_tab
    const size_t elems_pp = elems - 128ull * 32;   // no view stage
    if (cudaMalloc(reinterpret_cast<void**>(&m->wstream_pp), elems_pp * 2) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&m->view_tab), sizeof(float) * (128 * 27 + 128)) != cudaSuccess)
      return cleanup(fail(R2L_ERR_CUDA, "r2l_nerf_create: cudaMalloc failed"));
    uint16_t* d2 = reinterpret_cast<uint16_t*>(m->wstream_pp);
    size_t o2 = 0;
    for (int l = 0; l < 8 && rc == R2L_OK; ++l) {
      rc = pack_bias(pts_b[l], 256, 1.0f, d2 + o2, m->bf16, st, 1);
      o2 += 256ull * 16;
      if (rc != R2L_OK) break;
      if (l == 0) {
        rc = pack_layer(pts_w[0], 63, 256, 63, 64, &kmap0, 1.0f, d2 + o2, m->bf16, st, scratch, 1);
        o2 += 256ull * 64;
      } else if (l == 5) {
        rc = pack_layer(pts_w[5], 319, 256, 319, 320, &kmap5, 1.0f, d2 + o2, m->bf16, st, scratch, 1);
        o2 += 256ull * 320;
      } else {
        rc = pack_layer(pts_w[l], 256, 256, 256, 256, nullptr, 1.0f, d2 + o2, m->bf16, st, scratch, 1);
        o2 += 256ull * 256;
      }
    }
    if (rc == R2L_OK) {
      rc = pack_bias(feature_b, 256, 1.0f, d2 + o2, m->bf16, st, 1);
      o2 += 256ull * 16;
    }
    if (rc == R2L_OK) {
      rc = pack_layer(feature_w, 256, 256, 256, 256, nullptr, 1.0f, d2 + o2, m->bf16, st, scratch, 1);
      o2 += 256ull * 256;
    }
    if (rc == R2L_OK) {
      rc = pack_layer(views_w, 283, 128, 256, 256, nullptr, 1.0f, d2 + o2, m->bf16, st, scratch, 1);
      o2 += 128ull * 256;
    }
    if (rc != R2L_OK) return cleanup(rc);
    if (o2 != elems_pp) return cleanup(fail(R2L_ERR_INVALID, "r2l_nerf_create: internal pp stream size mismatch"));
    cudaError_t e2 = cudaMemcpy2DAsync(m->view_tab, 27 * 4, views_w + 256, 283 * 4, 27 * 4, 128, cudaMemcpyDeviceToDevice, st);
    if (e2 == cudaSuccess)
      e2 = cudaMemcpyAsync(m->view_tab + 128 * 27, views_b, 128 * 4, cudaMemcpyDeviceToDevice, st);
    if (e2 != cudaSuccess) return cleanup(fail(R2L_ERR_CUDA, "r2l_nerf_create: %s", cudaGetErrorString(e2)));
    m->pp_maps = new NerfPpMaps();
    rc = encode_rows_map(&m->pp_maps->m16, m->wstream_pp, elems_pp * 2, 32);
    if (rc == R2L_OK) rc = encode_rows_map(&m->pp_maps->m8, m->wstream_pp, elems_pp * 2, 16);
    if (rc == R2L_OK) rc = encode_rows_map(&m->pp_maps->m4, m->wstream_pp, elems_pp * 2, 8);
    if (rc != R2L_OK) return cleanup(rc);
  }
  if (cudaMalloc(reinterpret_cast<void**>(&m->far_wt), nerf_far_weight_bytes()) != cudaSuccess)
    return cleanup(fail(R2L_ERR_CUDA, "r2l_nerf_create: cudaMalloc failed"));
  rc = nerf_far_pack(pts_w, pts_b, alpha_w, m->far_wt, st);
  if (rc != R2L_OK) return cleanup(rc);
  cudaError_t e = cudaMemcpyAsync(m->aux + kNerfAuxAlphaW, alpha_w, 256 * 4, cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(m->aux + kNerfAuxRgbW, rgb_w, 384 * 4, cudaMemcpyDeviceToDevice, st);
  m->head_w = new NerfHeadW();
  if (e == cudaSuccess) e = cudaMemcpyAsync(m->head_w->alpha_w, alpha_w, 256 * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(m->head_w->rgb_w, rgb_w, 384 * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&m->alpha_b, alpha_b, 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(m->rgb_b, rgb_b, 12, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return cleanup(fail(R2L_ERR_CUDA, "r2l_nerf_create: %s", cudaGetErrorString(e)));
  rc = pack_max_check(m->bf16, st, "r2l_nerf_create");
  if (rc != R2L_OK) return cleanup(rc);
  *out_handle = m;
  return cleanup(R2L_OK);
}

static int check_dbg(Mlp* m, const char* who) {
  if (m->dbg_host && m->dbg_host->flag != 0)
    return fail(R2L_ERR_DEVICE_TRAP, "%s: kernel watchdog fired earlier (block %u thread %u barrier %u parity %u)", who,
                m->dbg_host->block, m->dbg_host->thread, m->dbg_host->barrier_id, m->dbg_host->parity);
  if (m->dbg_host && m->dbg_host->aux0 != 0) {
    // reported once, then cleared: the handle itself is healthy, the earlier RESULT was not
    const unsigned int row = m->dbg_host->aux1;
    m->dbg_host->aux0 = 0;
    return fail(R2L_ERR_RANGE, "%s: an earlier launch on this handle produced non-finite outputs (first seen at row "
                "%u): an activation left the range of %s operands%s", who, row, m->bf16 ? "bf16" : "fp16",
                m->bf16 ? "" : " (65504) — use precision 'bf16' or 'fp32' for this model");
  }
  return R2L_OK;
}

static int nerf_run_mlp(Mlp* m, NerfParams& p, cudaStream_t st);

// The MLP kernel, then (ray-sample calls with the fix-up enabled) the fp32 re-evaluation of the flagged far samples.
static int nerf_run(Mlp* m, NerfParams& p, cudaStream_t st) {
  const bool far = m->far_mode != 0 && p.embedded == nullptr && p.S > 0;
  if (far) {
    const long long n_rays = p.n_rows / p.S;
    R2L_CHECK_ARG(n_rays < (1LL << 31), "r2l_nerf_forward: too many rays for one call");
    if (n_rays > m->far_cap) {   // workspace grown on demand, like the view-bias workspace
      R2L_CUDA(cudaStreamSynchronize(st));
      if (m->far_ws) cudaFree(m->far_ws);
      m->far_ws = nullptr;
      m->far_cap = 0;
      R2L_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->far_ws), sizeof(int) * (2 + static_cast<size_t>(n_rays))));
      R2L_CUDA(cudaMemsetAsync(m->far_ws, 0, 2 * sizeof(int), st));
      m->far_cap = n_rays;
    }
    R2L_CUDA(cudaMemsetAsync(m->far_ws, 0, sizeof(int), st));
    p.far_list = m->far_ws + 2;
    p.far_count = m->far_ws;
    p.far_cap = static_cast<int>(m->far_cap);
    p.far_abs = m->far_abs;
    p.far_rel = m->far_rel;
  }
  int rc = nerf_run_mlp(m, p, st);
  if (rc != R2L_OK || !far) return rc;
  if (p.comp_ring == nullptr)
    return nerf_far_fixup_launch(m->far_wt, m->alpha_b, p.far_list, p.far_count, p.far_cap, m->far_ws + 1, p, p.raw, 0, st);
  // fused compositing: the flagged rays' rows sit in comp_far_raw (list order): patch sigma there, composite them again
  const int cap = p.far_cap < p.comp_far_cap ? p.far_cap : p.comp_far_cap;
  rc = nerf_far_fixup_launch(m->far_wt, m->alpha_b, p.far_list, p.far_count, cap, m->far_ws + 1, p,
                             reinterpret_cast<float*>(p.comp_far_raw), 1, st);
  if (rc != R2L_OK) return rc;
  return raw2outputs_list_launch(p.far_list, p.far_count, cap, p.S, p.comp_far_raw, p.z_vals, p.rays_d, p.d_stride,
                                 p.white_bkgd, p.o_rgb, p.o_disp, p.o_acc, p.o_weights, p.o_depth, st);
}

static int nerf_run_mlp(Mlp* m, NerfParams& p, cudaStream_t st) {
  p.wstream = m->wstream;
  p.alpha_w = m->aux + kNerfAuxAlphaW;
  p.rgb_w = m->aux + kNerfAuxRgbW;
  p.alpha_b = m->alpha_b;
  for (int i = 0; i < 3; ++i) p.rgb_b[i] = m->rgb_b[i];
  const long long n_tiles = (p.n_rows + kTileM - 1) / kTileM;
  R2L_CHECK_ARG(n_tiles < (1LL << 31), "r2l_nerf_forward: too many samples for one call");
  p.n_tiles = static_cast<int>(n_tiles);
  p.dbg = m->dbg_dev;
  if (m->nerf_pp) {
    R2L_CHECK_ARG(n_tiles < (1LL << 29), "r2l_nerf_forward: too many samples for one call");
    // per-ray view bias into the handle's workspace (grown on demand: the only allocation a forward can make),
    // then the ping-pong kernel: units of 4 tiles (2 per CTA of a pair)
    const bool emb = p.embedded != nullptr;
    const long long n_rays = emb ? p.n_rows : p.n_rows / p.S;
    if (n_rays > m->vb_cap) {
      R2L_CUDA(cudaStreamSynchronize(st));
      if (m->vb_ws) cudaFree(m->vb_ws);
      m->vb_ws = nullptr;
      m->vb_cap = 0;
      R2L_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->vb_ws), sizeof(float) * 128 * static_cast<size_t>(n_rays)));
      m->vb_cap = n_rays;
    }
    int rc = nerf_view_bias_launch(n_rays, emb ? p.embedded + 63 : p.viewdirs, emb ? p.emb_stride : p.v_stride,
                                   emb ? 1 : 0, m->view_tab, m->view_tab + 128 * 27, m->vb_ws, st);
    if (rc != R2L_OK) return rc;
    p.vb = m->vb_ws;
    p.wstream = m->wstream_pp;
    // CTA c owns tiles [c*T, (c+1)*T).  T is a multiple of `gran` tiles: 2 (the two tiles a CTA runs side by side),
    // and with fused compositing also of the tiles a compositing group spans, so every group stays inside one CTA.
    long long gran = 2;
    if (p.comp_ring != nullptr) {
      const long long grows = 32LL * p.comp_K;            // rows of a group: 256 or 192
      long long a = grows, b = kTileM;
      while (b) { const long long r = a % b; a = b; b = r; }
      const long long g_tiles = grows / a;                // tiles until group and tile boundaries coincide: 2 or 3
      gran = (g_tiles % 2 == 0) ? g_tiles : 2 * g_tiles;
    }
    const long long chunks = (n_tiles + gran - 1) / gran;
    const long long max_ctas = sm_count() & ~1;
    long long ctas = chunks < max_ctas ? chunks : max_ctas;
    ctas += ctas & 1;
    p.tiles_per_cta = static_cast<int>(((chunks + ctas - 1) / ctas) * gran);
    // CTAs (pairs) that would start past the last tile are not launched
    long long used = (n_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
    used += used & 1;
    return nerf_mlp_pp_launch(m->bf16, p, *m->pp_maps, *m->head_w, static_cast<int>(used), st);
  }
  const int grid = static_cast<int>(n_tiles < sm_count() ? n_tiles : sm_count());
  return nerf_mlp_launch(m->bf16, p, grid, st);
}

// Fused positional encoding + NeRF MLP on n_rays*S samples: raw[r, s, :] = NeRF(embed(o_r + d_r z_rs), embed(v_r)).
int r2l_nerf_forward(void* handle, long long n_rays, int S, const float* rays_o, long long o_stride,
                     const float* rays_d, long long d_stride, const float* viewdirs, long long v_stride,
                     const float* z_vals, float* raw, void* stream) {
  Mlp* m = static_cast<Mlp*>(handle);
  R2L_CHECK_ARG(m != nullptr && m->kind == 0, "r2l_nerf_forward: not a NeRF handle");
  R2L_CHECK_ARG(n_rays >= 0 && S > 0, "r2l_nerf_forward: bad sizes");
  if (n_rays == 0) return R2L_OK;
  R2L_CHECK_ARG(rays_o && rays_d && viewdirs && z_vals && raw, "r2l_nerf_forward: null pointer");
  R2L_CHECK_ARG((reinterpret_cast<uintptr_t>(raw) & 15) == 0, "r2l_nerf_forward: raw must be 16-byte aligned");
  int rc = check_dbg(m, "r2l_nerf_forward");
  if (rc != R2L_OK) return rc;
  NerfParams p{};
  p.rays_o = rays_o;
  p.rays_d = rays_d;
  p.viewdirs = viewdirs;
  p.o_stride = o_stride;
  p.d_stride = d_stride;
  p.v_stride = v_stride;
  p.z_vals = z_vals;
  p.S = S;
  p.n_rows = n_rays * S;
  p.n_rays = n_rays;
  p.raw = raw;
  return nerf_run(m, p, static_cast<cudaStream_t>(stream));
}

int r2l_raw2outputs(long long n_rays, int S, const float* raw, const float* z_vals, const float* rays_d,
                    long long d_stride, const float* noise, int white_bkgd, float* rgb_map, float* disp_map,
                    float* acc_map, float* weights, float* depth_map, void* stream);   // composite.cu

// render_rays' inner half in one call (main.py:707-709 / 738-741): raw = network(points), then raw2outputs(raw) ->
// rgb_map [n,3], disp_map, acc_map, depth_map [n] (nullable), weights [n,S] (nullable).  No noise (raw_noise_std = 0).
// Fused route (ping-pong kernel, S in {64,128,192,256}): the MLP kernel composites the rays itself, raw [n,S,4] never
// exists; otherwise MLP -> workspace -> r2l_raw2outputs.  Both routes give the same bits.
int r2l_nerf_render(void* handle, long long n_rays, int S, const float* rays_o, long long o_stride,
                    const float* rays_d, long long d_stride, const float* viewdirs, long long v_stride,
                    const float* z_vals, int white_bkgd, float* rgb_map, float* disp_map, float* acc_map,
                    float* weights, float* depth_map, void* stream) {
  Mlp* m = static_cast<Mlp*>(handle);
  R2L_CHECK_ARG(m != nullptr && m->kind == 0, "r2l_nerf_render: not a NeRF handle");
  R2L_CHECK_ARG(n_rays >= 0 && S >= 2 && d_stride >= 3, "r2l_nerf_render: bad sizes");
  if (n_rays == 0) return R2L_OK;
  R2L_CHECK_ARG(rays_o && rays_d && viewdirs && z_vals && rgb_map, "r2l_nerf_render: null pointer");
  R2L_CHECK_ARG(n_rays < (1LL << 31) / 4, "r2l_nerf_render: too many rays for one call");
  int rc = check_dbg(m, "r2l_nerf_render");
  if (rc != R2L_OK) return rc;
  auto st = static_cast<cudaStream_t>(stream);
  NerfParams p{};
  p.rays_o = rays_o;
  p.rays_d = rays_d;
  p.viewdirs = viewdirs;
  p.o_stride = o_stride;
  p.d_stride = d_stride;
  p.v_stride = v_stride;
  p.z_vals = z_vals;
  p.S = S;
  p.n_rows = n_rays * S;
  p.n_rays = n_rays;
  int K = 0, RPW = 0;
  const bool fused = m->comp_mode != 0 && m->nerf_pp && fused_composite_shape(S, &K, &RPW);
  if (!fused) {
    const size_t need = static_cast<size_t>(n_rays) * S * 16;
    if (need > m->raw_ws_bytes) {
      R2L_CUDA(cudaStreamSynchronize(st));
      if (m->raw_ws) cudaFree(m->raw_ws);
      m->raw_ws = nullptr;
      m->raw_ws_bytes = 0;
      R2L_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->raw_ws), need));
      m->raw_ws_bytes = need;
    }
    p.raw = m->raw_ws;
    rc = nerf_run(m, p, st);
    if (rc != R2L_OK) return rc;
    return r2l_raw2outputs(n_rays, S, m->raw_ws, z_vals, rays_d, d_stride, nullptr, white_bkgd, rgb_map, disp_map,
                           acc_map, weights, depth_map, stream);
  }
  if (m->comp_ring == nullptr) {
    const size_t rows = static_cast<size_t>(sm_count()) * 512;
    R2L_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->comp_ring), rows * sizeof(float4)));
    R2L_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->comp_aux), rows * sizeof(int)));
  }
  // compact copy of the flagged rays' rows: room for a quarter of the rays (random-init nets flag ~2 % in the coarse
  // pass, trained ones far fewer); rays beyond it keep the 16-bit sigma and show up in r2l_nerf_far_count
  long long far_rows = n_rays / 4;
  if (far_rows < 4096) far_rows = n_rays < 4096 ? n_rays : 4096;
  if (m->far_mode != 0) {
    const size_t need = static_cast<size_t>(far_rows) * S * sizeof(float4);
    if (need > m->comp_far_bytes) {
      R2L_CUDA(cudaStreamSynchronize(st));
      if (m->comp_far_raw) cudaFree(m->comp_far_raw);
      m->comp_far_raw = nullptr;
      m->comp_far_bytes = 0;
      R2L_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->comp_far_raw), need));
      m->comp_far_bytes = need;
    }
  }
  p.comp_ring = m->comp_ring;
  p.comp_aux = m->comp_aux;
  p.comp_far_raw = m->comp_far_raw;
  p.comp_far_cap = static_cast<int>(far_rows);
  p.comp_K = K;
  p.comp_RPW = RPW;
  p.white_bkgd = white_bkgd;
  p.o_rgb = rgb_map;
  p.o_disp = disp_map;
  p.o_acc = acc_map;
  p.o_depth = depth_map;
  p.o_weights = weights;
  return nerf_run(m, p, st);
}

// 0 (default): r2l_nerf_render takes the two-step route (MLP -> raw workspace -> raw2outputs); 1: fused where it applies
int r2l_nerf_render_mode(void* handle, int fused) {
  Mlp* m = static_cast<Mlp*>(handle);
  R2L_CHECK_ARG(m != nullptr && m->kind == 0, "r2l_nerf_render_mode: not a NeRF handle");
  R2L_CHECK_ARG(fused == 0 || fused == 1, "r2l_nerf_render_mode: mode must be 0 or 1");
  m->comp_mode = fused;
  return R2L_OK;
}

// NeRF.forward(x) API path: x [M, ldx] holds 63 embedded-point features followed by 27
// embedded-view features per row (model/nerf_raybased.py:377-401).  out [M, 4] = (rgb, sigma).
int r2l_nerf_forward_embedded(void* handle, long long M, const float* x, long long ldx, float* out, void* stream) {
  Mlp* m = static_cast<Mlp*>(handle);
  R2L_CHECK_ARG(m != nullptr && m->kind == 0, "r2l_nerf_forward_embedded: not a NeRF handle");
  R2L_CHECK_ARG(M >= 0 && ldx >= 90, "r2l_nerf_forward_embedded: bad sizes");
  if (M == 0) return R2L_OK;
  R2L_CHECK_ARG(x && out, "r2l_nerf_forward_embedded: null pointer");
  R2L_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0, "r2l_nerf_forward_embedded: out must be 16-byte aligned");
  int rc = check_dbg(m, "r2l_nerf_forward_embedded");
  if (rc != R2L_OK) return rc;
  // On this path every row is its own "ray" for the per-row view-branch bias workspace of the ping-pong kernel
  // (512 B per row), so large inputs are processed in blocks of 2^20 rows (a multiple of the 512-row unit).
  const long long block = m->nerf_pp ? (1LL << 20) : M;
  for (long long off = 0; off < M; off += block) {
    NerfParams p{};
    p.S = 1;
    p.n_rows = (M - off < block) ? (M - off) : block;
    p.raw = out + off * 4;
    p.embedded = x + off * ldx;
    p.emb_stride = ldx;
    rc = nerf_run(m, p, static_cast<cudaStream_t>(stream));
    if (rc != R2L_OK) return rc;
  }
  return R2L_OK;
}

// Far-sample sigma fix-up control (nerf_far.cu).  mode 0: off — raw holds the 16-bit kernels' own sigma everywhere;
// mode 1 (default): r2l_nerf_forward flags the rays whose last sample has |sigma| < max(abs_band, rel_band * sum |w_a|
// relu(h7)) and re-evaluates those points in fp32.  abs_band / rel_band < 0 keep the current values.
int r2l_nerf_far_fixup(void* handle, int mode, double abs_band, double rel_band) {
  Mlp* m = static_cast<Mlp*>(handle);
  R2L_CHECK_ARG(m != nullptr && m->kind == 0, "r2l_nerf_far_fixup: not a NeRF handle");
  R2L_CHECK_ARG(mode == 0 || mode == 1, "r2l_nerf_far_fixup: mode must be 0 or 1");
  m->far_mode = mode;
  if (abs_band >= 0.0) m->far_abs = static_cast<float>(abs_band);
  if (rel_band >= 0.0) m->far_rel = static_cast<float>(rel_band);
  return R2L_OK;
}

// Number of rays the last r2l_nerf_forward on `stream` flagged (synchronises the stream): diagnostics / tests.
int r2l_nerf_far_count(void* handle, long long* out_count, void* stream) {
  Mlp* m = static_cast<Mlp*>(handle);
  R2L_CHECK_ARG(m != nullptr && m->kind == 0 && out_count != nullptr, "r2l_nerf_far_count: bad arguments");
  *out_count = 0;
  if (m->far_ws == nullptr) return R2L_OK;
  int v = 0;
  R2L_CUDA(cudaMemcpyAsync(&v, m->far_ws + 1, sizeof(int), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
  R2L_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  *out_count = v;
  return R2L_OK;
}

// Profiling hook: r2l_nerf_forward + per-CTA cycle counters prof[n_CTAs][8] (device int64): [0] MMA thread total,
// [1] waiting for A groups, [2] waiting for weight stages, [3]/[5] WG0/WG1 waiting for accumulators,
// [4]/[6] WG0/WG1 epilogue work, [7] MMA thread waiting for the encoder.
int r2l_nerf_profile(void* handle, long long n_rays, int S, const float* rays_o, long long o_stride,
                     const float* rays_d, long long d_stride, const float* viewdirs, long long v_stride,
                     const float* z_vals, float* raw, long long* prof, void* stream) {
  Mlp* m = static_cast<Mlp*>(handle);
  R2L_CHECK_ARG(m != nullptr && m->kind == 0, "r2l_nerf_profile: not a NeRF handle");
  R2L_CHECK_ARG(n_rays > 0 && S > 0 && rays_o && rays_d && viewdirs && z_vals && raw && prof,
                "r2l_nerf_profile: bad arguments");
  R2L_CHECK_ARG(S >= 64 && S % 32 == 0, "r2l_nerf_profile: the profiling instantiation covers S >= 64, S %% 32 == 0");
  NerfParams p{};
  p.rays_o = rays_o;
  p.rays_d = rays_d;
  p.viewdirs = viewdirs;
  p.o_stride = o_stride;
  p.d_stride = d_stride;
  p.v_stride = v_stride;
  p.z_vals = z_vals;
  p.S = S;
  p.n_rows = n_rays * S;
  p.n_rays = n_rays;
  p.raw = raw;
  p.prof = prof;
  { const char* e = getenv("R2L_PROF_MODE"); p.prof_mode = (e != nullptr) ? atoi(e) : 0; }
  return nerf_run(m, p, static_cast<cudaStream_t>(stream));
}

// ------------------------------- R2L ----------------------------------------------
// head_w [256, n_points*63] (reference PositionalEmbedder feature order, model:198-208),
// w1[b]/w2[b] [256,256] = body.{b}.body.{0,2}.weight (model:443-465), tail_w [3,256].
int r2l_resmlp_create(void** out_handle, int dtype, int n_points, int n_blocks, const float* head_w,
                      const float* head_b, const float* const* w1, const float* const* b1, const float* const* w2,
                      const float* const* b2, double res_scale, const float* tail_w, const float* tail_b,
                      int sigmoid_out, int outer_skip, void* stream) {
  R2L_CHECK_ARG(out_handle != nullptr, "r2l_resmlp_create: null out_handle");
  R2L_CHECK_ARG(dtype == 0 || dtype == 1, "r2l_resmlp_create: dtype must be 0 (fp16) or 1 (bf16)");
  R2L_CHECK_ARG(n_points >= 4 && n_points % 4 == 0 && n_points <= 256,
                "r2l_resmlp_create: n_points must be a multiple of 4 in [4,256] (got %d)", n_points);
  R2L_CHECK_ARG(n_blocks >= 1 && n_blocks <= 1024, "r2l_resmlp_create: bad n_blocks");
  R2L_CHECK_ARG(head_w && head_b && w1 && b1 && w2 && b2 && tail_w && tail_b, "r2l_resmlp_create: null pointer");
  auto st = static_cast<cudaStream_t>(stream);
  Mlp* m = new Mlp();
  m->kind = 1;
  m->bf16 = dtype == 1;
  m->pair = pair_mode_default();
  m->r2l_pp = (m->pair == 1) ? r2l_pp_default() : 0;
  m->n_points = n_points;
  m->n_blocks = n_blocks;
  m->sigmoid_out = sigmoid_out;
  m->outer_skip = outer_skip;
  std::vector<int*> scratch;
  auto cleanup = [&](int rc) {
    cudaStreamSynchronize(st);
    for (int* s : scratch) cudaFree(s);
    if (rc != R2L_OK) destroy(m);
    return rc;
  };
  const int K_head = n_points * 64;
  // stage stream in consumption order; every layer is [K=16 bias stage][K/32 weight stages]
  const size_t elems = 256ull * 16 + 256ull * K_head + static_cast<size_t>(n_blocks) * 2 * (256 * 16 + 256 * 256);
  m->wbytes = elems * 2;
  const size_t aux_floats = 768;
  if (cudaMalloc(reinterpret_cast<void**>(&m->wstream), m->wbytes) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&m->aux), sizeof(float) * aux_floats) != cudaSuccess)
    return cleanup(fail(R2L_ERR_CUDA, "r2l_resmlp_create: cudaMalloc failed"));
  int rc = alloc_debug(m);
  if (rc == R2L_OK) rc = pack_max_reset(st);
  if (rc != R2L_OK) return cleanup(rc);
  // head K order: point block s, in-block index i (see encode_point_block) <- reference column (3s+c)*21 + f'
  std::vector<int> kmap(K_head);
  for (int s = 0; s < n_points; ++s) {
    for (int i = 0; i < 64; ++i) {
      int ref = -1;
      if (i < 3) {
        ref = (3 * s + i) * 21 + 20;
      } else if (i < 63) {
        const int f = (i - 3) / 6, rem = (i - 3) % 6;
        ref = (rem < 3) ? ((3 * s + rem) * 21 + f) : ((3 * s + rem - 3) * 21 + 10 + f);
      }
      kmap[s * 64 + i] = ref;
    }
  }
  uint16_t* dst = reinterpret_cast<uint16_t*>(m->wstream);
  size_t off = 0;
  const int pr = m->pair;
  rc = pack_bias(head_b, 256, 1.0f, dst + off, m->bf16, st, pr);
  off += 256ull * 16;
  if (rc == R2L_OK)
    rc = pack_layer(head_w, static_cast<long long>(n_points) * 63, 256, n_points * 63, K_head, &kmap, 1.0f, dst + off,
                    m->bf16, st, scratch, pr);
  off += 256ull * K_head;
  for (int b = 0; b < n_blocks && rc == R2L_OK; ++b) {
    if (!(w1[b] && b1[b] && w2[b] && b2[b])) rc = fail(R2L_ERR_INVALID, "r2l_resmlp_create: null block %d", b);
  }
  if (rc == R2L_OK) {
    // the 2 * n_blocks body layers in one launch (was four launches per block)
    const float** d_ptrs = nullptr;
    if (cudaMalloc(reinterpret_cast<void**>(&d_ptrs), sizeof(float*) * 4 * n_blocks) != cudaSuccess)
      return cleanup(fail(R2L_ERR_CUDA, "r2l_resmlp_create: cudaMalloc failed"));
    scratch.push_back(reinterpret_cast<int*>(d_ptrs));
    const float* const* groups[4] = {w1, b1, w2, b2};
    cudaError_t e1 = cudaSuccess;
    for (int g = 0; g < 4 && e1 == cudaSuccess; ++g)
      e1 = cudaMemcpyAsync(d_ptrs + g * n_blocks, groups[g], sizeof(float*) * n_blocks, cudaMemcpyHostToDevice, st);
    if (e1 != cudaSuccess) return cleanup(fail(R2L_ERR_CUDA, "r2l_resmlp_create: %s", cudaGetErrorString(e1)));
    pack_r2l_body_kernel<<<dim3(64, 2 * n_blocks), 256, 0, st>>>(d_ptrs, n_blocks, static_cast<float>(res_scale), dst + off,
                                                                 m->bf16 ? 1 : 0, pr);
    count_launch();
    if (cudaGetLastError() != cudaSuccess) rc = fail(R2L_ERR_CUDA, "r2l_resmlp_create: body pack launch failed");
    off += static_cast<size_t>(2 * n_blocks) * (256ull * 16 + 256ull * 256);
  }
  if (rc != R2L_OK) return cleanup(rc);
  if (m->pair == 1) {
    m->r2l_maps = new R2lPairMaps();
    rc = encode_rows_map(&m->r2l_maps->m16, m->wstream, m->wbytes, 32);
    if (rc == R2L_OK) rc = encode_rows_map(&m->r2l_maps->m4, m->wstream, m->wbytes, 8);
    if (rc != R2L_OK) return cleanup(rc);
  }
  cudaError_t e = cudaMemcpyAsync(m->aux, tail_w, 768 * 4, cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(m->b_tail, tail_b, 12, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return cleanup(fail(R2L_ERR_CUDA, "r2l_resmlp_create: %s", cudaGetErrorString(e)));
  rc = pack_max_check(m->bf16, st, "r2l_resmlp_create");
  if (rc != R2L_OK) return cleanup(rc);
  *out_handle = m;
  return cleanup(R2L_OK);
}

static int resmlp_run(Mlp* m, long long n_rays, const float* pts, long long pts_stride, const float* embedded,
                      long long emb_stride, float* rgb, cudaStream_t st, float* dbg_acc = nullptr,
                      float* dbg_x0 = nullptr, void* dbg_a = nullptr, long long* prof = nullptr,
                      float* const* peers = nullptr, int n_peer = 0, long long peer_row0 = 0,
                      const R2lParams* cam = nullptr) {
  R2lParams p{};
  if (cam != nullptr) {
    p.cam = cam->cam, p.cam_z = cam->cam_z, p.cam_H = cam->cam_H, p.cam_W = cam->cam_W;
    p.cam_focal = cam->cam_focal, p.cam_ray0 = cam->cam_ray0;
  }
  for (int g = 0; g < n_peer; ++g) p.rgb_peer[g] = peers[g];
  p.n_peer = n_peer;
  p.peer_row0 = peer_row0;
  p.wstream = m->wstream;
  p.w_tail = m->aux;
  for (int i = 0; i < 3; ++i) p.b_tail[i] = m->b_tail[i];
  p.n_blocks = m->n_blocks;
  p.n_points = m->n_points;
  p.pts = pts;
  p.pts_stride = pts_stride;
  p.n_rays = n_rays;
  p.rgb = rgb;
  p.sigmoid_out = m->sigmoid_out;
  p.outer_skip = m->outer_skip;
  const long long n_tiles = (n_rays + kTileM - 1) / kTileM;
  R2L_CHECK_ARG(n_tiles < (1LL << 31), "r2l_resmlp_forward: too many rays for one call");
  p.n_tiles = static_cast<int>(n_tiles);
  p.dbg = m->dbg_dev;
  p.embedded = embedded;
  p.emb_stride = emb_stride;
  p.dbg_head_acc = dbg_acc;
  p.dbg_head_x0 = dbg_x0;
  (void)dbg_a;
  p.prof = prof;
  int grid;
  if (m->r2l_pp && dbg_acc == nullptr) {
    const long long n_units = (n_rays + 255) / 256;
    const long long max_pairs = sm_count() / 2;
    grid = static_cast<int>(2 * (n_units < max_pairs ? n_units : max_pairs));
    return r2l_mlp_pp_launch(m->bf16, p, *m->r2l_maps, grid, st);
  }
  if (m->pair) {
    const long long n_units = (n_tiles + 1) / 2;
    const long long max_pairs = sm_count() / 2;
    grid = static_cast<int>(2 * (n_units < max_pairs ? n_units : max_pairs));
  } else {
    grid = static_cast<int>(n_tiles < sm_count() ? n_tiles : sm_count());
  }
  return r2l_mlp_launch(m->bf16, m->pair == 1, p, m->r2l_maps, grid, st);
}

// Fused PositionalEmbedder + NeRF_v3_2 on sampled points: pts [n_rays, n_points*3] -> rgb [n_rays, 3].
int r2l_resmlp_forward(void* handle, long long n_rays, const float* pts, long long pts_stride, float* rgb,
                       void* stream) {
  Mlp* m = static_cast<Mlp*>(handle);
  R2L_CHECK_ARG(m != nullptr && m->kind == 1, "r2l_resmlp_forward: not an R2L handle");
  R2L_CHECK_ARG(n_rays >= 0 && pts_stride >= 3LL * m->n_points, "r2l_resmlp_forward: bad sizes");
  if (n_rays == 0) return R2L_OK;
  R2L_CHECK_ARG(pts && rgb, "r2l_resmlp_forward: null pointer");
  int rc = check_dbg(m, "r2l_resmlp_forward");
  if (rc != R2L_OK) return rc;
  return resmlp_run(m, n_rays, pts, pts_stride, nullptr, 0, rgb, static_cast<cudaStream_t>(stream));
}

// PointSampler.sample_test + PositionalEmbedder + NeRF_v3_2 in ONE kernel (main.py:297-309): the head generates its
// rays from the pixel index (c2w [n_poses][3][4] device, z_vals [n_sample] device = PointSampler.z_vals) — no pts
// tensor.  Renders rays [ray0, ray0 + n_rays) of the pose-major range [n_poses][H*W] into rgb [n_rays][3], or, when
// peer_frames != NULL, into every peer's frame buffer at rows row0.. (see r2l_resmlp_forward_gather).
int r2l_resmlp_render(void* handle, int n_poses, int H, int W, double focal, const float* c2w, const float* z_vals,
                      int n_sample, long long ray0, long long n_rays, float* rgb, float* const* peer_frames,
                      int n_peers, long long row0, void* stream) {
  Mlp* m = static_cast<Mlp*>(handle);
  R2L_CHECK_ARG(m != nullptr && m->kind == 1, "r2l_resmlp_render: not an R2L handle");
  R2L_CHECK_ARG(n_poses >= 1 && H > 0 && W > 0 && focal != 0.0, "r2l_resmlp_render: bad camera");
  R2L_CHECK_ARG(n_sample == m->n_points, "r2l_resmlp_render: the sampler draws %d points per ray, the model takes %d",
                n_sample, m->n_points);
  R2L_CHECK_ARG(ray0 >= 0 && n_rays >= 0 && ray0 + n_rays <= static_cast<long long>(n_poses) * H * W,
                "r2l_resmlp_render: rays [%lld, %lld) outside %d x %d x %d", ray0, ray0 + n_rays, n_poses, H, W);
  if (n_rays == 0) return R2L_OK;
  R2L_CHECK_ARG(c2w && z_vals, "r2l_resmlp_render: null pointer");
  if (peer_frames != nullptr) {
    R2L_CHECK_ARG(n_peers >= 1 && n_peers <= kMaxPeers && row0 >= 0, "r2l_resmlp_render: 1..%d peers", kMaxPeers);
    for (int g = 0; g < n_peers; ++g)
      R2L_CHECK_ARG(peer_frames[g] != nullptr, "r2l_resmlp_render: peer %d has no frame buffer", g);
  } else {
    R2L_CHECK_ARG(rgb != nullptr, "r2l_resmlp_render: null pointer");
    n_peers = 0;
  }
  int rc = check_dbg(m, "r2l_resmlp_render");
  if (rc != R2L_OK) return rc;
  R2lParams cam{};
  cam.cam = c2w, cam.cam_z = z_vals, cam.cam_H = H, cam.cam_W = W;
  cam.cam_focal = static_cast<float>(focal), cam.cam_ray0 = ray0;
  return resmlp_run(m, n_rays, nullptr, 0, nullptr, 0, rgb, static_cast<cudaStream_t>(stream), nullptr, nullptr, nullptr,
                    nullptr, peer_frames, n_peers, row0, &cam);
}

// r2l_resmlp_forward whose tail stores this rank's rows [row0, row0 + n_rays) into EVERY peer's frame buffer
// (peer_frames: HOST array of n_peers device pointers to [>= row0 + n_rays][3] fp32 buffers mapped for peer access,
// this GPU's own included): the compute step and the all-gather of the ray-sharded frame in one kernel.
int r2l_resmlp_forward_gather(void* handle, long long n_rays, const float* pts, long long pts_stride,
                              float* const* peer_frames, int n_peers, long long row0, void* stream) {
  Mlp* m = static_cast<Mlp*>(handle);
  R2L_CHECK_ARG(m != nullptr && m->kind == 1, "r2l_resmlp_forward_gather: not an R2L handle");
  R2L_CHECK_ARG(n_rays >= 0 && pts_stride >= 3LL * m->n_points && row0 >= 0, "r2l_resmlp_forward_gather: bad sizes");
  R2L_CHECK_ARG(n_peers >= 1 && n_peers <= kMaxPeers, "r2l_resmlp_forward_gather: 1..%d peers (got %d)", kMaxPeers,
                n_peers);
  R2L_CHECK_ARG(peer_frames != nullptr, "r2l_resmlp_forward_gather: null pointer");
  for (int g = 0; g < n_peers; ++g)
    R2L_CHECK_ARG(peer_frames[g] != nullptr, "r2l_resmlp_forward_gather: peer %d has no frame buffer", g);
  if (n_rays == 0) return R2L_OK;
  R2L_CHECK_ARG(pts, "r2l_resmlp_forward_gather: null pointer");
  int rc = check_dbg(m, "r2l_resmlp_forward_gather");
  if (rc != R2L_OK) return rc;
  return resmlp_run(m, n_rays, pts, pts_stride, nullptr, 0, nullptr, static_cast<cudaStream_t>(stream), nullptr, nullptr,
                    nullptr, nullptr, peer_frames, n_peers, row0);
}

// NeRF_v3_2.forward(x) API path: x [n_rays, ldx] is the reference-layout embedding (n_points*63 features).
int r2l_resmlp_forward_embedded(void* handle, long long n_rays, const float* x, long long ldx, float* rgb,
                                void* stream) {
  Mlp* m = static_cast<Mlp*>(handle);
  R2L_CHECK_ARG(m != nullptr && m->kind == 1, "r2l_resmlp_forward_embedded: not an R2L handle");
  R2L_CHECK_ARG(n_rays >= 0 && ldx >= 63LL * m->n_points, "r2l_resmlp_forward_embedded: bad sizes");
  if (n_rays == 0) return R2L_OK;
  R2L_CHECK_ARG(x && rgb, "r2l_resmlp_forward_embedded: null pointer");
  int rc = check_dbg(m, "r2l_resmlp_forward_embedded");
  if (rc != R2L_OK) return rc;
  return resmlp_run(m, n_rays, nullptr, 0, x, ldx, rgb, static_cast<cudaStream_t>(stream));
}

// Debug hook (tests only): like r2l_resmlp_forward, additionally dumping the head layer's raw accumulators and
// x0 = relu(acc + b_head) as [ceil(n_rays/128)*128, 256] fp32.
int r2l_resmlp_debug_head(void* handle, long long n_rays, const float* pts, long long pts_stride, float* rgb,
                          float* head_acc, float* head_x0, void* head_a, void* stream) {
  Mlp* m = static_cast<Mlp*>(handle);
  R2L_CHECK_ARG(m != nullptr && m->kind == 1, "r2l_resmlp_debug_head: not an R2L handle");
  R2L_CHECK_ARG(n_rays > 0 && pts && rgb && head_acc && head_x0, "r2l_resmlp_debug_head: bad arguments");
  return resmlp_run(m, n_rays, pts, pts_stride, nullptr, 0, rgb, static_cast<cudaStream_t>(stream), head_acc, head_x0, head_a);
}

// Profiling hook: r2l_resmlp_forward that also fills per-CTA cycle counters prof[n_CTAs][8] (device int64):
// [0] MMA thread total, [1] MMA waiting for A operands, [2] MMA waiting for weight stages,
// [3]/[5] WG0/WG1 waiting for accumulators, [4]/[6] WG0/WG1 epilogue work, [7] WG0 encode time.
int r2l_resmlp_profile(void* handle, long long n_rays, const float* pts, long long pts_stride, float* rgb,
                       long long* prof, void* stream) {
  Mlp* m = static_cast<Mlp*>(handle);
  R2L_CHECK_ARG(m != nullptr && m->kind == 1, "r2l_resmlp_profile: not an R2L handle");
  R2L_CHECK_ARG(n_rays > 0 && pts && rgb && prof, "r2l_resmlp_profile: bad arguments");
  return resmlp_run(m, n_rays, pts, pts_stride, nullptr, 0, rgb, static_cast<cudaStream_t>(stream), nullptr, nullptr,
                    nullptr, prof);
}

// Debug hook (tests only): copy the packed 16-bit weight stream to host memory; returns its size via *bytes.
int r2l_mlp_debug_wstream(void* handle, void* out_host, unsigned long long capacity, unsigned long long* bytes) {
  Mlp* m = static_cast<Mlp*>(handle);
  R2L_CHECK_ARG(m != nullptr && bytes != nullptr, "r2l_mlp_debug_wstream: bad arguments");
  *bytes = m->wbytes;
  if (out_host != nullptr) {
    R2L_CHECK_ARG(capacity >= m->wbytes, "r2l_mlp_debug_wstream: buffer too small");
    R2L_CUDA(cudaMemcpy(out_host, m->wstream, m->wbytes, cudaMemcpyDeviceToHost));
  }
  return R2L_OK;
}

int r2l_mlp_destroy(void* handle) {
  destroy(static_cast<Mlp*>(handle));
  return R2L_OK;
}

// 0 = healthy; R2L_ERR_DEVICE_TRAP: out8 = the watchdog record of the first barrier that timed out; R2L_ERR_RANGE: a
// finished launch produced non-finite outputs (out8[5] = 1, out8[6] = row): an activation left the 16-bit operand range.
// Reads the handle's mapped record without synchronising: call it after the stream has been synchronised.
int r2l_mlp_status(void* handle, unsigned int* out8) {
  Mlp* m = static_cast<Mlp*>(handle);
  R2L_CHECK_ARG(m != nullptr, "r2l_mlp_status: null handle");
  if (out8 != nullptr && m->dbg_host != nullptr) memcpy(out8, m->dbg_host, sizeof(DebugBuf));
  if (m->dbg_host && m->dbg_host->flag) return R2L_ERR_DEVICE_TRAP;
  return (m->dbg_host && m->dbg_host->aux0) ? R2L_ERR_RANGE : R2L_OK;
}

}  // extern "C"
