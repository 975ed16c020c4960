// NeRF MLP (W256 x D8 + heads), "ping-pong" schedule on a CTA pair.  Same math as mlp_nerf.cu
// (model/nerf_raybased.py:377-401, main.py:65-87); what changes is how the epilogue latency is hidden.
//
// mlp_nerf.cu runs ONE 128-sample tile per CTA and lets layer l+1's MMAs chase layer l's epilogue: the tensor pipe
// still idles ~900 cycles per layer (accumulator commit -> warps wake -> tcgen05.ld -> cvt/st -> fences -> arrive ->
// issuer wakes), 30 % of the time.  Here every CTA owns TWO tiles T0, T1, each with its own activation buffer and
// ONE 256-column TMEM accumulator, and the tensor pipe alternates:  T0.L0, T1.L0, T0.L1, T1.L1, ...  While tile
// T1's layer executes (2176 tensor cycles), tile T0's epilogue converts its accumulator in place into T0's
// activations, so the next T0 layer finds its operands ready and the pipe never waits for an epilogue.
//
// What makes it fit:
//   * CTA pair (tcgen05.mma.cta_group::2, M = 256 = one tile of each CTA): every CTA holds only its N-half of a
//     weight stage (16 KiB), so two activation buffers (128 KiB) + one point block (16 KiB) + a 4-slot weight ring
//     (64 KiB) fit in 227 KiB.
//   * T0 and T1 SHARE every weight stage: a stage is loaded once, used by T0's layer and ~2176 cycles later by T1's,
//     and released when both issuers have committed it.  The four slots hold one whole 256x256 layer, so a slot has
//     ~1660 cycles to be refilled (commit -> producer -> L2 -> shared -> peer relay: measured ~2000 cycles when a
//     3-slot ring fed each tile separately, which left the tensor pipe waiting for weights 25-35 % of the time) and
//     the L2 -> shared-memory traffic is 16 B/cycle/SM, a quarter of the single-CTA kernel's.
//   * TWO issuer threads, one per tile, running concurrently: a thread needs ~56 cycles per tcgen05.mma and ~100 per
//     tcgen05.commit / mbarrier wait (scratch/ubench/mmaissue.cu), ~2600 cycles per tile-layer (measured) — more than
//     the 2176 tensor cycles, so one thread cannot feed the pipe.  Both walk the SAME stage sequence, T1 behind T0.
//   * The peer CTA's half stages arrive by cp.async.bulk.tensor...cta_group::2, which may signal the LEADER's
//     mbarrier: the issuers see both halves of a stage on one barrier, no relay thread.
//   * ONE point block P: the encoder warps re-encode a tile's points right before each of its two uses (steps 0 and
//     5) — sin/cos of 128 points is ~2 k issue slots, the encoder warps are otherwise idle.
//   * The view branch is a per-ray fp32 bias vb (computed by nerf_view_bias_kernel into a caller workspace) added in
//     the last epilogue instead of an extra K-stage: it needs no shared memory.
//   * Tried and measured slower on the same box (160000 x 192 samples: 29.7-30.1 ms for this version): a 5-slot ring
//     paid for by writing the point block over columns [0,64) of the tile's own activation buffer (skip layer with
//     its point stage last): 30.3 ms; the same with a one-directional lag token for issuer T1: 30.6-30.7 ms; with
//     symmetric turn tokens (strict alternation): 32.4 ms — one issuer at a time issues a stage every ~590 cycles,
//     slower than the pipe's 512, so two issuers working concurrently (and the tiles in near lockstep) win.  In the
//     schedule the issuers settle into, the ring is not what limits a tile: its chain issue -> drain behind the other
//     tile's queued MMAs -> epilogue (which waits for the other tile's epilogue: same 8 warps) -> wake-ups is.
//   * No per-group chasing barriers: a tile-layer starts when the tile's previous epilogue has signalled a_done[t]
//     (one barrier, 16 warp arrivals), which also closes every TMEM / shared-memory write-after-read hazard.
// Per tile: 10 steps (see mlp_nerf.cu): 0 = W0 P, 1-4, 5 = W5 [P, h], 6, 7 (+sigma), 8 = feature, 9 = views (N 128).
#include "common.cuh"
#include "mlp_params.cuh"
#include "composite_math.cuh"
#include "mlp_tc.cuh"
#include <type_traits>

namespace r2l {

constexpr int kPpThreads = 480;
constexpr int kPpProducerWarp = 12;
constexpr int kPpMmaWarp0 = 13;   // leader: issuer of tile T0; peer: relays weight arrivals to the leader
constexpr int kPpMmaWarp1 = 14;   // leader: issuer of tile T1
constexpr int kPpRing = 4;
constexpr int kPpBiasRing = 2;
// The peer CTA's half stages are copied with cp.async.bulk.tensor...cta_group::2, whose completion may be signalled
// on the LEADER's mbarrier (a plain cp.async.bulk with a remote barrier is a launch failure: tried), so the leader's
// issuers see both halves on one barrier with no relay hop (~450 cycles off the refill latency).
constexpr bool kPpDirect = true;
#ifndef R2L_PP_TURNS
#define R2L_PP_TURNS 0
#endif
// 1: the two issuers take turns, one tile-layer each (strict T0, T1, T0, ... alternation on the tensor pipe).  Measured
// SLOWER (32.0 vs 29.8 ms per 160000x192 pass): with strict alternation a slot is refilled only ~2900 cycles after
// the leading tile started its layer, and the refill latency (~2400 cycles) then paces the whole loop; left free, the
// issuers settle ~1200 cycles apart, weights are prefetched a layer ahead and the tensor pipe idles only between
// layers (~20 %, in-kernel timeline: scratch/prof_pp.py, R2L_PROF_MODE=5).
constexpr bool kPpTurns = R2L_PP_TURNS != 0;
// 2: ONE-directional: only issuer T1 waits (until T0 has issued its tile-layer), T0 never waits for T1 — a minimum
// lag for T1 without making the tiles mutually exclusive.
constexpr bool kPpLagOnly = R2L_PP_TURNS == 2;
constexpr uint32_t kPpStageB = kStageBytes / 2;       // 16 KiB: this CTA's N-half of a K=64 stage
constexpr uint32_t kPpBiasB = kBiasStageBytes / 2;    // 4 KiB
constexpr uint32_t kPpLbo256 = 128 * 16;              // 128 B-rows per CTA (N = 256)
constexpr uint32_t kPpLbo128 = 64 * 16;               // 64 B-rows per CTA (N = 128)
// shared memory map
constexpr int kPpOffA = 0;                                            // A[tile]
constexpr int kPpOffP = kPpOffA + 2 * kABufBytes;                     // P (one block, re-encoded per use)
constexpr int kPpOffOnes = kPpOffP + kPBlockBytes;
constexpr int kPpOffRing = kPpOffOnes + kOnesBytes;
constexpr int kPpOffBiasRing = kPpOffRing + kPpRing * kPpStageB;
constexpr int kPpOffAlphaW = kPpOffBiasRing + kPpBiasRing * kPpBiasB;   // 256 floats
constexpr int kPpOffRgbW = kPpOffAlphaW + 256 * 4;                    // 3*128 floats
constexpr int kPpOffPart = kPpOffRgbW + 384 * 4;                      // 128 x float4
constexpr int kPpOffAbs = kPpOffPart + 128 * 16;                      // 128 floats: WG1's half of sum |w_a| relu(h7)
constexpr int kPpOffBars = kPpOffAbs + 128 * 4;
constexpr int kPpNumBars = 2 * kPpRing + 2 * kPpBiasRing + 2 + 2 + 2 + 2 + 2 + 1 + 2;
static_assert(2 * 2 * 128 * 4 <= (256 + 384) * 4, "view-bias rows must fit the old head-weight region");
constexpr int kPpOffTmem = kPpOffBars + kPpNumBars * 8;
constexpr int kPpSmemBytes = kPpOffTmem + 16;
static_assert(kPpSmemBytes <= 227 * 1024, "NeRF ping-pong kernel shared memory exceeds 227 KiB");
static_assert(kPpOffRing % 1024 == 0, "weight ring must stay 1 KiB aligned");

constexpr int kCompRingRows = 512;   // rows of a CTA's staging ring (two units)

// Fused compositing (NerfParams::comp_ring): one warp composites group `gidx` of this CTA — comp_RPW consecutive rays =
// 32 * comp_K staged rows — with the arithmetic, the sample partition and the scan order of raw2outputs_blocked_kernel
// (composite.cu; shared definitions in composite_math.cuh), so a fused frame equals MLP + r2l_raw2outputs bit for
// bit.  Compact on purpose (two passes over the lane's samples instead of register arrays, run-time K): it is called
// once per unit and the ping-pong kernel is sensitive to code size.
static __device__ __noinline__ void pp_composite_group(const NerfParams& p, const float4* ring, const int* aux,
                                                       unsigned cta_unit0, int gidx, int lane) {
  const int K = p.comp_K, LPR = 32 / p.comp_RPW, S = p.S;
  const int sl = lane & (LPR - 1), rl = lane / LPR;
  const unsigned n_rays = static_cast<unsigned>(p.n_rays);
  const unsigned grow0 = static_cast<unsigned>(gidx) * 32u * static_cast<unsigned>(K);   // CTA-local row of the group
  // first ray of the group, on 32-row units (S is a multiple of 32; tile counts are < 2^29: host)
  const unsigned ray = (cta_unit0 + static_cast<unsigned>(gidx) * K) / (static_cast<unsigned>(S) >> 5) + rl;
  const bool ray_ok = ray < n_rays;
  const long long rr = ray_ok ? ray : n_rays - 1;
  const float* dp = p.rays_d + rr * p.d_stride;
  const float dnorm = comp_dnorm(__ldg(dp), __ldg(dp + 1), __ldg(dp + 2));
  const int base = sl * K;
  const unsigned r0 = grow0 + static_cast<unsigned>(lane * K);
  const float* zrow = p.z_vals + rr * S + base;
  // Every load of the lane's samples is issued up front (one L2 round trip, not one per sample) and the loops are
  // unrolled for K = 8 with `j < K` predicates, so everything stays in registers: with run-time loops the arrays lived
  // in local memory, and with the 227 KB shared-memory carve-out the L1 is too small to hold them — the compositor
  // then took long enough to delay the point blocks these warps owe the issuers (2 ms per frame, A/B on one box).
  float4 rv[8];
  float zr[9];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    rv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    zr[j] = 0.0f;
    if (j < K) {
      rv[j] = __ldcg(&ring[(r0 + j) & (kCompRingRows - 1)]);
      zr[j] = __ldg(zrow + j);
    }
  }
  const float z_after = (base + K < S) ? __ldg(zrow + K) : 0.0f;   // first sample of the next lane
  zr[8] = z_after;
  // far slot of this lane's ray (written by the thread that owned the ray's last sample)
  int fs = -1;
  if (p.far_list != nullptr) {
    if (sl == LPR - 1) fs = __ldcg(&aux[(grow0 + static_cast<unsigned>(rl * S + S - 1)) & (kCompRingRows - 1)]);
    fs = __shfl_sync(0xffffffffu, fs, rl * LPR + LPR - 1);
    if (!ray_ok || fs >= p.comp_far_cap) fs = -1;
  }
  float alpha[8];
  double excl_in[8];
  double run = 1.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float zn = (j + 1 == K) ? z_after : zr[j + 1];
    const float a = comp_alpha(rv[j].w, comp_dist(zr[j], zn, base + j == S - 1, dnorm));
    alpha[j] = a;
    excl_in[j] = run;
    if (ray_ok && j < K) run *= static_cast<double>(__fadd_rn(__fsub_rn(1.0f, a), 1e-10f));
  }
  double pr = run;   // inclusive scan of the lane totals, segmented by ray
  for (int o = 1; o < LPR; o <<= 1) {
    const double q = shfl_up_f64(pr, o);
    if (sl >= o) pr *= q;
  }
  double excl = shfl_up_f64(pr, 1);
  if (sl == 0) excl = 1.0;
  float ar = 0.f, ag = 0.f, ab = 0.f, adepth = 0.f, aacc = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float T = static_cast<float>(excl * excl_in[j]);
    const float wj = __fmul_rn(alpha[j], T);
    if (ray_ok && j < K) {
      const float4 v = rv[j];
      const float sr = fast_sigmoid(v.x), sg = fast_sigmoid(v.y), sb = fast_sigmoid(v.z);
      ar = __fadd_rn(ar, __fmul_rn(wj, sr));
      ag = __fadd_rn(ag, __fmul_rn(wj, sg));
      ab = __fadd_rn(ab, __fmul_rn(wj, sb));
      adepth = __fadd_rn(adepth, __fmul_rn(wj, zr[j]));
      aacc = __fadd_rn(aacc, wj);
      if (p.o_weights != nullptr) p.o_weights[rr * S + base + j] = wj;
      if (fs >= 0) p.comp_far_raw[static_cast<long long>(fs) * S + base + j] = v;
    }
  }
  for (int o = LPR / 2; o > 0; o >>= 1) {
    ar += __shfl_xor_sync(0xffffffffu, ar, o);
    ag += __shfl_xor_sync(0xffffffffu, ag, o);
    ab += __shfl_xor_sync(0xffffffffu, ab, o);
    adepth += __shfl_xor_sync(0xffffffffu, adepth, o);
    aacc += __shfl_xor_sync(0xffffffffu, aacc, o);
  }
  if (sl == 0 && ray_ok) {
    if (p.white_bkgd) {
      const float bg = __fsub_rn(1.0f, aacc);
      ar = __fadd_rn(ar, bg);
      ag = __fadd_rn(ag, bg);
      ab = __fadd_rn(ab, bg);
    }
    p.o_rgb[3 * rr] = ar;
    p.o_rgb[3 * rr + 1] = ag;
    p.o_rgb[3 * rr + 2] = ab;
    if (p.o_depth != nullptr) p.o_depth[rr] = adepth;
    if (p.o_acc != nullptr) p.o_acc[rr] = aacc;
    if (p.o_disp != nullptr) p.o_disp[rr] = comp_disp(adepth, aacc);
  }
}

__device__ __forceinline__ int pp_main_stages(int step) { return step == 0 ? 1 : (step == 5 ? 5 : 4); }

// PROF = true: the instantiation behind r2l_nerf_profile (in-kernel cycle counters / event timeline); the production
// kernel carries none of that code (code size, see the epilogue).
// UNI = true (every render call: ray samples with S a multiple of 32 and S >= 64): the 32 rows of a warp belong to ONE
// ray and a tile holds at most two rays, so the encoder warps stage the tile's (at most two) per-ray view-bias rows
// in shared memory ahead of time and the last epilogue reads them with broadcast LDS.128 — every thread fetching its
// own 16 x float4 from L2 after the accumulator wait put ~2300 cycles into that epilogue (in-kernel trace), and the
// two step-9 epilogues are what the unit boundary waits for.  UNI = false: general S and the NeRF.forward(x) rows
// (every row its own bias row, loaded from global memory).
// COMP = true (needs UNI; r2l_nerf_render's fused route): the kernel composites the rays itself (NerfParams::comp_ring).
// Its own instantiation: the compositor is ~26 KB of code, and the kernel without it is measurably faster per frame
// (these kernels are sensitive to code size, see the epilogue), so the plain forward does not carry it.
template <bool BF16, bool PROF, bool UNI, bool COMP>
__global__ void __launch_bounds__(kPpThreads, 1)
nerf_mlp_pp_kernel(const __grid_constant__ NerfParams p, const __grid_constant__ NerfPpMaps maps,
                   const __grid_constant__ NerfHeadW hw) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* const sA = smem + kPpOffA;
  uint8_t* const sP = smem + kPpOffP;
  uint8_t* const sOnes = smem + kPpOffOnes;
  uint8_t* const sRing = smem + kPpOffRing;
  uint8_t* const sBiasRing = smem + kPpOffBiasRing;
  // [tile][ray of the tile 0|1][128] fp32 view-bias rows (UNI); the head weights that used to live here come from the
  // constant bank now (NerfHeadW)
  float* const sVb = reinterpret_cast<float*>(smem + kPpOffAlphaW);
  float4* const sPart = reinterpret_cast<float4*>(smem + kPpOffPart);
  float* const sAbs = reinterpret_cast<float*>(smem + kPpOffAbs);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + kPpOffBars);
  uint64_t* const w_full = bars;                    // leader: own producer + the peer's relay
  uint64_t* const w_empty = w_full + kPpRing;       // both issuers have committed the stage (2 arrivals, both CTAs)
  uint64_t* const b_full = w_empty + kPpRing;
  uint64_t* const b_empty = b_full + kPpBiasRing;
  uint64_t* const d_full = b_empty + kPpBiasRing;   // [tile]: accumulator complete (commit, both CTAs)
  uint64_t* const a_done = d_full + 2;              // [tile]: leader; 16 epilogue warps are done with the tile's layer
  uint64_t* const p_ready = a_done + 2;             // [tile]: leader; 8 encoder warps have written P for tile t
  uint64_t* const p_free = p_ready + 2;             // [tile]: tile t's MMAs that read P have completed (commit, both CTAs)
  uint64_t* const turn = p_free + 2;                // [tile]: leader; the other issuer has issued its tile-layer
  uint64_t* const vb_ready = turn + 2;              // THIS CTA: its 4 encoder warps have staged the unit's view-bias rows
  uint64_t* const raw_ready = vb_ready + 1;         // [tile] THIS CTA (fused compositing): WG0's 4 warps have staged the tile's (rgb, sigma) rows
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + kPpOffTmem);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader (issues the MMAs)
  // CTA c owns the CONTIGUOUS tiles [c*T, (c+1)*T) (T even): a unit = tiles tile0 + 2*unit + {0, 1} of BOTH CTAs of
  // the pair, so a ray's samples stay inside one CTA (what the fused compositing needs).  The pair runs while its
  // leader (the lower tile range) has tiles left; rows past the end are clamped / masked as before.
  const long long tile0 = static_cast<long long>(blockIdx.x) * p.tiles_per_cta;
  const long long lead_left = static_cast<long long>(p.n_tiles) - static_cast<long long>(blockIdx.x & ~1u) * p.tiles_per_cta;
  const int n_units = lead_left <= 0 ? 0 : static_cast<int>(lead_left >= p.tiles_per_cta ? p.tiles_per_cta / 2 : (lead_left + 1) / 2);
  constexpr int unit0 = 0, unit_step = 1;

  // ---- one-time setup ----
  write_ones_block<BF16>(sOnes, threadIdx.x, kPpThreads);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kPpRing; ++i) {
      mbar_init(&w_full[i], (rank == 0 && !kPpDirect) ? 2 : 1);
      mbar_init(&w_empty[i], 2);
    }
    for (int i = 0; i < kPpBiasRing; ++i) {
      mbar_init(&b_full[i], (rank == 0 && !kPpDirect) ? 2 : 1);
      mbar_init(&b_empty[i], 2);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&d_full[i], 1);
      mbar_init(&a_done[i], 16);
      mbar_init(&p_ready[i], 8);
      mbar_init(&p_free[i], 1);
      mbar_init(&turn[i], 1);
    }
    mbar_init(vb_ready, 4);
    mbar_init(&raw_ready[0], 4);
    mbar_init(&raw_ready[1], 4);
    mbar_fence_init();
  }
  fence_proxy_async_smem();
  if (warp == kPpMmaWarp0) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers are initialised before anybody signals them
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // debug timeline (r2l_nerf_profile with R2L_PROF_MODE=5): CTA 0 records (tag, clock) pairs for its 4th unit
  const bool tracing = PROF && p.prof != nullptr && p.prof_mode == 5 && blockIdx.x == 0;
  // (the prof buffer must hold 3 x 400 + 8 int64 in this mode; each role appends to its own region, no atomics)
  int n_trace = 0;
  auto trace = [&](uint32_t it_, long long tag) {
    if (tracing && it_ == 3 && n_trace < 200) {
      const int role = (tag / 100000 >= 3) ? 2 : static_cast<int>((tag / 10000) % 10);
      p.prof[8 + role * 400 + 2 * n_trace] = tag;
      p.prof[9 + role * 400 + 2 * n_trace] = clock64();
      ++n_trace;
    }
  };

  // ===================== MMA issuer of tile t (leader CTA; one thread per tile) =====================
  // Both issuers walk the same stage sequence (g, gb).  Every barrier wait below is unambiguous although only the
  // phase PARITY is visible: a stage's refill needs BOTH issuers' commits, so neither can be a whole phase away
  // from the barrier it waits on.
  auto mma_issuer = [&](const int t) {
    const uint32_t idesc256 = make_idesc_f16(BF16, 2 * kTileM, 256);
    const uint32_t idesc128 = make_idesc_f16(BF16, 2 * kTileM, 128);
    const uint32_t aA = smem_u32(sA) + t * kABufBytes;
    const uint32_t aP = smem_u32(sP);
    const uint32_t aOnes = smem_u32(sOnes);
    const uint32_t aRing = smem_u32(sRing);
    const uint32_t aBiasRing = smem_u32(sBiasRing);
    const uint32_t d = tmem_base + 256u * t;
    uint32_t g = 0, gb = 0;        // position in the weight / bias stage streams
    uint32_t par_done = 0;         // parity of the next a_done[t] phase
    uint32_t par_p = 0;            // parity of the next p_ready[t] phase
    uint32_t n_turn = 0;           // turns taken so far
    const bool prof = PROF && p.prof != nullptr && p.prof_mode != 5;
    long long t_a = 0, t_w = 0, t_p = 0, t_b = 0;
    const long long t_start = prof ? clock64() : 0;
    auto next_w = [&]() -> uint32_t {
      const uint32_t slot = g % kPpRing;
      const long long c0 = prof ? clock64() : 0;
      mbar_wait(&w_full[slot], (g / kPpRing) & 1, p.dbg, 220 + slot + 30 * t);
      if (prof) t_w += clock64() - c0;
      tc_fence_after_sync();
      return slot;
    };
    auto wait_p = [&]() {
      const long long c0 = prof ? clock64() : 0;
      mbar_wait(&p_ready[t], par_p, p.dbg, 200 + t);
      if (prof) t_p += clock64() - c0;
      par_p ^= 1u;
      tc_fence_after_sync();
    };
    uint32_t it = 0;
    for (int unit = unit0; unit < n_units; unit += unit_step, ++it) {
      for (int step = 0; step < 10; ++step) {
        // Everything that does not depend on the tile's previous epilogue is waited for FIRST (bias stage, first
        // weight stage, point block): the issuer is idle until a_done anyway, and the ~100 cycles of each barrier wait
        // then stay off the a_done -> first MMA path.
        uint32_t bslot = 0;
        if (step < 9) {
          bslot = gb % kPpBiasRing;
          const long long c0 = prof ? clock64() : 0;
          mbar_wait(&b_full[bslot], (gb / kPpBiasRing) & 1, p.dbg, 240 + bslot + 30 * t);
          if (prof) t_b += clock64() - c0;
        }
        if (step == 0 || step == 5) wait_p();
        uint32_t slot = next_w();
        if (kPpTurns && !(t == 0 && (n_turn == 0 || kPpLagOnly))) {
          // my turn: the other issuer has issued all MMAs of its tile-layer (keeps the two tiles' MMA phases apart on
          // the in-order tensor pipe, so a tile's accumulator completes ~2176 cycles after its first MMA, not ~4000)
          const long long c0 = prof ? clock64() : 0;
          mbar_wait(&turn[t], (t == 0 ? n_turn - 1 : n_turn) & 1u, p.dbg, 270 + t);
          if (prof) t_w += clock64() - c0;
        }
        ++n_turn;
        // the tile's previous epilogue is done: A[t] holds this layer's input and D[t] may be overwritten
        if (!(it == 0 && step == 0)) {
          const long long c0 = prof ? clock64() : 0;
          mbar_wait(&a_done[t], par_done, p.dbg, 210 + t);
          if (prof) t_a += clock64() - c0;
          par_done ^= 1u;
        }
        trace(it, 200000 + t * 10000 + step * 100);   // a_done seen: the step may start
        tc_fence_after_sync();
        if (step < 9) {
          // ---- bias step (fresh accumulator), [point-block stage], 4 activation stages
          issue_bias_stage<true>(d, aOnes, aBiasRing + bslot * kPpBiasB, kPpLbo256, idesc256, true);
          umma_commit_pair(&b_empty[bslot]);
          ++gb;
          if (step == 0 || step == 5) {
            trace(it, 100000 + t * 10000 + step * 100 + 9);
            issue_stage<4, true>(d, aP, aRing + slot * kPpStageB, kPpLbo256, idesc256, false);
            if (kPpTurns && step == 0 && !(kPpLagOnly && t == 1)) mbar_arrive(&turn[1 - t]);
            umma_commit_pair(&w_empty[slot]);
            umma_commit_pair(&p_free[t]);
            ++g;
            if (step == 5) slot = next_w();
          }
          if (step > 0) {
            for (int st = 0; st < 4; ++st) {
              if (st > 0) slot = next_w();
              trace(it, 100000 + t * 10000 + step * 100 + st);
              issue_stage<4, true>(d, aA + st * kGroupBytes, aRing + slot * kPpStageB, kPpLbo256, idesc256, false);
              // step 5 has 5 stages for a 4-slot ring: its last stage can only be loaded after the OTHER tile has
              // used the first one, so the turn is passed one stage early there
              if (kPpTurns && st == (step == 5 ? 2 : 3) && !(kPpLagOnly && t == 1)) mbar_arrive(&turn[1 - t]);
              umma_commit_pair(&w_empty[slot]);
              ++g;
            }
          }
        } else {
          // ---- step 9: view branch, N = 128 (per-ray bias added by the epilogue)
          for (int st = 0; st < 4; ++st) {
            if (st > 0) slot = next_w();
            issue_stage<4, true>(d, aA + st * kGroupBytes, aRing + slot * kPpStageB, kPpLbo128, idesc128, st == 0);
            if (kPpTurns && st == 3 && !(kPpLagOnly && t == 1)) mbar_arrive(&turn[1 - t]);
            umma_commit_pair(&w_empty[slot]);
            ++g;
          }
        }
        umma_commit_pair(&d_full[t]);
      }
    }
    if (prof) {
      long long* o = p.prof + blockIdx.x * 8;
      if (t == 0) {
        o[0] = clock64() - t_start;   // issuer T0: total
        o[1] = t_a;                   // waiting for the tile's previous epilogue
        o[2] = t_w + t_b;             // waiting for weight / bias stages
        o[7] = t_p;                   // waiting for the encoder
      } else if (p.prof_mode == 1) {  // issuer T1's split instead of T0's
        o[0] = clock64() - t_start;
        o[1] = t_a;
        o[2] = t_w + t_b;
        o[7] = t_p;
      }
    }
  };

  if (warp == kPpProducerWarp) {
    // ===================== weight producer: this CTA's N-half of every stage, once per unit (both tiles) ==========
    if (lane == 0) {
      uint32_t g = 0, gb = 0;
      auto push = [&](const uint8_t* src, uint32_t full_bytes) {
        const uint32_t half = full_bytes / 2;
        const uint32_t slot = g % kPpRing;
        mbar_wait(&w_empty[slot], ((g / kPpRing) & 1) ^ 1, p.dbg, 100 + slot, 8);
        if (kPpDirect) {
          // the peer's half stage completes its bytes directly on the leader's barrier: no relay hop
          if (rank == 0) {
            mbar_expect_tx(&w_full[slot], full_bytes);
            bulk_g2s(sRing + slot * kPpStageB, src, half, &w_full[slot]);
          } else {
            // rows of 512 bytes: row index of this half stage in the stream
            const int row = static_cast<int>((src + half - p.wstream) >> 9);
            tma2d_g2s_pair_bar(sRing + slot * kPpStageB, half == kPpStageB ? &maps.m16 : &maps.m8, 0, row,
                               mapa_u32(&w_full[slot], 0));
          }
        } else {
          mbar_expect_tx(&w_full[slot], half);
          bulk_g2s(sRing + slot * kPpStageB, src + rank * half, half, &w_full[slot]);
        }
        ++g;
      };
      auto push_bias = [&](const uint8_t* src) {
        const uint32_t slot = gb % kPpBiasRing;
        mbar_wait(&b_empty[slot], ((gb / kPpBiasRing) & 1) ^ 1, p.dbg, 120 + slot, 8);
        if (kPpDirect) {
          if (rank == 0) {
            mbar_expect_tx(&b_full[slot], 2 * kPpBiasB);
            bulk_g2s(sBiasRing + slot * kPpBiasB, src, kPpBiasB, &b_full[slot]);
          } else {
            const int row = static_cast<int>((src + kPpBiasB - p.wstream) >> 9);
            tma2d_g2s_pair_bar(sBiasRing + slot * kPpBiasB, &maps.m4, 0, row, mapa_u32(&b_full[slot], 0));
          }
        } else {
          mbar_expect_tx(&b_full[slot], kPpBiasB);
          bulk_g2s(sBiasRing + slot * kPpBiasB, src + rank * kPpBiasB, kPpBiasB, &b_full[slot]);
        }
        ++gb;
      };
      for (int unit = unit0; unit < n_units; unit += unit_step) {
        const uint8_t* src = p.wstream;
        for (int step = 0; step < 10; ++step) {
          const int nm = pp_main_stages(step);
          const uint32_t sb = step < 9 ? kStageBytes : kStageBytes / 2;
          if (step < 9) {
            push_bias(src);
            src += kBiasStageBytes;
          }
          for (int i = 0; i < nm; ++i) {
            push(src, sb);
            src += sb;
          }
        }
      }
    }
  } else if (warp == kPpMmaWarp0) {
    if (rank != 0) {
      // ===================== peer CTA: relay "my half stage has landed" to the leader's barriers =====================
      if (lane == 0 && !kPpDirect) {
        uint32_t g = 0, gb = 0;
        for (int unit = unit0; unit < n_units; unit += unit_step) {
          for (int step = 0; step < 10; ++step) {
            const int nm = pp_main_stages(step);
            if (step < 9) {
              const uint32_t slot = gb % kPpBiasRing;
              mbar_wait(&b_full[slot], (gb / kPpBiasRing) & 1, p.dbg, 170 + slot, 8);
              mbar_arrive_cluster(mapa_u32(&b_full[slot], 0));
              ++gb;
            }
            for (int i = 0; i < nm; ++i) {
              const uint32_t slot = g % kPpRing;
              mbar_wait(&w_full[slot], (g / kPpRing) & 1, p.dbg, 150 + slot, 8);
              mbar_arrive_cluster(mapa_u32(&w_full[slot], 0));
              ++g;
            }
          }
        }
      }
    }
  }
  if (warp == kPpMmaWarp0 || warp == kPpMmaWarp1) {
    // ONE call site: the issuer code exists once, the tile is a run-time argument (code size, see the epilogue)
    if (rank == 0 && lane == 0) mma_issuer(warp - kPpMmaWarp0);
  } else if (warp == kPpProducerWarp) {
    // (handled above)
  } else if (warp >= 8) {
    // ===================== encoder warpgroup: the single point block, re-encoded before each use =====================
    // Order of uses per unit: T0 step 0, T1 step 0, T0 step 5, T1 step 5.  Before overwriting P the encoder waits for
    // the PREVIOUS use's MMAs (p_free of the other tile).
    const int row = (warp & 3) * 32 + lane;
    uint32_t n_enc = 0;   // uses encoded so far; use n belongs to tile n & 1
    // Fused compositing: the groups (comp_RPW rays = 32 * comp_K rows) whose last row unit v staged, one per warp.
    // The ring rows read here are rewritten by unit v + 2's last epilogues (v + 1's for the carried part of a
    // straddling ray), which come after the step-5 MMAs that wait for the block these warps encode NEXT: no extra
    // barrier is needed for the reuse.
    auto composite_unit = [&](int v) {
      mbar_wait(&raw_ready[0], v & 1u, p.dbg, 520, 8);
      mbar_wait(&raw_ready[1], v & 1u, p.dbg, 521, 8);
      const int grows = 32 * p.comp_K;
      const int g_first = (2 * kTileM * v) / grows, g_end = (2 * kTileM * (v + 1)) / grows;
      for (int g = g_first + (warp & 3); g < g_end; g += 4)
        pp_composite_group(p, p.comp_ring + static_cast<size_t>(blockIdx.x) * kCompRingRows,
                           p.comp_aux + static_cast<size_t>(blockIdx.x) * kCompRingRows,
                           static_cast<unsigned>(tile0) * (kTileM / 32), g, lane);
    };
    for (int unit = unit0; unit < n_units; unit += unit_step) {
      for (int use = 0; use < 4; ++use, ++n_enc) {
        const int t = use & 1;
        const long long tile = tile0 + 2 * unit + t;
        long long gr = tile * kTileM + row;
        if (gr >= p.n_rows) gr = p.n_rows - 1;
        // previous use n_enc-1 was tile 1-t's ((n_enc-1)/2)-th use: its p_free phase index is (n_enc-1)/2
        const uint32_t prev_par = ((n_enc - 1) >> 1) & 1u;
        if (p.embedded != nullptr) {
          // API path: the caller already embedded the points (63 features per row)
          const float* x = p.embedded + gr * p.emb_stride;
          float e[64];
#pragma unroll
          for (int i = 0; i < 64; ++i) e[i] = (i < 63) ? __ldg(x + i) : 0.0f;
          if (n_enc > 0) mbar_wait(&p_free[1 - t], prev_par, p.dbg, 500 + t, 4);
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            uint4 q;
            q.x = pack2<BF16>(e[8 * ch + 0], e[8 * ch + 1]);
            q.y = pack2<BF16>(e[8 * ch + 2], e[8 * ch + 3]);
            q.z = pack2<BF16>(e[8 * ch + 4], e[8 * ch + 5]);
            q.w = pack2<BF16>(e[8 * ch + 6], e[8 * ch + 7]);
            *reinterpret_cast<uint4*>(sP + ch * kChunkBytes + row * 16) = q;
          }
        } else {
          const long long ray = gr / p.S;
          const float z = __ldg(p.z_vals + gr);
          const float* o = p.rays_o + ray * p.o_stride;
          const float* dd = p.rays_d + ray * p.d_stride;
          const float px = __fadd_rn(__ldg(o + 0), __fmul_rn(__ldg(dd + 0), z));
          const float py = __fadd_rn(__ldg(o + 1), __fmul_rn(__ldg(dd + 1), z));
          const float pz = __fadd_rn(__ldg(o + 2), __fmul_rn(__ldg(dd + 2), z));
          // encode into registers FIRST, wait for the block to be free, then only store: the ~800 cycles of sin / cos
          // work stay off the p_free -> p_ready path (T1's steps 0 and 5 start that much earlier after T0's)
          uint4 q[8];
          encode_point_packed<BF16>(px, py, pz, q);
          if (n_enc > 0) mbar_wait(&p_free[1 - t], prev_par, p.dbg, 500 + t, 4);
          store_point_block(sP, row, q);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) lane_arrive<true>(&p_ready[t]);
        // Fused compositing of the PREVIOUS unit's rays: this point (T1's first block of the unit is encoded) is
        // reached right after T0's step 0 of this unit, i.e. when the previous unit's last epilogues are finishing, and
        // the next thing these warps owe anybody (T0's step-5 block) is five layers away.
        if (COMP && use == 1 && unit > 0) composite_unit(unit - 1);
      }
      if (UNI) {
        // the unit's view-bias rows: thread `row` copies column `row` of the (at most two) rays of each tile.  The
        // previous unit's step-9 epilogues have finished reading the buffer: this point is only reached after T0's
        // step-5 MMAs of THIS unit (p_free), which the same epilogue warps fed after finishing the previous unit.
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const long long r_first = (tile0 + 2 * unit + t) * kTileM;
          long long ra = r_first, rb = r_first + kTileM - 1;
          if (ra >= p.n_rows) ra = p.n_rows - 1;
          if (rb >= p.n_rows) rb = p.n_rows - 1;
          sVb[(2 * t + 0) * 128 + row] = __ldg(p.vb + (ra / p.S) * 128 + row);
          sVb[(2 * t + 1) * 128 + row] = __ldg(p.vb + (rb / p.S) * 128 + row);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(vb_ready);
      }
    }
    if (COMP && n_units > 0) composite_unit(n_units - 1);
  } else {
    // ===================== epilogue warpgroups (both tiles, alternating) =====================
    const int wg = warp >> 2;                       // owns the 32-column pieces wg, wg+2, wg+4, wg+6
    const int row = (warp & 3) * 32 + lane;         // tile row == TMEM lane
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const bool prof = PROF && p.prof != nullptr && (threadIdx.x & 127) == 0 && p.prof_mode != 5;
    uint32_t eit = 0;
    long long t_d = 0;
    const long long t_start = prof ? clock64() : 0;
    uint32_t par_d = 0;   // bit tile: parity of the next d_full phase
    auto wait_d = [&](int t, uint32_t id) {
      const long long cd = prof ? clock64() : 0;
      mbar_wait(&d_full[t], (par_d >> t) & 1u, p.dbg, id, 4);
      if (prof) t_d += clock64() - cd;
      par_d ^= 1u << t;
      tc_fence_after_sync();
    };
    // this warp's four 32-column pieces (columns 32*wg + 64*i) of accumulator D[t], software pipelined
    auto for_pieces = [&](int t, auto&& f) {
      uint32_t va[32], vb[32];
      const uint32_t c0 = 32 * wg;
      const uint32_t base = lane_taddr + 256 * t + c0;
      tmem_ld32(base, va);
      tmem_ld_wait();
      tmem_ld32(base + 64, vb);
      f(std::integral_constant<int, 0>{}, c0, va);
      tmem_ld_wait();
      tmem_ld32(base + 128, va);
      f(std::integral_constant<int, 1>{}, c0 + 64, vb);
      tmem_ld_wait();
      tmem_ld32(base + 192, vb);
      f(std::integral_constant<int, 2>{}, c0 + 128, va);
      tmem_ld_wait();
      f(std::integral_constant<int, 3>{}, c0 + 192, vb);
    };
    for (int unit = unit0; unit < n_units; unit += unit_step, ++eit) {
      float sigma_part0 = 0.0f, sigma_part1 = 0.0f;   // per tile (scalars: t is a run-time index)
      float abs_part0 = 0.0f, abs_part1 = 0.0f;       // sum |w_a| relu(h7): scale of the far-sample guard band (nerf_far.cu)
      for (int step = 0; step < 10; ++step) {
        // t is a RUN-TIME loop variable on purpose: unrolled, the epilogue code of both tiles (4 for_pieces variants
        // x 2) made the kernel 131 KB of SASS; A/B on one box showed every step slowing down once the kernel grew
        // past ~128 KB (instruction cache), so code size is kept in check here
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
          uint8_t* const a_row = sA + t * kABufBytes + row * 16;
          wait_d(t, 300 + step * 2 + t);
          if (threadIdx.x == 0) trace(eit, 300000 + t * 10000 + step * 100);   // epilogue of (t, step) starts
          if (step < 9) {
            if (step == 7) {
              float sp = 0.0f, sa = 0.0f;
              // alpha_linear on the fp32 accumulators; the weight index is a compile-time constant per (WG, piece, i)
              auto sigma_dot = [&](auto wg_c) {
                constexpr int WG = decltype(wg_c)::value;
                for_pieces(t, [&](auto piece_c, uint32_t col0, uint32_t (&v)[32]) {
                  constexpr int C0 = 32 * WG + 64 * decltype(piece_c)::value;
                  store_sub<BF16, true>(v, a_row + (col0 >> 5) * kSubBytes);
#pragma unroll
                  for (int i = 0; i < 32; ++i) {
                    const float h = relu_nan(__uint_as_float(v[i]));
                    sp = fmaf(hw.alpha_w[C0 + i], h, sp);
                    if ((i & 3) == 0) sa = fmaf(fabsf(hw.alpha_w[C0 + i]), h, sa);   // every 4th term: a scale estimate
                  }
                });
              };
              if (wg == 0) sigma_dot(std::integral_constant<int, 0>{}); else sigma_dot(std::integral_constant<int, 1>{});
              if (t == 0) sigma_part0 = sp, abs_part0 = sa; else sigma_part1 = sp, abs_part1 = sa;
            } else if (step == 8) {
              for_pieces(t, [&](auto, uint32_t col0, uint32_t (&v)[32]) {
                store_sub<BF16, false>(v, a_row + (col0 >> 5) * kSubBytes);
              });
            } else {
              for_pieces(t, [&](auto, uint32_t col0, uint32_t (&v)[32]) {
                store_sub<BF16, true>(v, a_row + (col0 >> 5) * kSubBytes);
              });
            }
            warp_arrive<true>(&a_done[t], lane);
            if (threadIdx.x == 0) trace(eit, 400000 + t * 10000 + step * 100);   // ... and has ended (warp 0)
          } else {
            // step 9: view branch (N = 128 -> D[t] columns [0,128)); this warp owns columns 32*wg and 64 + 32*wg.
            // The TMEM loads are issued FIRST: the index arithmetic below (64-bit divisions) runs under their latency.
            uint32_t va[32], vb[32];
            tmem_ld32(lane_taddr + 256 * t + 32 * wg, va);
            tmem_ld32(lane_taddr + 256 * t + 64 + 32 * wg, vb);
            const long long tile = tile0 + 2 * unit + t;
            const long long g_row = tile * kTileM + row;
            const bool valid = g_row < p.n_rows;
            long long ray;
            const float4* vb4;
            if (UNI) {
              // this warp's 32 rows share one ray: the tile's first or second (rows of a tile straddle at most 2 rays).
              // S is a multiple of 32 here, so ray = (row / 32) / (S / 32) on 32-row units (tile counts are < 2^29: host)
              const unsigned s32 = static_cast<unsigned>(p.S) >> 5;
              unsigned u_w = 4u * static_cast<unsigned>(tile) + static_cast<unsigned>(warp & 3), u_t = 4u * static_cast<unsigned>(tile);
              const unsigned u_last = static_cast<unsigned>((p.n_rows - 1) >> 5);
              u_w = min(u_w, u_last), u_t = min(u_t, u_last);
              const unsigned ray_w = u_w / s32;
              ray = ray_w;
              vb4 = reinterpret_cast<const float4*>(sVb + (2 * t + (ray_w != u_t / s32 ? 1 : 0)) * 128);
              if (t == 0) mbar_wait(vb_ready, eit & 1u, p.dbg, 330);   // staged long ago; one phase per unit
            } else {
              ray = (valid ? g_row : (p.n_rows - 1)) / p.S;
              vb4 = reinterpret_cast<const float4*>(p.vb + ray * 128);
            }
            float r = 0.f, gch = 0.f, b = 0.f;
            {
              tmem_ld_wait();
              // D[t] is drained: the next unit's step 0 may overwrite it.  (No proxy fence: this epilogue wrote no
              // shared-memory operand.)
              tc_fence_before_sync();
              __syncwarp();
              if (lane == 0) lane_arrive<true>(&a_done[t]);
              // rgb_linear on relu(accumulator + per-ray view bias); weights from the constant bank (static indices)
              auto rgb_dot = [&](auto wg_c) {
                constexpr int WG = decltype(wg_c)::value;
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                  const float4 bb = UNI ? vb4[8 * WG + i4] : __ldg(vb4 + 8 * WG + i4);
                  const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    constexpr int dummy = 0;
                    (void)dummy;
                    const int n = 32 * WG + 4 * i4 + k;
                    const float x = relu_nan(__uint_as_float(va[4 * i4 + k]) + bv[k]);
                    r = fmaf(hw.rgb_w[n], x, r);
                    gch = fmaf(hw.rgb_w[128 + n], x, gch);
                    b = fmaf(hw.rgb_w[256 + n], x, b);
                  }
                }
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                  const float4 bb = UNI ? vb4[16 + 8 * WG + i4] : __ldg(vb4 + 16 + 8 * WG + i4);
                  const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const int n = 64 + 32 * WG + 4 * i4 + k;
                    const float x = relu_nan(__uint_as_float(vb[4 * i4 + k]) + bv[k]);
                    r = fmaf(hw.rgb_w[n], x, r);
                    gch = fmaf(hw.rgb_w[128 + n], x, gch);
                    b = fmaf(hw.rgb_w[256 + n], x, b);
                  }
                }
              };
              if (wg == 0) rgb_dot(std::integral_constant<int, 0>{}); else rgb_dot(std::integral_constant<int, 1>{});
            }
            const float sigma_mine = t == 0 ? sigma_part0 : sigma_part1, abs_mine = t == 0 ? abs_part0 : abs_part1;
            if (wg == 1) {
              sPart[row] = make_float4(r, gch, b, sigma_mine);
              sAbs[row] = abs_mine;
              named_bar_arrive(1, 256);
              named_bar_sync(2, 256);   // WG0 has consumed sPart
            } else {
              named_bar_sync(1, 256);
              const float4 o1 = sPart[row];
              const float a1 = sAbs[row];
              named_bar_arrive(2, 256);
              if (valid) {
                float4 o;
                o.x = r + o1.x + p.rgb_b[0];
                o.y = gch + o1.y + p.rgb_b[1];
                o.z = b + o1.z + p.rgb_b[2];
                o.w = sigma_mine + o1.w + p.alpha_b;
                note_nonfinite(p.dbg, o.x + o.y + o.z + o.w, g_row);
                const int fslot = nerf_far_flag(p, g_row, ray, o.w, 4.0f * (abs_mine + a1));
                if (COMP) {
                  // fused compositing: stage the row in this CTA's ring (global memory, L2 resident)
                  const size_t slot = static_cast<size_t>(blockIdx.x) * kCompRingRows +
                                      ((2u * kTileM * static_cast<unsigned>(unit) + kTileM * t + row) & (kCompRingRows - 1));
                  p.comp_ring[slot] = o;
                  if (g_row - ray * p.S == p.S - 1) p.comp_aux[slot] = fslot;
                } else {
                  reinterpret_cast<float4*>(p.raw)[g_row] = o;
                }
              }
              if (COMP) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&raw_ready[t]);
              }
            }
          }
        }
      }
    }
    if (prof) {
      long long* o = p.prof + blockIdx.x * 8 + 3 + wg * 2;
      o[0] = t_d;                          // WG: waiting for accumulators
      o[1] = (clock64() - t_start) - t_d;  // WG: epilogue work (everything else)
    }
  }

  // ---- teardown ----
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();   // both CTAs are done with their TMEM and with each other's barriers
  if (warp == kPpMmaWarp0) {
    tc_fence_after_sync();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// Per-ray view-branch bias: vb[ray][n] = bv[n] + sum_j Wvd[n][j] * embed4(viewdir)[j]   (fp32)
// embed4 = NeRF Embedder with L = 4 on the unit view direction (27 features, helpers:24-74); the view half of
// views_linears[0] is constant along a ray, so it is evaluated once per ray (model/nerf_raybased.py:390-394).
// One thread per output neuron n: its 27 weights stay in registers, a block works on 32 rays per iteration (96 threads
// embed one (ray, coordinate) each; every ray's 128 outputs leave as one coalesced 512-byte row).  The accumulation
// order (bias first, features ascending) is that of the earlier 8-rays-per-iteration version: same bits.
constexpr int kVbRays = 32;
__global__ void __launch_bounds__(128)
nerf_view_bias_kernel(long long n_rays, const float* __restrict__ viewdirs, long long v_stride, int pre_embedded,
                      const float* __restrict__ wvd /*[128][27]*/, const float* __restrict__ bv, float* __restrict__ vb) {
  __shared__ __align__(16) float s_e[kVbRays][28];
  const int n = threadIdx.x;
  float w[27];
#pragma unroll
  for (int j = 0; j < 27; ++j) w[j] = __ldg(wvd + n * 27 + j);
  const float b = bv[n];
  if (threadIdx.x < kVbRays) s_e[threadIdx.x][27] = 0.0f;
  for (long long r0 = static_cast<long long>(blockIdx.x) * kVbRays; r0 < n_rays;
       r0 += static_cast<long long>(gridDim.x) * kVbRays) {
    __syncthreads();
    if (pre_embedded) {
      // rows already hold the 27 embedded view features
      for (int it = threadIdx.x; it < kVbRays * 27; it += 128) {
        const int rr = it / 27, j = it % 27;
        const long long ray = r0 + rr;
        if (ray < n_rays) s_e[rr][j] = viewdirs[ray * v_stride + j];
      }
    } else if (threadIdx.x < kVbRays * 3) {
      // one (ray, coordinate) per thread: identity + 4 sincos
      const int rr = threadIdx.x / 3, c = threadIdx.x % 3;
      const long long ray = r0 + rr;
      if (ray < n_rays) {
        const float v = viewdirs[ray * v_stride + c];
        s_e[rr][c] = v;
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          float sn, co;
          sincosf(v * static_cast<float>(1 << f), &sn, &co);
          s_e[rr][3 + 6 * f + c] = sn;
          s_e[rr][3 + 6 * f + 3 + c] = co;
        }
      }
    }
    __syncthreads();
    const int n_here = static_cast<int>(n_rays - r0 < kVbRays ? n_rays - r0 : kVbRays);
    for (int rr = 0; rr < n_here; ++rr) {
      const float4* e4 = reinterpret_cast<const float4*>(s_e[rr]);
      float acc = b;
#pragma unroll
      for (int q = 0; q < 7; ++q) {
        const float4 e = e4[q];   // broadcast: every thread reads the same row
        acc = fmaf(w[4 * q + 0], e.x, acc);
        acc = fmaf(w[4 * q + 1], e.y, acc);
        acc = fmaf(w[4 * q + 2], e.z, acc);
        if (4 * q + 3 < 27) acc = fmaf(w[4 * q + 3], e.w, acc);
      }
      vb[(r0 + rr) * 128 + n] = acc;
    }
  }
}

template <bool BF16, bool PROF, bool UNI, bool COMP = false>
int launch_nerf_pp(const NerfParams& p, const NerfPpMaps& maps, const NerfHeadW& hw, int grid, cudaStream_t st) {
  R2L_CUDA(cudaFuncSetAttribute(nerf_mlp_pp_kernel<BF16, PROF, UNI, COMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPpSmemBytes));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(kPpThreads);
  cfg.dynamicSmemBytes = kPpSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  R2L_CUDA(cudaLaunchKernelEx(&cfg, nerf_mlp_pp_kernel<BF16, PROF, UNI, COMP>, p, maps, hw));
  count_launch();
  return R2L_OK;
}

// grid must be even (CTA pairs); weights packed in the pair layout WITHOUT the view stage (mlp_api.cu)
int nerf_mlp_pp_launch(bool bf16, const NerfParams& p, const NerfPpMaps& maps, const NerfHeadW& hw, int grid,
                       cudaStream_t st) {
  const bool uni = p.embedded == nullptr && p.S >= 64 && (p.S % 32) == 0;
  if (p.comp_ring != nullptr) {
    if (!uni || p.prof != nullptr) return fail(R2L_ERR_INVALID, "nerf_mlp_pp_launch: fused compositing needs ray samples with S %% 32 == 0");
    return bf16 ? launch_nerf_pp<true, false, true, true>(p, maps, hw, grid, st)
                : launch_nerf_pp<false, false, true, true>(p, maps, hw, grid, st);
  }
  if (p.prof != nullptr)   // the profiling hooks exist for the render form only
    return bf16 ? launch_nerf_pp<true, true, true>(p, maps, hw, grid, st) : launch_nerf_pp<false, true, true>(p, maps, hw, grid, st);
  if (uni) return bf16 ? launch_nerf_pp<true, false, true>(p, maps, hw, grid, st) : launch_nerf_pp<false, false, true>(p, maps, hw, grid, st);
  return bf16 ? launch_nerf_pp<true, false, false>(p, maps, hw, grid, st) : launch_nerf_pp<false, false, false>(p, maps, hw, grid, st);
}

int nerf_view_bias_launch(long long n_rays, const float* viewdirs, long long v_stride, int pre_embedded,
                          const float* wvd, const float* bv, float* vb, cudaStream_t st) {
  long long blocks = (n_rays + kVbRays - 1) / kVbRays;
  const long long cap = static_cast<long long>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  nerf_view_bias_kernel<<<static_cast<int>(blocks), 128, 0, st>>>(n_rays, viewdirs, v_stride, pre_embedded, wvd, bv, vb);
  R2L_LAUNCH_CHECK();
  return R2L_OK;
}

}  // namespace r2l
