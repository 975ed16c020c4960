/* r2l_b200 — C ABI of the B200-native per-ray rendering hot path (R2L / NeRF).
 *
 * Drop-in boundary.  The reference (MingSun-Tse/Efficient-NeRF) has no FFI or operator
 * registry: its boundary is the Python call surface of utils/run_nerf_raybased_helpers.py,
 * model/nerf_raybased.py and the render functions in main.py.  Each entry point below is the
 * device-side body of one of those functions; the Python mirror in efficient-nerf_b200/
 * (same names and signatures as the reference) allocates outputs with torch and calls these
 * through ctypes.  See INTEGRATION.md for the binding a maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; r2l_last_error() gives the text
 *     (thread-local).  No exceptions cross the ABI.  There is NO CPU fallback.
 *     Codes: 1 invalid argument, 2 CUDA error, 3 unsupported configuration, 4 device trap (kernel watchdog),
 *     5 range (a weight or an activation does not fit the 16-bit operand type, see r2l_mlp_status).
 *   - all pointers are DEVICE pointers to fp32 data unless stated otherwise; tensors are
 *     row-major and contiguous, `*_stride` arguments are row strides in ELEMENTS.
 *   - `stream` is a cudaStream_t (as void*); calls are asynchronous on that stream.
 *   - the library never allocates on the hot path: workspaces are passed in by the caller;
 *     only the *_create functions allocate (packed weights).  Exception: a NeRF handle owns three
 *     workspaces that GROW to the largest call seen (per-ray view bias, flagged-ray list, and for
 *     r2l_nerf_render the raw / staging buffers); growing synchronises the stream once.
 *   - a model handle serves ONE call at a time: calls on the same handle from several streams or
 *     threads must be serialised by the caller (they share the handle's workspaces and its status
 *     record).  Different handles are independent.
 */
#ifndef R2L_B200_H_
#define R2L_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

const char* r2l_last_error(void);
int r2l_abi_version(void);
/* Number of CUDA kernels this library has launched so far in this process (bench.py's gpu_launches). */
long long r2l_kernel_launches(void);

/* ---- rays --------------------------------------------------------------------------------- */

/* get_rays(H, W, focal, c2w) -> rays_o, rays_d  [H*W, 3]
 * replaces utils/run_nerf_raybased_helpers.py:231-257 (and get_rays_np :428-441).
 * c2w: device [3,4] row-major. */
int r2l_get_rays(int H, int W, double focal, const float* c2w, float* rays_o, float* rays_d, void* stream);

/* ndc_rays(H, W, focal, near, rays_o, rays_d) -> (o, d)  [n, 3]
 * replaces utils/run_nerf_raybased_helpers.py:260-279. */
int r2l_ndc_rays(long long n, int H, int W, double focal, double near, const float* rays_o, const float* rays_d,
                 float* out_o, float* out_d, void* stream);

/* viewdirs = rays_d / |rays_d|   replaces main.py:148-157. */
int r2l_normalize_dirs(long long n, const float* dirs, long long stride, float* out, void* stream);

/* z_vals = near*(1-t) + far*t (or lindisp), optional stratified perturbation with caller-supplied
 * t_rand [n,S] (may be NULL).  replaces main.py:676-699.  near/far: per-ray, stride nf_stride. */
int r2l_z_vals(long long n, int S, const float* near, const float* far, long long nf_stride, const float* t_vals,
               int lindisp, const float* t_rand, float* z_out, void* stream);

/* pts[r,s,:] = o_r + d_r * z_rs   replaces main.py:701,733 and model/nerf_raybased.py:114-126
 * (PointSampler.sample_train; z_stride = 0 shares one z row between all rays). */
int r2l_points_from_rays(long long n, int S, const float* rays_o, long long o_stride, const float* rays_d,
                         long long d_stride, const float* z, long long z_stride, float* pts, void* stream);

/* PointSampler.sample_test(c2w) -> pts [H*W, S*3]   replaces model/nerf_raybased.py:76-102. */
int r2l_point_sample(int H, int W, double focal, const float* c2w, const float* z_vals, int S, float* pts,
                     void* stream);

/* The same for a batch of poses in one launch: c2w [n_poses][3][4] -> pts [n_poses][H*W][S*3] (the test loop of
 * main.py:297-301 renders one pose per call; batching poses lets one fused-MLP launch fill whole waves of tiles). */
int r2l_point_sample_batch(int n_poses, int H, int W, double focal, const float* c2w, const float* z_vals, int S,
                           float* pts, void* stream);

/* Pluecker ray representation out [n,6] = [rays_d, cross(rays_o, rays_d)]; o_stride == 0 broadcasts one origin.
 * replaces PointSampler.sample_train_plucker / sample_test_plucker, model/nerf_raybased.py:170-190 (main.py:296-298). */
int r2l_plucker(long long n, const float* rays_o, long long o_stride, const float* rays_d, long long d_stride,
                float* out, void* stream);

/* ---- positional encoding ------------------------------------------------------------------ */

/* layout 0: Embedder.embed (utils/run_nerf_raybased_helpers.py:24-74), x [rows, D] -> [rows, D*(1+2L)]
 * layout 1: PositionalEmbedder.__call__ (model/nerf_raybased.py:191-208), x [rows, D] -> [rows, D*(2L+1)] */
int r2l_embed(long long rows, int D, int L, int include_input, int layout, const float* x, float* out,
              void* stream);

/* ---- compositing and hierarchical sampling ------------------------------------------------ */

/* raw2outputs(raw, z_vals, rays_d, noise, white_bkgd) -> rgb_map [n,3], disp_map [n], acc_map [n],
 * weights [n,S], depth_map [n] (any output may be NULL).  replaces main.py:556-621 and its twins
 * (utils/create_data.py:335-402, model/nerf_raybased.py:226-295, helpers:77-144).
 * noise: optional [n,S] already scaled by raw_noise_std. */
int r2l_raw2outputs(long long n_rays, int S, const float* raw, const float* z_vals, const float* rays_d,
                    long long d_stride, const float* noise, int white_bkgd, float* rgb_map, float* disp_map,
                    float* acc_map, float* weights, float* depth_map, void* stream);

/* sample_pdf(bins [n,nb], weights [n,nb-1], u) -> samples [n,Ni] (+ optional searchsorted indices,
 * int64 [n,Ni]).  replaces utils/run_nerf_raybased_helpers.py:283-330.  u: [Ni] shared by all rays
 * (det=True linspace) or [n,Ni] (u_per_ray=1); built by the caller exactly like the reference does
 * on the host.  Bin indices are bit-exact with the reference's CPU path. */
int r2l_sample_pdf(long long n_rays, int nb, int Ni, const float* bins, long long bins_stride,
                   const float* weights, long long w_stride, const float* u, int u_per_ray, float* samples,
                   long long* inds_out, void* stream);

/* z_out = sort(cat[za, zb]) per ray; z_std = std(zb, unbiased=False) (may be NULL).
 * replaces main.py:730-732 and :750. */
int r2l_merge_sorted(long long n_rays, int na, int nbv, const float* za, const float* zb, float* z_out,
                     float* z_std, void* stream);

/* Fused hierarchical sampling for 64 coarse samples; u: shared table [Ni] (u_per_ray=0), per-ray variates [n,Ni]
 * (u_per_ray=1), or a shared table the caller guarantees to be ASCENDING (u_per_ray=2: det=True's linspace, every
 * render — the search-free kernel; same samples / inds / merged depths); Ni = 64 or 128:
 *   z_out [n, 64+Ni] = sort(cat[z_vals, sample_pdf(.5*(z[1:]+z[:-1]), weights[:,1:-1], Ni, u)]), z_std [n] (may be NULL)
 * = main.py:720-733 and :750 in one launch; optional samples [n,Ni] / inds int64 [n,Ni].  z_vals, weights: [n,64]
 * contiguous.  Bit-identical to r2l_sample_pdf followed by r2l_merge_sorted. */
int r2l_hier_sample(long long n_rays, int n_coarse, int Ni, const float* z_vals, const float* weights, const float* u,
                    int u_per_ray, float* z_out, float* z_std, float* samples, long long* inds_out, void* stream);

/* ---- image metrics of the render_path caller ------------------------------------ */

/* abs_err [n_img][n_per_img] = |a - b| (may be NULL; main.py:331) and sum_sq [n_img] (device double, zeroed by the
 * call) = sum (a-b)^2, one pass.  img2mse (helpers:19) = sum_sq / n_per_img; mse2psnr (helpers:20) = -10 log10(mse).
 * replaces the per-frame error / PSNR stage of render_path, main.py:330-332, 384-390. */
int r2l_image_error(int n_img, long long n_per_img, const float* a, const float* b, float* abs_err, double* sum_sq,
                    void* stream);

/* SSIM of utils/ssim_torch.py:27-54 as called from main.py:46,333-335: a, b [n_img][H][W][3] fp32 (image i at
 * a + i*img_stride floats); taps = HOST pointer to the 11 normalised Gaussian taps (ssim_torch.py:10-16);
 * sum_out [n_img] (device double, zeroed by the call) = sum of the SSIM map over the 3 channels;
 * SSIM_i = sum_out[i] / (3 H W). */
int r2l_ssim(int n_img, int H, int W, const float* a, const float* b, long long img_stride, const float* taps,
             double* sum_out, void* stream);

/* LDR-FLIP of frame stacks (utils/flip_loss.py:69-140 as called from main.py:370-379): test, ref [n_img][H][W][3]
 * fp32, every value mapped v -> v*scale + offset on load (main.py:366-368 rescales both stacks to [-1,1] first).
 * consts: HOST pointer to 26 floats (sRGB->XYZ matrix A[9], its inverse[9], illuminant A*1 [3], cmax, qc, qf, pc, pt);
 * taps: DEVICE pointer to the CSF filters A / RG / BY [3][(2 r_csf+1)^2] followed by the edge / point x-filters
 * [2][(2 r_feat+1)^2], computed by the caller as the reference computes them; workspace: DEVICE [n_img][2][3][H][W]
 * floats.  flip_map (optional) [n_img][H][W]; sum_out [n_img] double (zeroed by the call): mean = sum / (H W). */
int r2l_flip(int n_img, int H, int W, const float* test, const float* ref, long long img_stride, double scale_test,
             double offset_test, double scale_ref, double offset_ref, const float* consts, int r_csf, int r_feat,
             const float* taps, float* workspace, float* flip_map, double* sum_out, void* stream);

/* ---- MLPs ---------------------------------------------------------------------------------- */

/* Y = act(X W^T + b [+ R]) in fp32 on CUDA cores (act: 0 none, 1 relu, 2 sigmoid); the
 * precision="fp32" path of NeRF / ResMLP / NeRF_v3_2 forward (model/nerf_raybased.py:377-401,
 * 461-465, 539-544) for any architecture.  K, ldx, ldw multiples of 4. */
int r2l_linear_fp32(long long M, int N, int K, const float* X, long long ldx, const float* W, long long ldw,
                    const float* bias, float* Y, long long ldy, int act, const float* R, long long ldr,
                    void* stream);

/* Packed NeRF (D=8, W=256, input_ch=63, input_ch_views=27, skips=[4], use_viewdirs) for the fused
 * tcgen05 kernel.  Weights: reference nn.Linear layout, device fp32 (model/nerf_raybased.py:357-372).
 * dtype: 0 = fp16 operands, 1 = bf16 operands (fp32 accumulate either way). */
int r2l_nerf_create(void** out_handle, int dtype, const float* const* pts_w, const float* const* pts_b,
                    const float* views_w, const float* views_b, const float* feature_w, const float* feature_b,
                    const float* alpha_w, const float* alpha_b, const float* rgb_w, const float* rgb_b,
                    void* stream);

/* run_network + NeRF.forward fused (main.py:65-87, model/nerf_raybased.py:377-401):
 * raw[r,s,:] = NeRF(embed(o_r + d_r z_rs), embed(viewdir_r)); raw [n_rays,S,4] = (rgb, sigma). */
int r2l_nerf_forward(void* handle, long long n_rays, int S, const float* rays_o, long long o_stride,
                     const float* rays_d, long long d_stride, const float* viewdirs, long long v_stride,
                     const float* z_vals, float* raw, void* stream);

/* render_rays' inner half in ONE call (main.py:707-709 coarse, :738-741 fine): raw = network(points o + d z), then
 * raw2outputs(raw, z_vals, rays_d, raw_noise_std = 0, white_bkgd) (main.py:556-621) ->
 * rgb_map [n_rays,3], disp_map / acc_map / depth_map [n_rays] (each may be NULL), weights [n_rays,S] (may be NULL).
 * Two routes, same bits: r2l_nerf_forward into a workspace of the handle + r2l_raw2outputs (default), or — fused mode,
 * CTA-pair kernel, S in {64,128,192,256} — the MLP kernel composites the rays itself and raw [n_rays,S,4] is never
 * written (tests/test_gpu_frames.py).  The far-sample fix-up applies as in r2l_nerf_forward. */
int r2l_nerf_render(void* handle, long long n_rays, int S, const float* rays_o, long long o_stride,
                    const float* rays_d, long long d_stride, const float* viewdirs, long long v_stride,
                    const float* z_vals, int white_bkgd, float* rgb_map, float* disp_map, float* acc_map,
                    float* weights, float* depth_map, void* stream);
/* fused = 0 (default; environment R2L_NERF_FUSED=1 changes it): two-step route; 1: fused where it applies.  Fused
 * saves the raw round trip (1.3 GB of HBM traffic per 400x400 frame) but is ~1 % slower per frame. */
int r2l_nerf_render_mode(void* handle, int fused);

/* Far-sample sigma fix-up (main.py:578-581, 598: the last sample's 1e10 interval makes its alpha a step function of
 * sign(sigma)).  mode 1 (default): r2l_nerf_forward flags the rays whose last sample has
 * |sigma| < max(abs_band, rel_band * sum_i |alpha_w_i| relu(h7_i)) and re-evaluates those points in fp32 (bit-identical
 * to the precision="fp32" path) before returning raw; mode 0: off.  Negative bands keep the current values
 * (defaults 1e-4, 2^-10). */
int r2l_nerf_far_fixup(void* handle, int mode, double abs_band, double rel_band);
/* Rays flagged by the last r2l_nerf_forward enqueued on `stream` (synchronises it). */
int r2l_nerf_far_count(void* handle, long long* out_count, void* stream);

/* NeRF.forward(x [M, >=90]) -> [M,4]   (model/nerf_raybased.py:377-401). */
int r2l_nerf_forward_embedded(void* handle, long long M, const float* x, long long ldx, float* out, void* stream);

/* Profiling hook: r2l_nerf_forward + per-CTA cycle counters prof[n_CTAs][8] (device int64): MMA-thread total /
 * wait-for-A / wait-for-weights, WG0 and WG1 wait-for-accumulator / epilogue work, MMA wait-for-encoder. */
int r2l_nerf_profile(void* handle, long long n_rays, int S, const float* rays_o, long long o_stride,
                     const float* rays_d, long long d_stride, const float* viewdirs, long long v_stride,
                     const float* z_vals, float* raw, long long* prof, void* stream);

/* Packed NeRF_v3_2 with ResMLP body (model/nerf_raybased.py:443-544): head Linear(n_points*63, 256),
 * n_blocks x ResMLP(256, n_learnable=2), tail Linear(256, 3) [+ Sigmoid]. */
int r2l_resmlp_create(void** out_handle, int dtype, int n_points, int n_blocks, const float* head_w,
                      const float* head_b, const float* const* w1, const float* const* b1, const float* const* w2,
                      const float* const* b2, double res_scale, const float* tail_w, const float* tail_b,
                      int sigmoid_out, int outer_skip, void* stream);

/* model(positional_embedder(pts)) fused (main.py:297-309): pts [n_rays, n_points*3] -> rgb [n_rays,3]. */
int r2l_resmlp_forward(void* handle, long long n_rays, const float* pts, long long pts_stride, float* rgb,
                       void* stream);

/* The whole R2L frame in ONE kernel: PointSampler.sample_test (model/nerf_raybased.py:94-102) + PositionalEmbedder +
 * NeRF_v3_2 (main.py:297-309).  The head generates its rays from the pixel index: c2w [n_poses][3][4] (device),
 * z_vals [n_sample] (device; PointSampler's linspace(near, far)); renders rays [ray0, ray0+n_rays) of the pose-major
 * range [n_poses][H*W] into rgb [n_rays][3] — or, with peer_frames != NULL (n_peers <= 8), into every GPU's frame
 * buffer at rows row0.. like r2l_resmlp_forward_gather.  Bit-identical to r2l_point_sample + r2l_resmlp_forward. */
int r2l_resmlp_render(void* handle, int n_poses, int H, int W, double focal, const float* c2w, const float* z_vals,
                      int n_sample, long long ray0, long long n_rays, float* rgb, float* const* peer_frames,
                      int n_peers, long long row0, void* stream);

/* Ray-sharded frame (the reference's nn.DataParallel gather, main.py:37-42): r2l_resmlp_forward whose tail
 * stores this rank's rows [row0, row0+n_rays) into EVERY GPU's frame buffer by peer-to-peer stores — compute and
 * all-gather in one kernel.  peer_frames: HOST array of n_peers (<= 8) device pointers to [>= row0+n_rays][3] fp32
 * buffers mapped for peer access (symmetric memory), this GPU's own included.  The caller publishes the frame with a
 * cross-GPU barrier on the same stream. */
int r2l_resmlp_forward_gather(void* handle, long long n_rays, const float* pts, long long pts_stride,
                              float* const* peer_frames, int n_peers, long long row0, void* stream);

/* NeRF_v3_2.forward(x [n_rays, n_points*63]) -> [n_rays,3]   (model/nerf_raybased.py:539-544). */
int r2l_resmlp_forward_embedded(void* handle, long long n_rays, const float* x, long long ldx, float* rgb,
                                void* stream);

/* Debug hook for the tests: r2l_resmlp_forward that also dumps the head layer's accumulators and x0
 * ([ceil(n_rays/128)*128, 256] fp32 each). */
int r2l_resmlp_debug_head(void* handle, long long n_rays, const float* pts, long long pts_stride, float* rgb,
                          float* head_acc, float* head_x0, void* head_a, void* stream);

/* Profiling hook: r2l_resmlp_forward + per-CTA cycle counters prof[n_CTAs][8] (device int64): MMA-thread total /
 * wait-for-A / wait-for-weights, WG0 and WG1 wait-for-accumulator / epilogue work, WG0 encode time. */
int r2l_resmlp_profile(void* handle, long long n_rays, const float* pts, long long pts_stride, float* rgb,
                       long long* prof, void* stream);

/* Debug hook for the tests: copy the packed 16-bit weight stage stream to HOST memory. */
int r2l_mlp_debug_wstream(void* handle, void* out_host, unsigned long long capacity, unsigned long long* bytes);

int r2l_mlp_destroy(void* handle);

/* 0 = healthy; 4 (device trap) if a kernel watchdog fired; 5 (range) if a finished launch produced non-finite outputs,
 * i.e. an activation left the range of the 16-bit operands (fp16: 65504).  The activation converts do not saturate, so
 * an overflow propagates as inf / NaN to the output row and is detected there instead of rendering a silently clamped
 * colour; the next forward call on the handle fails with the same code (once).  Weights are range-checked by
 * r2l_nerf_create / r2l_resmlp_create.  Reads the mapped record without synchronising.  out8: optional 8 x uint32. */
int r2l_mlp_status(void* handle, unsigned int* out8);

/* Unit-test probe: D [128,N] = A [128,K] x W [N,K]^T with 16-bit operands on tcgen05. */
int r2l_tc_gemm_probe(int dtype, int N, int K, const float* A, const float* W, float* D, int swap_lbo_sbo,
                      void* stream);

/* Unit-test probe of the CTA-pair path: D [256,N] = A [256,K] x W [N,K]^T on tcgen05.mma.cta_group::2. */
int r2l_tc_gemm_probe_pair(int dtype, int N, int K, const float* A, const float* W, float* D, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* R2L_B200_H_ */
