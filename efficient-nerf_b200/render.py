"""B200-native stand-in for the render glue of the reference's main.py / utils/create_data.py:

  batchify       main.py:51-62        run_network   main.py:65-87
  batchify_rays  main.py:90-104       render        main.py:107-186
  raw2outputs    main.py:556-621      render_rays   main.py:624-756
  (utils/create_data.py:41-176, 335-544 are identical copies whose render_rays also returns
   'depth_map' — pass return_depth=True, or use `render_rays_create_data`.)

`render_rays` keeps the reference signature.  When `network_fn` / `network_fine` are this package's
`NeRF` modules on the tensor-core path and the ray batch carries view directions, the whole
coarse -> composite -> sample_pdf -> merge -> fine -> composite chain runs as fused CUDA kernels
(positional encoding never touches HBM, sample_pdf never leaves the device) and
`network_query_fn`'s Python chunking is bypassed.  Otherwise `network_query_fn` is called like the
reference does.  Random draws (perturb, raw_noise_std, non-deterministic u) come from the CPU
generator in the reference's order (t_rand -> coarse noise -> u -> fine noise) and may be injected.
"""
import numpy as np
import torch

from . import _lib
from .nerf_raybased import NeRF
from .run_nerf_raybased_helpers import _dev  # noqa: E402
from .run_nerf_raybased_helpers import hier_sample, hier_sample_supported  # noqa: E402
from .run_nerf_raybased_helpers import (get_rays, ndc_rays, normalize_dirs, raw2outputs, sample_pdf, merge_sorted,
                                        _host_noise, _make_u)

# rays processed per launch when nothing stochastic forces the reference's chunking
MAX_RAYS_PER_LAUNCH = 1 << 20


def batchify(fn, chunk):
    """Constructs a version of 'fn' that applies to smaller batches (main.py:51-62)."""
    if chunk is None:
        return fn

    def ret(inputs):
        return torch.cat([fn(inputs[i:i + chunk]) for i in range(0, inputs.shape[0], chunk)], 0)

    return ret


def run_network(inputs, viewdirs, fn, embed_fn, embeddirs_fn, netchunk=1024 * 64):
    """Prepares inputs and applies network 'fn' (main.py:65-87)."""
    inputs_flat = torch.reshape(inputs, [-1, inputs.shape[-1]])
    embedded = embed_fn(inputs_flat)
    if viewdirs is not None:
        input_dirs = viewdirs[:, None].expand(inputs.shape)
        input_dirs_flat = torch.reshape(input_dirs, [-1, input_dirs.shape[-1]])
        embedded_dirs = embeddirs_fn(input_dirs_flat)
        embedded = torch.cat([embedded, embedded_dirs], -1)
    outputs_flat = batchify(fn, netchunk)(embedded)
    outputs = torch.reshape(outputs_flat, list(inputs.shape[:-1]) + [outputs_flat.shape[-1]])
    return outputs


def _z_vals(near, far, t_vals, lindisp, t_rand):
    """near/far [N,1] device tensors -> z_vals [N, S] (main.py:676-699)."""
    N = near.shape[0]
    S = t_vals.shape[0]
    dev = near.device
    nf = torch.cat([near.reshape(N, 1), far.reshape(N, 1)], -1).contiguous()
    z = torch.empty((N, S), dtype=torch.float32, device=dev)
    tr = _lib.as_f32_cuda(t_rand, dev, "t_rand") if t_rand is not None else None
    with torch.cuda.device(dev):
        _lib.call("r2l_z_vals", N, S, _lib.ptr(nf), _lib.ptr(nf[:, 1:]), 2, _lib.ptr(t_vals), int(bool(lindisp)),
                  _lib.ptr(tr), _lib.ptr(z), _lib.stream_ptr(dev))
    return z


def _points(rays_o, rays_d, z_vals):
    N, S = z_vals.shape
    pts = torch.empty((N, S, 3), dtype=torch.float32, device=z_vals.device)
    with torch.cuda.device(z_vals.device):
        _lib.call("r2l_points_from_rays", N, S, _lib.ptr(rays_o), rays_o.stride(0), _lib.ptr(rays_d),
                  rays_d.stride(0), _lib.ptr(z_vals), S, _lib.ptr(pts), _lib.stream_ptr(z_vals.device))
    return pts


def _unwrap(net):
    """nn.DataParallel / main.py's MyDataParallel -> the wrapped module.  The reference wraps its models whenever it is
    not --render_only (main.py:472-479) and create_data.py always does (:298-299); here one process drives one GPU,
    so the wrapper's scatter / replicate / gather is bypassed and the fused kernels see the real module."""
    return net.module if isinstance(net, torch.nn.DataParallel) else net


def _fused_ok(net, viewdirs):
    net = _unwrap(net)
    return (isinstance(net, NeRF) and viewdirs is not None and net.precision != "fp32"
            and net.supports_tensor_core_path())


_LINSPACE_CACHE = {}


def _host_linspace_on(dev, n):
    """torch.linspace(0, 1, n) computed on the HOST like the reference does (its fp32 values differ from a device
    linspace in the last bit) and uploaded once per (device, n)."""
    key = (dev.type, dev.index, int(n))
    t = _LINSPACE_CACHE.get(key)
    if t is None:
        t = torch.linspace(0., 1., steps=int(n)).to(dev)
        _LINSPACE_CACHE[key] = t
    return t


def render_rays(ray_batch, network_fn, network_query_fn, N_samples, retraw=False, lindisp=False, perturb=0.,
                N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0., verbose=False, pytest=False,
                return_depth=False, t_rand=None, u=None, noise0=None, noise1=None, return_debug=False):
    """Volumetric rendering of a batch of rays (main.py:624-756).  Returns the reference's dict."""
    rb = _lib.as_f32_cuda(ray_batch, name="ray_batch")
    dev = rb.device
    N_rays = rb.shape[0]
    rays_o, rays_d = rb[:, 0:3], rb[:, 3:6]
    viewdirs = rb[:, -3:] if rb.shape[-1] > 8 else None
    near, far = rb[:, 6:7], rb[:, 7:8]

    t_vals = _host_linspace_on(dev, N_samples)  # host linspace, uploaded once per (device, N) (main.py:676)
    if perturb > 0. and t_rand is None:
        if pytest:
            np.random.seed(0)
            t_rand = torch.Tensor(np.random.rand(N_rays, N_samples))
        else:
            t_rand = torch.rand((N_rays, N_samples))  # CPU generator (main.py:691)
    z_vals = _z_vals(near, far, t_vals, lindisp, t_rand if perturb > 0. else None)

    network_fn, network_fine = _unwrap(network_fn), _unwrap(network_fine)

    def query(z, net):
        if _fused_ok(net, viewdirs):
            return net.forward_samples(rays_o, rays_d, viewdirs, z)
        return network_query_fn(_points(rays_o, rays_d, z), viewdirs, net)

    # deterministic renders that do not hand `raw` out: network + raw2outputs in ONE library call (r2l_nerf_render;
    # in its fused mode the NeRF kernel composites the rays itself); same bits as the two-step route below
    one_call = (raw_noise_std == 0. and noise0 is None and noise1 is None and not retraw and not return_debug
                and rays_d.stride(-1) == 1)

    def shade(z, net, noise, S_noise, want_weights):
        if one_call and _fused_ok(net, viewdirs):
            rgb, disp, acc, w, depth = net.render_samples(rays_o, rays_d, viewdirs, z, white_bkgd, want_weights)
            return None, rgb, disp, acc, w, depth
        raw_ = query(z, net)
        if noise is None and raw_noise_std > 0.:
            noise = _host_noise((N_rays, S_noise), raw_noise_std, pytest)
        return (raw_,) + tuple(raw2outputs(raw_, z, rays_d, raw_noise_std, white_bkgd, pytest=pytest, noise=noise))

    raw, rgb_map, disp_map, acc_map, weights, depth_map = shade(z_vals, network_fn, noise0, N_samples,
                                                                N_importance > 0)
    debug = {}
    if N_importance > 0:
        rgb_map_0, disp_map_0, acc_map_0 = rgb_map, disp_map, acc_map
        u_sorted = None
        if u is None:
            if perturb == 0. and not pytest:
                u = _host_linspace_on(dev, N_importance)   # det=True table (helpers:292), cached on the device
                u_sorted = True
            else:
                u, _ = _make_u([N_rays], N_importance, det=(perturb == 0.), pytest=pytest)
        if hier_sample_supported(z_vals, weights, N_importance, u):
            # every deterministic render: mids, inverse-CDF sampling, sorted merge and z_std in one kernel
            z_all, z_std, z_samples, inds = hier_sample(z_vals, weights, N_importance, u, want_samples=return_debug,
                                                        want_inds=return_debug, u_sorted=u_sorted)
            if return_debug:
                debug.update(z_vals0=z_vals, weights0=weights, raw0=raw, inds=inds, z_samples=z_samples)
            z_vals = z_all
        else:
            z_vals_mid = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
            if return_debug:
                z_samples, inds = sample_pdf(z_vals_mid, weights[..., 1:-1], N_importance, u=u, return_inds=True)
                debug.update(z_vals0=z_vals, weights0=weights, raw0=raw, inds=inds, z_samples=z_samples)
            else:
                z_samples = sample_pdf(z_vals_mid, weights[..., 1:-1], N_importance, u=u)
            z_vals, z_std = merge_sorted(z_vals, z_samples, want_std=True)
        run_fn = network_fn if network_fine is None else network_fine
        raw, rgb_map, disp_map, acc_map, weights, depth_map = shade(z_vals, run_fn, noise1, N_samples + N_importance,
                                                                    False)

    ret = {'rgb_map': rgb_map, 'disp_map': disp_map, 'acc_map': acc_map}
    if return_depth:
        ret['depth_map'] = depth_map
    if retraw:
        ret['raw'] = raw
    if N_importance > 0:
        ret['rgb0'] = rgb_map_0
        ret['disp0'] = disp_map_0
        ret['acc0'] = acc_map_0
        ret['z_std'] = z_std
    if return_debug:
        debug.update(z_vals=z_vals, weights=weights)
        ret['debug'] = debug
    return ret


def render_rays_create_data(ray_batch, network_fn, network_query_fn, N_samples, **kwargs):
    """utils/create_data.py:405-544 flavour: same as render_rays plus 'depth_map'."""
    kwargs['return_depth'] = True
    return render_rays(ray_batch, network_fn, network_query_fn, N_samples, **kwargs)


def _is_stochastic(kwargs):
    return kwargs.get('perturb', 0.) > 0. or kwargs.get('raw_noise_std', 0.) > 0.


def batchify_rays(rays_flat, chunk=1024 * 32, **kwargs):
    """Render rays in minibatches (main.py:90-104).  Results do not depend on `chunk`; when nothing
    stochastic pins the reference's draw order the batch is rendered in as few launches as possible."""
    if not _is_stochastic(kwargs):
        chunk = max(int(chunk), MAX_RAYS_PER_LAUNCH)
    all_ret = {}
    for i in range(0, rays_flat.shape[0], chunk):
        ret = render_rays(rays_flat[i:i + chunk], **kwargs)
        for k in ret:
            all_ret.setdefault(k, []).append(ret[k])
    return {k: (v[0] if len(v) == 1 else torch.cat(v, 0)) for k, v in all_ret.items()}


def render(H, W, focal, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0., far=1., use_viewdirs=False,
           c2w_staticcam=None, **kwargs):
    """Render a full image (c2w) or a ray batch (main.py:107-186).
    Returns [rgb_map, disp_map, acc_map, extras-dict] with the reference's shapes."""
    if c2w is not None:
        rays_o, rays_d = get_rays(H, W, focal, c2w)
    else:
        rays_o, rays_d = rays
        rays_o = _lib.as_f32_cuda(rays_o, name="rays_o")
        rays_d = _lib.as_f32_cuda(rays_d, rays_o.device, "rays_d")
    viewdirs = None
    if use_viewdirs:
        viewdirs = rays_d
        if c2w_staticcam is not None:
            rays_o, rays_d = get_rays(H, W, focal, c2w_staticcam)
        viewdirs = normalize_dirs(viewdirs)  # [N, 3]
    sh = rays_d.shape
    if ndc:
        rays_o, rays_d = ndc_rays(H, W, focal, 1., rays_o, rays_d)
    rays_o = torch.reshape(rays_o, [-1, 3]).float()
    rays_d = torch.reshape(rays_d, [-1, 3]).float()
    near_t, far_t = near * torch.ones_like(rays_d[..., :1]), far * torch.ones_like(rays_d[..., :1])
    rays_flat = torch.cat([rays_o, rays_d, near_t, far_t], -1)
    if use_viewdirs:
        rays_flat = torch.cat([rays_flat, viewdirs], -1)
    all_ret = batchify_rays(rays_flat, chunk, **kwargs)
    for k in all_ret:
        if k == 'debug':
            continue
        k_sh = list(sh[:-1]) + list(all_ret[k].shape[1:])
        all_ret[k] = torch.reshape(all_ret[k], k_sh)
    k_extract = ['rgb_map', 'disp_map', 'acc_map']
    ret_list = [all_ret[k] for k in k_extract]
    ret_dict = {k: all_ret[k] for k in all_ret if k not in k_extract}
    return ret_list + [ret_dict]


def render_r2l(model, point_sampler, c2w, positional_embedder=None):
    """R2L branch of render_path (main.py:285-325): one frame = one un-chunked forward.
    Uses the fused encode+MLP kernel when the model supports it; returns rgb [H*W, 3]."""
    c = c2w if isinstance(c2w, torch.Tensor) else torch.as_tensor(c2w)
    # a stack of poses [P, 3, 4] renders P frames with ONE launch -> [P*H*W, 3]
    if model.precision != "fp32" and model.supports_tensor_core_path():
        L = positional_embedder.L if positional_embedder is not None else 10
        if L == 10 and point_sampler.n_sample * 63 == model.input_dim:
            return model.render_poses(point_sampler, c)      # rays generated inside the fused kernel
    pts = point_sampler.sample_test_batch(c, lazy=False) if c.dim() == 3 else point_sampler._sample(c)
    if positional_embedder is None:
        raise ValueError("positional_embedder is required for the non-fused path")
    return model(positional_embedder(pts))


class GraphedR2L:
    """The R2L frame (ray generation + encode + MLP: one kernel) captured ONCE into a CUDA graph and replayed per
    pose: one graph launch per frame (or per stack of `n_poses` frames) instead of the kernel launch plus the
    Python glue around it — what matters when a frame is sharded over many GPUs and takes a few hundred
    microseconds.  `g(c2w)` copies the pose(s) into the graph's static input and replays; the returned tensor is the
    graph's static output [n_poses*H*W, 3] (overwritten by the next call)."""

    def __init__(self, model, point_sampler, n_poses=1, positional_embedder=None, rows=None, frame=None):
        dev = point_sampler.z_vals.device
        self.n_poses = int(n_poses)
        self.rows = rows
        self.frame = frame      # sharding.PeerFrame: the MLP kernel stores its tiles into every GPU's frame buffer
        if frame is not None:
            if self.n_poses != 1:
                raise ValueError("a PeerFrame holds one frame")
            self.rows = rows = (frame.row0, frame.row1)
        self.c2w = torch.zeros((self.n_poses, 3, 4), dtype=torch.float32, device=dev)
        self.c2w[:, 0, 0] = self.c2w[:, 1, 1] = self.c2w[:, 2, 2] = 1.
        model.packed_handle() if getattr(model, "precision", "fp32") != "fp32" else None

        def run():
            if self.rows is None:
                return render_r2l(model, point_sampler, self.c2w if self.n_poses > 1 else self.c2w[0],
                                  positional_embedder)
            # one rank's ray block; with a frame the tiles go to every GPU (caller: frame.publish() after the replay)
            return model.render_poses(point_sampler, self.c2w, rows=self.rows, frame=self.frame)

        with torch.no_grad():
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):      # warm-up outside capture (lazy initialisation, allocator)
                for _ in range(2):
                    run()
            torch.cuda.current_stream(dev).wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            k0 = _lib.kernel_launches()
            with torch.cuda.graph(self.graph):
                self.out = run()
            self.kernels_per_replay = _lib.kernel_launches() - k0   # library kernels recorded in the graph

    def __call__(self, c2w):
        c = c2w if isinstance(c2w, torch.Tensor) else torch.as_tensor(c2w)
        c = c.reshape(-1, c.shape[-2], 4)[:, :3, :4]
        if c.shape[0] != self.n_poses:
            raise ValueError(f"this graph renders {self.n_poses} pose(s) per replay, got {c.shape[0]}")
        self.c2w.copy_(c, non_blocking=True)
        self.graph.replay()
        return self.out


def render_path(render_poses, hwf, chunk, render_kwargs, gt_imgs=None, savedir=None, render_factor=0,
                model_name="nerf", point_sampler=None, positional_embedder=None, learn_depth=False):
    """The test loop around the hot path (main.py:189-400): render every pose, stack the frames, and — with ground
    truth — the reference's per-frame metrics, computed on the device (`metrics.image_errors`: one error/MSE pass
    and one SSIM pass for the whole stack instead of a host round trip per frame).

    model_name == 'nerf' renders through `render` (main.py:283-290); anything else is the R2L branch
    `model(positional_embedder(point_sampler.sample_test(c2w)))` (main.py:292-321).  Returns (rgbs [N,H,W,3],
    disps, misc) with misc['test_loss', 'test_psnr', 'test_psnr_v2', 'test_ssim', 'test_flip', 'errors'] as in
    main.py:384-394.
    misc['test_flip'] is the reference's FLIP stage (main.py:370-379) on the device (csrc/flip.cu).
    Out of scope: PNG writing (`savedir` must be None), LPIPS (a pretrained network; misc has no
    'test_lpips'), `given_render_path_rays`."""
    from . import metrics
    if savedir is not None:
        raise NotImplementedError("render_path: image files are written by the caller (PNG I/O is out of scope)")
    H, W, focal = hwf
    if render_factor != 0:
        H, W, focal = int(H / render_factor), int(W / render_factor), focal / render_factor
    net = render_kwargs["network_fn"]
    was_training = getattr(net, "training", False)
    if hasattr(net, "eval"):
        net.eval()
    rgbs, disps = [], []
    with torch.no_grad():
        for c2w in render_poses:
            c2w = _lib.as_f32_cuda(c2w, name="c2w")[:3, :4]
            if model_name in ("nerf",):
                rgb, disp, acc, _ = render(H, W, focal, chunk=chunk, c2w=c2w, **render_kwargs)
            else:
                if point_sampler is None:
                    raise ValueError("render_path: the R2L branch needs point_sampler")
                out = render_r2l(net, point_sampler, c2w, positional_embedder)
                rgb = out[:, :3] if learn_depth else out
                rgb = rgb.reshape(H, W, 3)
                disp = rgb   # placeholder, as in the reference (main.py:321)
            rgbs.append(rgb)
            disps.append(disp)
        rgbs = torch.stack(rgbs, 0) if rgbs else torch.zeros((0, H, W, 3), device=_dev())
        disps = torch.stack(disps, 0) if disps else torch.zeros((0, H, W), device=_dev())
        misc = {}
        if gt_imgs is not None and rgbs.shape[0] > 0:
            gt = _lib.as_f32_cuda(gt_imgs, rgbs.device, "gt_imgs")[:, :H, :W, :].contiguous()
            m = metrics.image_errors(rgbs, gt)
            test_loss = m["mse"].double().mean().to(torch.float32)   # equal-size frames: mean of means == global mean
            misc["test_loss"] = test_loss
            misc["test_psnr"] = metrics.mse2psnr(test_loss)
            misc["test_psnr_v2"] = m["psnr"].mean()
            misc["test_ssim"] = m["ssim"].mean()
            # FLIP of the stacks rescaled to [-1, 1] (each by its own min / max), exactly what main.py:362-379 computes:
            # it reuses the LPIPS inputs; the sRGB decode inside FLIP then clamps the negative half to 0
            lim = torch.stack(torch.aminmax(rgbs) + torch.aminmax(gt)).double().tolist()
            sc = [(2. / (hi - lo), -lo * 2. / (hi - lo) - 1.) for lo, hi in (lim[0:2], lim[2:4])]
            _, flips = metrics.flip_map(rgbs, gt, scale=sc, want_map=False)
            misc["test_flip"] = flips.mean()
            misc["errors"] = m["errors"]
    if was_training and hasattr(net, "train"):
        net.train()
    return rgbs, disps, misc

