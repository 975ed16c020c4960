"""B200-native stand-in for the hot-path functions of the reference's
utils/run_nerf_raybased_helpers.py: same names, signatures and return conventions,
bodies run as hand-written sm_100a CUDA through the C ABI (include/r2l_b200.h).

  get_rays      helpers:231-257      ndc_rays     helpers:260-279
  Embedder      helpers:24-56        get_embedder helpers:59-74
  raw2outputs   helpers:77-144       sample_pdf   helpers:283-330

Outputs live on the CUDA device of the inputs.  Inputs given on the host (the reference's
render_rays calls sample_pdf with `.cpu()` tensors, main.py:723-727) are uploaded, computed on
the GPU and the result is returned on the host again, so call sites keep working unchanged.
"""

import numpy as np
import torch
import torch.nn as nn

from . import _lib

to8b = lambda x: (255 * np.clip(x if isinstance(x, np.ndarray) else x.detach().cpu().numpy(), 0, 1)).astype(np.uint8)
img2mse = lambda x, y: torch.mean((x - y)**2)
mse2psnr = lambda x: -10. * torch.log(x) / torch.log(torch.tensor([10.], device=x.device))


def _dev():
    _lib.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def _c2w34(c2w, device):
    c = _lib.as_f32_cuda(c2w, device if isinstance(device, torch.device) else None, "c2w")
    if c.dim() != 2 or c.shape[0] < 3 or c.shape[1] < 4:
        raise ValueError(f"c2w must be [>=3, >=4], got {tuple(c.shape)}")
    return c[:3, :4].contiguous()


# ----------------------------------------------------------------------------- rays
def get_rays(H, W, focal, c2w, trans_origin='', focal_scale=1):
    """Pixel rays of a pinhole camera: returns rays_o, rays_d of shape [H, W, 3] (helpers:231-257)."""
    if trans_origin:
        raise NotImplementedError("trans_origin is outside the accelerated hot path (unused by the render configs)")
    focal = float(focal) * focal_scale
    dev = c2w.device if isinstance(c2w, torch.Tensor) and c2w.is_cuda else _dev()
    c = _c2w34(c2w, dev)
    H, W = int(H), int(W)
    rays_o = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
    rays_d = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.call("r2l_get_rays", H, W, focal, _lib.ptr(c), _lib.ptr(rays_o), _lib.ptr(rays_d), _lib.stream_ptr(dev))
    return rays_o, rays_d


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """Normalised-device-coordinate warp for forward-facing scenes (helpers:260-279)."""
    ro = _lib.as_f32_cuda(rays_o, name="rays_o")
    rd = _lib.as_f32_cuda(rays_d, ro.device, "rays_d")
    if ro.shape != rd.shape or ro.shape[-1] != 3:
        raise ValueError("rays_o / rays_d must both be [..., 3]")
    oo, od = torch.empty_like(ro), torch.empty_like(rd)
    n = ro.numel() // 3
    with torch.cuda.device(ro.device):
        _lib.call("r2l_ndc_rays", n, int(H), int(W), float(focal), float(near), _lib.ptr(ro), _lib.ptr(rd),
                  _lib.ptr(oo), _lib.ptr(od), _lib.stream_ptr(ro.device))
    return oo, od


def normalize_dirs(rays_d):
    """viewdirs = rays_d / |rays_d| (main.py:148-157) -> [N, 3]."""
    d = _lib.as_f32_cuda(rays_d, name="rays_d").reshape(-1, 3)
    out = torch.empty_like(d)
    with torch.cuda.device(d.device):
        _lib.call("r2l_normalize_dirs", d.shape[0], _lib.ptr(d), 3, _lib.ptr(out), _lib.stream_ptr(d.device))
    return out


# ----------------------------------------------------------------------------- positional encoding
def _embed(x, L, include_input, layout):
    xin = _lib.as_f32_cuda(x, name="inputs")
    D = xin.shape[-1]
    rows = xin.numel() // D if D > 0 else 0
    out_dim = D * (2 * L + (1 if include_input else 0))
    out = torch.empty(tuple(xin.shape[:-1]) + (out_dim,), dtype=torch.float32, device=xin.device)
    if rows > 0:
        with torch.cuda.device(xin.device):
            _lib.call("r2l_embed", rows, D, L, int(bool(include_input)), layout, _lib.ptr(xin), _lib.ptr(out),
                      _lib.stream_ptr(xin.device))
    return out


class Embedder:
    """Positional encoding (section 5.1), NeRF feature order (helpers:24-56)."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs
        self.create_embedding_fn()

    def create_embedding_fn(self):
        kw = self.kwargs
        d = kw['input_dims']
        if not kw.get('log_sampling', True):
            raise NotImplementedError("only log_sampling=True frequency bands are accelerated")
        fns = kw.get('periodic_fns', [torch.sin, torch.cos])
        if list(fns) != [torch.sin, torch.cos]:
            raise NotImplementedError("periodic_fns must be [torch.sin, torch.cos]")
        if kw['num_freqs'] > 0 and kw['max_freq_log2'] != kw['num_freqs'] - 1:
            raise NotImplementedError("frequency bands must be 2**[0..num_freqs-1]")
        self.include_input = bool(kw['include_input'])
        self.num_freqs = int(kw['num_freqs'])
        self.out_dim = d * (2 * self.num_freqs + (1 if self.include_input else 0))

    def embed(self, inputs):
        return _embed(inputs, self.num_freqs, self.include_input, 0)


def get_embedder(multires, i=0):
    if i == -1:
        return nn.Identity(), 3
    embedder_obj = Embedder(include_input=True, input_dims=3, max_freq_log2=multires - 1, num_freqs=multires,
                            log_sampling=True, periodic_fns=[torch.sin, torch.cos])
    embed = lambda x, eo=embedder_obj: eo.embed(x)
    embed.multires = multires  # lets the fused renderer recognise the standard encoders
    return embed, embedder_obj.out_dim


# ----------------------------------------------------------------------------- compositing
def _host_noise(shape, raw_noise_std, pytest):
    """Sigma noise drawn exactly like the reference: CPU generator (main.py:589-596)."""
    if pytest:
        np.random.seed(0)
        noise = torch.Tensor(np.random.rand(*list(shape)) * raw_noise_std)
    else:
        noise = torch.randn(shape) * raw_noise_std
    return noise


def raw2outputs(raw, z_vals, rays_d, raw_noise_std=0, white_bkgd=False, pytest=False, verbose=False, noise=None,
                **_ignored):
    """Alpha compositing (main.py:556-621).  Returns rgb_map, disp_map, acc_map, weights, depth_map."""
    on_host = isinstance(raw, torch.Tensor) and not raw.is_cuda
    r = _lib.as_f32_cuda(raw, name="raw")
    dev = r.device
    if r.dim() != 3 or r.shape[-1] != 4:
        raise ValueError(f"raw must be [N_rays, N_samples, 4], got {tuple(r.shape)}")
    N, S = r.shape[0], r.shape[1]
    z = _lib.as_f32_cuda(z_vals, dev, "z_vals")
    if tuple(z.shape) != (N, S):
        z = z.expand(N, S).contiguous()
    d = _lib.as_f32_cuda(rays_d, dev, "rays_d").reshape(N, 3)
    if noise is None and raw_noise_std > 0.:
        noise = _host_noise((N, S), raw_noise_std, pytest)
    nz = _lib.as_f32_cuda(noise, dev, "noise").reshape(N, S) if noise is not None else None
    rgb_map = torch.empty((N, 3), dtype=torch.float32, device=dev)
    disp_map = torch.empty((N,), dtype=torch.float32, device=dev)
    acc_map = torch.empty((N,), dtype=torch.float32, device=dev)
    weights = torch.empty((N, S), dtype=torch.float32, device=dev)
    depth_map = torch.empty((N,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.call("r2l_raw2outputs", N, S, _lib.ptr(r), _lib.ptr(z), _lib.ptr(d), 3, _lib.ptr(nz),
                  int(bool(white_bkgd)), _lib.ptr(rgb_map), _lib.ptr(disp_map), _lib.ptr(acc_map), _lib.ptr(weights),
                  _lib.ptr(depth_map), _lib.stream_ptr(dev))
    outs = (rgb_map, disp_map, acc_map, weights, depth_map)
    return tuple(o.cpu() for o in outs) if on_host else outs


# ----------------------------------------------------------------------------- hierarchical sampling
def _make_u(shape_prefix, N_samples, det, pytest):
    """The uniform variates, built on the host exactly like the reference (helpers:291-307)."""
    if det:
        u = torch.linspace(0., 1., steps=N_samples)
        per_ray = False
    else:
        u = torch.rand(list(shape_prefix) + [N_samples])
        per_ray = True
    if pytest:
        np.random.seed(0)
        if det:
            u = torch.Tensor(np.linspace(0., 1., N_samples))
            per_ray = False
        else:
            u = torch.Tensor(np.random.rand(*(list(shape_prefix) + [N_samples])))
            per_ray = True
    return u, per_ray


def sample_pdf(bins, weights, N_samples, det=False, pytest=False, u=None, return_inds=False, u_sorted=None):
    """Inverse-CDF sampling of the coarse weights (helpers:283-330).

    bins [N, nb], weights [N, nb-1] -> samples [N, N_samples]; searchsorted indices are bit-exact with
    the reference's CPU path.  `u` may be injected ([N_samples] or [N, N_samples]).  A shared table that is ascending
    (det=True's linspace; `u_sorted`: None = look at host tables, a device table is taken at the caller's word) runs
    the search-free kernel — same samples and indices.
    """
    on_host = isinstance(bins, torch.Tensor) and not bins.is_cuda
    dev = bins.device if isinstance(bins, torch.Tensor) and bins.is_cuda else _dev()
    b = bins if (isinstance(bins, torch.Tensor) and bins.is_cuda and bins.dtype == torch.float32
                 and bins.stride(-1) == 1 and bins.dim() == 2) else _lib.as_f32_cuda(bins, dev, "bins")
    w = weights if (isinstance(weights, torch.Tensor) and weights.is_cuda and weights.dtype == torch.float32
                    and weights.stride(-1) == 1 and weights.dim() == 2
                    and weights.device == dev) else _lib.as_f32_cuda(weights, dev, "weights")
    if b.requires_grad or w.requires_grad:
        b, w = b.detach(), w.detach()
    if b.dim() != 2 or w.dim() != 2 or w.shape[0] != b.shape[0] or w.shape[1] != b.shape[1] - 1:
        raise ValueError(f"bins [N, nb] / weights [N, nb-1] expected, got {tuple(b.shape)} / {tuple(w.shape)}")
    N, nb = b.shape
    if u is None:
        u, per_ray = _make_u([N], N_samples, det, pytest)
        if det and not pytest and u_sorted is None:
            u_sorted = True
    else:
        per_ray = (u.dim() == 2)
    if not per_ray and u_sorted is None:
        u_sorted = isinstance(u, torch.Tensor) and (not u.is_cuda) and bool((u[1:] >= u[:-1]).all())
    u = _lib.as_f32_cuda(u, dev, "u")
    if per_ray and tuple(u.shape) != (N, N_samples):
        raise ValueError("per-ray u must be [N, N_samples]")
    samples = torch.empty((N, N_samples), dtype=torch.float32, device=dev)
    inds = torch.empty((N, N_samples), dtype=torch.int64, device=dev) if return_inds else None
    with torch.cuda.device(dev):
        _lib.call("r2l_sample_pdf", N, nb, int(N_samples), _lib.ptr(b), b.stride(0), _lib.ptr(w), w.stride(0),
                  _lib.ptr(u), 1 if per_ray else (2 if u_sorted else 0), _lib.ptr(samples), _lib.ptr(inds),
                  _lib.stream_ptr(dev))
    if on_host:
        samples = samples.cpu()
        inds = inds.cpu() if inds is not None else None
    return (samples, inds) if return_inds else samples


def hier_sample_supported(z_vals, weights, N_importance, u):
    """True when the fused kernel applies: 64 coarse samples, 64 / 128 fine samples, `u` either ONE table [Ni] shared
    by all rays (det=True: every rendering configuration) or per-ray variates [N, Ni] (stochastic renders)."""
    if not (isinstance(z_vals, torch.Tensor) and z_vals.is_cuda and z_vals.dim() == 2 and z_vals.shape[1] == 64
            and z_vals.dtype == torch.float32 and weights.shape == z_vals.shape and weights.is_cuda
            and weights.dtype == torch.float32 and N_importance in (64, 128) and isinstance(u, torch.Tensor)):
        return False
    if u.dim() == 1:
        return u.shape[0] == N_importance
    return u.dim() == 2 and tuple(u.shape) == (z_vals.shape[0], N_importance)


def hier_sample(z_vals, weights, N_importance, u, want_samples=False, want_inds=False, u_sorted=None):
    """Fused main.py:720-733,750: z_vals_mid -> sample_pdf(weights[..., 1:-1], u) -> sort(cat[z_vals, z_samples]) and
    std(z_samples), one launch.  Returns (z_all [N, 64+Ni], z_std [N], z_samples | None, inds | None); bit-identical to
    sample_pdf + merge_sorted (tested).  `u_sorted`: a shared table [Ni] that is ascending (det=True's linspace) runs
    the search-free kernel; None = look (host tables only — a device table is not read back, the caller says so)."""
    if u.dim() == 1 and u_sorted is None:
        u_sorted = (not u.is_cuda) and bool((u[1:] >= u[:-1]).all())
    z = z_vals.contiguous()
    w = weights.contiguous()
    dev = z.device
    N = z.shape[0]
    ud = _lib.as_f32_cuda(u, dev, "u")
    out = torch.empty((N, 64 + N_importance), dtype=torch.float32, device=dev)
    z_std = torch.empty((N,), dtype=torch.float32, device=dev)
    samples = torch.empty((N, N_importance), dtype=torch.float32, device=dev) if want_samples else None
    inds = torch.empty((N, N_importance), dtype=torch.int64, device=dev) if want_inds else None
    with torch.cuda.device(dev):
        _lib.call("r2l_hier_sample", N, 64, int(N_importance), _lib.ptr(z), _lib.ptr(w), _lib.ptr(ud),
                  1 if ud.dim() == 2 else (2 if u_sorted else 0), _lib.ptr(out), _lib.ptr(z_std), _lib.ptr(samples),
                  _lib.ptr(inds),
                  _lib.stream_ptr(dev))
    return out, z_std, samples, inds


def merge_sorted(z_vals, z_samples, want_std=False):
    """z = sort(cat[z_vals, z_samples], -1) (main.py:730-732); optionally std(z_samples) (main.py:750)."""
    za = _lib.as_f32_cuda(z_vals, name="z_vals")
    zb = _lib.as_f32_cuda(z_samples, za.device, "z_samples")
    N, na, nbv = za.shape[0], za.shape[1], zb.shape[1]
    out = torch.empty((N, na + nbv), dtype=torch.float32, device=za.device)
    z_std = torch.empty((N,), dtype=torch.float32, device=za.device) if want_std else None
    with torch.cuda.device(za.device):
        _lib.call("r2l_merge_sorted", N, na, nbv, _lib.ptr(za), _lib.ptr(zb), _lib.ptr(out), _lib.ptr(z_std),
                  _lib.stream_ptr(za.device))
    return (out, z_std) if want_std else out


# ----------------------------------------------------------------------------- the rest of the module's call surface
def get_rays_np(H, W, focal, c2w):
    """helpers:428-441: the numpy twin of get_rays (dataset preprocessing).  Computed by the same device kernel and
    copied back, so both forms agree; returns numpy arrays like the reference."""
    c = c2w if isinstance(c2w, torch.Tensor) else torch.as_tensor(np.asarray(c2w, dtype=np.float32))
    f = float(focal.item() if hasattr(focal, "item") else focal)
    ro, rd = get_rays(H, W, f, c)
    return ro.cpu().numpy(), rd.cpu().numpy()


to_array = lambda x: x if isinstance(x, np.ndarray) else x.data.cpu().numpy()   # helpers:16


def to_tensor(x):
    """helpers:14-15: onto the module's device (here: the current CUDA device)."""
    return x.to(_dev()) if isinstance(x, torch.Tensor) else torch.Tensor(x).to(_dev())


def undataparallel(input):
    """helpers:408-425."""
    from .compat import undataparallel as _u
    return _u(input)


def load_weights_v2(model, ckpt, key):
    """helpers:362-382: load ckpt[key] when model and weights agree on the DataParallel `module.` prefix."""
    model_dp = any(name.startswith('module.') for name, _ in model.named_modules())
    state_dict = ckpt[key]
    weights_dp = any(k.startswith('module.') for k in state_dict)
    if model_dp == weights_dp:
        model.load_state_dict(state_dict)
    else:
        raise NotImplementedError


def load_weights(model, ckpt_path, key):
    """helpers:347-359 (without smilelogging's path check): strips `module.` and loads ckpt[key]."""
    from .compat import load_checkpoint
    ckpt = load_checkpoint(ckpt_path)
    model.load_state_dict(undataparallel(dict(ckpt[key])) if not isinstance(ckpt[key], torch.nn.Module) else
                          ckpt[key].state_dict())
    return ckpt_path, ckpt


def __getattr__(name):
    # batchify / run_network live in render.py (their main.py twins, :51-104); the reference's helpers module has
    # them too (helpers:147-183), so `from utils.run_nerf_raybased_helpers import run_network` keeps working
    if name in ("batchify", "run_network"):
        from . import render
        return getattr(render, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")

