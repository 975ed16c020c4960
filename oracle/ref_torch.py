"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's per-ray rendering path.

This is the ORACLE the CUDA path is checked against.  It is NOT part of the product: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import it; the
product (efficient-nerf_b200/) never does and fails loudly without its CUDA library.

Every function restates one reference function in plain functional PyTorch on CPU tensors and cites
the file:line it follows (paths relative to the reference repo).  The arithmetic of the reference
lives in PyTorch's ATen CPU kernels (pinned torch==1.9.0, requirements.txt:18; this image runs
torch 2.11.0), so the oracle deliberately uses the same eager fp32 tensor ops in the same order —
that is what makes it bit-comparable with the reference (see oracle/make_golden.py, which imports the
real reference from /root/reference, checks oracle == reference bit-for-bit and writes tests/golden/).
oracle/sample_pdf_np.py additionally restates sample_pdf at the level of individual roundings.

Parity pinning: the reference ships no tests or golden vectors, so the fixtures under
tests/golden/ are outputs of the reference's own code run in the build container.
"""
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------------- synthetic inputs
def pose_spherical(theta, phi, radius):
    """Camera-to-world matrix on a sphere (dataset/load_blender.py:10-28)."""
    def trans_t(t):
        return torch.Tensor([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, t], [0, 0, 0, 1]]).float()

    def rot_phi(a):
        return torch.Tensor([[1, 0, 0, 0], [0, np.cos(a), -np.sin(a), 0], [0, np.sin(a), np.cos(a), 0],
                             [0, 0, 0, 1]]).float()

    def rot_theta(a):
        return torch.Tensor([[np.cos(a), 0, -np.sin(a), 0], [0, 1, 0, 0], [np.sin(a), 0, np.cos(a), 0],
                             [0, 0, 0, 1]]).float()

    m = trans_t(radius)
    m = rot_phi(phi / 180. * np.pi) @ m
    m = rot_theta(theta / 180. * np.pi) @ m
    m = torch.Tensor([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]) @ m
    return m


LEGO = dict(H=400, W=400, focal=555.5555155968841, near=2., far=6.)       # main.py:927-931, half_res
LEGO_800 = dict(H=800, W=800, focal=1111.1110311937682, near=2., far=6.)
FERN = dict(H=378, W=504, focal=407.5658, near=0., far=1.)                # load_llff.py:131,446; NDC (main.py:914-919)


def nerf_state_dicts(seed=0):
    """Random-init coarse and fine NeRF weights exactly as create_nerf builds them (main.py:426-445):
    torch.manual_seed(seed); NeRF(8,256,63,27,5,[4],True) twice, default nn.Linear init, in the
    reference's module construction order (model/nerf_raybased.py:357-375)."""
    torch.manual_seed(seed)

    def one():
        sd = {}
        dims = [(63, 256)] + [(256, 256)] * 4 + [(319, 256)] + [(256, 256)] * 2
        for i, (k, n) in enumerate(dims):
            l = nn.Linear(k, n)
            sd[f'pts_linears.{i}.weight'], sd[f'pts_linears.{i}.bias'] = l.weight.detach(), l.bias.detach()
        for name, (k, n) in (('views_linears.0', (283, 128)), ('feature_linear', (256, 256)),
                             ('alpha_linear', (256, 1)), ('rgb_linear', (128, 3))):
            l = nn.Linear(k, n)
            sd[f'{name}.weight'], sd[f'{name}.bias'] = l.weight.detach(), l.bias.detach()
        return sd

    return one(), one()


def r2l_args(netdepth=88, netwidth=256, use_residual=True):
    """The namespace NeRF_v3_2 reads (model/nerf_raybased.py:486-535,543) for the README.md:51 command."""
    return SimpleNamespace(netdepth=netdepth, netwidth=netwidth, layerwise_netwidths='', act='relu', linear_tail=False,
                           use_residual=use_residual,
                           trial=SimpleNamespace(inact='relu', outact='none', body_arch='resmlp', n_block=-1,
                                                 res_scale=1., n_learnable=2))


def r2l_state_dict(seed=0, netdepth=88, netwidth=256, input_dim=1008):
    """Random-init NeRF_v3_2 (resmlp) weights with the reference's RNG consumption: head, then the
    discarded plain-MLP body (model:503-505), then the ResMLP blocks (model:511-524), then the tail."""
    torch.manual_seed(seed)
    sd = {}
    W = netwidth
    l = nn.Linear(input_dim, W)
    sd['head.0.weight'], sd['head.0.bias'] = l.weight.detach(), l.bias.detach()
    for _ in range(1, netdepth - 1):
        nn.Linear(W, W)  # built and thrown away by the reference
    for b in range((netdepth - 2) // 2):
        for j in (0, 2):
            l = nn.Linear(W, W)
            sd[f'body.{b}.body.{j}.weight'], sd[f'body.{b}.body.{j}.bias'] = l.weight.detach(), l.bias.detach()
    l = nn.Linear(W, 3)
    sd['tail.0.weight'], sd['tail.0.bias'] = l.weight.detach(), l.bias.detach()
    return sd


# --------------------------------------------------------------------------------- rays
def get_rays(H, W, focal, c2w):
    """utils/run_nerf_raybased_helpers.py:231-257 (trans_origin='' branch)."""
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing='ij')
    i, j = i.t(), j.t()
    dirs = torch.stack([(i - W * .5) / focal, -(j - H * .5) / focal, -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs.unsqueeze(-2) * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """utils/run_nerf_raybased_helpers.py:260-279."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    o0 = -1. / (W / (2. * focal)) * rays_o[..., 0] / rays_o[..., 2]
    o1 = -1. / (H / (2. * focal)) * rays_o[..., 1] / rays_o[..., 2]
    o2 = 1. + 2. * near / rays_o[..., 2]
    d0 = -1. / (W / (2. * focal)) * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2])
    d1 = -1. / (H / (2. * focal)) * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2])
    d2 = -2. * near / rays_o[..., 2]
    return torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)


def pack_rays(rays_o, rays_d, near, far, H=None, W=None, focal=None, ndc=False, use_viewdirs=True):
    """Ray batch [N, 8|11] = (o, d, near, far, viewdirs) as `render` builds it (main.py:143-175)."""
    viewdirs = None
    if use_viewdirs:
        viewdirs = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
        viewdirs = torch.reshape(viewdirs, [-1, 3]).float()
    if ndc:
        rays_o, rays_d = ndc_rays(H, W, focal, 1., rays_o, rays_d)
    rays_o = torch.reshape(rays_o, [-1, 3]).float()
    rays_d = torch.reshape(rays_d, [-1, 3]).float()
    n, f = near * torch.ones_like(rays_d[..., :1]), far * torch.ones_like(rays_d[..., :1])
    rays = torch.cat([rays_o, rays_d, n, f], -1)
    if use_viewdirs:
        rays = torch.cat([rays, viewdirs], -1)
    return rays


# --------------------------------------------------------------------------------- encodings
def embed_nerf(x, L):
    """Embedder.embed with include_input, log-sampled bands 2^0..2^(L-1) (helpers:24-56)."""
    out = [x]
    for freq in 2.**torch.linspace(0., L - 1, steps=L):
        out += [torch.sin(x * freq), torch.cos(x * freq)]
    return torch.cat(out, -1)


def embed_r2l(x, L, include_input=True):
    """PositionalEmbedder.__call__ (model/nerf_raybased.py:198-208)."""
    w = 2**torch.linspace(0, L - 1, steps=L)
    y = x[..., None] * w
    y = torch.cat([torch.sin(y), torch.cos(y)], dim=-1)
    if include_input:
        y = torch.cat([y, x.unsqueeze(dim=-1)], dim=-1)
    return y.view(y.shape[0], -1)


def point_sample(H, W, focal, n_sample, near, far, c2w):
    """PointSampler.__init__ + sample_test (model/nerf_raybased.py:76-102) -> [H*W, n_sample*3]."""
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing='ij')
    i, j = i.t(), j.t()
    dirs = torch.stack([(i - W * .5) / focal, -(j - H * .5) / focal, -torch.ones_like(i)], dim=-1)
    t_vals = torch.linspace(0., 1., steps=n_sample)
    z_vals = near * (1 - t_vals) + far * (t_vals)
    z_test = z_vals[None, :].expand(H * W, n_sample)
    rays_d = torch.sum(dirs.unsqueeze(dim=-2) * c2w[:3, :3], dim=-1).view(-1, 3)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_test[..., :, None]
    return pts.view(pts.shape[0], -1)


# --------------------------------------------------------------------------------- networks
def nerf_forward(sd, x, skips=(4,), D=8):
    """NeRF.forward with use_viewdirs=True (model/nerf_raybased.py:377-401); x [M, 63+27] -> [M, 4]."""
    input_pts, input_views = torch.split(x, [63, 27], dim=-1)
    h = input_pts
    for i in range(D):
        h = F.relu(F.linear(h, sd[f'pts_linears.{i}.weight'], sd[f'pts_linears.{i}.bias']))
        if i in skips:
            h = torch.cat([input_pts, h], -1)
    alpha = F.linear(h, sd['alpha_linear.weight'], sd['alpha_linear.bias'])
    feature = F.linear(h, sd['feature_linear.weight'], sd['feature_linear.bias'])
    h = torch.cat([feature, input_views], -1)
    h = F.relu(F.linear(h, sd['views_linears.0.weight'], sd['views_linears.0.bias']))
    rgb = F.linear(h, sd['rgb_linear.weight'], sd['rgb_linear.bias'])
    return torch.cat([rgb, alpha], -1)


def run_network(pts, viewdirs, sd, netchunk=1024 * 64):
    """run_network + batchify (main.py:51-87) with the L=10 / L=4 encoders of create_nerf (main.py:413-418)."""
    flat = torch.reshape(pts, [-1, pts.shape[-1]])
    emb = embed_nerf(flat, 10)
    dirs = viewdirs[:, None].expand(pts.shape)
    emb = torch.cat([emb, embed_nerf(torch.reshape(dirs, [-1, 3]), 4)], -1)
    out = torch.cat([nerf_forward(sd, emb[i:i + netchunk]) for i in range(0, emb.shape[0], netchunk)], 0)
    return torch.reshape(out, list(pts.shape[:-1]) + [out.shape[-1]])


def r2l_forward(sd, x, n_blocks=43, use_residual=True, res_scale=1.):
    """NeRF_v3_2.forward with the resmlp body (model/nerf_raybased.py:539-544, 461-465)."""
    x = F.relu(F.linear(x, sd['head.0.weight'], sd['head.0.bias']))
    y = x
    for b in range(n_blocks):
        h = F.relu(F.linear(y, sd[f'body.{b}.body.0.weight'], sd[f'body.{b}.body.0.bias']))
        y = F.linear(h, sd[f'body.{b}.body.2.weight'], sd[f'body.{b}.body.2.bias']).mul(res_scale) + y
    x = y + x if use_residual else y
    return torch.sigmoid(F.linear(x, sd['tail.0.weight'], sd['tail.0.bias']))


# --------------------------------------------------------------------------------- compositing / sampling
def raw2outputs(raw, z_vals, rays_d, noise=None, white_bkgd=False):
    """main.py:556-621 (noise: the already-scaled sigma noise tensor or None)."""
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.Tensor([1e10]).expand(dists[..., :1].shape)], -1)
    dists = dists * torch.norm(rays_d[..., None, :], dim=-1)
    rgb = torch.sigmoid(raw[..., :3])
    sigma = raw[..., 3] + (noise if noise is not None else 0.)
    alpha = 1. - torch.exp(-F.relu(sigma) * dists)
    weights = alpha * torch.cumprod(torch.cat([torch.ones((alpha.shape[0], 1)), 1. - alpha + 1e-10], -1), -1)[:, :-1]
    rgb_map = torch.sum(weights[..., None] * rgb, -2)
    depth_map = torch.sum(weights * z_vals, -1)
    disp_map = 1. / torch.max(1e-10 * torch.ones_like(depth_map), depth_map / torch.sum(weights, -1))
    acc_map = torch.sum(weights, -1)
    if white_bkgd:
        rgb_map = rgb_map + (1. - acc_map[..., None])
    return rgb_map, disp_map, acc_map, weights, depth_map


def sample_pdf(bins, weights, u):
    """utils/run_nerf_raybased_helpers.py:283-330 with the variates `u` ([N, Ni]) passed in.
    Returns (samples, inds)."""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.max(torch.zeros_like(inds - 1), inds - 1)
    above = torch.min((cdf.shape[-1] - 1) * torch.ones_like(inds), inds)
    inds_g = torch.stack([below, above], -1)
    shape = [inds_g.shape[0], inds_g.shape[1], cdf.shape[-1]]
    cdf_g = torch.gather(cdf.unsqueeze(1).expand(shape), 2, inds_g)
    bins_g = torch.gather(bins.unsqueeze(1).expand(shape), 2, inds_g)
    denom = cdf_g[..., 1] - cdf_g[..., 0]
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_g[..., 0]) / denom
    return bins_g[..., 0] + t * (bins_g[..., 1] - bins_g[..., 0]), inds


def make_u(n_rays, n_importance, det):
    """helpers:291-296: linspace (det) or CPU-generator uniforms, expanded to [N, Ni]."""
    if det:
        return torch.linspace(0., 1., steps=n_importance).expand([n_rays, n_importance])
    return torch.rand([n_rays, n_importance])


# --------------------------------------------------------------------------------- render_rays
def render_rays(ray_batch, sd_coarse, sd_fine, N_samples=64, N_importance=128, perturb=0., lindisp=False,
                white_bkgd=True, raw_noise_std=0., netchunk=1024 * 64, t_rand=None, u=None, noise0=None, noise1=None):
    """main.py:624-756 with retraw=True.  Random tensors are drawn from the CPU generator in the
    reference's order (t_rand, coarse noise, u, fine noise) unless injected.  Also returns every
    intermediate the parity tests compare."""
    N_rays = ray_batch.shape[0]
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    viewdirs = ray_batch[:, -3:]
    bounds = torch.reshape(ray_batch[..., 6:8], [-1, 1, 2])
    near, far = bounds[..., 0], bounds[..., 1]
    t_vals = torch.linspace(0., 1., steps=N_samples)
    if not lindisp:
        z_vals = near * (1. - t_vals) + far * (t_vals)
    else:
        z_vals = 1. / (1. / near * (1. - t_vals) + 1. / far * (t_vals))
    z_vals = z_vals.expand([N_rays, N_samples])
    if perturb > 0.:
        mids = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
        upper = torch.cat([mids, z_vals[..., -1:]], -1)
        lower = torch.cat([z_vals[..., :1], mids], -1)
        if t_rand is None:
            t_rand = torch.rand(z_vals.shape)
        z_vals = lower + (upper - lower) * t_rand
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
    raw0 = run_network(pts, viewdirs, sd_coarse, netchunk)
    if raw_noise_std > 0. and noise0 is None:
        noise0 = torch.randn(raw0[..., 3].shape) * raw_noise_std
    rgb0, disp0, acc0, weights0, depth0 = raw2outputs(raw0, z_vals, rays_d, noise0, white_bkgd)
    out = dict(z_vals0=z_vals, raw0=raw0, weights0=weights0, rgb0=rgb0, disp0=disp0, acc0=acc0, depth0=depth0)
    if N_importance > 0:
        z_mid = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
        if u is None:
            u = make_u(N_rays, N_importance, det=(perturb == 0.))
        z_samples, inds = sample_pdf(z_mid, weights0[..., 1:-1], u)
        z_samples = z_samples.detach()
        z_vals, _ = torch.sort(torch.cat([z_vals, z_samples], -1), -1)
        pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
        raw = run_network(pts, viewdirs, sd_fine if sd_fine is not None else sd_coarse, netchunk)
        if raw_noise_std > 0. and noise1 is None:
            noise1 = torch.randn(raw[..., 3].shape) * raw_noise_std
        rgb_map, disp_map, acc_map, weights, depth_map = raw2outputs(raw, z_vals, rays_d, noise1, white_bkgd)
        out.update(u=u, inds=inds, z_samples=z_samples, z_vals=z_vals, raw=raw, weights=weights, rgb_map=rgb_map,
                   disp_map=disp_map, acc_map=acc_map, depth_map=depth_map,
                   z_std=torch.std(z_samples, dim=-1, unbiased=False))
    else:
        out.update(rgb_map=rgb0, disp_map=disp0, acc_map=acc0, depth_map=depth0, raw=raw0, weights=weights0,
                   z_vals=z_vals)
    return out


def render_r2l(sd, H, W, focal, near, far, c2w, n_sample=16, L=10, n_blocks=43, rows=None):
    """R2L branch of render_path (main.py:297-309): rgb [H*W, 3] (or the selected rows only)."""
    pts = point_sample(H, W, focal, n_sample, near, far, c2w)
    if rows is not None:
        pts = pts[rows]
    return r2l_forward(sd, embed_r2l(pts, L), n_blocks=n_blocks)


def psnr(a, b):
    return float(-10. * torch.log10(torch.mean((a - b)**2)))


# ----------------------------------------------------------------------------------------------------
# Image metrics of render_path (main.py:330-335, 384-391)
# ----------------------------------------------------------------------------------------------------
def img2mse(x, y):
    """utils/run_nerf_raybased_helpers.py:19."""
    return torch.mean((x - y)**2)


def mse2psnr(x):
    """utils/run_nerf_raybased_helpers.py:20."""
    return -10. * torch.log(x) / torch.log(torch.Tensor([10.]))


def ssim_window(window_size=11, sigma=1.5, channel=3):
    """utils/ssim_torch.py:10-25: 1-D Gaussian (python-double exp into an fp32 tensor, normalised in fp32), its outer
    product, expanded to one filter per channel."""
    from math import exp
    g = torch.Tensor([exp(-(x - window_size // 2)**2 / float(2 * sigma**2)) for x in range(window_size)])
    g = (g / g.sum()).unsqueeze(1)
    w2 = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, window_size, window_size).contiguous()


def ssim(img, ref, window_size=11):
    """utils/ssim_torch.py:27-54 through main.py:46: img, ref [C, H, W] -> scalar (size_average=True)."""
    import torch.nn.functional as F
    a, b = img.unsqueeze(0), ref.unsqueeze(0)
    ch = a.shape[1]
    w = ssim_window(window_size, 1.5, ch).type_as(a)
    pad = window_size // 2
    mu1 = F.conv2d(a, w, padding=pad, groups=ch)
    mu2 = F.conv2d(b, w, padding=pad, groups=ch)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(a * a, w, padding=pad, groups=ch) - mu1_sq
    s2 = F.conv2d(b * b, w, padding=pad, groups=ch) - mu2_sq
    s12 = F.conv2d(a * b, w, padding=pad, groups=ch) - mu1_mu2
    C1, C2 = 0.01**2, 0.03**2
    m = ((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))
    return m.mean()


# ----------------------------------------------------------------------------------------------------
# Remaining PointSampler / PositionalEmbedder forms (model/nerf_raybased.py:128-190, 210-223)
# ----------------------------------------------------------------------------------------------------
def sample_train_cnnstyle(z_vals, rays_o, rays_d, perturb, t_rand=None):
    """model:149-168 (= sample_train2 :128-147): rays [n_img, ph, pw, 3]; one offset t_rand [n_img] per image."""
    z = z_vals[None, None, None, :].expand(*rays_o.shape[:3], z_vals.shape[0])
    if perturb > 0.:
        mids = .5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mids, z[..., -1:]], dim=-1)
        lower = torch.cat([z[..., :1], mids], dim=-1)
        t = t_rand[:, None, None, None].expand_as(z)
        z = lower + (upper - lower) * t
    return rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]


def plucker(rays_o, rays_d):
    """model:170-176: [rays_d, cross(rays_o, rays_d)]."""
    return torch.cat([rays_d, torch.cross(rays_o, rays_d, dim=-1)], dim=-1)


def sample_test_plucker(H, W, focal, c2w):
    """model:178-190."""
    ro, rd = get_rays(H, W, focal, c2w)
    rd = rd.reshape(-1, 3)
    return plucker(c2w[:3, -1].expand(rd.shape), rd)


def embed_cnnstyle(x, L, include_input=True):
    """model:210-223: [..., dim] -> [..., dim, 2L+1]."""
    w = 2**torch.linspace(0, L - 1, steps=L)
    y = x[..., :, None] * w
    y = torch.cat([torch.sin(y), torch.cos(y)], dim=-1)
    if include_input:
        y = torch.cat([y, x.unsqueeze(dim=-1)], dim=-1)
    return y
