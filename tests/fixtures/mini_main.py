"""A miniature caller written for the drop-in tests (NOT reference code): it binds the hot path through the
reference's import paths and option/logging packages the way main.py does (main.py:12-24, 297-309, 407-453), defines
deliberately naive module-level glue that `dropin.patch_glue` must replace, and renders one R2L and one NeRF frame."""
import numpy as np
import torch

from model.nerf_raybased import NeRF, NeRF_v3_2, PositionalEmbedder, PointSampler
from utils.run_nerf_raybased_helpers import get_rays, get_embedder, to8b
from smilelogging import Logger
from smilelogging import argparser as parser
from smilelogging.utils import update_args, get_n_params_, get_n_flops_

parser.add_argument('--config', is_config_file=True)
parser.add_argument('--out', type=str, required=True)
parser.add_argument('--res', type=int, default=40)
parser.add_argument('--netdepth', type=int, default=8)
parser.add_argument('--netwidth', type=int, default=256)
parser.add_argument('--n_sample_per_ray', type=int, default=16)
parser.add_argument('--multires', type=int, default=10)
parser.add_argument('--use_residual', action='store_true')
parser.add_argument('--white_bkgd', action='store_true')
parser.add_argument('--act', type=str, default='relu')
parser.add_argument('--layerwise_netwidths', type=str, default='')
parser.add_argument('--linear_tail', action='store_true')
parser.add_argument('--trial.ON', action='store_true')
parser.add_argument('--trial.body_arch', type=str, default='mlp')
parser.add_argument('--trial.res_scale', type=float, default=1.)
parser.add_argument('--trial.n_learnable', type=int, default=2)
parser.add_argument('--trial.inact', type=str, default='relu')
parser.add_argument('--trial.outact', type=str, default='none')
parser.add_argument('--trial.n_block', type=int, default=-1)
args = update_args(parser.parse_args())
logger = Logger(args)
device = torch.device('cuda')
CALLS = []


def render_rays(ray_batch, network_fn, network_query_fn, N_samples, **kwargs):
    CALLS.append('naive render_rays')          # must never run: the launcher rebinds this name
    raise AssertionError('the module-level glue was not patched')


def render(H, W, focal, chunk=1024 * 32, rays=None, c2w=None, **kwargs):
    CALLS.append('naive render')
    raise AssertionError('the module-level glue was not patched')


def render_func(model, pose):                      # main.py:401-404
    with torch.no_grad():
        return model(positional_embedder(point_sampler.sample_test(pose)))


def train():
    global positional_embedder, point_sampler
    H = W = args.res
    focal = 555.5555155968841 * W / 400.
    c2w = torch.tensor([[1., 0., 0., 0.], [0., .5, -.8660254, -3.4641016], [0., .8660254, .5, 2.]], device=device)
    torch.manual_seed(0)
    positional_embedder = PositionalEmbedder(L=args.multires)
    point_sampler = PointSampler(H, W, focal, args.n_sample_per_ray, 2., 6.)
    model = NeRF_v3_2(args, args.n_sample_per_ray * 3 * positional_embedder.embed_dim, 3).to(device)
    logger.info(f'params {get_n_params_(model)} flops {get_n_flops_(model, count_adds=False)}')
    with torch.no_grad():
        rgb_r2l = model(positional_embedder(point_sampler.sample_test(c2w))).view(H, W, 3)
    embed_fn, input_ch = get_embedder(10, 0)
    embeddirs_fn, input_ch_views = get_embedder(4, 0)
    coarse = NeRF(D=8, W=256, input_ch=input_ch, output_ch=5, skips=[4], input_ch_views=input_ch_views,
                  use_viewdirs=True).to(device)
    fine = NeRF(D=8, W=256, input_ch=input_ch, output_ch=5, skips=[4], input_ch_views=input_ch_views,
                use_viewdirs=True).to(device)
    with torch.no_grad():
        rgb, disp, acc, extras = render(H, W, focal, chunk=32768, c2w=c2w, network_fn=coarse, network_fine=fine,
                                        network_query_fn=None, N_samples=64, N_importance=128, perturb=0.,
                                        raw_noise_std=0., white_bkgd=args.white_bkgd, use_viewdirs=True, ndc=False,
                                        near=2., far=6.)
    # main.py:1125-1132 (--benchmark): torch's Timer imports render_func from __main__, with a 4x4 pose
    import torch.utils.benchmark as benchmark
    pose44 = torch.cat([c2w, torch.tensor([[0., 0., 0., 1.]], device=device)], 0)
    timer = benchmark.Timer(stmt='render_func(model, pose)', setup='from __main__ import render_func',
                            globals={'model': model, 'pose': pose44})
    timed = timer.timeit(3)
    assert torch.equal(render_func(model, pose44).view(H, W, 3), rgb_r2l)
    logger.info(f'render_func: {timed.mean * 1e3:.3f} ms')
    np.savez(args.out, r2l=rgb_r2l.cpu().numpy(), nerf=rgb.cpu().numpy(), frame8=to8b(rgb_r2l),
             lazy=type(positional_embedder(point_sampler.sample_test(c2w))).__name__,
             lazy_pts=type(point_sampler.sample_test(c2w)).__name__, calls=np.array(CALLS, dtype=str),
             trial=args.trial.body_arch)
