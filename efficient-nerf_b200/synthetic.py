"""Synthetic inputs of the BASELINE configs: camera constants, test poses and seeded random-init models.

There are no datasets or checkpoints in the build environment, so benchmarks and smoke tests use the reference's
architectures with PyTorch's default nn.Linear initialisation under a fixed seed, constructed in the reference's order
(main.py:426-445: coarse, then fine) — the modules of this package consume the RNG exactly like the reference's
(tests/test_host_logic.py), so `seeded_*` yields the weights the oracle's `*_state_dict(seed)` helpers restate.
"""
from types import SimpleNamespace

import torch

from .create_data import pose_spherical   # noqa: F401  (dataset/load_blender.py:10-28)
from .nerf_raybased import NeRF, NeRF_v3_2

LEGO = dict(H=400, W=400, focal=555.5555155968841, near=2., far=6.)       # main.py:927-931, half_res
LEGO_800 = dict(H=800, W=800, focal=1111.1110311937682, near=2., far=6.)
FERN = dict(H=378, W=504, focal=407.5658, near=0., far=1.)                # load_llff.py:131,446; NDC (main.py:914-919)


def test_pose(k, n_poses=200):
    """Pose k of the synthetic test orbit: pose_spherical(-180 + 360 k / n, -30, 4)[:3, :4]."""
    return pose_spherical(-180. + 360. * (k % n_poses) / n_poses, -30., 4.)[:3, :4].contiguous()


def r2l_args(netdepth=88, netwidth=256, use_residual=True):
    """The option namespace NeRF_v3_2 reads (model/nerf_raybased.py:486-535,543) for README.md:51's R2L command."""
    return SimpleNamespace(netdepth=netdepth, netwidth=netwidth, layerwise_netwidths='', act='relu', linear_tail=False,
                           use_residual=use_residual,
                           trial=SimpleNamespace(inact='relu', outact='none', body_arch='resmlp', n_block=-1,
                                                 res_scale=1., n_learnable=2))


def seeded_nerf_pair(seed=0, precision="fp16", device="cuda"):
    """(coarse, fine) NeRF W256 D8 with view directions, random init under torch.manual_seed(seed)."""
    torch.manual_seed(seed)
    coarse = NeRF(8, 256, 63, 27, 5, [4], True, precision=precision)
    fine = NeRF(8, 256, 63, 27, 5, [4], True, precision=precision)
    return coarse.to(device).eval(), fine.to(device).eval()


def seeded_r2l(seed=0, precision="fp16", device="cuda", netdepth=88, netwidth=256, input_dim=1008):
    """R2L NeRF_v3_2 (ResMLP W256 D88, 16 points per ray), random init under torch.manual_seed(seed)."""
    torch.manual_seed(seed)
    net = NeRF_v3_2(r2l_args(netdepth, netwidth), input_dim, 3, precision=precision)
    return net.to(device).eval()
