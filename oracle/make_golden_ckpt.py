"""TEST INFRASTRUCTURE ONLY — a small checkpoint written by the REAL reference classes, in the reference's own
format (main.py:1516-1542: state dicts + the pickled R2L module), for the checkpoint-ingestion tests.

    python oracle/make_golden_ckpt.py  ->  tests/golden/ckpt_r2l_reference.tar, ckpt_nerf_reference.npz
The R2L net is W256 with 4 points per ray and ONE ResMLP block (netdepth 4): the smallest architecture the fused
kernel covers (0.8 MB).  `args.trial` is a utils.EmptyClass, as smilelogging's update_args leaves it.
"""
import argparse
import os
import sys

import numpy as np
import torch

REF = os.environ.get("R2L_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    sys.path.insert(0, REF)
    import model.nerf_raybased as M
    import utils
    torch.autograd.set_detect_anomaly(False)
    trial = utils.EmptyClass()
    trial.ON, trial.body_arch, trial.n_block, trial.n_learnable = True, 'resmlp', -1, 2
    trial.res_scale, trial.inact, trial.outact = 1., 'relu', 'none'
    args = argparse.Namespace(netdepth=4, netwidth=256, layerwise_netwidths='', act='relu', linear_tail=False,
                              use_residual=True, trial=trial)
    torch.manual_seed(123)
    net = M.NeRF_v3_2(args, 4 * 3 * 21, 3).eval()
    with torch.no_grad():
        for p in net.parameters():          # "trained" weights: move away from the seeded init
            p.add_(0.01 * torch.randn_like(p))
    to_save = {'global_step': 1234, 'best_psnr': 30.5, 'best_psnr_step': 1000,
               'network_fn_state_dict': net.state_dict(), 'optimizer_state_dict': {}, 'network_fn': net}
    path = os.path.join(OUT, "ckpt_r2l_reference.tar")
    torch.save(to_save, path)
    pts = (torch.rand(64, 12) * 2 - 1) * 3
    with torch.no_grad():
        rgb = net(M.PositionalEmbedder(L=10)(pts))
    # DataParallel-style NeRF state dict (create_data.py wraps the teacher in nn.DataParallel: `module.` prefix)
    torch.manual_seed(7)
    nerf = M.NeRF(8, 256, 63, 27, 5, [4], True).eval()
    sd = {('module.' + k): v.numpy() for k, v in nerf.state_dict().items()}
    x = torch.cat([torch.randn(32, 63).clamp(-1, 1), torch.randn(32, 27).clamp(-1, 1)], -1)
    with torch.no_grad():
        out = nerf(x)
    np.savez(os.path.join(OUT, "ckpt_reference_io.npz"), pts=pts.numpy(), rgb=rgb.numpy(), nerf_x=x.numpy(),
             nerf_out=out.numpy(), **{"nerf_sd::" + k: v for k, v in sd.items()})
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
