"""Golden vectors for the device FLIP (efficient_nerf_b200.metrics.FLIP) from the REFERENCE's own utils/flip_loss.py.

The reference's FLIP is CUDA-only (`.cuda()` / device='cuda' throughout), so unlike the other fixtures this one is
generated on a GPU box, from the unmodified file staged under baseline/_ref:

    gpurun -- 'python oracle/make_golden_flip.py gpurun_out/flip.npz'      # then: cp gpurun_out/flip.npz tests/golden/

TEST INFRASTRUCTURE ONLY.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def load_reference_flip():
    from oracle import ref_real as R
    root = R.reference_root()
    spec = importlib.util.spec_from_file_location("_r2l_ref_flip_loss", os.path.join(root, "utils", "flip_loss.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_inputs(seed=0, N=2, H=48, W=56):
    """Image pairs [N, H, W, 3] in [0, 1]: smooth gradients, hard edges, isolated points and noise."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    ref = torch.stack([.5 + .5 * torch.sin(6.3 * (xx + k * yy) + k) for k in range(3)], -1)[None].repeat(N, 1, 1, 1)
    ref[:, H // 4:H // 2, W // 3:2 * W // 3] = torch.tensor([.9, .1, .2])
    ref[1, ::7, ::5] = 1.
    test = (ref + .08 * torch.randn(ref.shape, generator=g)).clamp(0, 1)
    test[0, H // 2:, :W // 4] = test[0, H // 2:, :W // 4].flip(-1)
    test[1, 5:9, 5:9] = 0.
    return test.contiguous(), ref.contiguous()


def main(out):
    mod = load_reference_flip()
    flip = mod.FLIP()
    test, ref = make_inputs()
    ppd = 0.7 * (3840 / 0.7) * (np.pi / 180)
    res = {}
    with torch.no_grad():
        for name, (s, o) in (("unit", (1., 0.)), ("rescaled", (2., -1.))):   # main.py:366-368 feeds [-1, 1] stacks
            a = (test * s + o).permute(0, 3, 1, 2).cuda()
            b = (ref * s + o).permute(0, 3, 1, 2).cuda()
            m = flip.compute_flip(a, b, ppd)          # main.py:377: compute_flip(rec, ref, ppd)
            res["map_" + name] = m[:, 0].cpu().numpy()
            res["mean_" + name] = float(flip.forward(a, b))
    np.savez(out, test=test.numpy(), ref=ref.numpy(), ppd=ppd, torch=torch.__version__, **res)
    print("wrote", out, {k: (v.shape if hasattr(v, "shape") else v) for k, v in res.items()})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "flip.npz"))
