"""Golden vectors for the image metrics (tests/golden/metrics.npz), produced by the REAL reference:
utils/ssim_torch.py::ssim behind the wrapper of main.py:46, and img2mse / mse2psnr of
utils/run_nerf_raybased_helpers.py:19-20.  Also pins oracle/ref_torch.py's restatement bit-for-bit.
Run in the build container only:  python oracle/make_golden_metrics.py          TEST INFRASTRUCTURE ONLY."""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("R2L_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_torch as O  # noqa: E402


def main():
    sys.path.insert(0, REF)
    import utils.run_nerf_raybased_helpers as Hh
    from utils.ssim_torch import ssim as ssim_
    torch.autograd.set_detect_anomaly(False)
    sys.path.remove(REF)
    ssim = lambda img, ref: ssim_(torch.unsqueeze(img, 0), torch.unsqueeze(ref, 0))   # main.py:46
    torch.manual_seed(0)
    out = {}
    with torch.no_grad():
        for name, (H, W) in dict(a=(48, 80), b=(33, 37), c=(100, 64)).items():
            # a smooth "rendered" image and a perturbed "ground truth" in [0, 1], HWC like render_path's tensors
            yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
            gt = torch.stack([0.5 + 0.5 * torch.sin(6 * xx + 3 * yy), xx * yy, 0.5 + 0.5 * torch.cos(9 * yy)], -1)
            gt = (gt + 0.05 * torch.rand(H, W, 3)).clamp(0, 1)
            rgb = (gt + 0.1 * torch.randn(H, W, 3) * (xx[..., None] > 0.5)).clamp(0, 1)
            s = ssim(rgb.permute(2, 0, 1), gt.permute(2, 0, 1))
            mse = Hh.img2mse(rgb, gt)
            psnr = Hh.mse2psnr(mse)
            assert torch.equal(O.ssim(rgb.permute(2, 0, 1), gt.permute(2, 0, 1)), s), "oracle ssim != reference"
            assert torch.equal(O.img2mse(rgb, gt), mse) and torch.equal(O.mse2psnr(mse), psnr)
            print(f"  oracle == reference (bit-exact): ssim/mse/psnr {name} {H}x{W}: ssim {float(s):.6f} psnr {float(psnr):.4f}")
            out.update({f"rgb_{name}": rgb.numpy(), f"gt_{name}": gt.numpy(), f"ssim_{name}": s.numpy(),
                        f"mse_{name}": mse.numpy(), f"psnr_{name}": psnr.numpy(),
                        f"err_{name}": (rgb - gt).abs().numpy()})
    np.savez(os.path.join(ROOT, "tests", "golden", "metrics.npz"), **out)
    print("written tests/golden/metrics.npz")


if __name__ == "__main__":
    main()
