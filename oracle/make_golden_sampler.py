"""Golden vectors for the remaining PointSampler / PositionalEmbedder forms (tests/golden/sampler_extra.npz), produced
by the REAL reference classes (model/nerf_raybased.py:76-223); also pins oracle/ref_torch.py bit-for-bit.
Run in the build container only:  python oracle/make_golden_sampler.py          TEST INFRASTRUCTURE ONLY."""
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
    import model.nerf_raybased as M
    torch.autograd.set_detect_anomaly(False)
    sys.path.remove(REF)
    H, W, focal = 24, 36, 50.
    ps = M.PointSampler(H, W, focal, 8, 2., 6.)
    c2w = O.pose_spherical(25., -40., 4.)[:3, :4]
    torch.manual_seed(0)
    out = dict(c2w=c2w.numpy(), H=H, W=W, focal=focal)
    with torch.no_grad():
        ro = torch.randn(3, 5, 7, 3)
        rd = torch.randn(3, 5, 7, 3)
        t_rand = torch.rand(3)
        torch.manual_seed(123)
        p1 = ps.sample_train_cnnstyle(ro, rd, perturb=1.)      # draws torch.rand(3) on the CPU generator
        torch.manual_seed(123)
        tr = torch.rand(3)
        p0 = ps.sample_train_cnnstyle(ro, rd, perturb=0.)
        torch.manual_seed(123)
        p2 = ps.sample_train2(ro, rd, perturb=1.)
        assert torch.equal(p1, p2)
        assert torch.equal(O.sample_train_cnnstyle(ps.z_vals, ro, rd, 1., tr), p1)
        assert torch.equal(O.sample_train_cnnstyle(ps.z_vals, ro, rd, 0.), p0)
        pl = ps.sample_train_plucker(ro.reshape(-1, 3), rd.reshape(-1, 3))
        plt = ps.sample_test_plucker(c2w)
        assert torch.equal(O.plucker(ro.reshape(-1, 3), rd.reshape(-1, 3)), pl)
        assert torch.equal(O.sample_test_plucker(H, W, focal, c2w), plt)
        pe = M.PositionalEmbedder(L=5)
        x = torch.randn(2, 3, 4, 6, 3)
        e = pe.embed_cnnstyle(x)
        assert torch.equal(pe.embed(x), e) and torch.equal(O.embed_cnnstyle(x, 5), e)
        print("  oracle == reference (bit-exact): sample_train_cnnstyle / sample_train2 / plucker / embed_cnnstyle")
        out.update(ro=ro.numpy(), rd=rd.numpy(), t_rand=tr.numpy(), pts_perturb=p1.numpy(), pts_det=p0.numpy(),
                   plucker_train=pl.numpy(), plucker_test=plt.numpy(), x=x.numpy(), embed_L5=e.numpy())
    np.savez(os.path.join(ROOT, "tests", "golden", "sampler_extra.npz"), **out)
    print("written tests/golden/sampler_extra.npz")


if __name__ == "__main__":
    main()
