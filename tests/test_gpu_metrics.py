"""GPU parity tests of the render_path caller: device-side PSNR / SSIM / error maps (metrics.cu) against the fixture
written by the reference's own code (oracle/make_golden_metrics.py) and the CPU oracle, and `render_path` itself.
Tolerances: error map bit-exact; MSE 1e-6 relative (double accumulation vs ATen's fp32 pairwise mean); PSNR 1e-4 dB;
SSIM 2e-6 absolute (separable fp32 Gaussian vs the reference's 2-D fp32 window: summation order only)."""
import numpy as np
import pytest
import torch

from conftest import t

pytestmark = pytest.mark.gpu


def test_metrics_golden(E, golden):
    g = golden("metrics")
    for n in "abc":
        f = lambda k: float(np.asarray(g[f"{k}_{n}"]).reshape(-1)[0])
        rgb, gt = t(g[f"rgb_{n}"]).cuda(), t(g[f"gt_{n}"]).cuda()
        m = E.metrics.image_errors(rgb, gt)
        assert torch.equal(m["errors"].cpu(), t(g[f"err_{n}"]))
        assert abs(float(m["mse"]) - f("mse")) <= 1e-6 * f("mse"), (float(m["mse"]), f("mse"))
        assert abs(float(m["psnr"]) - f("psnr")) <= 1e-4
        assert abs(float(m["ssim"]) - f("ssim")) <= 2e-6, (n, float(m["ssim"]), f("ssim"))
        # the reference's call forms (main.py:46, 332-335)
        s = E.metrics.ssim(rgb.permute(2, 0, 1), gt.permute(2, 0, 1))
        assert abs(float(s) - f("ssim")) <= 2e-6
        p = E.metrics.mse2psnr(E.metrics.img2mse(rgb, gt))
        assert abs(float(p) - f("psnr")) <= 1e-4


@pytest.mark.parametrize("H,W", [(1, 1), (11, 5), (32, 32), (33, 65), (400, 400), (378, 504)])
def test_metrics_vs_oracle_batch(E, O, H, W):
    """Ragged tiles, images smaller than the window, BASELINE frame sizes; a batch of frames in one call."""
    torch.manual_seed(H * 1000 + W)
    N = 3 if H * W > 10000 else 5
    gt = torch.rand(N, H, W, 3)
    rgb = (gt + 0.2 * torch.randn(N, H, W, 3) * torch.rand(N, 1, 1, 1)).clamp(0, 1)
    rgb[0] = gt[0]   # identical pair: mse 0, psnr inf, ssim 1
    m = E.metrics.image_errors(rgb.cuda(), gt.cuda())
    for i in range(N):
        ref_ssim = float(O.ssim(rgb[i].permute(2, 0, 1), gt[i].permute(2, 0, 1)))
        ref_mse = float(O.img2mse(rgb[i], gt[i]))
        assert abs(float(m["ssim"][i]) - ref_ssim) <= 5e-6, (i, float(m["ssim"][i]), ref_ssim)
        assert abs(float(m["mse"][i]) - ref_mse) <= 1e-6 * ref_mse + 1e-12
    assert torch.equal(m["errors"].cpu(), (rgb - gt).abs())
    assert float(m["mse"][0]) == 0.0 and float(m["psnr"][0]) == float("inf") and abs(float(m["ssim"][0]) - 1.0) < 1e-6
    e = E.metrics.image_errors(torch.zeros(0, H, W, 3).cuda(), torch.zeros(0, H, W, 3).cuda())
    assert e["mse"].shape == (0,) and e["ssim"].shape == (0,)


def test_render_path_r2l_and_nerf(E, O):
    """render_path (main.py:189-400): frames stacked, metrics against ground truth on the device; the R2L branch and
    the NeRF branch give the frames their single-frame entry points give, and misc matches the oracle's metrics."""
    H = W = 40
    focal = O.LEGO["focal"] * W / O.LEGO["W"]
    poses = [O.pose_spherical(th, -30., 4.) for th in (-90., 0., 45.)]
    sd = O.r2l_state_dict(0)
    net = E.NeRF_v3_2(O.r2l_args(), 1008, 3, precision="fp16")
    net.load_state_dict(sd)
    net = net.cuda().eval()
    ps = E.PointSampler(H, W, focal, 16, 2., 6.)
    with torch.no_grad():
        frames = torch.stack([E.render_r2l(net, ps, p[:3, :4].cuda()).reshape(H, W, 3) for p in poses], 0)
    torch.manual_seed(1)
    gt = (frames.cpu() + 0.05 * torch.randn(3, H, W, 3)).clamp(0, 1)
    rgbs, disps, misc = E.render_path(poses, (H, W, focal), 32768, dict(network_fn=net, perturb=0.), gt_imgs=gt,
                                      model_name="nerf_v3.2", point_sampler=ps)
    assert torch.equal(rgbs, frames) and rgbs.shape == (3, H, W, 3) and disps.shape == rgbs.shape
    ref_psnr = torch.stack([O.mse2psnr(O.img2mse(rgbs[i].cpu(), gt[i])) for i in range(3)]).mean()
    ref_ssim = torch.stack([O.ssim(rgbs[i].cpu().permute(2, 0, 1), gt[i].permute(2, 0, 1)) for i in range(3)]).mean()
    assert abs(float(misc["test_psnr_v2"]) - float(ref_psnr)) < 1e-4
    assert abs(float(misc["test_ssim"]) - float(ref_ssim)) < 5e-6
    assert abs(float(misc["test_loss"]) - float(O.img2mse(rgbs.cpu(), gt))) < 1e-9
    assert abs(float(misc["test_psnr"]) - float(O.mse2psnr(O.img2mse(rgbs.cpu(), gt)))) < 1e-4
    assert misc["errors"].shape == (3, H, W, 3)
    # test_flip = FLIP of the stacks rescaled to [-1, 1] like main.py:362-379 does (it reuses the LPIPS inputs)
    assert 0. < float(misc["test_flip"]) < 1.
    import os
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "main.py")):
        from oracle.make_golden_flip import load_reference_flip
        rescale = lambda x, ymin, ymax: (ymax - ymin) / (x.max() - x.min()) * (x - x.min()) + ymin   # main.py:364-365
        rec, ref = rescale(rgbs.permute(0, 3, 1, 2), -1, 1), rescale(gt.cuda().permute(0, 3, 1, 2), -1, 1)
        with torch.no_grad():
            want = load_reference_flip().FLIP().compute_flip(rec, ref, 0.7 * (3840 / 0.7) * (np.pi / 180)).mean()
        assert abs(float(misc["test_flip"]) - float(want)) < 2e-5
    # NeRF branch
    sdc, sdf = O.nerf_state_dicts(0)
    coarse = E.NeRF(8, 256, 63, 27, 5, [4], True, precision="fp16")
    fine = E.NeRF(8, 256, 63, 27, 5, [4], True, precision="fp16")
    coarse.load_state_dict(sdc), fine.load_state_dict(sdf)
    coarse, fine = coarse.cuda().eval(), fine.cuda().eval()
    kw = dict(network_query_fn=None, perturb=0., N_importance=128, network_fine=fine, N_samples=64, network_fn=coarse,
              use_viewdirs=True, white_bkgd=True, raw_noise_std=0., ndc=False, lindisp=False, near=2., far=6.)
    rgbs, disps, misc = E.render_path(poses[:2], (H, W, focal), 32768, kw)
    with torch.no_grad():
        one, d1, _, _ = E.render_image(H, W, focal, chunk=32768, c2w=poses[1][:3, :4].cuda(), **kw)
    assert rgbs.shape == (2, H, W, 3) and disps.shape == (2, H, W) and misc == {}
    assert torch.equal(rgbs[1], one) and torch.equal(disps[1], d1, ) or torch.allclose(disps[1], d1, equal_nan=True)


# ------------------------------------------------------------------ FLIP (utils/flip_loss.py, main.py:370-379)
def test_flip_matches_the_reference_fixture(E, golden):
    """tests/golden/flip.npz was produced on a B200 by the REFERENCE's own FLIP (oracle/make_golden_flip.py)."""
    g = golden("flip")
    test, ref = t(g["test"]).cuda(), t(g["ref"]).cuda()
    for name, sc in (("unit", (1., 0.)), ("rescaled", (2., -1.))):
        m, mean = E.metrics.flip_map(test, ref, float(g["ppd"]), scale=sc)
        want = t(g["map_" + name])
        assert float((m.cpu() - want).abs().max()) < 2e-4, name
        assert abs(float(mean.mean()) - float(g["mean_" + name])) < 2e-5
    # the module form, reference layout [N, 3, H, W] (main.py:362-377)
    f = E.metrics.FLIP()
    a, b = test.permute(0, 3, 1, 2), ref.permute(0, 3, 1, 2)
    m = f.compute_flip(a, b, f.pixels_per_degree)
    assert tuple(m.shape) == (2, 1, 48, 56)
    assert abs(float(f(a, b)) - float(g["mean_unit"])) < 2e-5
    assert float(E.metrics.flip_map(test, test)[1].abs().max()) == 0.        # identical images: FLIP = 0


def test_flip_live_against_the_staged_reference(E):
    """When baseline/_ref travelled to the box: the reference's own FLIP module, run here, on full-size frames."""
    import os
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
    if not os.path.exists(os.path.join(root, "utils", "flip_loss.py")):
        pytest.skip("baseline/_ref is not staged")
    from oracle.make_golden_flip import load_reference_flip
    ref_flip = load_reference_flip().FLIP()
    torch.manual_seed(3)
    gt = torch.rand(3, 100, 100, 3, device="cuda")
    gt = torch.nn.functional.interpolate(gt.permute(0, 3, 1, 2), size=(400, 400), mode="bilinear").permute(0, 2, 3, 1).contiguous()
    img = (gt + .05 * torch.randn_like(gt)).clamp(0, 1)
    with torch.no_grad():
        want = ref_flip.compute_flip(img.permute(0, 3, 1, 2), gt.permute(0, 3, 1, 2), ref_flip.pixels_per_degree)
        got, mean = E.metrics.flip_map(img, gt)
    assert float((got - want[:, 0]).abs().max()) < 2e-4
    assert abs(float(mean.mean()) - float(want.mean())) < 1e-5
