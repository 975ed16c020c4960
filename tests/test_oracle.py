"""CPU tests: the oracle (oracle/ref_torch.py, oracle/sample_pdf_np.py) against the golden vectors that
oracle/make_golden.py produced by running the reference's own code."""
import numpy as np
import torch

from conftest import t


def same(a, b):
    a = a.detach().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    return a.shape == np.asarray(b).shape and np.array_equal(a, b, equal_nan=True)


def test_rays(golden, O):
    g = golden("rays_small")
    ro, rd = O.get_rays(int(g["H"]), int(g["W"]), float(g["focal"]), t(g["c2w"]))
    assert same(ro, g["rays_o"]) and same(rd, g["rays_d"])
    no, nd = O.ndc_rays(int(g["H"]), int(g["W"]), float(g["focal"]), 1., ro, rd)
    assert same(no, g["ndc_o"]) and same(nd, g["ndc_d"])
    for name in ("lego", "fern"):
        g = golden("rays_" + name)
        ro, rd = O.get_rays(int(g["H"]), int(g["W"]), float(g["focal"]), t(g["c2w"]))
        assert same(rd.reshape(-1, 3)[g["idx"]], g["rays_d"])
        if name == "fern":
            no, nd = O.ndc_rays(int(g["H"]), int(g["W"]), float(g["focal"]), 1., ro, rd)
            assert same(no.reshape(-1, 3)[g["idx"]], g["ndc_o"]) and same(nd.reshape(-1, 3)[g["idx"]], g["ndc_d"])


def test_encodings(golden, O):
    g = golden("embed")
    x = t(g["x"])
    assert same(O.embed_nerf(x, 10), g["nerf_L10"])
    assert same(O.embed_nerf(x / x.norm(dim=-1, keepdim=True), 4), g["nerf_L4"])
    assert same(O.embed_r2l(t(g["pts"]), 10), g["r2l_L10"])
    assert same(O.point_sample(6, 8, 11.1, 16, 2., 6., t(g["ps_c2w"])), g["ps_pts"])


def test_sample_pdf_torch_and_numpy_restatements(golden, O):
    from oracle.sample_pdf_np import sample_pdf_np
    g = golden("sample_pdf")
    bins, w = t(g["bins"]), t(g["weights"])
    u = torch.linspace(0., 1., 128).expand(bins.shape[0], 128)
    s, i = O.sample_pdf(bins, w, u)
    assert same(s, g["samples_det"]) and same(i, g["inds_det"])
    s, i, _ = sample_pdf_np(g["bins"], g["weights"], u.numpy())
    assert same(s, g["samples_det"]) and same(i, g["inds_det"])
    s, i = O.sample_pdf(bins, w, t(g["u_rnd"]))
    assert same(s, g["samples_rnd"]) and same(i, g["inds_rnd"])
    s, i, _ = sample_pdf_np(g["bins"], g["weights"], g["u_rnd"])
    assert same(s, g["samples_rnd"]) and same(i, g["inds_rnd"])


def test_sample_pdf_numpy_sum_order_matches_aten():
    from oracle.sample_pdf_np import aten_inner_sum_f32
    torch.manual_seed(0)
    for n in (1, 3, 5, 7, 8, 9, 31, 62, 63, 126, 200, 511):
        w = torch.rand(2000, n)**4 + 1e-5
        assert np.array_equal(torch.sum(w, -1).numpy(), aten_inner_sum_f32(w.numpy())), n


def test_raw2outputs(golden, O):
    g = golden("raw2outputs")
    out = O.raw2outputs(t(g["raw"]), t(g["z"]), t(g["d"]), None, True)
    for a, k in zip(out, ("rgb", "disp", "acc", "weights", "depth")):
        assert same(a, g[k]), k
    assert np.isnan(g["disp"]).sum() == 1  # the all-empty ray: 0/0 propagates through torch.max


def test_weight_init_checksums(golden, O):
    g = golden("nerf_render_lego")
    sdc, sdf = O.nerf_state_dicts(0)
    for sd, key in ((sdc, "cks_coarse"), (sdf, "cks_fine")):
        for k, v in g[key]:
            assert float(sd[k].double().abs().sum()) == float(v), k
    sd = O.r2l_state_dict(0)
    for k, v in golden("r2l_lego")["cks"]:
        assert float(sd[k].double().abs().sum()) == float(v), k


def close(a, b, atol=1e-5):
    a = a.detach().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    return a.shape == np.asarray(b).shape and np.allclose(a, b, rtol=0, atol=atol, equal_nan=True)


def test_nerf_render_lego(golden, O):
    """End to end (MLP outputs depend on the host BLAS blocking -> 1e-5), sample_pdf stage bit-exact."""
    g = golden("nerf_render_lego")
    sdc, sdf = O.nerf_state_dicts(0)
    batch = O.pack_rays(t(g["rays_o"]), t(g["rays_d"]), 2., 6.)
    with torch.no_grad():
        o = O.render_rays(batch, sdc, sdf, 64, 128, white_bkgd=True)
    for k in ("rgb_map", "acc_map", "rgb0", "acc0", "z_std", "raw", "weights0"):
        assert close(o[k], g[k]), k
    # the bit-exact stage: the reference's coarse weights in -> identical bin indices out
    z0 = t(g["z_vals"])  # merged depths are not needed here; rebuild the coarse mids from near/far
    t_vals = torch.linspace(0., 1., 64)
    zc = (2. * (1. - t_vals) + 6. * t_vals).expand(64, 64)
    mids = .5 * (zc[..., 1:] + zc[..., :-1])
    u = torch.linspace(0., 1., 128).expand(64, 128)
    s, inds = O.sample_pdf(mids, t(g["weights0"])[..., 1:-1], u)
    assert same(inds, g["inds"])
    merged, _ = torch.sort(torch.cat([zc, s], -1), -1)
    assert same(merged, z0)


def test_r2l(golden, O):
    g = golden("r2l_lego")
    sd = O.r2l_state_dict(0)
    with torch.no_grad():
        pts = O.point_sample(400, 400, O.LEGO["focal"], 16, 2., 6., t(g["c2w"]))[g["idx"]]
        assert same(pts, g["pts"])
        rgb = O.r2l_forward(sd, O.embed_r2l(pts, 10))
    assert close(rgb, g["rgb"], 1e-5)


def test_image_metrics(golden, O):
    """ssim / img2mse / mse2psnr restatements against the fixture written by the reference's own code."""
    g = golden("metrics")
    for n in "abc":
        rgb, gt = t(g[f"rgb_{n}"]), t(g[f"gt_{n}"])
        f = lambda k: float(np.asarray(g[f"{k}_{n}"]).reshape(-1)[0])
        assert float(O.ssim(rgb.permute(2, 0, 1), gt.permute(2, 0, 1))) == f("ssim")
        mse = O.img2mse(rgb, gt)
        assert float(mse) == f("mse") and float(O.mse2psnr(mse)) == f("psnr")


def test_sampler_extra_forms(golden, O):
    """CNN-style point sampling, Pluecker rays and the unflattened embedding against the reference's own outputs."""
    g = golden("sampler_extra")
    H, W, focal = int(g["H"]), int(g["W"]), float(g["focal"])
    z = 2. * (1 - torch.linspace(0., 1., 8)) + 6. * torch.linspace(0., 1., 8)
    ro, rd = t(g["ro"]), t(g["rd"])
    assert torch.equal(O.sample_train_cnnstyle(z, ro, rd, 1., t(g["t_rand"])), t(g["pts_perturb"]))
    assert torch.equal(O.sample_train_cnnstyle(z, ro, rd, 0.), t(g["pts_det"]))
    assert torch.equal(O.plucker(ro.reshape(-1, 3), rd.reshape(-1, 3)), t(g["plucker_train"]))
    assert torch.equal(O.sample_test_plucker(H, W, focal, t(g["c2w"])), t(g["plucker_test"]))
    assert torch.equal(O.embed_cnnstyle(t(g["x"]), 5), t(g["embed_L5"]))


def test_search_free_identities_of_the_det_kernels():
    """The counting identities behind hier_sample_det_kernel / sample_pdf_det_kernel (marks + running max instead of
    per-sample bisections, rank merge instead of a sort), restated in numpy, against searchsorted / sort — on the
    golden weights, on sparse / flat / tied pdfs and on ascending tables that are not a linspace."""
    from oracle.sample_pdf_np import sample_pdf_np, inds_by_marks, merge_by_ranks
    rng = np.random.RandomState(0)
    N = 300
    t_vals = np.linspace(0., 1., 64, dtype=np.float32)
    z = np.broadcast_to(2. * (1. - t_vals) + 6. * t_vals, (N, 64)).astype(np.float32).copy()
    z[N // 2:] = np.sort(rng.uniform(2., 6., (N - N // 2, 64)).astype(np.float32), -1)
    z[3, 10:20] = z[3, 10]                                       # tied coarse depths
    mids = (.5 * (z[:, 1:] + z[:, :-1])).astype(np.float32)
    w = (rng.rand(N, 62) ** 8).astype(np.float32)
    w[::3] = rng.rand(N, 62)[::3]
    w[5] = 0.                                                    # flat pdf
    w[7, 10:50] = 0.                                             # long run of equal cdf entries
    w[9] = 0.
    w[9, 31] = 1.                                                # all mass in one bin: samples pile up
    tables = [np.linspace(0., 1., 128, dtype=np.float32), np.linspace(0., 1., 64, dtype=np.float32),
              np.sort(rng.rand(128)).astype(np.float32), np.sort(np.round(rng.rand(128) * 8) / 8).astype(np.float32),
              np.zeros(64, np.float32), np.ones(128, np.float32)]
    for u in tables:
        s, inds, cdf = sample_pdf_np(mids, w, u)
        assert np.array_equal(inds_by_marks(cdf, u), inds)
        assert (np.diff(s, axis=-1) >= 0).all()                  # ascending u -> ascending samples
        merged = merge_by_ranks(z, s, inds)
        assert np.array_equal(merged, np.sort(np.concatenate([z, s], -1), -1))
