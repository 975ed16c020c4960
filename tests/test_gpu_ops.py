"""GPU parity tests (through the Python mirror -> ctypes -> C ABI -> CUDA) for the non-MLP hot-path ops.

Bar: bit-exact for rays / NDC / stratified depths / point sampling / sample_pdf (indices AND samples) /
sorted merge; for ops that call sin/cos/exp the tolerance is stated at the assert (CUDA's sinf/cosf/expf
are within 2 ulp of the reference's Sleef kernels, never bit-identical)."""
import numpy as np
import pytest
import torch

from conftest import t

pytestmark = pytest.mark.gpu


def cu(x):
    return (x if isinstance(x, torch.Tensor) else t(x)).cuda()


def exact(a, b, what=""):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if not np.array_equal(a, b, equal_nan=True):
        bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
        raise AssertionError(f"{what}: {bad.sum()} / {a.size} elements differ, max|d|={np.nanmax(np.abs(a - b)[bad])}")


def close(a, b, atol, what=""):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert np.array_equal(np.isnan(a), np.isnan(b)), f"{what}: NaN pattern differs"
    d = np.nanmax(np.abs(a - b)) if a.size else 0.0
    assert d <= atol, f"{what}: max|d|={d} > {atol}"


# ------------------------------------------------------------------ rays
def test_get_rays_and_ndc_bit_exact(E, O, golden):
    g = golden("rays_small")
    H, W, f = int(g["H"]), int(g["W"]), float(g["focal"])
    ro, rd = E.get_rays(H, W, f, cu(g["c2w"]))
    exact(ro, g["rays_o"], "rays_o"), exact(rd, g["rays_d"], "rays_d")
    no, nd = E.ndc_rays(H, W, f, 1., ro, rd)
    exact(no, g["ndc_o"], "ndc o"), exact(nd, g["ndc_d"], "ndc d")
    # full frames of the BASELINE cameras against the oracle run live
    for cam, c2w in ((O.LEGO, O.pose_spherical(77., -30., 4.)[:3, :4]), (O.LEGO_800, O.pose_spherical(-10., -60., 4.)[:3, :4]),
                     (O.FERN, t(golden("rays_fern")["c2w"]))):
        ro, rd = E.get_rays(cam["H"], cam["W"], cam["focal"], c2w.cuda())
        o_ro, o_rd = O.get_rays(cam["H"], cam["W"], cam["focal"], c2w)
        exact(rd, o_rd, "full-frame rays_d"), exact(ro, o_ro.contiguous(), "full-frame rays_o")
    g = golden("rays_fern")
    ro, rd = E.get_rays(int(g["H"]), int(g["W"]), float(g["focal"]), cu(g["c2w"]))
    no, nd = E.ndc_rays(int(g["H"]), int(g["W"]), float(g["focal"]), 1., ro, rd)
    exact(no.reshape(-1, 3)[g["idx"]], g["ndc_o"], "fern ndc o"), exact(nd.reshape(-1, 3)[g["idx"]], g["ndc_d"], "fern ndc d")
    # numpy / host c2w are accepted like the reference does
    ro2, _ = E.get_rays(6, 5, 7.5, g["c2w"])
    assert ro2.is_cuda and tuple(ro2.shape) == (6, 5, 3)


def test_viewdirs_and_ray_packing(E, O):
    torch.manual_seed(0)
    d = torch.randn(1000, 3)
    v = E.normalize_dirs(d.cuda())
    ref = d / torch.norm(d, dim=-1, keepdim=True)
    close(v, ref, 1.2e-7, "viewdirs (1 ulp: |d| reduction order)")


def test_z_vals_and_perturb_bit_exact(E, O):
    from efficient_nerf_b200.render import _z_vals
    torch.manual_seed(1)
    N, S = 300, 64
    near = torch.full((N, 1), 2.) + torch.rand(N, 1)
    far = torch.full((N, 1), 6.) + torch.rand(N, 1)
    t_vals = torch.linspace(0., 1., S)
    t_rand = torch.rand(N, S)
    for lindisp in (False, True):
        for tr in (None, t_rand):
            z = _z_vals(near.cuda(), far.cuda(), t_vals.cuda(), lindisp, tr)
            if not lindisp:
                zr = near * (1. - t_vals) + far * t_vals
            else:
                zr = 1. / (1. / near * (1. - t_vals) + 1. / far * t_vals)
            if tr is not None:
                mids = .5 * (zr[..., 1:] + zr[..., :-1])
                upper = torch.cat([mids, zr[..., -1:]], -1)
                lower = torch.cat([zr[..., :1], mids], -1)
                zr = lower + (upper - lower) * tr
            exact(z, zr, f"z_vals lindisp={lindisp} perturb={tr is not None}")


def test_point_sampler_bit_exact(E, O, golden):
    g = golden("embed")
    ps = E.PointSampler(6, 8, 11.1, 16, 2., 6.)
    exact(ps.sample_test(cu(g["ps_c2w"])), g["ps_pts"], "PointSampler.sample_test")
    c2w = O.pose_spherical(12., -33., 4.)[:3, :4]
    ps = E.PointSampler(400, 400, O.LEGO["focal"], 16, 2., 6.)
    exact(ps.sample_test(c2w.cuda()), O.point_sample(400, 400, O.LEGO["focal"], 16, 2., 6., c2w), "sample_test 400x400")
    # sample_train without / with injected perturbation
    ro, rd = O.get_rays(20, 20, 30., c2w)
    ro, rd = ro.reshape(-1, 3).contiguous(), rd.reshape(-1, 3).contiguous()
    ps = E.PointSampler(20, 20, 30., 16, 2., 6.)
    z = (2. * (1 - torch.linspace(0., 1., 16)) + 6. * torch.linspace(0., 1., 16))
    ref = (ro[..., None, :] + rd[..., None, :] * z[None, :, None]).reshape(400, -1)
    exact(ps.sample_train(ro.cuda(), rd.cuda(), 0.), ref, "sample_train")
    tr = torch.rand(400, 16)
    zz = z[None].expand(400, 16)
    mids = .5 * (zz[..., 1:] + zz[..., :-1])
    zp = torch.cat([zz[..., :1], mids], -1) + (torch.cat([mids, zz[..., -1:]], -1) - torch.cat([zz[..., :1], mids], -1)) * tr
    ref = (ro[..., None, :] + rd[..., None, :] * zp[..., :, None]).reshape(400, -1)
    exact(ps.sample_train(ro.cuda(), rd.cuda(), 1., t_rand=tr), ref, "sample_train perturbed")


# ------------------------------------------------------------------ encodings
def test_embedders(E, O, golden):
    g = golden("embed")
    x = cu(g["x"])
    e10, d10 = E.get_embedder(10, 0)
    e4, d4 = E.get_embedder(4, 0)
    assert (d10, d4) == (63, 27)
    # tolerance 2e-6 absolute: arguments reach 2^9*5 rad; sinf/cosf (2 ulp) vs Sleef (1 ulp) on |values| <= 1,
    # the identity columns are copied exactly
    close(e10(x), g["nerf_L10"], 2e-6, "Embedder L=10")
    exact(e10(x)[:, :3], g["x"], "identity part")
    xn = x / x.norm(dim=-1, keepdim=True)
    close(e4(xn), O.embed_nerf(xn.cpu(), 4), 2e-6, "Embedder L=4")
    pe = E.PositionalEmbedder(L=10)
    assert pe.embed_dim == 21
    close(pe(cu(g["pts"])), g["r2l_L10"], 2e-6, "PositionalEmbedder")
    # larger, ragged sizes and leading dims
    torch.manual_seed(5)
    x = (torch.rand(3, 1001, 3) * 2 - 1) * 6
    out = e10(x.cuda())
    assert tuple(out.shape) == (3, 1001, 63)
    close(out, O.embed_nerf(x, 10), 2e-6, "Embedder ragged")
    p = (torch.rand(777, 48) * 2 - 1) * 6
    close(pe(p.cuda()), O.embed_r2l(p, 10), 2e-6, "PositionalEmbedder ragged")
    assert tuple(e10(torch.zeros(0, 3).cuda()).shape) == (0, 63)
    i_fn, i_dim = E.get_embedder(10, -1)
    assert i_dim == 3


# ------------------------------------------------------------------ compositing
def test_raw2outputs_golden_and_nan_propagation(E, golden):
    g = golden("raw2outputs")
    out = E.raw2outputs(cu(g["raw"]), cu(g["z"]), cu(g["d"]), 0, True)
    # tolerance 2e-6: expf (2 ulp) vs Sleef inside alpha and sigmoid, fp32 sum order of <=64 terms
    for a, k in zip(out, ("rgb", "disp", "acc", "weights", "depth")):
        close(a, g[k], 2e-6 if k != "disp" else 2e-6 * 4, f"raw2outputs {k}")
    assert torch.isnan(out[1]).sum().item() == 1  # the empty ray: disp = 1/max(1e-10, 0/0) = NaN like torch.max


@pytest.mark.parametrize("S", [2, 31, 64, 100, 128, 160, 192, 200, 256, 257])   # every blocked-kernel K, full and ragged; 257: generic kernel
def test_raw2outputs_vs_oracle(E, O, S):
    torch.manual_seed(S)
    N = 513
    raw = torch.randn(N, S, 4)
    raw[:, :, 3] *= 3.
    z = torch.sort(torch.rand(N, S) * 4 + 2, -1)[0]
    d = torch.randn(N, 3)
    noise = torch.randn(N, S) * 0.5
    for wb in (False, True):
        for nz in (None, noise):
            ref = O.raw2outputs(raw, z, d, nz, wb)
            out = E.raw2outputs(raw.cuda(), z.cuda(), d.cuda(), 1.0 if nz is not None else 0., wb, noise=nz)
            for a, b, k in zip(out, ref, ("rgb", "disp", "acc", "weights", "depth")):
                tol = 5e-6 if k not in ("disp",) else 5e-5
                close(a, b, tol, f"raw2outputs S={S} {k}")
    out = E.raw2outputs(torch.zeros(0, S, 4).cuda(), torch.zeros(0, S).cuda(), torch.zeros(0, 3).cuda())
    assert out[0].shape == (0, 3) and out[3].shape == (0, S)


# ------------------------------------------------------------------ sample_pdf
def test_sample_pdf_golden_bit_exact(E, golden):
    g = golden("sample_pdf")
    s, i = E.sample_pdf(cu(g["bins"]), cu(g["weights"]), 128, det=True, return_inds=True)
    exact(i, g["inds_det"], "inds det"), exact(s, g["samples_det"], "samples det")
    s, i = E.sample_pdf(cu(g["bins"]), cu(g["weights"]), 64, u=cu(g["u_rnd"]), return_inds=True)
    exact(i, g["inds_rnd"], "inds random"), exact(s, g["samples_rnd"], "samples random")
    # host tensors in -> host tensors out (the reference's call site passes .cpu() tensors, main.py:723)
    s = E.sample_pdf(t(g["bins"]), t(g["weights"]), 128, det=True)
    assert not s.is_cuda
    exact(s, g["samples_det"], "host round trip")


def test_sample_pdf_reference_weights_bit_exact(E, O, golden):
    """The north-star gate: fed the reference's coarse weights, bin indices are bit-exact."""
    g = golden("nerf_render_lego")
    t_vals = torch.linspace(0., 1., 64)
    zc = (2. * (1. - t_vals) + 6. * t_vals).expand(64, 64)
    mids = (.5 * (zc[..., 1:] + zc[..., :-1])).contiguous()
    w0 = cu(g["weights0"])
    s, i = E.sample_pdf(mids.cuda(), w0[..., 1:-1], 128, det=True, return_inds=True)  # strided view, no copy
    exact(i, g["inds"], "inds from reference weights")
    merged, _ = E.merge_sorted(zc.contiguous().cuda(), s, want_std=True)
    exact(merged, g["z_vals"], "merged fine depths")


@pytest.mark.parametrize("nb", [2, 5, 8, 9, 17, 63, 64, 65, 129, 300])
def test_sample_pdf_vs_oracle_ragged(E, O, nb):
    torch.manual_seed(nb)
    N = 4001
    bins = torch.sort(torch.rand(N, nb) * 4 + 2, -1)[0]
    w = torch.rand(N, nb - 1)**8
    w[::3] = torch.rand(N, nb - 1)[::3]
    w[5] = 0.
    for Ni in (1, 64, 128, 200):
        u = torch.linspace(0., 1., Ni).expand(N, Ni) if Ni > 1 else torch.full((N, 1), 0.5)
        sr, ir = O.sample_pdf(bins, w, u)
        s, i = E.sample_pdf(bins.cuda(), w.cuda(), Ni, u=u.contiguous().cuda(), return_inds=True)
        exact(i, ir, f"inds nb={nb} Ni={Ni}"), exact(s, sr, f"samples nb={nb} Ni={Ni}")
    ur = torch.rand(N, 96)
    sr, ir = O.sample_pdf(bins, w, ur)
    s, i = E.sample_pdf(bins.cuda(), w.cuda(), 96, u=ur.cuda(), return_inds=True)
    exact(i, ir, "inds random"), exact(s, sr, "samples random")
    assert E.sample_pdf(torch.zeros(0, nb).cuda(), torch.zeros(0, nb - 1).cuda(), 8, det=True).shape == (0, 8)


def test_sample_pdf_full_frame_properties(E):
    """BASELINE size (160 000 rays): sortedness, range and idempotence properties."""
    torch.manual_seed(0)
    N = 160000
    bins = torch.sort(torch.rand(N, 63, device="cuda") * 4 + 2, -1)[0]
    w = torch.rand(N, 62, device="cuda")**6
    s, i = E.sample_pdf(bins, w, 128, det=True, return_inds=True)
    assert bool((s[:, 1:] >= s[:, :-1]).all())             # det u is increasing -> samples non-decreasing
    assert bool((i[:, 1:] >= i[:, :-1]).all()) and int(i.min()) >= 1 and int(i.max()) <= 63
    assert bool((s >= bins[:, :1]).all()) and bool((s <= bins[:, -1:]).all())
    s2 = E.sample_pdf(bins, w, 128, det=True)
    assert torch.equal(s, s2)                               # deterministic


@pytest.mark.parametrize("Ni", [64, 128])
def test_hier_sample_fused_equals_unfused(E, Ni):
    """r2l_hier_sample (mids + sample_pdf + sorted merge + z_std in one kernel) against the separate kernels, bit for
    bit: stratified z, peaky / flat / all-zero weights (ties among the samples), and rows whose z is NOT ascending
    (the in-kernel bitonic fallback)."""
    torch.manual_seed(Ni)
    N = 5000
    t_vals = torch.linspace(0., 1., 64)
    z = (2. * (1. - t_vals) + 6. * t_vals).expand(N, 64).clone()
    mids = .5 * (z[:, 1:] + z[:, :-1])
    upper, lower = torch.cat([mids, z[:, -1:]], -1), torch.cat([z[:, :1], mids], -1)
    z[N // 2:] = (lower + (upper - lower) * torch.rand(N, 64))[N // 2:]          # stratified rows
    w = torch.rand(N, 64)**8
    w[::3] = torch.rand(N, 64)[::3]
    w[5] = 0.                      # flat pdf
    w[7, 10:50] = 0.               # long zero run: many equal cdf entries
    z[11] = z[11].flip(-1)         # descending z: fallback path
    z[12, 20], z[12, 40] = z[12, 40].clone(), z[12, 20].clone()
    u = torch.linspace(0., 1., Ni)
    zc, wc = z.cuda(), w.cuda()
    assert E.run_nerf_raybased_helpers.hier_sample_supported(zc, wc, Ni, u)
    assert not E.run_nerf_raybased_helpers.hier_sample_supported(zc[:, :32], wc[:, :32], Ni, u)    # 32 coarse samples
    z_all, z_std, smp, inds = E.run_nerf_raybased_helpers.hier_sample(zc, wc, Ni, u, want_samples=True, want_inds=True)
    mids_c = (.5 * (zc[..., 1:] + zc[..., :-1])).contiguous()
    s_ref, i_ref = E.sample_pdf(mids_c, wc[..., 1:-1], Ni, u=u, return_inds=True, u_sorted=False)   # bisection kernel
    m_ref, std_ref = E.merge_sorted(zc, s_ref, want_std=True)
    exact(smp, s_ref, "fused samples"), exact(inds, i_ref, "fused inds")
    exact(z_all, m_ref, "fused merged depths"), close(z_std, std_ref, 5e-7, "fused z_std")   # other summation order
    exact(z_all, torch.sort(torch.cat([z, s_ref.cpu()], -1), -1)[0], "fused vs torch.sort")
    # the same table without the "ascending" promise runs the bisection kernel: identical results
    g_all, g_std, g_smp, g_inds = E.run_nerf_raybased_helpers.hier_sample(zc, wc, Ni, u.cuda(), want_samples=True,
                                                                          want_inds=True, u_sorted=False)
    exact(g_smp, smp, "search-free vs bisection samples"), exact(g_inds, inds, "search-free vs bisection inds")
    exact(g_all, z_all, "search-free vs bisection merged depths"), close(g_std, z_std, 5e-7, "search-free vs bisection z_std")
    # ascending tables that are NOT a linspace (ties, clustered, constant), and weights that break the cdf
    # (negative -> non-monotone, NaN, inf): the search-free kernel must still equal the separate kernels
    w2 = wc.clone()
    w2[20, 30] = -0.5
    w2[21, 5] = float("nan")
    w2[22, 40] = float("inf")
    w2[23] = -1e-5                      # every w + 1e-5 == 0: total 0, pdf NaN
    tables = [torch.sort(torch.rand(Ni))[0], torch.sort(torch.round(torch.rand(Ni) * 8) / 8)[0], torch.zeros(Ni),
              torch.ones(Ni), torch.sort(torch.rand(Ni) ** 6)[0], u]
    for k, ut in enumerate(tables):
        ww = w2 if k == len(tables) - 1 else wc
        a_all, a_std, a_smp, a_inds = E.run_nerf_raybased_helpers.hier_sample(zc, ww, Ni, ut, want_samples=True,
                                                                              want_inds=True)
        s_ref, i_ref = E.sample_pdf(mids_c, ww[..., 1:-1], Ni, u=ut, return_inds=True, u_sorted=False)
        m_ref, std_ref = E.merge_sorted(zc, s_ref, want_std=True)
        exact(a_smp, s_ref, f"table {k} samples"), exact(a_inds, i_ref, f"table {k} inds")
        # the stand-alone sample_pdf takes the search-free kernel for an ascending host table: same bits
        s_det, i_det = E.sample_pdf(mids_c, ww[..., 1:-1], Ni, u=ut, return_inds=True)
        exact(s_det, s_ref, f"table {k} sample_pdf search-free samples")
        exact(i_det, i_ref, f"table {k} sample_pdf search-free inds")
        s_str = E.sample_pdf(mids_c, ww[:, 1:-1], Ni, u=ut)          # strided weights view, no inds
        exact(s_str, s_ref, f"table {k} sample_pdf search-free, strided weights")
        ok = torch.isfinite(s_ref).all(-1)          # rows with NaN samples: sort order of NaNs is unspecified
        exact(a_all[ok], m_ref[ok], f"table {k} merged depths")
        close(a_std[ok], std_ref[ok], 1e-6, f"table {k} z_std")
    # stochastic variates: per-ray [N, Ni] (create_data) and an unsorted shared table -> samples come out unordered,
    # the kernel sorts them alone and rank-merges; rows 11 / 12 (z not ascending) take the full bitonic sort
    for uu in (torch.rand(N, Ni), torch.rand(Ni)):
        assert E.run_nerf_raybased_helpers.hier_sample_supported(zc, wc, Ni, uu)
        z_all, z_std, smp, inds = E.run_nerf_raybased_helpers.hier_sample(zc, wc, Ni, uu, want_samples=True,
                                                                          want_inds=True)
        s_ref, i_ref = E.sample_pdf(mids_c, wc[..., 1:-1], Ni, u=uu, return_inds=True)
        m_ref, std_ref = E.merge_sorted(zc, s_ref, want_std=True)
        exact(smp, s_ref, "fused samples (stochastic u)"), exact(inds, i_ref, "fused inds (stochastic u)")
        exact(z_all, m_ref, "fused merged depths (stochastic u)"), exact(z_std, std_ref, "fused z_std (stochastic u)")
    e = E.run_nerf_raybased_helpers.hier_sample(zc[:0], wc[:0], Ni, u)
    assert e[0].shape == (0, 64 + Ni) and e[1].shape == (0,)


def test_merge_sorted(E):
    torch.manual_seed(0)
    for na, nbv in ((64, 128), (64, 64), (16, 7), (0, 33), (100, 300)):
        a = torch.sort(torch.rand(1000, na) * 4 + 2, -1)[0]
        b = torch.rand(1000, nbv) * 4 + 2
        out, std = E.merge_sorted(a.cuda(), b.cuda(), want_std=True)
        exact(out, torch.sort(torch.cat([a, b], -1), -1)[0], f"merge {na}+{nbv}")
        close(std, torch.std(b, dim=-1, unbiased=False), 1e-6, "z_std")
        # both inputs ascending (every deterministic render): the rank-merge fast path; with heavy ties
        bs = torch.sort(b, -1)[0]
        out, std = E.merge_sorted(a.cuda(), bs.cuda(), want_std=True)
        exact(out, torch.sort(torch.cat([a, bs], -1), -1)[0], f"sorted merge {na}+{nbv}")
        close(std, torch.std(bs, dim=-1, unbiased=False), 1e-6, "z_std sorted")
        aq = torch.sort(torch.round(a * 4) / 4, -1)[0]
        bq = torch.sort(torch.round(b * 4) / 4, -1)[0]
        out, _ = E.merge_sorted(aq.cuda(), bq.cuda(), want_std=True)
        exact(out, torch.sort(torch.cat([aq, bq], -1), -1)[0], f"tied merge {na}+{nbv}")


# ------------------------------------------------------------------ remaining PointSampler / PositionalEmbedder forms
def test_sampler_extra_forms_golden(E, golden):
    """sample_train_cnnstyle / sample_train2 (one stratification offset per image), Pluecker rays
    (sample_train_plucker / sample_test_plucker, the --plucker branch of render_path) and the unflattened
    embed / embed_cnnstyle against the outputs of the reference's own classes (oracle/make_golden_sampler.py)."""
    g = golden("sampler_extra")
    H, W, focal = int(g["H"]), int(g["W"]), float(g["focal"])
    ps = E.PointSampler(H, W, focal, 8, 2., 6.)
    ro, rd = cu(g["ro"]), cu(g["rd"])
    close(ps.sample_train_cnnstyle(ro, rd, 1., t_rand=t(g["t_rand"])), g["pts_perturb"], 2e-6, "cnnstyle perturb")
    close(ps.sample_train2(ro, rd, 0.), g["pts_det"], 2e-6, "cnnstyle det")
    assert ps.sample_train_cnnstyle(ro, rd, 0.).shape == (3, 5, 7, 8, 3)
    exact(ps.sample_train_plucker(ro.reshape(-1, 3), rd.reshape(-1, 3)), g["plucker_train"], "plucker train")
    close(ps.sample_test_plucker(cu(g["c2w"])), g["plucker_test"], 2e-6, "plucker test")
    pe = E.PositionalEmbedder(5)
    e = pe.embed_cnnstyle(cu(g["x"]))
    assert e.shape == (2, 3, 4, 6, 3, 11)
    close(e, g["embed_L5"], 2e-6, "embed_cnnstyle")
    close(pe.embed(cu(g["x"])), g["embed_L5"], 2e-6, "embed")

