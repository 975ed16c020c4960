"""Compositing fused into the NeRF kernel (r2l_nerf_render, csrc/mlp_nerf_pp.cu compositor) against the two-step route
(r2l_nerf_forward -> raw [N,S,4] -> r2l_raw2outputs, main.py:707-709 / 738-741 + 556-621).

The fused route stages (rgb, sigma) rows in a per-CTA ring, composites every ray inside the MLP kernel with the same
sample partition and scan order as raw2outputs_blocked_kernel, and re-composites the rays whose far sample the fp32
fix-up patched.  The bar is BIT-IDENTITY of every output (rgb_map, disp_map, acc_map, depth_map, weights) — which makes
the whole-frame census of test_gpu_frames.py (run on the two-step route) hold for the fused route as well.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def load_nerf(E, sd, precision="fp16"):
    net = E.NeRF(8, 256, 63, 27, 5, [4], True, precision=precision)
    net.load_state_dict(sd)
    return net.cuda().eval()


def rays(n, S, seed, near=2., far=6.):
    g = torch.Generator().manual_seed(seed)
    o = (torch.randn(n, 3, generator=g) * 0.3 + torch.tensor([0., 0., 4.])).cuda()
    d = torch.randn(n, 3, generator=g)
    d = (d / d.norm(dim=-1, keepdim=True) * (0.7 + 0.6 * torch.rand(n, 1, generator=g))).cuda()
    v = d / d.norm(dim=-1, keepdim=True)
    # ascending, irregular depths (what the fine pass sees)
    z = torch.sort(near + (far - near) * torch.rand(n, S, generator=g), dim=-1)[0].cuda()
    return o, d, v, z


def same(a, b):
    return a.shape == b.shape and bool(torch.equal(torch.nan_to_num(a, nan=-7.), torch.nan_to_num(b, nan=-7.)))


def both_routes(E, net, o, d, v, z, white):
    with torch.no_grad():
        net.set_fused_compositing(True)
        fused = net.render_samples(o, d, v, z, white, want_weights=True)
        n_flag = net.far_flagged()
        net.set_fused_compositing(False)
        twostep = net.render_samples(o, d, v, z, white, want_weights=True)
        raw = net.forward_samples(o, d, v, z)
        plain = E.raw2outputs(raw, z, d, 0., white)
        net.set_fused_compositing(True)
    torch.cuda.synchronize()
    return fused, twostep, plain, n_flag


@pytest.mark.parametrize("S", [64, 128, 192, 256])
@pytest.mark.parametrize("n", [1, 7, 1000, 40001])
def test_fused_compositing_is_bit_identical_to_the_two_step_route(E, O, S, n):
    sdc, _ = O.nerf_state_dicts(0)
    net = load_nerf(E, sdc)
    o, d, v, z = rays(n, S, seed=100 * S + n)
    for white in (False, True):
        fused, twostep, plain, n_flag = both_routes(E, net, o, d, v, z, white)
        names = ("rgb_map", "disp_map", "acc_map", "weights", "depth_map")
        for name, f, t, pl in zip(names, fused, twostep, plain):
            assert same(f, t), (name, S, n, white, float((f - t).abs().max()))
            assert same(t, pl), (name, "two-step vs forward_samples + raw2outputs")
    if n >= 1000:
        assert n_flag > 0     # random-init nets flag ~2 % of the rays: the re-composited rays are part of the comparison


def test_fused_route_launches_no_raw2outputs_kernel(E, O):
    sdc, _ = O.nerf_state_dicts(0)
    net = load_nerf(E, sdc)
    o, d, v, z = rays(5000, 64, seed=3)
    with torch.no_grad():
        net.render_samples(o, d, v, z, True)            # packs the weights, grows the workspaces
        net.set_fused_compositing(True)
        k0 = E._lib.kernel_launches()
        net.render_samples(o, d, v, z, True)
        k_fused = E._lib.kernel_launches() - k0
        net.set_fused_compositing(False)
        net.render_samples(o, d, v, z, True)
        k0 = E._lib.kernel_launches()
        net.render_samples(o, d, v, z, True)
        k_two = E._lib.kernel_launches() - k0
        net.set_fused_compositing(True)
    # view bias + MLP (+ compositor inside) + far fix-up + the flagged rays' composite  vs  view bias + MLP + fix-up + raw2outputs
    assert k_fused == 4 and k_two == 4, (k_fused, k_two)


def test_whole_frame_fused_equals_two_step(E, O):
    """render() of a 400x400 lego frame, coarse 64 + fine 192 samples: every returned map bit-identical, and the
    fine pass sees bit-identical depths (weights0 -> hier_sample) because rgb0 / z_std agree."""
    sdc, sdf = O.nerf_state_dicts(0)
    coarse, fine = load_nerf(E, sdc), load_nerf(E, sdf)
    cam = O.LEGO
    c2w = O.pose_spherical(-60., -30., 4.)[:3, :4].cuda()
    kw = dict(network_query_fn=None, perturb=0., N_importance=128, network_fine=fine, N_samples=64, network_fn=coarse,
              use_viewdirs=True, white_bkgd=True, raw_noise_std=0., ndc=False, near=2., far=6.)
    out = {}
    with torch.no_grad():
        for mode in (True, False):
            coarse.set_fused_compositing(mode), fine.set_fused_compositing(mode)
            rgb, disp, acc, extras = E.render_image(cam["H"], cam["W"], cam["focal"], chunk=32768, c2w=c2w, **kw)
            out[mode] = dict(rgb=rgb, disp=disp, acc=acc, **{k: extras[k] for k in ("rgb0", "disp0", "acc0", "z_std")})
        coarse.set_fused_compositing(True), fine.set_fused_compositing(True)
    for k in out[True]:
        assert same(out[True][k], out[False][k]), k
    assert out[True]["rgb"].shape == (cam["H"], cam["W"], 3)


def test_fern_ndc_frame_fused_equals_two_step(E, O):
    sdc, sdf = O.nerf_state_dicts(0)
    coarse, fine = load_nerf(E, sdc), load_nerf(E, sdf)
    cam = O.FERN
    c2w = torch.tensor([[0.99, 0.01, -0.1, 0.3], [-0.02, 0.995, 0.05, -0.2], [0.1, -0.05, 0.99, 0.1]]).cuda()
    kw = dict(network_query_fn=None, perturb=0., N_importance=64, network_fine=fine, N_samples=64, network_fn=coarse,
              use_viewdirs=True, white_bkgd=False, raw_noise_std=0., ndc=True, near=0., far=1.)
    out = {}
    with torch.no_grad():
        for mode in (True, False):
            coarse.set_fused_compositing(mode), fine.set_fused_compositing(mode)
            rgb, disp, acc, extras = E.render_image(cam["H"], cam["W"], cam["focal"], chunk=32768, c2w=c2w, **kw)
            out[mode] = dict(rgb=rgb, disp=disp, acc=acc, rgb0=extras["rgb0"], z_std=extras["z_std"])
        coarse.set_fused_compositing(True), fine.set_fused_compositing(True)
    for k in out[True]:
        assert same(out[True][k], out[False][k]), k


@pytest.mark.parametrize("S", [2, 33, 65, 96, 160])
def test_render_samples_other_sample_counts_take_the_two_step_route(E, O, S):
    """Sample counts without a fused shape: r2l_nerf_render falls back to MLP -> workspace -> raw2outputs in BOTH modes."""
    sdc, _ = O.nerf_state_dicts(0)
    net = load_nerf(E, sdc)
    o, d, v, z = rays(777, S, seed=S)
    fused, twostep, plain, _ = both_routes(E, net, o, d, v, z, True)
    for f, t, pl in zip(fused, twostep, plain):
        assert same(f, t) and same(t, pl)


def test_fused_compositing_without_the_far_fixup_and_without_optional_outputs(E, O):
    sdc, _ = O.nerf_state_dicts(0)
    net = load_nerf(E, sdc)
    o, d, v, z = rays(3001, 192, seed=11)
    with torch.no_grad():
        net.set_far_fixup(False)
        fused, twostep, plain, n_flag = both_routes(E, net, o, d, v, z, False)
        for f, t in zip(fused, twostep):
            assert same(f, t)
        # no weights wanted (the fine pass): the other maps do not change
        net.set_fused_compositing(True)
        rgb, disp, acc, w, depth = net.render_samples(o, d, v, z, False, want_weights=False)
        assert w is None and same(rgb, fused[0]) and same(disp, fused[1]) and same(acc, fused[2]) and same(depth, fused[4])
        net.set_far_fixup(True)
        net.set_fused_compositing(False)


def test_render_samples_empty_batch(E, O):
    sdc, _ = O.nerf_state_dicts(0)
    net = load_nerf(E, sdc)
    o, d, v, z = rays(0, 64, seed=1)
    with torch.no_grad():
        rgb, disp, acc, w, depth = net.render_samples(o, d, v, z, True)
    assert rgb.shape == (0, 3) and w.shape == (0, 64) and depth.shape == (0,)
