"""fp16 operand range: weights are range-checked when packed; activations that leave the fp16
range turn into inf / NaN (the converts do not saturate), reach the output row, are detected there and reported as
R2L_ERR_RANGE by r2l_mlp_status and by the next call — never a silently clamped colour.  Inside the range, nets with
large high-frequency first-layer weights (what trained NeRFs look like) still track the fp32 path."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def nerf(E, O, precision, seed=0, mutate=None):
    sd = {k: v.clone() for k, v in O.nerf_state_dicts(seed)[0].items()}
    if mutate is not None:
        mutate(sd)
    net = E.NeRF(8, 256, 63, 27, 5, [4], True, precision=precision)
    net.load_state_dict(sd)
    return net.cuda().eval()


def rays(E, O, n=4096):
    c2w = O.pose_spherical(30., -30., 4.)[:3, :4].cuda()
    ro, rd = E.get_rays(400, 400, O.LEGO["focal"], c2w)
    idx = torch.arange(0, 160000, 160000 // n)[:n].cuda()
    ro, rd = ro.reshape(-1, 3)[idx].contiguous(), rd.reshape(-1, 3)[idx].contiguous()
    z = (torch.linspace(0., 1., 64).cuda() * 4. + 2.)[None].expand(n, 64).contiguous()
    return ro, rd, E.normalize_dirs(rd), z


def fp32_raw(E, net32, ro, rd, vd, z):
    pts = (ro[:, None] + rd[:, None] * z[..., None]).reshape(-1, 3)
    x = torch.cat([E.get_embedder(10, 0)[0](pts), E.get_embedder(4, 0)[0](vd[:, None].expand(-1, 64, -1).reshape(-1, 3))], -1)
    return net32(x).reshape(z.shape[0], 64, 4)


@pytest.mark.parametrize("gain", [1.0, 2.4])
def test_large_first_layer_weights_stay_close_to_fp32(E, O, gain):
    """pts_linears.0 rescaled to magnitude 50 (x400 over the random init): first hidden layer in the hundreds.  With
    PyTorch's default init every later layer shrinks the signal by ~0.41, so `gain` = 2.4 on layers 1..7 keeps the
    activations in the hundreds all the way to the heads (sigma ~ 1e2)."""
    def big_first(sd):
        w = sd["pts_linears.0.weight"]
        sd["pts_linears.0.weight"] = w * (50. / float(w.abs().max()))
        for l in range(1, 8):
            sd[f"pts_linears.{l}.weight"] = sd[f"pts_linears.{l}.weight"] * gain
    n16, n32 = nerf(E, O, "fp16", mutate=big_first), nerf(E, O, "fp32", mutate=big_first)
    ro, rd, vd, z = rays(E, O)
    with torch.no_grad():
        n16.set_far_fixup(False)
        raw = n16.forward_samples(ro, rd, vd, z)
        ref = fp32_raw(E, n32, ro, rd, vd, z)
    code, _ = n16.range_status()
    assert code == 0
    scale = float(ref.abs().max())
    err = float((raw - ref).abs().max())
    print(f"\n[range] |W0| = 50, gain {gain}: outputs up to {scale:.2f}, max |fused - fp32| = {err:.3e} "
          f"({err / scale:.2e} of the output scale)")
    assert err <= 2e-3 * max(1., scale)
    assert scale > (20. if gain > 1. else 0.1)
    rgb16, rgb32 = torch.sigmoid(raw[..., :3]), torch.sigmoid(ref[..., :3])
    d = float((rgb16 - rgb32).abs().max())
    print(f"        sigmoid(rgb): max |fused - fp32| = {d:.2e}")
    # gain 1: the 2e-3 gate holds.  gain 2.4 (pre-sigmoid colours ~ +-100, every layer in the hundreds): the 11-bit
    # operands give ~1e-3 of the output SCALE, which the sigmoid's slope turns into ~1e-2 near its middle — the regime a
    # split-operand (hi + lo) mode would be for; recorded here, not hidden
    assert d <= (2e-3 if gain == 1.0 else 3e-2)


def test_activation_overflow_is_reported_not_clamped(E, O):
    """every point layer x32: activations pass 65504 after a few layers"""
    def x32(sd):
        for k in sd:
            if k.startswith("pts_linears") and k.endswith("weight"):
                sd[k] = sd[k] * 32.
    n16 = nerf(E, O, "fp16", mutate=x32)
    ro, rd, vd, z = rays(E, O, 1024)
    with torch.no_grad():
        raw = n16.forward_samples(ro, rd, vd, z)
    assert not bool(torch.isfinite(raw).all())                 # inf / NaN, not 65504-clamped garbage
    code, rec = n16.range_status()
    assert code == 5 and rec[5] == 1
    with pytest.raises(RuntimeError, match="left the range of fp16"):
        with torch.no_grad():
            n16.forward_samples(ro, rd, vd, z)
    # reported once: the handle stays usable (the caller may have fixed its inputs)
    with torch.no_grad():
        n16.forward_samples(ro, rd, vd, z)
    # bf16 operands have fp32's range: the same model runs (at bf16 precision)
    nb = nerf(E, O, "bf16", mutate=x32)
    with torch.no_grad():
        rawb = nb.forward_samples(ro, rd, vd, z)
    assert bool(torch.isfinite(rawb).all()) and nb.range_status()[0] == 0


def test_r2l_overflow_is_reported(E, O):
    sd = {k: v.clone() for k, v in O.r2l_state_dict(0).items()}
    for k in sd:
        if k.endswith("weight") and k.startswith("body"):
            sd[k] = sd[k] * 8.
    net = E.NeRF_v3_2(O.r2l_args(), 1008, 3, precision="fp16")
    net.load_state_dict(sd)
    net = net.cuda().eval()
    ps = E.PointSampler(400, 400, O.LEGO["focal"], 16, 2., 6.)
    c2w = O.pose_spherical(0., -30., 4.)[:3, :4].cuda()
    with torch.no_grad():
        rgb = net.render_poses(ps, c2w, rows=(0, 4096))
    code, _ = net.range_status()
    assert code == 5
    with pytest.raises(RuntimeError, match="left the range"):
        with torch.no_grad():
            net.render_poses(ps, c2w, rows=(0, 4096))
    del rgb


def test_weights_beyond_fp16_are_rejected_at_pack_time(E, O):
    def huge(sd):
        sd["pts_linears.3.weight"][5, 7] = 1.0e5
    with pytest.raises(RuntimeError, match="does not fit fp16"):
        nerf(E, O, "fp16", mutate=huge).packed_handle()
    assert nerf(E, O, "bf16", mutate=huge).packed_handle() is not None
    def nan(sd):
        sd["pts_linears.1.bias"][0] = float("nan")
    with pytest.raises(RuntimeError, match="does not fit fp16"):
        nerf(E, O, "fp16", mutate=nan).packed_handle()
