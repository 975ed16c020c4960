"""GPU parity tests for the MLP kernels (tcgen05 fused NeRF / R2L and the fp32 CUDA-core path) and the
fused render path, against the CPU oracle and the golden vectors produced by the reference's own code.

Tolerances (north_star): rgb maps within 2e-3 max-abs of the reference, PSNR delta <= 0.05 dB;
sample_pdf bin indices bit-exact when fed the reference's weights (tests/test_gpu_ops.py)."""
import numpy as np
import pytest
import torch

from conftest import t

pytestmark = pytest.mark.gpu

RGB_TOL = 2e-3


def maxabs(a, b):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu() if isinstance(b, torch.Tensor) else t(b)
    return float((a - b).abs().max())


def load_nerf(E, sd, precision):
    net = E.NeRF(8, 256, 63, 27, 5, [4], True, precision=precision)
    net.load_state_dict(sd)
    return net.cuda().eval()


def load_r2l(E, O, sd, precision):
    net = E.NeRF_v3_2(O.r2l_args(), 1008, 3, precision=precision)
    net.load_state_dict(sd)
    return net.cuda().eval()


# ------------------------------------------------------------------ building blocks
@pytest.mark.parametrize("dtype", [0, 1])
@pytest.mark.parametrize("N,K", [(256, 64), (256, 256), (128, 320), (128, 256), (32, 32), (64, 96)])
def test_tcgen05_gemm_probe(E, dtype, N, K):
    """One tcgen05 GEMM with the operand layout / descriptors the fused kernels use."""
    torch.manual_seed(N + K)
    A = torch.randn(128, K, device="cuda")
    W = torch.randn(N, K, device="cuda")
    D = torch.full((128, N), float("nan"), device="cuda")
    E._lib.call("r2l_tc_gemm_probe", dtype, N, K, E._lib.ptr(A), E._lib.ptr(W), E._lib.ptr(D), 0, E._lib.stream_ptr())
    torch.cuda.synchronize()
    cast = torch.bfloat16 if dtype == 1 else torch.float16
    ref = A.to(cast).double() @ W.to(cast).double().t()
    err = float((D.double() - ref).abs().max())
    assert err < 1e-3 * K**0.5, f"probe N={N} K={K} dtype={dtype}: max err {err}"


@pytest.mark.parametrize("dtype", [0, 1])
@pytest.mark.parametrize("N,K", [(256, 64), (256, 256), (128, 256), (128, 32), (64, 96)])
def test_tcgen05_gemm_probe_cta_pair(E, dtype, N, K):
    """tcgen05.mma.cta_group::2: M = 256 over two CTAs, each supplying its A rows and one N-half of B."""
    torch.manual_seed(N * 3 + K)
    A = torch.randn(256, K, device="cuda")
    W = torch.randn(N, K, device="cuda")
    D = torch.full((256, N), float("nan"), device="cuda")
    E._lib.call("r2l_tc_gemm_probe_pair", dtype, N, K, E._lib.ptr(A), E._lib.ptr(W), E._lib.ptr(D), E._lib.stream_ptr())
    torch.cuda.synchronize()
    cast = torch.bfloat16 if dtype == 1 else torch.float16
    ref = A.to(cast).double() @ W.to(cast).double().t()
    err = float((D.double() - ref).abs().max())
    assert err < 1e-3 * K**0.5, f"pair probe N={N} K={K} dtype={dtype}: max err {err}"


def test_linear_fp32(E):
    from efficient_nerf_b200.nerf_raybased import _linear_fp32
    torch.manual_seed(0)
    for M, N, K in ((1000, 256, 63), (513, 128, 283), (300, 3, 128), (257, 1, 256), (4096, 256, 1008)):
        x, w, b = torch.randn(M, K), torch.randn(N, K) / K**0.5, torch.randn(N)
        r = torch.randn(M, N)
        for act, fn in ((0, lambda v: v), (1, torch.relu), (2, torch.sigmoid)):
            y = _linear_fp32(x.cuda(), w.cuda(), b.cuda(), act)
            assert maxabs(y, fn(torch.nn.functional.linear(x.double(), w.double(), b.double())).float()) < 2e-5
        y = _linear_fp32(x.cuda(), w.cuda(), b.cuda(), 0, residual=r.cuda(), scale=0.5)
        assert maxabs(y, (torch.nn.functional.linear(x, w, b) * 0.5 + r)) < 2e-5


# ------------------------------------------------------------------ NeRF
def test_module_parity_with_reference_layout(E, O):
    """Same state_dict keys / shapes / seeded init as the reference modules (checkpoints are drop-in)."""
    sdc, sdf = O.nerf_state_dicts(0)
    torch.manual_seed(0)
    c = E.NeRF(8, 256, 63, 27, 5, [4], True)
    f = E.NeRF(8, 256, 63, 27, 5, [4], True)
    for net, sd in ((c, sdc), (f, sdf)):
        msd = net.state_dict()
        assert set(msd) == set(sd)
        assert all(torch.equal(msd[k], sd[k]) for k in sd)
    sd = O.r2l_state_dict(0)
    torch.manual_seed(0)
    r = E.NeRF_v3_2(O.r2l_args(), 1008, 3)
    msd = r.state_dict()
    assert set(msd) == set(sd) and all(torch.equal(msd[k], sd[k]) for k in sd)


def test_nerf_forward_fp32_path(E, O):
    sdc, _ = O.nerf_state_dicts(0)
    net = load_nerf(E, sdc, "fp32")
    torch.manual_seed(1)
    x = torch.cat([O.embed_nerf((torch.rand(700, 3) * 2 - 1) * 4, 10),
                   O.embed_nerf(torch.nn.functional.normalize(torch.randn(700, 3), dim=-1), 4)], -1)
    with torch.no_grad():
        ref = O.nerf_forward(sdc, x)
        out = net(x.cuda())
    assert maxabs(out, ref) < 1e-5


@pytest.mark.parametrize("precision,tol", [("fp16", 2e-3), ("bf16", 2e-2)])
def test_nerf_forward_tensor_core(E, O, precision, tol):
    """NeRF.forward(x) API path on tcgen05 vs the fp32 oracle (raw network outputs, pre-compositing)."""
    sdc, _ = O.nerf_state_dicts(0)
    net = load_nerf(E, sdc, precision)
    torch.manual_seed(2)
    for M in (128, 1000, 129):
        x = torch.cat([O.embed_nerf((torch.rand(M, 3) * 2 - 1) * 4, 10),
                       O.embed_nerf(torch.nn.functional.normalize(torch.randn(M, 3), dim=-1), 4)], -1)
        with torch.no_grad():
            ref = O.nerf_forward(sdc, x)
            out = net(x.cuda())
        torch.cuda.synchronize()
        assert out.shape == ref.shape
        assert maxabs(out, ref) < tol, (precision, M, maxabs(out, ref))


def test_nerf_forward_embedded_large_input_is_blocked(E, O):
    """NeRF.forward(x) on more rows than one 2^20-row block of the API path: same values as the rows alone."""
    sdc, _ = O.nerf_state_dicts(0)
    net = load_nerf(E, sdc, "fp16")
    torch.manual_seed(4)
    M = (1 << 20) + 777
    x = torch.randn(M, 90, device="cuda") * 0.5
    with torch.no_grad():
        y = net(x)
        assert y.shape == (M, 4) and bool(torch.isfinite(y).all())
        for lo, hi in ((0, 300), ((1 << 20) - 100, (1 << 20) + 300), (M - 200, M)):
            assert torch.equal(y[lo:hi], net(x[lo:hi].contiguous())), (lo, hi)


def test_nerf_fused_encode_matches_api_path(E, O):
    """forward_samples (encoding fused in-kernel) == forward(embedded) on the same points."""
    sdc, _ = O.nerf_state_dicts(0)
    net = load_nerf(E, sdc, "fp16")
    torch.manual_seed(3)
    N, S = 37, 64
    c2w = O.pose_spherical(40., -30., 4.)[:3, :4]
    ro, rd = O.get_rays(400, 400, O.LEGO["focal"], c2w)
    idx = torch.arange(0, 160000, 4321)[:N]
    ro, rd = ro.reshape(-1, 3)[idx].contiguous(), rd.reshape(-1, 3)[idx].contiguous()
    vd = rd / rd.norm(dim=-1, keepdim=True)
    z = torch.sort(torch.rand(N, S) * 4 + 2, -1)[0]
    with torch.no_grad():
        raw = net.forward_samples(ro.cuda(), rd.cuda(), vd.cuda(), z.cuda())
        pts = ro[:, None, :] + rd[:, None, :] * z[..., None]
        ref = O.run_network(pts, vd, sdc)
    assert maxabs(raw, ref) < 2e-3


def test_nerf_pingpong_kernel_vs_single_cta_kernel(E, O, monkeypatch):
    """The CTA-pair ping-pong kernel (default) against the single-CTA chasing kernel (R2L_NERF_PP=0) on the same
    inputs: same 16-bit operands, fp32 accumulation, so they agree far inside the 2e-3 gate (the view branch is
    evaluated in fp32 instead of 16 bit in the ping-pong kernel); ragged sizes exercise partial units / tiles.
    At full-frame size the ping-pong kernel must be bit-reproducible run to run (two concurrent issuer threads and a
    shared weight ring: a protocol race would show up as run-to-run differences)."""
    sdc, _ = O.nerf_state_dicts(0)
    monkeypatch.setenv("R2L_NERF_PP", "0")
    single = load_nerf(E, sdc, "fp16")
    h0 = single.packed_handle()
    monkeypatch.setenv("R2L_NERF_PP", "1")
    pp = load_nerf(E, sdc, "fp16")
    h1 = pp.packed_handle()
    assert h0 is not h1
    c2w = O.pose_spherical(10., -30., 4.)[:3, :4].cuda()
    ro, rd = E.get_rays(400, 400, O.LEGO["focal"], c2w)
    ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
    vd = E.normalize_dirs(rd)
    torch.manual_seed(5)
    with torch.no_grad():
        for N, S in ((1, 64), (3, 43), (517, 64), (2049, 192)):   # 1 row short of / beyond unit boundaries
            z = torch.sort(torch.rand(N, S, device="cuda") * 4 + 2, -1)[0]
            a = single.forward_samples(ro[:N], rd[:N], vd[:N], z)
            b = pp.forward_samples(ro[:N], rd[:N], vd[:N], z)
            assert a.shape == b.shape == (N, S, 4)
            assert maxabs(a, b) < 5e-4, (N, S, maxabs(a, b))
        N, S = 160000, 192
        z = torch.sort(torch.rand(N, S, device="cuda") * 4 + 2, -1)[0]
        ref = pp.forward_samples(ro, rd, vd, z).clone()
        assert bool(torch.isfinite(ref).all())
        for _ in range(3):
            assert torch.equal(pp.forward_samples(ro, rd, vd, z), ref)
        assert maxabs(single.forward_samples(ro, rd, vd, z), ref) < 5e-4


def render_golden(E, O, golden, name, precision):
    g = golden(name)
    sdc, sdf = O.nerf_state_dicts(0)
    coarse, fine = load_nerf(E, sdc, precision), load_nerf(E, sdf, precision)
    fern = name.endswith("fern")
    cam = O.FERN if fern else O.LEGO
    rays = torch.stack([t(g["rays_o"]), t(g["rays_d"])], 0).cuda()
    kw = dict(network_query_fn=None, perturb=0., N_importance=64 if fern else 128, network_fine=fine, N_samples=64,
              network_fn=coarse, use_viewdirs=True, white_bkgd=not fern, raw_noise_std=0., ndc=fern, lindisp=False,
              near=0. if fern else 2., far=1. if fern else 6., retraw=True)
    with torch.no_grad():
        rgb, disp, acc, extras = E.render_image(cam["H"], cam["W"], cam["focal"], chunk=32768, rays=rays, **kw)
    return g, rgb, disp, acc, extras


@pytest.mark.parametrize("name", ["nerf_render_lego", "nerf_render_fern"])
def test_render_rays_fused_vs_reference(E, O, golden, name):
    """Full hierarchical render (coarse 64 + fine 128|64) vs the reference's outputs: rgb within 2e-3."""
    g, rgb, disp, acc, extras = render_golden(E, O, golden, name, "fp16")
    assert set(extras) >= {"raw", "rgb0", "disp0", "acc0", "z_std"}
    assert maxabs(rgb, g["rgb_map"]) <= RGB_TOL, maxabs(rgb, g["rgb_map"])
    assert maxabs(extras["rgb0"], g["rgb0"]) <= RGB_TOL
    assert maxabs(acc, g["acc_map"]) <= RGB_TOL
    assert maxabs(extras["z_std"], g["z_std"]) <= 5e-3
    # PSNR delta against a common target (the reference frame of the OTHER config's pixels is not available
    # here; use a fixed pseudo-target so that the delta is defined): <= 0.05 dB
    torch.manual_seed(0)
    target = torch.rand_like(t(g["rgb_map"]))
    assert abs(O.psnr(rgb.cpu(), target) - O.psnr(t(g["rgb_map"]), target)) <= 0.05


def test_render_rays_fp32_path_matches_reference_tightly(E, O, golden):
    """precision='fp32' goes through network_query_fn (run_network + Embedder + NeRF.forward on CUDA cores):
    every stage then matches the reference to ~1e-5, which pins the glue (z_vals, compositing, sample_pdf,
    merge) independently of tensor-core rounding."""
    g = golden("nerf_render_lego")
    sdc, sdf = O.nerf_state_dicts(0)
    coarse, fine = load_nerf(E, sdc, "fp32"), load_nerf(E, sdf, "fp32")
    embed_fn, _ = E.get_embedder(10, 0)
    embeddirs_fn, _ = E.get_embedder(4, 0)
    nq = lambda i, v, f: E.run_network(i, v, f, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn, netchunk=65536)
    rays = torch.stack([t(g["rays_o"]), t(g["rays_d"])], 0).cuda()
    kw = dict(network_query_fn=nq, perturb=0., N_importance=128, network_fine=fine, N_samples=64, network_fn=coarse,
              use_viewdirs=True, white_bkgd=True, raw_noise_std=0., ndc=False, lindisp=False, near=2., far=6.,
              retraw=True)
    with torch.no_grad():
        rgb, disp, acc, extras = E.render_image(400, 400, O.LEGO["focal"], chunk=32768, rays=rays, **kw)
    assert maxabs(extras["rgb0"], g["rgb0"]) < 2e-5
    assert maxabs(rgb, g["rgb_map"]) < 1e-4
    assert maxabs(acc, g["acc_map"]) < 1e-4


def test_render_stochastic_path_with_injected_draws(E, O, golden):
    """create_data flavour (perturb=1, raw_noise_std=1): same CPU-generator draws as the reference."""
    g = golden("nerf_render_lego")
    gp = golden("nerf_render_perturb")
    sdc, sdf = O.nerf_state_dicts(0)
    coarse, fine = load_nerf(E, sdc, "fp16"), load_nerf(E, sdf, "fp16")
    batch = O.pack_rays(t(g["rays_o"])[:16], t(g["rays_d"])[:16], 2., 6.)
    torch.manual_seed(int(gp["seed"]))
    with torch.no_grad():
        ret = E.render_rays_create_data(batch.cuda(), coarse, None, 64, perturb=1., N_importance=128,
                                        network_fine=fine, white_bkgd=True, raw_noise_std=1.0)
    assert "depth_map" in ret
    # sigma noise of std 1 on random-init (tiny) densities dominates the image, so agreement here
    # shows that the draws were consumed in the reference's order
    assert maxabs(ret["rgb_map"], gp["rgb_map"]) <= 5e-3
    assert maxabs(ret["z_std"], gp["z_std"]) <= 5e-3


def test_render_chunk_invariance_and_tail_tiles(E, O):
    sdc, sdf = O.nerf_state_dicts(0)
    coarse, fine = load_nerf(E, sdc, "fp16"), load_nerf(E, sdf, "fp16")
    c2w = O.pose_spherical(10., -30., 4.)[:3, :4]
    ro, rd = O.get_rays(400, 400, O.LEGO["focal"], c2w)
    idx = torch.arange(0, 160000, 531)[:301]   # 301 rays: 64*301 and 192*301 are not multiples of 128
    batch = O.pack_rays(ro.reshape(-1, 3)[idx], rd.reshape(-1, 3)[idx], 2., 6.).cuda()
    kw = dict(network_query_fn=None, N_samples=64, N_importance=128, network_fine=fine, white_bkgd=True)
    with torch.no_grad():
        a = E.render_rays(batch, coarse, **kw)
        b1 = E.render_rays(batch[:100], coarse, **kw)
        b2 = E.render_rays(batch[100:], coarse, **kw)
    for k in ("rgb_map", "disp_map", "acc_map", "rgb0", "z_std"):
        assert torch.equal(a[k], torch.cat([b1[k], b2[k]], 0)), k   # independent of chunking, bit for bit
    assert bool(torch.isfinite(a["rgb_map"]).all())


# ------------------------------------------------------------------ R2L
def test_r2l_fp32_path(E, O, golden):
    g = golden("r2l_lego")
    sd = O.r2l_state_dict(0)
    net = load_r2l(E, O, sd, "fp32")
    pe = E.PositionalEmbedder(L=10)
    with torch.no_grad():
        rgb = net(pe(t(g["pts"]).cuda()))
    assert maxabs(rgb, g["rgb"]) < 2e-5


@pytest.mark.parametrize("precision,tol", [("fp16", RGB_TOL), ("bf16", 1.5e-2)])
def test_r2l_fused_vs_reference(E, O, golden, precision, tol):
    """PointSampler -> (fused PositionalEmbedder + W256 D88 ResMLP) vs the reference's rgb."""
    g = golden("r2l_lego")
    sd = O.r2l_state_dict(0)
    net = load_r2l(E, O, sd, precision)
    pts = t(g["pts"]).cuda()
    with torch.no_grad():
        rgb = net.forward_points(pts)
        torch.cuda.synchronize()
        assert maxabs(rgb, g["rgb"]) <= tol, (precision, maxabs(rgb, g["rgb"]))
        # API path: model(positional_embedder(pts)) with a materialised embedding, and the lazy handle
        pe = E.PositionalEmbedder(L=10)
        rgb_api = net(pe(pts))
        assert maxabs(rgb_api, g["rgb"]) <= tol
        rgb_lazy = net(E.PositionalEmbedder(L=10, lazy=True)(pts))
        assert torch.equal(rgb_lazy, rgb)
        # ragged batch (tail tile) and a single ray
        rgb_r = net.forward_points(pts[:77])
        assert torch.equal(rgb_r, rgb[:77])
        assert torch.equal(net.forward_points(pts[5:6]), rgb[5:6])


@pytest.mark.parametrize("n_points", [4, 8, 16])
def test_r2l_head_accumulators_and_packed_weights(E, O, n_points, monkeypatch):
    """Debug hooks: the packed weight stream is the expected permutation, and the head layer's raw tcgen05
    accumulators equal the fp16-rounded reference product (K = 64 per point, accumulated over chunks).
    Pinned to the single-CTA stream layout (R2L_PAIR=0); the CTA-pair layout (two N-halves per stage) is covered by
    test_tcgen05_gemm_probe_cta_pair and by every R2L parity test, which run the default pair kernel."""
    import ctypes
    monkeypatch.setenv("R2L_PAIR", "0")
    L = E._lib
    torch.manual_seed(n_points)
    net = E.NeRF_v3_2(O.r2l_args(netdepth=6), n_points * 63, 3, precision="fp16").cuda().eval()
    pts = ((torch.rand(200, n_points * 3) * 2 - 1) * 4).cuda()
    h = net.packed_handle()
    rgb = torch.empty(200, 3, device="cuda")
    acc = torch.zeros(256, 256, device="cuda")
    x0 = torch.zeros(256, 256, device="cuda")
    L.call("r2l_resmlp_debug_head", h.h, 200, L.ptr(pts), pts.stride(0), L.ptr(rgb), L.ptr(acc), L.ptr(x0), None,
           L.stream_ptr())
    torch.cuda.synchronize()
    x = O.embed_r2l(pts.cpu(), 10)
    W, b = net.head[0].weight.detach().cpu(), net.head[0].bias.detach().cpu()
    # the bias is folded into the MMA (hi/lo 16-bit split x 1.0), so the accumulators already include it
    acc_ref = (torch.nn.functional.linear(x.half().double(), W.half().double()) + b.double()).float()
    assert maxabs(acc[:200], acc_ref) < 2e-3
    assert maxabs(x0[:200], torch.relu(acc_ref)) < 2e-3
    with torch.no_grad():
        assert maxabs(rgb, net._forward_fp32(x.cuda())) < 2e-3
    n = ctypes.c_ulonglong(0)
    L.call("r2l_mlp_debug_wstream", h.h, None, 0, ctypes.byref(n))
    # stream = per layer [K=16 bias stage][K/32 weight stages]: head + 2 blocks x 2 layers
    assert n.value == 2 * (256 * 16 + 256 * n_points * 64 + 2 * 2 * (256 * 16 + 256 * 256))
    buf = np.zeros(n.value // 2, dtype=np.float16)
    L.call("r2l_mlp_debug_wstream", h.h, buf.ctypes.data_as(ctypes.c_void_p), n.value, ctypes.byref(n))
    K = n_points * 64
    bias_stage = torch.from_numpy(buf[:256 * 16]).float().reshape(2, 256, 8)     # [k-chunk][n][k%8]
    assert maxabs(bias_stage[0, :, 0] + bias_stage[0, :, 1], b) < 1e-6              # hi + lo == bias
    assert float(bias_stage[0, :, 2:].abs().max()) == 0. and float(bias_stage[1].abs().max()) == 0.
    Wp = torch.from_numpy(buf[256 * 16:256 * 16 + 256 * K]).float().reshape(K // 64, 8, 256, 8).permute(2, 0, 1, 3)
    Wp = Wp.reshape(256, K)
    for s_ in range(n_points):          # block order: k = 64 s + i ; i<3 identity, then (sin, cos) per frequency
        for c in range(3):
            assert torch.equal(Wp[:, 64 * s_ + c], W[:, (3 * s_ + c) * 21 + 20].half().float())
            assert torch.equal(Wp[:, 64 * s_ + 3 + 6 * 4 + c], W[:, (3 * s_ + c) * 21 + 4].half().float())
            assert torch.equal(Wp[:, 64 * s_ + 3 + 6 * 4 + 3 + c], W[:, (3 * s_ + c) * 21 + 14].half().float())
        assert float(Wp[:, 64 * s_ + 63].abs().max()) == 0.


def test_r2l_full_frame_properties(E, O):
    """BASELINE size: one 400x400 frame in one un-chunked forward; sharded == unsharded."""
    sd = O.r2l_state_dict(0)
    net = load_r2l(E, O, sd, "fp16")
    c2w = O.pose_spherical(-180., -30., 4.)[:3, :4].cuda()
    ps = E.PointSampler(400, 400, O.LEGO["focal"], 16, 2., 6.)
    with torch.no_grad():
        rgb = E.render_r2l(net, ps, c2w)
        assert tuple(rgb.shape) == (160000, 3)
        assert bool(torch.isfinite(rgb).all()) and float(rgb.min()) >= 0. and float(rgb.max()) <= 1.
        pts = ps.sample_test(c2w)
        parts = []
        for r in range(8):
            s0, s1 = E.sharding.shard_rays(160000, r, 8)
            parts.append(net.forward_points(pts[s0:s1]))
        assert torch.equal(torch.cat(parts, 0), rgb)
        ref = O.render_r2l(sd, 400, 400, O.LEGO["focal"], 2., 6., c2w.cpu(), rows=torch.arange(0, 160000, 1601))
    assert maxabs(rgb[::1601], ref) <= RGB_TOL


def test_nerf_full_frame_properties(E, O):
    """BASELINE config 1 at full size (160 000 rays, 64 + 128 samples)."""
    sdc, sdf = O.nerf_state_dicts(0)
    coarse, fine = load_nerf(E, sdc, "fp16"), load_nerf(E, sdf, "fp16")
    c2w = O.pose_spherical(-180., -30., 4.)[:3, :4].cuda()
    kw = dict(network_query_fn=None, perturb=0., N_importance=128, network_fine=fine, N_samples=64, network_fn=coarse,
              use_viewdirs=True, white_bkgd=True, raw_noise_std=0., ndc=False, near=2., far=6.)
    with torch.no_grad():
        rgb, disp, acc, extras = E.render_image(400, 400, O.LEGO["focal"], chunk=32768, c2w=c2w, **kw)
    assert tuple(rgb.shape) == (400, 400, 3) and tuple(extras["z_std"].shape) == (400, 400)
    assert bool(torch.isfinite(rgb).all()) and float(rgb.min()) >= -1e-6 and float(rgb.max()) <= 1. + 1e-5
    assert float(acc.max()) <= 1. + 1e-5


# ------------------------------------------------------------------ stress / larger configs
def test_fused_kernels_are_deterministic_under_repetition(E, O):
    """The persistent kernels synchronise through ~40 mbarriers; a protocol race would show up as run-to-run
    differences.  Full-size launches repeated back to back must be bit-identical."""
    sd = O.r2l_state_dict(0)
    net = load_r2l(E, O, sd, "fp16")
    ps = E.PointSampler(400, 400, O.LEGO["focal"], 16, 2., 6.)
    pts = ps.sample_test(O.pose_spherical(33., -30., 4.)[:3, :4].cuda())
    sdc, _ = O.nerf_state_dicts(0)
    nerf = load_nerf(E, sdc, "fp16")
    ro, rd = E.get_rays(400, 400, O.LEGO["focal"], O.pose_spherical(33., -30., 4.)[:3, :4].cuda())
    ro, rd = ro.reshape(-1, 3)[:40000].contiguous(), rd.reshape(-1, 3)[:40000].contiguous()
    vd = E.normalize_dirs(rd)
    torch.manual_seed(5)
    z = torch.sort(torch.rand(40000, 192, device="cuda") * 4 + 2, -1)[0]
    with torch.no_grad():
        ref = net.forward_points(pts)
        for _ in range(30):
            assert torch.equal(net.forward_points(pts), ref)
        raw = nerf.forward_samples(ro, rd, vd, z)
        for _ in range(6):
            assert torch.equal(nerf.forward_samples(ro, rd, vd, z), raw)
    assert bool(torch.isfinite(raw).all())


def test_r2l_800x800_frame(E, O):
    """BASELINE configs[3]: lego_noview_800x800 — 640 000 rays in one un-chunked forward."""
    sd = O.r2l_state_dict(0)
    net = load_r2l(E, O, sd, "fp16")
    cam = O.LEGO_800
    c2w = O.pose_spherical(120., -30., 4.)[:3, :4]
    ps = E.PointSampler(cam["H"], cam["W"], cam["focal"], 16, 2., 6.)
    rows = torch.arange(0, 640000, 6397)
    with torch.no_grad():
        rgb = E.render_r2l(net, ps, c2w.cuda())
        ref = O.render_r2l(sd, cam["H"], cam["W"], cam["focal"], 2., 6., c2w, rows=rows)
    assert tuple(rgb.shape) == (640000, 3) and bool(torch.isfinite(rgb).all())
    assert maxabs(rgb[rows.cuda()], ref) <= RGB_TOL


def test_nerf_800x800_rays_vs_oracle(E, O):
    """BASELINE configs[3]: lego_800x800 NeRF — a strided subset of the 800x800 frame against the oracle, and the
    whole frame for finiteness / range (640 000 rays x 192 samples in one launch)."""
    sdc, sdf = O.nerf_state_dicts(0)
    coarse, fine = load_nerf(E, sdc, "fp16"), load_nerf(E, sdf, "fp16")
    cam = O.LEGO_800
    c2w = O.pose_spherical(-60., -30., 4.)[:3, :4]
    kw = dict(network_query_fn=None, perturb=0., N_importance=128, network_fine=fine, N_samples=64, network_fn=coarse,
              use_viewdirs=True, white_bkgd=True, raw_noise_std=0., ndc=False, near=2., far=6.)
    ro, rd = O.get_rays(cam["H"], cam["W"], cam["focal"], c2w)
    idx = torch.arange(0, 640000, 10007)
    batch = O.pack_rays(ro.reshape(-1, 3)[idx], rd.reshape(-1, 3)[idx], 2., 6.)
    with torch.no_grad():
        ref = O.render_rays(batch, sdc, sdf, 64, 128, white_bkgd=True)
        rgb, disp, acc, extras = E.render_image(cam["H"], cam["W"], cam["focal"], chunk=32768, c2w=c2w.cuda(), **kw)
    assert tuple(rgb.shape) == (800, 800, 3) and bool(torch.isfinite(rgb).all())
    assert maxabs(rgb.reshape(-1, 3)[idx.cuda()], ref["rgb_map"]) <= RGB_TOL


def test_r2l_pair_kernel_vs_single_cta_kernel(E, O, monkeypatch):
    """The CTA-pair R2L kernel (default) against the single-CTA kernel (R2L_PAIR=0): the same math on the same 16-bit
    operands (M = 256 vs 128 tiles only), so they agree to accumulation-order noise; ragged ray counts cover a
    missing peer tile; full-frame reruns must be bit-reproducible."""
    sd = O.r2l_state_dict(0)
    monkeypatch.setenv("R2L_PAIR", "0")
    single = load_r2l(E, O, sd, "fp16")
    single.packed_handle()
    monkeypatch.setenv("R2L_PAIR", "1")
    pair = load_r2l(E, O, sd, "fp16")
    pair.packed_handle()
    ps = E.PointSampler(400, 400, O.LEGO["focal"], 16, 2., 6.)
    pts = ps.sample_test(O.pose_spherical(20., -30., 4.)[:3, :4].cuda())
    with torch.no_grad():
        for n in (1, 127, 129, 257, 160000):
            a, b = single.forward_points(pts[:n]), pair.forward_points(pts[:n])
            assert a.shape == b.shape == (n, 3) and maxabs(a, b) < 2e-4, (n, maxabs(a, b))
        ref = pair.forward_points(pts).clone()
        for _ in range(3):
            assert torch.equal(pair.forward_points(pts), ref)


def test_r2l_pose_batch_equals_single_poses(E, O):
    """PointSampler.sample_test_batch / render_r2l with a stack of poses: bit-identical to one pose per call."""
    sd = O.r2l_state_dict(0)
    net = load_r2l(E, O, sd, "fp16")
    Hh, Ww = 30, 50   # 1500 rays per frame: not a multiple of the 128-ray tile, so frames share tiles
    ps = E.PointSampler(Hh, Ww, 60., 16, 2., 6.)
    poses = torch.stack([O.pose_spherical(th, -30., 4.)[:3, :4] for th in (-120., -10., 75.)], 0).cuda()
    with torch.no_grad():
        pb = ps.sample_test_batch(poses)
        p1 = torch.cat([ps.sample_test(p) for p in poses], 0)
        assert pb.shape == (3 * Hh * Ww, 48) and torch.equal(pb, p1)
        fb = E.render_r2l(net, ps, poses)
        f1 = torch.cat([E.render_r2l(net, ps, p) for p in poses], 0)
    assert fb.shape == (3 * Hh * Ww, 3) and torch.equal(fb, f1)
    assert ps.sample_test_batch(poses[:0]).shape == (0, 48)


def test_r2l_cuda_graph_replay_equals_eager(E, O):
    """GraphedR2L: the sampler + fused MLP captured into a CUDA graph; replays with new poses reproduce the eager
    frames bit for bit (1 pose and a stack of 3 poses per replay)."""
    sd = O.r2l_state_dict(0)
    net = load_r2l(E, O, sd, "fp16")
    Hh, Ww = 40, 56
    ps = E.PointSampler(Hh, Ww, 70., 16, 2., 6.)
    poses = [O.pose_spherical(th, -30., 4.)[:3, :4].cuda() for th in (-150., -20., 15., 99., 140., 171.)]
    with torch.no_grad():
        g1 = E.GraphedR2L(net, ps, 1)
        for p in poses[:3]:
            assert torch.equal(g1(p), E.render_r2l(net, ps, p))
        g3 = E.GraphedR2L(net, ps, 3)
        for k in (0, 3):
            stack = torch.stack(poses[k:k + 3], 0)
            assert torch.equal(g3(stack), E.render_r2l(net, ps, stack))
        with pytest.raises(ValueError):
            g3(poses[0])



def test_r2l_ray_generation_fused_into_the_kernel(E, O):
    """r2l_resmlp_render (rays from pixel index + pose inside the MLP kernel, no points tensor) against
    r2l_point_sample + r2l_resmlp_forward, bit for bit: whole frames whose size is not a multiple of the 128-ray tile,
    pose stacks, arbitrary row ranges, 4x4 poses, and the reference call pattern through the lazy handles."""
    sd = O.r2l_state_dict(0)
    net = load_r2l(E, O, sd, "fp16")
    Hh, Ww = 37, 41
    ps = E.PointSampler(Hh, Ww, 55., 16, 2., 6.)
    poses = torch.stack([O.pose_spherical(th, -30., 4.)[:3, :4] for th in (-120., -10., 75.)], 0).cuda()
    n = Hh * Ww
    with torch.no_grad():
        for p in poses:
            want = net.forward_points(ps.sample_test(p))
            assert torch.equal(net.render_poses(ps, p), want)
            p44 = torch.cat([p, torch.tensor([[0., 0., 0., 1.]], device="cuda")], 0)
            assert torch.equal(net.render_poses(ps, p44), want)
        want3 = net.forward_points(ps.sample_test_batch(poses))
        assert torch.equal(net.render_poses(ps, poses), want3)
        for r0, r1 in ((0, 1), (1000, 1517), (n - 5, n + 300), (2 * n + 7, 3 * n), (5, 5)):
            assert torch.equal(net.render_poses(ps, poses, rows=(r0, r1)), want3[r0:r1])
        with pytest.raises(ValueError):
            net.render_poses(ps, poses, rows=(0, 3 * n + 1))
        with pytest.raises(ValueError):
            net.render_poses(E.PointSampler(Hh, Ww, 55., 8, 2., 6.), poses)
        # main.py:297-309 through the lazy sampler / embedder: one kernel, same bits; the handles still behave like tensors
        lps = E.PointSampler(Hh, Ww, 55., 16, 2., 6., lazy=True)
        emb = E.PositionalEmbedder(10, lazy=True)
        lp = lps.sample_test(poses[1])
        assert type(lp).__name__ == "LazyPoints" and tuple(lp.shape) == (n, 48)
        k0 = E._lib.kernel_launches()
        out = net(emb(lp))
        assert E._lib.kernel_launches() - k0 == 1
        assert torch.equal(out, net.forward_points(ps.sample_test(poses[1])))
        assert torch.equal(torch.add(lp, 0.), ps.sample_test(poses[1])) and torch.equal(lp[5:9], ps.sample_test(poses[1])[5:9])
        assert torch.equal(lp.view(n, 16, 3), ps.sample_test2(poses[1]))
        assert torch.equal(E.PositionalEmbedder(10)(lps.sample_test(poses[1])), E.PositionalEmbedder(10)(ps.sample_test(poses[1])))
        assert torch.equal(emb(lps.sample_test(poses[1])).materialize(), E.PositionalEmbedder(10)(ps.sample_test(poses[1])))


def test_r2l_pingpong_kernel_matches_the_default(E, O, monkeypatch):
    """mlp_r2l_pp.cu (R2L_PP=1: two 64-row tiles per CTA, tcgen05.mma.cta_group::2 with M = 128) runs the same MMAs in
    the same K order as the default pair kernel; only the fp32 tail sum (256 -> 3) is associated differently (four
    partial sums per row instead of two): equal to a few ulp, on full tiles, ragged tails and one ray."""
    sd = O.r2l_state_dict(0)
    ps = E.PointSampler(400, 400, O.LEGO["focal"], 16, 2., 6.)
    c2w = O.pose_spherical(75., -30., 4.)[:3, :4].cuda()
    outs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("R2L_PP", mode)      # read when the handle is created
        net = load_r2l(E, O, sd, "fp16")
        with torch.no_grad():
            pts = ps.sample_test(c2w)
            outs[mode] = [net.render_poses(ps, c2w, rows=(0, 40000)), net.forward_points(pts[1000:1000 + 777]),
                          net.forward_points(pts[5:6]), net(E.PositionalEmbedder(10)(pts[:300]))]
        torch.cuda.synchronize()
        assert net.range_status()[0] == 0
    for a, b in zip(outs["0"], outs["1"]):
        assert a.shape == b.shape and float((a - b).abs().max()) <= 1e-6
    ref = O.render_r2l(sd, 400, 400, O.LEGO["focal"], 2., 6., c2w.cpu(), rows=torch.arange(0, 40000, 401))
    assert maxabs(outs["1"][0][::401], ref) <= RGB_TOL
