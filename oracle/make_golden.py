"""Generate tests/golden/*.npz from the REAL reference (/root/reference) and pin the oracle to it.

Run in the build container (the reference is not present on the GPU box):
    python oracle/make_golden.py
What it does
  1. imports the reference's own model/nerf_raybased.py and utils/run_nerf_raybased_helpers.py
     unmodified, and AST-extracts batchify/run_network/batchify_rays/render/raw2outputs/render_rays
     from main.py (main.py itself cannot be imported: smilelogging is a dangling symlink);
  2. runs them on small seeded inputs and on strided subsets of the BASELINE configs;
  3. asserts that oracle/ref_torch.py (and oracle/sample_pdf_np.py) reproduce every output
     BIT-FOR-BIT — this is the "oracle pinned against the reference" check;
  4. writes the reference outputs as fixtures (a few hundred KB) with torch version / CPU capability.
TEST INFRASTRUCTURE ONLY.
"""
import ast
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

REF = os.environ.get("R2L_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import ref_torch as O  # noqa: E402
from oracle.sample_pdf_np import sample_pdf_np  # noqa: E402


def load_reference():
    sys.path.insert(0, REF)
    import model.nerf_raybased as M
    import utils.run_nerf_raybased_helpers as Hh
    torch.autograd.set_detect_anomaly(False)
    src = open(os.path.join(REF, "main.py")).read()
    ns = dict(torch=torch, np=np, F=F, device=torch.device("cpu"), to_tensor=Hh.to_tensor, get_rays=Hh.get_rays,
              ndc_rays=Hh.ndc_rays, sample_pdf=Hh.sample_pdf, DEBUG=False)
    want = {"batchify", "run_network", "batchify_rays", "render", "raw2outputs", "render_rays"}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in want:
            exec(compile(ast.Module([node], []), "main.py", "exec"), ns)
    sys.path.remove(REF)
    return M, Hh, ns


def eq(a, b, what):
    a = a.detach().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    if a.shape != b.shape or not np.array_equal(a, b, equal_nan=True):
        raise AssertionError(f"oracle != reference for {what}: max|d|={np.nanmax(np.abs(a - b))}")
    print(f"  oracle == reference (bit-exact): {what}")


def checksum(sd):
    return {k: float(v.double().abs().sum()) for k, v in sd.items()}


def main():
    os.makedirs(OUT, exist_ok=True)
    M, Hh, ns = load_reference()
    meta = dict(torch=torch.__version__, numpy=np.__version__, cpu=torch.backends.cpu.get_cpu_capability())
    print(meta)
    g = {}
    with torch.no_grad():
        # ---------------- rays: small camera, full grid ----------------
        c2w = O.pose_spherical(30., -30., 4.)[:3, :4]
        H, W, focal = 20, 30, 41.66666
        ro, rd = Hh.get_rays(H, W, focal, c2w)
        o_ro, o_rd = O.get_rays(H, W, focal, c2w)
        eq(o_ro, ro, "get_rays rays_o"), eq(o_rd, rd, "get_rays rays_d")
        no, nd = Hh.ndc_rays(H, W, focal, 1., ro, rd)
        o_no, o_nd = O.ndc_rays(H, W, focal, 1., o_ro, o_rd)
        eq(o_no, no, "ndc_rays o"), eq(o_nd, nd, "ndc_rays d")
        np.savez(os.path.join(OUT, "rays_small.npz"), c2w=c2w.numpy(), H=H, W=W, focal=focal, rays_o=ro.numpy(),
                 rays_d=rd.numpy(), ndc_o=no.numpy(), ndc_d=nd.numpy())
        # ---------------- rays: strided subsets of the real cameras ----------------
        for name, cam, ndc in (("lego", O.LEGO, False), ("fern", O.FERN, True)):
            c2w = O.pose_spherical(-60., -30., 4.)[:3, :4] if name == "lego" else torch.tensor(
                [[0.99, 0.01, -0.1, 0.3], [-0.02, 0.995, 0.05, -0.2], [0.1, -0.05, 0.99, 0.1]], dtype=torch.float32)
            ro, rd = Hh.get_rays(cam["H"], cam["W"], cam["focal"], c2w)
            o_ro, o_rd = O.get_rays(cam["H"], cam["W"], cam["focal"], c2w)
            eq(o_rd, rd, f"get_rays {name} full frame")
            idx = np.arange(0, cam["H"] * cam["W"], 997)
            d = dict(c2w=c2w.numpy(), idx=idx, rays_o=ro.reshape(-1, 3).numpy()[idx], rays_d=rd.reshape(-1, 3).numpy()[idx],
                     **{k: v for k, v in cam.items()})
            if ndc:
                no, nd = Hh.ndc_rays(cam["H"], cam["W"], cam["focal"], 1., ro, rd)
                o_no, o_nd = O.ndc_rays(cam["H"], cam["W"], cam["focal"], 1., o_ro, o_rd)
                eq(o_no, no, "ndc fern o"), eq(o_nd, nd, "ndc fern d")
                d.update(ndc_o=no.reshape(-1, 3).numpy()[idx], ndc_d=nd.reshape(-1, 3).numpy()[idx])
            np.savez(os.path.join(OUT, f"rays_{name}.npz"), **d)
        # ---------------- encodings ----------------
        torch.manual_seed(1)
        x = (torch.rand(64, 3) * 2 - 1) * 5.0
        e10, _ = Hh.get_embedder(10, 0)
        e4, _ = Hh.get_embedder(4, 0)
        r10, r4 = e10(x), e4(x / x.norm(dim=-1, keepdim=True))
        eq(O.embed_nerf(x, 10), r10, "Embedder L=10"), eq(O.embed_nerf(x / x.norm(dim=-1, keepdim=True), 4), r4,
                                                          "Embedder L=4")
        pts = (torch.rand(8, 48) * 2 - 1) * 6.0
        pe = M.PositionalEmbedder(L=10)
        rr = pe(pts)
        eq(O.embed_r2l(pts, 10), rr, "PositionalEmbedder L=10")
        ps = M.PointSampler(6, 8, 11.1, 16, 2., 6.)
        c2w = O.pose_spherical(100., -45., 4.)[:3, :4]
        pst = ps.sample_test(c2w)
        eq(O.point_sample(6, 8, 11.1, 16, 2., 6., c2w), pst, "PointSampler.sample_test")
        np.savez(os.path.join(OUT, "embed.npz"), x=x.numpy(), nerf_L10=r10.numpy(), nerf_L4=r4.numpy(), pts=pts.numpy(),
                 r2l_L10=rr.numpy(), ps_c2w=c2w.numpy(), ps_pts=pst.numpy())
        # ---------------- sample_pdf ----------------
        torch.manual_seed(2)
        N = 48
        bins = torch.sort(torch.rand(N, 63) * 4 + 2, -1)[0]
        w = torch.rand(N, 62)**8
        w[::4] = torch.rand(N, 62)[::4]
        w[1] = 0.
        s_det = Hh.sample_pdf(bins, w, 128, det=True)
        u_det = torch.linspace(0., 1., 128).expand(N, 128)
        o_s, o_i = O.sample_pdf(bins, w, u_det)
        eq(o_s, s_det, "sample_pdf det")
        n_s, n_i, _ = sample_pdf_np(bins.numpy(), w.numpy(), u_det.numpy())
        eq(n_s, s_det, "sample_pdf det (numpy restatement)"), eq(n_i, o_i, "sample_pdf inds (numpy restatement)")
        torch.manual_seed(3)
        s_rnd = Hh.sample_pdf(bins, w, 64, det=False)
        torch.manual_seed(3)
        u_rnd = torch.rand(N, 64)
        o_s2, o_i2 = O.sample_pdf(bins, w, u_rnd)
        eq(o_s2, s_rnd, "sample_pdf random u")
        n_s2, n_i2, _ = sample_pdf_np(bins.numpy(), w.numpy(), u_rnd.numpy())
        eq(n_s2, s_rnd, "sample_pdf random (numpy)"), eq(n_i2, o_i2, "sample_pdf random inds (numpy)")
        np.savez(os.path.join(OUT, "sample_pdf.npz"), bins=bins.numpy(), weights=w.numpy(), samples_det=s_det.numpy(),
                 inds_det=o_i.numpy(), u_rnd=u_rnd.numpy(), samples_rnd=s_rnd.numpy(), inds_rnd=o_i2.numpy())
        # ---------------- raw2outputs ----------------
        torch.manual_seed(4)
        raw = torch.randn(40, 64, 4)
        raw[3, :, 3] = -1.0   # a ray with no density at all -> acc 0, disp NaN
        z = torch.sort(torch.rand(40, 64) * 4 + 2, -1)[0]
        d = torch.randn(40, 3)
        ref = ns["raw2outputs"](raw, z, d, 0, True)
        mine = O.raw2outputs(raw, z, d, None, True)
        for a, b, nme in zip(mine, ref, ("rgb", "disp", "acc", "weights", "depth")):
            eq(a, b, f"raw2outputs {nme}")
        np.savez(os.path.join(OUT, "raw2outputs.npz"), raw=raw.numpy(), z=z.numpy(), d=d.numpy(),
                 **{k: v.numpy() for k, v in zip(("rgb", "disp", "acc", "weights", "depth"), ref)})
        # ---------------- NeRF render (lego config, subset of a 400x400 view) ----------------
        torch.manual_seed(0)
        coarse = M.NeRF(8, 256, 63, 27, 5, [4], True).eval()
        fine = M.NeRF(8, 256, 63, 27, 5, [4], True).eval()
        sdc, sdf = O.nerf_state_dicts(0)
        for k in sdc:
            eq(sdc[k], coarse.state_dict()[k], f"coarse init {k}") if k.endswith("0.weight") else None
        assert all(torch.equal(sdc[k], coarse.state_dict()[k]) for k in sdc)
        assert all(torch.equal(sdf[k], fine.state_dict()[k]) for k in sdf)
        embed_fn, _ = Hh.get_embedder(10, 0)
        embeddirs_fn, _ = Hh.get_embedder(4, 0)
        nq = lambda i, v, f: ns["run_network"](i, v, f, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn, netchunk=65536)
        cam = O.LEGO
        c2w = O.pose_spherical(-180. + 360. * 0 / 1, -30., 4.)[:3, :4]
        ro, rd = Hh.get_rays(cam["H"], cam["W"], cam["focal"], c2w)
        idx = np.arange(100, 160000, 2531)[:64]
        rays = torch.stack([ro.reshape(-1, 3)[idx], rd.reshape(-1, 3)[idx]], 0)
        kw = dict(network_query_fn=nq, perturb=0., N_importance=128, network_fine=fine, N_samples=64,
                  network_fn=coarse, use_viewdirs=True, white_bkgd=True, raw_noise_std=0., ndc=False, lindisp=False,
                  near=2., far=6., retraw=True)
        rgb, disp, acc, extras = ns["render"](cam["H"], cam["W"], cam["focal"], chunk=32768, rays=rays, **kw)
        batch = O.pack_rays(rays[0], rays[1], 2., 6.)
        o = O.render_rays(batch, sdc, sdf, 64, 128, white_bkgd=True)
        eq(o["rgb_map"], rgb, "render rgb_map"), eq(o["disp_map"], disp, "render disp_map")
        eq(o["acc_map"], acc, "render acc_map"), eq(o["raw"], extras["raw"], "render raw(fine)")
        eq(o["rgb0"], extras["rgb0"], "render rgb0"), eq(o["z_std"], extras["z_std"], "render z_std")
        np.savez(os.path.join(OUT, "nerf_render_lego.npz"), c2w=c2w.numpy(), idx=idx, rays_o=rays[0].numpy(),
                 rays_d=rays[1].numpy(), rgb_map=rgb.numpy(), disp_map=disp.numpy(), acc_map=acc.numpy(),
                 raw=extras["raw"].numpy(), rgb0=extras["rgb0"].numpy(), disp0=extras["disp0"].numpy(),
                 acc0=extras["acc0"].numpy(), z_std=extras["z_std"].numpy(), inds=o["inds"].numpy(),
                 weights0=o["weights0"].numpy(), z_vals=o["z_vals"].numpy(),
                 cks_coarse=np.array(sorted(checksum(sdc).items()), dtype=object),
                 cks_fine=np.array(sorted(checksum(sdf).items()), dtype=object))
        # stochastic flavour (create_data uses the train kwargs: perturb=1), injected CPU-generator draws
        torch.manual_seed(7)
        kw2 = dict(kw, perturb=1., raw_noise_std=1.0)
        rgb2, disp2, acc2, ex2 = ns["render"](cam["H"], cam["W"], cam["focal"], chunk=32768, rays=rays[:, :16], **kw2)
        torch.manual_seed(7)
        o2 = O.render_rays(batch[:16], sdc, sdf, 64, 128, perturb=1., raw_noise_std=1.0, white_bkgd=True)
        eq(o2["rgb_map"], rgb2, "render perturb rgb_map"), eq(o2["z_std"], ex2["z_std"], "render perturb z_std")
        np.savez(os.path.join(OUT, "nerf_render_perturb.npz"), rgb_map=rgb2.numpy(), disp_map=disp2.numpy(),
                 acc_map=acc2.numpy(), z_std=ex2["z_std"].numpy(), seed=7)
        # fern-like NDC config (N_importance=64, white_bkgd False)
        cam = O.FERN
        c2w = torch.tensor([[0.99, 0.01, -0.1, 0.3], [-0.02, 0.995, 0.05, -0.2], [0.1, -0.05, 0.99, 0.1]])
        ro, rd = Hh.get_rays(cam["H"], cam["W"], cam["focal"], c2w)
        idxf = np.arange(50, cam["H"] * cam["W"], 6007)[:32]
        raysf = torch.stack([ro.reshape(-1, 3)[idxf], rd.reshape(-1, 3)[idxf]], 0)
        kwf = dict(kw, N_importance=64, white_bkgd=False, ndc=True, near=0., far=1.)
        rgbf, dispf, accf, exf = ns["render"](cam["H"], cam["W"], cam["focal"], chunk=32768, rays=raysf, **kwf)
        bf = O.pack_rays(raysf[0], raysf[1], 0., 1., cam["H"], cam["W"], cam["focal"], ndc=True)
        of = O.render_rays(bf, sdc, sdf, 64, 64, white_bkgd=False)
        eq(of["rgb_map"], rgbf, "render fern rgb_map"), eq(of["disp_map"], dispf, "render fern disp")
        np.savez(os.path.join(OUT, "nerf_render_fern.npz"), c2w=c2w.numpy(), idx=idxf, rays_o=raysf[0].numpy(),
                 rays_d=raysf[1].numpy(), rgb_map=rgbf.numpy(), disp_map=dispf.numpy(), acc_map=accf.numpy(),
                 rgb0=exf["rgb0"].numpy(), z_std=exf["z_std"].numpy())
        # ---------------- R2L ----------------
        args = O.r2l_args()
        torch.manual_seed(0)
        net = M.NeRF_v3_2(args, 1008, 3).eval()
        sd = O.r2l_state_dict(0)
        rsd = net.state_dict()
        assert set(sd) == set(rsd) and all(torch.equal(sd[k], rsd[k]) for k in sd), "R2L init mismatch"
        print("  oracle == reference (bit-exact): R2L random init (all 178 tensors)")
        cam = O.LEGO
        c2w = O.pose_spherical(-180., -30., 4.)[:3, :4]
        ps = M.PointSampler(cam["H"], cam["W"], cam["focal"], 16, 2., 6.)
        pts = ps.sample_test(c2w)
        idx = np.arange(7, 160000, 1259)[:128]
        rgb = net(M.PositionalEmbedder(L=10)(pts[idx]))
        eq(O.point_sample(cam["H"], cam["W"], cam["focal"], 16, 2., 6., c2w)[idx], pts[idx], "R2L pts")
        eq(O.r2l_forward(sd, O.embed_r2l(pts[idx], 10)), rgb, "R2L rgb")
        np.savez(os.path.join(OUT, "r2l_lego.npz"), c2w=c2w.numpy(), idx=idx, pts=pts[idx].numpy(), rgb=rgb.numpy(),
                 cks=np.array(sorted(checksum(sd).items()), dtype=object))
    np.savez(os.path.join(OUT, "meta.npz"), **{k: np.array(str(v)) for k, v in meta.items()})
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
