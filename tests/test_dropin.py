"""Drop-in launcher: the reference's unmodified entry points bound to the B200 path.

CPU: with the real reference present (this container only) main.py is loaded UNMODIFIED through the launcher, its
names must resolve to this package, its own option.py must parse the README command, and train() must run the
reference's data loading until the first device entry raises (no CPU fallback).  GPU (no reference on the box): a
miniature caller written for this test goes through the same launcher and must render exactly what direct calls do."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
MINI = os.path.join(ROOT, "tests", "fixtures", "mini_main.py")

R2L_CMD = ("--model_name R2L --config {ref}/configs/lego_noview.txt --n_sample_per_ray 16 --netwidth 256 --netdepth 88 "
           "--use_residual --cache_ignore data --trial.ON --trial.body_arch resmlp --render_only --render_test "
           "--testskip 1 --screen --project Test__R2L_W256D88__blender_lego")          # README.md:51 minus the ckpt

PROBE = r'''
import json, sys, traceback
sys.path.insert(0, {root!r})
from efficient_nerf_b200 import dropin, synthetic
import efficient_nerf_b200 as E
dropin.make_synthetic_blender({tmp!r} + "/scene", res=32)
argv = ({cmd!r} + " --datadir {tmp}/scene --experiments_dir {tmp}/Experiments").split()
mod = dropin.load_script({script!r}, argv)
out = dict(patched=mod.__dropin_patched__,
           bound={{n: getattr(mod, n).__module__ for n in ("NeRF", "NeRF_v3_2", "PointSampler", "get_rays", "sample_pdf",
                                                          "ndc_rays", "get_embedder", "render_rays", "render", "raw2outputs")}},
           lazy=mod.PositionalEmbedder.__name__, trial=vars(mod.args.trial), half_res=mod.args.half_res,
           N_importance=mod.args.N_importance, use_viewdirs=mod.args.use_viewdirs, white_bkgd=mod.args.white_bkgd,
           netdepth=mod.args.netdepth, chunk=mod.args.chunk)
import torch
torch.manual_seed(0)
net = mod.NeRF_v3_2(mod.args, 1008, 3)          # what create_nerf builds from the parsed options (main.py:455-461)
ref = synthetic.seeded_r2l(0, "fp16", device="cpu")
out["same_init"] = all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), ref.state_dict().values()))
out["n_keys"] = len(net.state_dict())
try:
    mod.train()
    out["raised"] = None
except RuntimeError as e:
    frames = [f.filename for f in traceback.extract_tb(e.__traceback__)]
    out["raised"] = str(e)
    out["through_reference"] = any(f.startswith({ref!r}) for f in frames)
    out["into_package"] = any("efficient-nerf_b200" in f or "efficient_nerf_b200" in f for f in frames)
print("PROBE" + json.dumps(out, default=str))
'''


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "main.py")), reason="the reference checkout is not present")
def test_unmodified_main_binds_to_the_package(tmp_path):
    code = PROBE.format(root=ROOT, tmp=str(tmp_path), cmd=R2L_CMD.format(ref=REF), script=os.path.join(REF, "main.py"),
                        ref=REF)
    r = subprocess.run([sys.executable, "-c", code], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith("PROBE")]
    assert line, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads(line[0][5:])
    assert out["patched"] == ["batchify", "run_network", "batchify_rays", "render", "raw2outputs", "render_rays"]
    assert all(m.startswith("efficient_nerf_b200") for m in out["bound"].values()), out["bound"]
    assert out["lazy"] == "_LazyPositionalEmbedder"
    assert out["trial"]["body_arch"] == "resmlp" and out["trial"]["ON"] is True and out["trial"]["n_learnable"] == 2
    assert (out["half_res"], out["N_importance"], out["use_viewdirs"], out["white_bkgd"]) == (True, 128, False, True)
    assert (out["netdepth"], out["chunk"]) == (88, 32768)
    assert out["same_init"] and out["n_keys"] == 176
    if not torch.cuda.is_available():
        # the reference's own train() loaded the synthetic scene and died at the first device entry of the package
        assert "no CPU fallback" in out["raised"] and out["through_reference"] and out["into_package"]


def test_config_parser_and_update_args(tmp_path):
    from efficient_nerf_b200 import dropin
    pkg, u = dropin._smilelogging_stub()
    p = pkg.argparser
    p.add_argument("--config", is_config_file=True)
    p.add_argument("--N_samples", type=int, default=1)
    p.add_argument("--use_viewdirs", action="store_true")
    p.add_argument("--white_bkgd", action="store_true")
    p.add_argument("--datadir", type=str, default="")
    p.add_argument("--trial.ON", action="store_true")
    p.add_argument("--trial.near", type=float, default=-1)
    p.add_argument("--other.ON", action="store_true")
    p.add_argument("--other.x", type=int, default=3)
    cfg = tmp_path / "c.txt"
    cfg.write_text("# comment\nN_samples = 64\n\nuse_viewdirs = False # inline comment\nwhite_bkgd = True\n"
                   "datadir = ./data/x\n")
    a = u.update_args(p.parse_args(["--config", str(cfg), "--datadir", "/cli/wins", "--trial.ON", "--screen",
                                    "--project", "P"]))
    assert (a.N_samples, a.use_viewdirs, a.white_bkgd, a.datadir) == (64, False, True, "/cli/wins")
    assert a.trial.near == -1 and a.trial.ON and not hasattr(a, "other") and not hasattr(a, "trial.near")
    assert a.project_name == "P" and a.screen_print
    assert u.strdict_to_dict("a:1,b:2", int) == {"a": 1, "b": 2}
    assert dropin.parse_expid_iter("Experiments/x_SERVER142-20210704-150540/weights/200000.tar") == \
        ("SERVER142-20210704-150540", "200000")
    assert dropin.parse_expid_iter("lego.tar") == ("Unknown", "Unknown")
    lin = torch.nn.Sequential(torch.nn.Linear(3, 5), torch.nn.ReLU(), torch.nn.Linear(5, 2))
    assert u.get_n_params_(lin) == 3 * 5 + 5 + 5 * 2 + 2 and u.get_n_flops_(lin) == 25


def test_synthetic_blender_scene_has_the_loader_format(tmp_path):
    from efficient_nerf_b200 import dropin
    from PIL import Image
    d = dropin.make_synthetic_blender(str(tmp_path / "s"), res=16, n_train=2, n_val=1, n_test=1)
    for split, n in (("train", 2), ("val", 1), ("test", 1)):
        meta = json.load(open(os.path.join(d, f"transforms_{split}.json")))
        assert len(meta["frames"]) == n and abs(.5 * 800 / np.tan(.5 * meta["camera_angle_x"]) - 1111.111) < 1e-2
        fr = meta["frames"][0]
        assert np.array(fr["transform_matrix"]).shape == (4, 4)
        assert np.asarray(Image.open(os.path.join(d, fr["file_path"] + ".png"))).shape == (16, 16, 4)


@pytest.mark.gpu
def test_launcher_renders_what_direct_calls_render(tmp_path, E):
    out = str(tmp_path / "out.npz")
    argv = ["--out", out, "--res", "40", "--netdepth", "88", "--use_residual", "--white_bkgd", "--trial.ON",
            "--trial.body_arch", "resmlp", "--experiments_dir", str(tmp_path / "Experiments")]
    r = subprocess.run([sys.executable, "-m", "efficient_nerf_b200.dropin", MINI] + argv, cwd=str(tmp_path),
                       env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    got = np.load(out)
    assert str(got["lazy"]) == "LazyEmbedding" and str(got["lazy_pts"]) == "LazyPoints" and got["calls"].size == 0 and str(got["trial"]) == "resmlp"
    # the same frames from direct package calls with the same seed and construction order
    H = W = 40
    focal = 555.5555155968841 * W / 400.
    c2w = torch.tensor([[1., 0., 0., 0.], [0., .5, -.8660254, -3.4641016], [0., .8660254, .5, 2.]], device="cuda")
    torch.manual_seed(0)
    net = E.NeRF_v3_2(E.synthetic.r2l_args(88, 256), 1008, 3).cuda()
    coarse = E.NeRF(8, 256, 63, 27, 5, [4], True).cuda()
    fine = E.NeRF(8, 256, 63, 27, 5, [4], True).cuda()
    with torch.no_grad():
        ps = E.PointSampler(H, W, focal, 16, 2., 6.)
        r2l = net.forward_points(ps.sample_test(c2w)).view(H, W, 3)
        rgb, _, _, _ = E.render.render(H, W, focal, chunk=32768, c2w=c2w, network_fn=coarse, network_fine=fine,
                                       network_query_fn=None, N_samples=64, N_importance=128, perturb=0.,
                                       raw_noise_std=0., white_bkgd=True, use_viewdirs=True, ndc=False, near=2., far=6.)
    assert np.array_equal(got["r2l"], r2l.cpu().numpy())
    assert np.array_equal(got["nerf"], rgb.cpu().numpy())
    assert got["frame8"].dtype == np.uint8 and got["frame8"].shape == (H, W, 3)


PROBE_CD = r'''
import json, sys
sys.path.insert(0, {root!r})
from efficient_nerf_b200 import dropin
dropin.make_synthetic_blender({tmp!r} + "/scene", res=32)
argv = ("--create_data rand --config {ref}/configs/lego.txt --n_pose_kd 100 --datadir {tmp}/scene "
        "--experiments_dir {tmp}/Experiments --project cd").split()
mod = dropin.load_script("{ref}/utils/create_data.py", argv)
out = dict(patched=mod.__dropin_patched__, render_rays=mod.render_rays.__name__, nerf=mod.NeRF.__module__,
           get_rays=mod.get_rays1.__module__, create_data=mod.args.create_data, n_pose_kd=mod.args.n_pose_kd)
try:
    dropin.load_script("{ref}/utils/create_data.py", argv + ["--focal_scale", "2"])
    out["focal_scale"] = "accepted"
except NotImplementedError as e:
    out["focal_scale"] = "rejected"
print("PROBE" + json.dumps(out, default=str))
'''


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "utils", "create_data.py")), reason="the reference checkout is not present")
def test_unmodified_create_data_binds_to_the_package(tmp_path):
    code = PROBE_CD.format(root=ROOT, tmp=str(tmp_path), ref=REF)
    r = subprocess.run([sys.executable, "-c", code], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith("PROBE")]
    assert line, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads(line[0][5:])
    assert out["patched"] == ["batchify", "run_network", "batchify_rays", "render", "raw2outputs", "render_rays"]
    assert out["render_rays"] == "render_rays_create_data"           # the flavour that also returns depth_map
    assert out["nerf"].startswith("efficient_nerf_b200") and out["get_rays"].startswith("efficient_nerf_b200")
    assert out["create_data"] == "rand" and int(out["n_pose_kd"]) == 100
    assert out["focal_scale"] == "rejected"                          # options the fused render would silently drop
