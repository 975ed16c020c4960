"""SURVEY §8f rank 1, end to end: the reference's UNMODIFIED main.py (staged byte for byte under baseline/_ref by
baseline/stage_reference.py, hashes pinned in baseline/MANIFEST.sha256) runs `--render_only --render_test` and
`--benchmark` through `python -m efficient_nerf_b200.dropin` on a B200, on a synthetic blender scene and seeded
checkpoints in the reference's format; the frames it writes must be bit-identical to direct package calls.

The CPU half checks the staging itself (manifest, loader, oracle == staged reference on a few rays)."""
import glob
import importlib.util
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "baseline", "_ref")
staged = pytest.mark.skipif(not os.path.exists(os.path.join(REFDIR, "main.py")),
                            reason="baseline/_ref is not staged (python baseline/stage_reference.py)")


def _stage_mod():
    spec = importlib.util.spec_from_file_location("stage_reference", os.path.join(ROOT, "baseline", "stage_reference.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@staged
def test_staged_reference_is_unmodified():
    st = _stage_mod()
    assert st.check(REFDIR) >= 70
    if os.path.exists("/root/reference/main.py"):          # build container: byte-identical to the checkout
        for rel in ("main.py", "option.py", "model/nerf_raybased.py", "utils/run_nerf_raybased_helpers.py",
                    "utils/create_data.py", "utils/flip_loss.py", "utils/ssim_torch.py", "dataset/load_blender.py"):
            with open(os.path.join(REFDIR, rel), "rb") as a, open(os.path.join("/root/reference", rel), "rb") as b:
                assert a.read() == b.read(), rel
    # nothing of it is tracked by git
    r = subprocess.run(["git", "ls-files", "baseline/_ref"], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode != 0 or r.stdout.strip() == ""


@staged
def test_oracle_equals_the_staged_reference(O):
    """the CPU baseline of bench.py (`cpu_baseline.kind = "reference"`) runs these modules; the oracle restates them"""
    from oracle import ref_real as R
    ref = R.load(REFDIR, device="cpu")
    c2w = O.pose_spherical(40., -30., 4.)[:3, :4]
    idx = torch.arange(0, 160000, 9973)
    with torch.no_grad():
        model, pe, ps = R.build_r2l(ref, 0)
        got = model(pe(ps.sample_test(c2w)[idx]))
        want = O.render_r2l(O.r2l_state_dict(0), 400, 400, O.LEGO["focal"], 2., 6., c2w, rows=idx)
        assert torch.equal(got, want)
        coarse, fine, kw = R.build_nerf(ref, 0)
        ro, rd = ref.Hh.get_rays(400, 400, O.LEGO["focal"], c2w)
        rays = (ro.reshape(-1, 3)[idx], rd.reshape(-1, 3)[idx])
        rgb, disp, acc, extras = ref.main["render"](400, 400, O.LEGO["focal"], chunk=32768, rays=rays, **kw)
        sdc, sdf = O.nerf_state_dicts(0)
        o = O.render_rays(O.pack_rays(rays[0], rays[1], 2., 6.), sdc, sdf, 64, 128, white_bkgd=True)
        assert torch.equal(rgb, o["rgb_map"]) and torch.equal(extras["rgb0"], o["rgb0"])


def _run_main(tmp_path, argv, timeout=900):
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0])
    r = subprocess.run([sys.executable, "-m", "efficient_nerf_b200.dropin", os.path.join(REFDIR, "main.py")] + argv,
                       cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=timeout)
    return r


def _frames(tmp_path):
    hits = glob.glob(str(tmp_path / "Experiments" / "*" / "gen_img" / "video_*.mp4.frames" / "frames.npy"))
    hits = [h for h in hits if "_error" not in h]
    assert len(hits) == 1, hits
    return np.load(hits[0])


@staged
@pytest.mark.gpu
def test_unmodified_main_render_only_r2l(tmp_path, E):
    """README.md:51's command (W256 D88 R2L, 400x400 test views) on the real script."""
    from efficient_nerf_b200 import dropin
    scene = dropin.make_synthetic_blender(str(tmp_path / "scene"), res=800, n_train=1, n_val=1, n_test=2)
    net = E.synthetic.seeded_r2l(0, "fp16")
    ckpt = str(tmp_path / "r2l_seed0.tar")
    torch.save({"global_step": 7, "network_fn_state_dict": {k: v.cpu() for k, v in net.state_dict().items()}}, ckpt)
    argv = (f"--model_name R2L --config {REFDIR}/configs/lego_noview.txt --n_sample_per_ray 16 --netwidth 256 --netdepth 88 "
            f"--use_residual --cache_ignore data --trial.ON --trial.body_arch resmlp --pretrained_ckpt {ckpt} "
            f"--render_only --render_test --testskip 1 --screen --project Test__R2L_W256D88__blender_lego "
            f"--datadir {scene} --experiments_dir {tmp_path}/Experiments").split()
    r = _run_main(tmp_path, argv)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    m = re.search(r"\[TEST\] TestPSNR ([\d.naninf-]+) TestPSNRv2 ([\d.naninf-]+) TestSSIM ([\d.naninf-]+) "
                  r"TestLPIPS ([\d.naninf-]+) TestFLIP ([\d.naninf-]+)", r.stdout)
    assert m, r.stdout[-3000:]
    psnr, psnr2, ssim, lp, flip = (float(x) for x in m.groups())
    assert np.isfinite([psnr, psnr2, ssim, flip]).all() and 0 < flip < 1 and 0 < psnr < 60
    got = _frames(tmp_path)                                   # to8b of the frames main.py rendered
    assert got.shape == (2, 400, 400, 3) and got.dtype == np.uint8
    # the same two test poses, direct package calls
    import json
    meta = json.load(open(os.path.join(scene, "transforms_test.json")))
    ps = E.PointSampler(400, 400, .5 * 400 / np.tan(.5 * meta["camera_angle_x"]), 16, 2., 6.)
    with torch.no_grad():
        for i, fr in enumerate(meta["frames"]):
            c2w = torch.tensor(fr["transform_matrix"], dtype=torch.float32)[:3, :4].cuda()
            rgb = net.render_poses(ps, c2w).view(400, 400, 3)
            assert np.array_equal(got[i], E.run_nerf_raybased_helpers.to8b(rgb)), i
    assert "Load pretrained ckpt successfully" in r.stdout


@staged
@pytest.mark.gpu
def test_unmodified_main_render_only_nerf(tmp_path, E):
    """configs/lego.txt (NeRF W256 D8, 64 + 128 samples, viewdirs, white background) on the real script."""
    from efficient_nerf_b200 import dropin
    scene = dropin.make_synthetic_blender(str(tmp_path / "scene"), res=160, n_train=1, n_val=1, n_test=2)
    coarse, fine = E.synthetic.seeded_nerf_pair(0, "fp16")
    ckpt = str(tmp_path / "nerf_seed0.tar")
    torch.save({"global_step": 7, "network_fn_state_dict": {k: v.cpu() for k, v in coarse.state_dict().items()},
                "network_fine_state_dict": {k: v.cpu() for k, v in fine.state_dict().items()}}, ckpt)
    argv = (f"--model_name nerf --config {REFDIR}/configs/lego.txt --pretrained_ckpt {ckpt} --render_only --render_test "
            f"--testskip 1 --screen --project Test__nerf__blender_lego --cache_ignore data "
            f"--datadir {scene} --experiments_dir {tmp_path}/Experiments").split()
    r = _run_main(tmp_path, argv)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert re.search(r"\[TEST\] TestPSNR", r.stdout), r.stdout[-3000:]
    got = _frames(tmp_path)
    assert got.shape == (2, 80, 80, 3)
    import json
    meta = json.load(open(os.path.join(scene, "transforms_test.json")))
    focal = .5 * 80 / np.tan(.5 * meta["camera_angle_x"])
    with torch.no_grad():
        for i, fr in enumerate(meta["frames"]):
            c2w = torch.tensor(fr["transform_matrix"], dtype=torch.float32)[:3, :4].cuda()
            rgb, _, _, _ = E.render.render(80, 80, focal, chunk=32768, c2w=c2w, network_fn=coarse, network_fine=fine,
                                           network_query_fn=None, N_samples=64, N_importance=128, perturb=0.,
                                           raw_noise_std=0., white_bkgd=True, use_viewdirs=True, ndc=False, near=2.,
                                           far=6.)
            assert np.array_equal(got[i], E.run_nerf_raybased_helpers.to8b(rgb)), i


@staged
@pytest.mark.gpu
def test_unmodified_main_benchmark(tmp_path, E):
    """main.py:1124-1133 `--benchmark`: torch.utils.benchmark.Timer around render_func(model, pose), 100 runs."""
    from efficient_nerf_b200 import dropin
    scene = dropin.make_synthetic_blender(str(tmp_path / "scene"), res=800, n_train=1, n_val=1, n_test=1)
    argv = (f"--model_name R2L --config {REFDIR}/configs/lego_noview.txt --n_sample_per_ray 16 --netwidth 256 --netdepth 88 "
            f"--use_residual --cache_ignore data --trial.ON --trial.body_arch resmlp --benchmark --screen "
            f"--project Bench__R2L --datadir {scene} --experiments_dir {tmp_path}/Experiments").split()
    r = _run_main(tmp_path, argv)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    m = re.search(r"render_func\(model, pose\).*?\n\s*([\d.]+)\s*(us|ms|s)\b", r.stdout, re.S)
    assert m, r.stdout[-3000:]
    ms = float(m.group(1)) * {"us": 1e-3, "ms": 1., "s": 1e3}[m.group(2)]
    print(f"[reference main.py --benchmark] render_func(model, pose): {ms:.3f} ms per 400x400 frame")
    assert ms < 20.                      # the reference's eager PyTorch path needs ~100 ms on a datacentre GPU
