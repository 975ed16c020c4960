"""CPU tests of the host-side logic: sharding maths, the world_size-2 gather path (gloo), module
construction / state_dict parity with the reference layout, lazy embedding handles, bench plumbing."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_rays_covers_everything_in_tile_multiples(E):
    S = E.sharding
    for n in (0, 1, 127, 128, 129, 160000, 190512, 640000):
        for ws in (1, 2, 3, 4, 8):
            spans = [S.shard_rays(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans[:-1], spans[1:]):
                assert a1 == b0 and a0 <= a1
            for a0, a1 in spans[:-1]:
                assert (a0 % 128 == 0 or a0 == n) and (a1 % 128 == 0 or a1 == n)
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) < 256 or n < 128 * ws
    assert S.shard_rays(160000, 7, 8) == (140032, 160000)


def test_shard_poses_round_robin(E):
    S = E.sharding
    all_idx = sorted(i for r in range(8) for i in S.shard_poses(200, r, 8))
    assert all_idx == list(range(200))
    assert S.shard_poses(5, 3, 4) == [3]


WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import efficient_nerf_b200 as E
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
n = 1000
full = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3)
s0, s1 = E.sharding.shard_rays(n)
out = E.sharding.gather_rays(full[s0:s1].clone(), n)
assert torch.equal(out, full), "gather_rays mismatch"
for n2 in (1000, 1024):   # ragged split (fallback path) and equal blocks (one all_gather_into_tensor)
    full2 = torch.arange(n2 * 3, dtype=torch.float32).reshape(n2, 3)
    a0, a1 = E.sharding.shard_rays(n2)
    buf = torch.full((n2, 3), -1.)
    res = E.sharding.gather_rays_into(full2[a0:a1].clone(), buf, n2)
    assert res is buf and torch.equal(buf, full2), "gather_rays_into mismatch"
mine = [torch.full((8, 3), float(i)) for i in E.sharding.shard_poses(5)]
frames = E.sharding.gather_frames(mine, 5)
assert [float(f[0, 0]) for f in frames] == [0., 1., 2., 3., 4.], "gather_frames mismatch"
dist.destroy_process_group()
print("ok", rank)
'''


def test_gather_paths_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


def test_modules_keep_the_reference_state_dict_layout(E, O):
    sdc, _ = O.nerf_state_dicts(0)
    torch.manual_seed(0)
    net = E.NeRF(8, 256, 63, 27, 5, [4], True)
    msd = net.state_dict()
    assert list(msd) == list(sdc)                       # same keys, same order
    assert all(torch.equal(msd[k], sdc[k]) for k in sdc)  # same seeded init => same RNG consumption
    assert net.supports_tensor_core_path()
    assert not E.NeRF(8, 128, 63, 27, 5, [4], True).supports_tensor_core_path()
    assert not E.NeRF(8, 256, 63, 27, 4, [4], False).supports_tensor_core_path()
    sd = O.r2l_state_dict(0)
    torch.manual_seed(0)
    r = E.NeRF_v3_2(O.r2l_args(), 1008, 3)
    assert set(r.state_dict()) == set(sd) and all(torch.equal(r.state_dict()[k], sd[k]) for k in sd)
    assert r.supports_tensor_core_path()
    assert sum(p.numel() for p in r.parameters()) == 5917187
    assert sum(p.numel() for p in net.parameters()) == 595844
    a = O.r2l_args(netdepth=10)
    a.trial.body_arch = 'mlp'
    assert not E.NeRF_v3_2(a, 1008, 3).supports_tensor_core_path()     # falls to the fp32 CUDA path


def test_lazy_embedding_shape_protocol(E):
    from efficient_nerf_b200.nerf_raybased import LazyEmbedding
    le = LazyEmbedding(torch.zeros(10, 48), 10, True)
    assert tuple(le.shape) == (10, 1008) and le.dtype == torch.float32


def test_precision_argument_validation(E):
    with pytest.raises(ValueError):
        E.NeRF(8, 256, 63, 27, 5, [4], True, precision="int8")


def test_bench_reference_arm_prints_contract_line():
    """bench.py --impl reference runs the reference's own files (baseline/_ref, staged by baseline/stage_reference.py;
    the oracle port only when no staged tree exists) on the host cores and prints one JSON line."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--workload", "r2l"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mrays/s" and line["value"] > 0
    staged = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "main.py")) or os.path.exists("/root/reference/main.py")
    assert line["cpu_baseline"]["kind"] == ("reference" if staged else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0


def test_mirror_exposes_the_reference_call_surface(E):
    """Every name the reference's two hot-path modules define (def / class / lambda at module level, and the methods
    of their classes) exists under the same name in the mirror — minus what DESIGN.md §7 declares out of scope
    (translate_origin*, parse_expid_iter, get_selected_coords, visualize_3d: training / plotting helpers;
    NeRF.load_weights_from_keras).  Lists taken from utils/run_nerf_raybased_helpers.py and model/nerf_raybased.py."""
    helpers = ["to_tensor", "to_array", "to8b", "img2mse", "mse2psnr", "Embedder", "get_embedder", "raw2outputs",
               "batchify", "run_network", "get_rays", "ndc_rays", "sample_pdf", "load_weights", "load_weights_v2",
               "undataparallel", "get_rays_np"]
    for n in helpers:
        assert hasattr(E.run_nerf_raybased_helpers, n), f"run_nerf_raybased_helpers.{n} missing"
    model = {"Embedder": ["create_embedding_fn", "embed"], "get_embedder": [], "raw2outputs": [], "batchify": [],
             "run_network": [], "get_activation": [], "NeRF": ["forward"], "ResMLP": ["forward"],
             "NeRF_v3_2": ["forward"],
             "PointSampler": ["sample_test", "sample_test2", "sample_train", "sample_train2", "sample_train_cnnstyle",
                              "sample_train_plucker", "sample_test_plucker"],
             "PositionalEmbedder": ["__call__", "embed", "embed_cnnstyle"]}
    for n, methods in model.items():
        assert hasattr(E.nerf_raybased, n), f"nerf_raybased.{n} missing"
        for m in methods:
            assert hasattr(getattr(E.nerf_raybased, n), m), f"nerf_raybased.{n}.{m} missing"
    import inspect
    sig = inspect.signature(E.nerf_raybased.raw2outputs)
    assert list(sig.parameters) == ["raw", "z_vals", "rays_d", "raw_noise_std", "white_bkgd", "pytest", "global_step",
                                    "print"]
    sig = inspect.signature(E.run_nerf_raybased_helpers.sample_pdf)
    assert list(sig.parameters)[:5] == ["bins", "weights", "N_samples", "det", "pytest"]
    for n in ("render", "render_rays", "batchify_rays", "batchify", "run_network", "render_path", "raw2outputs"):
        assert hasattr(E.render, n), f"render.{n} missing"


def test_synthetic_workloads_equal_the_oracle_seeds(E, O):
    """bench.py's GPU arm builds its models and poses from the package (never from oracle/); they must be the very
    weights / poses the CPU reference arm renders (oracle/ref_torch.py seeded state dicts)."""
    net = E.synthetic.seeded_r2l(0, "fp16", device="cpu")
    sd = O.r2l_state_dict(0)
    assert set(sd) == set(net.state_dict())
    assert all(torch.equal(net.state_dict()[k], v) for k, v in sd.items())
    coarse, fine = E.synthetic.seeded_nerf_pair(0, "fp16", device="cpu")
    sdc, sdf = O.nerf_state_dicts(0)
    assert all(torch.equal(coarse.state_dict()[k], v) for k, v in sdc.items())
    assert all(torch.equal(fine.state_dict()[k], v) for k, v in sdf.items())
    for k in (0, 3, 199, 200):
        assert torch.equal(E.synthetic.test_pose(k), O.pose_spherical(-180. + 360. * (k % 200) / 200., -30., 4.)[:3, :4])
    assert (E.synthetic.LEGO, E.synthetic.LEGO_800, E.synthetic.FERN) == (O.LEGO, O.LEGO_800, O.FERN)


def test_bench_gpu_arm_does_not_import_the_oracle():
    """Only the CPU legs of bench.py (cpu_baseline / --impl reference) may execute oracle/."""
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py")).read()
    body = src.split("def cpu_reference", 1)[1].split("\ndef ", 1)[0]
    assert "from oracle import ref_torch" in body and "from oracle import ref_real" in body
    assert src.count("from oracle") == body.count("from oracle") == 2      # nowhere else
