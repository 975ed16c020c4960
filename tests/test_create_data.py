"""create_data `rand` (BASELINE config 3): host logic on CPU against the reference's own random
sequence (tests/golden/create_data.npz, written by oracle/make_golden_create_data.py from the real reference), the
.npy wire format, resume and rank sharding; the GPU test renders through the fused kernels."""
import os

import numpy as np
import pytest
import torch

from conftest import t


def fake_render(H, W, focal, o, d):
    """deterministic stand-in for the teacher: rgb is a function of the ray, depth of the focal length"""
    rgb = torch.sigmoid(torch.stack([o[..., 0] + d[..., 0], o[..., 1] * d[..., 1], d[..., 2] * 3], -1))
    return rgb, torch.full(o.shape[:-1], float(focal) * 1e-3)


def run(E, tmp, **kw):
    from efficient_nerf_b200 import create_data as CD
    args = dict(teacher_fn=None, teacher_fine=None, datadir=str(tmp), n_pose_kd=4, H=8, W=8, focal=10., i_save=2,
                split_size=32, render_fn=fake_render, writer_threads=2)
    args.update(kw)
    return CD, CD.create_data_rand(**args)


def test_pose_and_focal_stream_matches_the_reference(E, golden):
    from efficient_nerf_b200 import create_data as CD
    g = golden("create_data")
    assert np.array_equal(CD.pose_spherical(37.5, -42.0, 4.0).numpy(), g["fixed_pose"])
    np.random.seed(0)
    for k in range(8):
        pose = CD.get_rand_pose()
        f = np.random.rand()
        assert np.array_equal(pose.numpy(), g["poses"][k]), k      # bit-exact float32 matrices
        assert f == g["fdraws"][k]
    assert np.array_equal(np.random.permutation(1000), g["perm1"])
    assert np.array_equal(np.random.permutation(1000), g["perm2"])


def test_wire_format_shuffle_and_reference_stream(E, tmp_path):
    np.random.seed(3)
    CD, written = run(E, tmp_path / "a")
    F = CD.files_per_group(8, 8, 2, 32)                       # 2 poses * 64 rays / 32 = 4 files per group
    assert F == 4 and written == list(range(1, 9))
    files = sorted((x for x in os.listdir(tmp_path / "a") if x.endswith(".npy")), key=lambda x: int(x[5:-4]))
    assert files == [f"data_{k}.npy" for k in range(1, 9)]
    arr = np.load(tmp_path / "a" / "data_1.npy")
    assert arr.dtype == np.float32 and arr.shape == (32, 9)
    # replay the reference's stream by hand: pose, focal, pose, focal, perm, perm per group
    np.random.seed(3)
    want = []
    for g in range(2):
        rows = []
        for _ in range(2):
            pose = CD.get_rand_pose()
            focal_ = 10. * (np.random.rand() + 1)
            o, d = CD._host_rays(8, 8, focal_, pose[:3, :4])
            rgb, _ = fake_render(8, 8, focal_, o, d)
            rows.append(torch.cat([o.reshape(-1, 3), d.reshape(-1, 3), rgb.reshape(-1, 3)], -1))
        data = torch.cat(rows, 0)
        ix1, ix2 = np.random.permutation(128), np.random.permutation(128)
        want.append(data[ix1][ix2].numpy())                   # exactly the reference's double shuffle
    got = CD.load_shards(str(tmp_path / "a"))
    assert np.array_equal(got, np.concatenate(want, 0))


def test_learn_depth_columns_and_tail_rows_dropped(E, tmp_path):
    np.random.seed(0)
    CD, written = run(E, tmp_path / "d", learn_depth='depth', split_size=48)   # 128 rows -> 2 files, 32 rows dropped
    assert len(written) == 4
    arr = CD.load_shards(str(tmp_path / "d"))
    assert arr.shape == (4 * 48, 10)
    np.random.seed(0)
    CD, _ = run(E, tmp_path / "s", learn_depth='surface')
    arr = CD.load_shards(str(tmp_path / "s"))
    assert arr.shape[1] == 12
    # surface point = o + d * depth (create_data.py:833-834): (surface - o) is parallel to d, scale = the pose's depth
    off = arr[:, 9:12] - arr[:, 0:3]
    assert np.abs(np.cross(off, arr[:, 3:6])).max() < 1e-5
    scale = (off * arr[:, 3:6]).sum(-1) / (arr[:, 3:6] ** 2).sum(-1)
    assert scale.min() > 0.0099 and scale.max() < 0.0201      # depth = focal_ * 1e-3, focal_ in [10, 20)


def test_resume_counts_existing_files_like_the_reference(E, tmp_path):
    np.random.seed(1)
    CD, w1 = run(E, tmp_path / "r", n_pose_kd=2)
    CD, w2 = run(E, tmp_path / "r", n_pose_kd=2)
    assert w1 == [1, 2, 3, 4] and w2 == [5, 6, 7, 8]          # create_data.py:790-796


def test_rank_sharding_is_disjoint_complete_and_rank_count_independent(E, tmp_path):
    CD, w_all = run(E, tmp_path / "one", n_pose_kd=8, world_size=1, stream="per_group", seed=11)
    CD, w0 = run(E, tmp_path / "two", n_pose_kd=8, rank=0, world_size=2, seed=11)
    CD, w1 = run(E, tmp_path / "two", n_pose_kd=8, rank=1, world_size=2, seed=11)
    assert sorted(w0 + w1) == w_all and not set(w0) & set(w1)
    assert w0 == [1, 2, 3, 4, 9, 10, 11, 12]                    # groups 0 and 2
    assert np.array_equal(CD.load_shards(str(tmp_path / "one")), CD.load_shards(str(tmp_path / "two")))
    # resume in per_group mode: finished groups are skipped
    CD, again = run(E, tmp_path / "two", n_pose_kd=8, rank=0, world_size=2, seed=11)
    assert again == []
    # a crash can leave a group's LAST file complete while earlier ranges are missing (parallel writers): only the
    # completion marker, written after every writer has finished, makes a group "done" 
    import glob
    import os
    marks = sorted(glob.glob(str(tmp_path / "two" / ".group_1_4.done")))
    assert len(marks) == 1
    os.remove(marks[0])
    os.remove(str(tmp_path / "two" / "data_2.npy"))
    CD, again = run(E, tmp_path / "two", n_pose_kd=8, rank=0, world_size=2, seed=11)
    assert again == [1, 2, 3, 4]
    assert np.array_equal(CD.load_shards(str(tmp_path / "one")), CD.load_shards(str(tmp_path / "two")))
    # the reference stream on two ranks: every rank replays the global np.random stream
    np.random.seed(5)
    CD, r_all = run(E, tmp_path / "ref1", n_pose_kd=8)
    np.random.seed(5)
    CD, r0 = run(E, tmp_path / "ref2", n_pose_kd=8, rank=0, world_size=2, stream="reference", resume=False)
    np.random.seed(5)
    CD, r1 = run(E, tmp_path / "ref2", n_pose_kd=8, rank=1, world_size=2, stream="reference", resume=False)
    assert sorted(r0 + r1) == r_all
    assert np.array_equal(CD.load_shards(str(tmp_path / "ref1")), CD.load_shards(str(tmp_path / "ref2")))


def test_argument_validation(E, tmp_path):
    with pytest.raises(ValueError):
        run(E, tmp_path / "x", n_pose_kd=3)
    with pytest.raises(ValueError):
        run(E, tmp_path / "x", stream="nope")


@pytest.mark.gpu
def test_create_data_rand_on_gpu_matches_direct_render(E, O, tmp_path):
    """Rows written by the sharded writer == (o, d, rgb) of the same poses rendered directly; fast_rng mode runs."""
    from efficient_nerf_b200 import create_data as CD
    sdc, sdf = O.nerf_state_dicts(0)
    nets = []
    for sd in (sdc, sdf):
        n = E.NeRF(8, 256, 63, 27, 5, [4], True, precision="fp16")
        n.load_state_dict(sd)
        nets.append(n.cuda().eval())
    H = W = 32
    focal = O.LEGO["focal"] * 32 / 400
    with torch.no_grad():
        np.random.seed(2)
        written = CD.create_data_rand(nets[0], nets[1], str(tmp_path / "g"), 4, H, W, focal, perturb=0., i_save=2,
                                      split_size=256, learn_depth='depth')
        assert written == list(range(1, 17))
        got = CD.load_shards(str(tmp_path / "g"))
        assert got.shape == (4 * H * W, 10) and np.isfinite(got).all()
        np.random.seed(2)
        want = []
        for g in range(2):
            rows = []
            for _ in range(2):
                pose = CD.get_rand_pose()
                focal_ = focal * (np.random.rand() + 1)
                ro, rd = E.get_rays(H, W, focal_, pose[:3, :4].cuda())
                rgb, _, _, ex = E.render_image(H, W, focal, chunk=32768, rays=(ro, rd), network_fn=nets[0],
                                               network_fine=nets[1], network_query_fn=None, N_samples=64,
                                               N_importance=128, perturb=0., white_bkgd=True, use_viewdirs=True,
                                               ndc=False, near=2., far=6., return_depth=True)
                rows.append(torch.cat([ro.reshape(-1, 3), rd.reshape(-1, 3), rgb.reshape(-1, 3),
                                       ex["depth_map"].reshape(-1, 1)], -1))
            data = torch.cat(rows, 0).cpu()
            ix1, ix2 = np.random.permutation(2 * H * W), np.random.permutation(2 * H * W)
            want.append(data[ix1][ix2].numpy())
        assert np.array_equal(got, np.concatenate(want, 0))
        # stochastic teacher (the reference's default perturb=1) with device-side draws
        w2 = CD.create_data_rand(nets[0], nets[1], str(tmp_path / "f"), 2, H, W, focal, perturb=1., i_save=2,
                                 split_size=256, fast_rng=True, stream="per_group", seed=3)
        arr = CD.load_shards(str(tmp_path / "f"))
        assert len(w2) == 8 and np.isfinite(arr).all() and arr[:, 6:9].min() >= 0. and arr[:, 6:9].max() <= 1.0 + 1e-5


@pytest.mark.gpu
def test_dataparallel_wrapped_teacher_takes_the_fused_path(E, O, tmp_path):
    """utils/create_data.py:298-299 ALWAYS wraps teacher_fn / teacher_fine in torch.nn.DataParallel (main.py wraps in
    MyDataParallel whenever it is not --render_only).  render_rays unwraps them, so the fused kernels run — same
    frame, bit for bit, and the same number of library kernels — instead of the embed + DataParallel.forward detour
    (ADVICE r1); a non-fused teacher (precision='fp32') still renders through the reference-style query function."""
    from efficient_nerf_b200 import create_data as CD
    sdc, sdf = O.nerf_state_dicts(0)
    nets = []
    for sd in (sdc, sdf):
        n = E.NeRF(8, 256, 63, 27, 5, [4], True, precision="fp16")
        n.load_state_dict(sd)
        nets.append(n.cuda().eval())
    wrapped = [torch.nn.DataParallel(n, device_ids=[0]) for n in nets]
    H = W = 24
    focal = O.LEGO["focal"] * W / 400
    c2w = O.pose_spherical(20., -30., 4.)[:3, :4].cuda()
    kw = dict(network_query_fn=None, N_samples=64, N_importance=128, perturb=0., white_bkgd=True, use_viewdirs=True,
              ndc=False, near=2., far=6.)
    with torch.no_grad():
        nets[0].packed_handle(), nets[1].packed_handle()      # weight packing launches its own kernels, once
        k0 = E._lib.kernel_launches()
        a, _, _, _ = E.render_image(H, W, focal, chunk=32768, c2w=c2w, network_fn=nets[0], network_fine=nets[1], **kw)
        k1 = E._lib.kernel_launches()
        b, _, _, _ = E.render_image(H, W, focal, chunk=32768, c2w=c2w, network_fn=wrapped[0], network_fine=wrapped[1], **kw)
        k2 = E._lib.kernel_launches()
    assert torch.equal(a, b) and (k2 - k1) == (k1 - k0)
    # the handles of the wrapped modules were reused, not re-packed per call
    h0 = nets[0].packed_handle()
    with torch.no_grad():
        E.render_image(H, W, focal, chunk=32768, c2w=c2w, network_fn=wrapped[0], network_fine=wrapped[1], **kw)
    assert nets[0].packed_handle() is h0
    # DataParallel replicas get their own (empty) handle cache
    rep = nets[0]._replicate_for_data_parallel()
    assert rep._packed == {} and rep._packed is not nets[0]._packed
    # create_data with wrapped teachers and with a teacher that is NOT on the tensor-core path
    np.random.seed(4)
    w1 = CD.create_data_rand(wrapped[0], wrapped[1], str(tmp_path / "dp"), 2, H, W, focal, perturb=0., i_save=2,
                             split_size=128)
    np.random.seed(4)
    w2 = CD.create_data_rand(nets[0], nets[1], str(tmp_path / "plain"), 2, H, W, focal, perturb=0., i_save=2,
                             split_size=128)
    assert w1 == w2 and np.array_equal(CD.load_shards(str(tmp_path / "dp")), CD.load_shards(str(tmp_path / "plain")))
    n32 = []
    for sd in (sdc, sdf):
        n = E.NeRF(8, 256, 63, 27, 5, [4], True, precision="fp32")
        n.load_state_dict(sd)
        n32.append(n.cuda().eval())
    np.random.seed(4)
    CD.create_data_rand(n32[0], n32[1], str(tmp_path / "fp32"), 2, H, W, focal, perturb=0., i_save=2, split_size=128)
    got, ref = CD.load_shards(str(tmp_path / "fp32")), CD.load_shards(str(tmp_path / "plain"))
    assert got.shape == ref.shape and np.abs(got[:, 6:9] - ref[:, 6:9]).max() <= 2e-3 and np.array_equal(got[:, :6], ref[:, :6])
