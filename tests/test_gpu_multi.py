"""Multi-GPU test of the sharded render path on real devices (NCCL); skipped on boxes with fewer than 2 GPUs (the
driver's 1-GPU tier), where tests/test_host_logic.py covers the same logic with gloo on CPU.
Two ranks each render their contiguous block of one R2L frame's rays and all-gather the tiles
(sharding.gather_rays_into); every rank must end up with exactly the single-GPU frame.  The same for the fused tile
gather (MLP kernel storing into every GPU's symmetric-memory frame buffer + barrier), eager and graph-replayed."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
rank = int(sys.argv[3])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=rank, world_size=2,
                        device_id=torch.device("cuda", rank))
import efficient_nerf_b200 as E
from oracle import ref_torch as O
H = W = 64
sd = O.r2l_state_dict(0)
net = E.NeRF_v3_2(O.r2l_args(), 1008, 3, precision="fp16"); net.load_state_dict(sd); net = net.cuda().eval()
ps = E.PointSampler(H, W, 80., 16, 2., 6.)
c2w = O.pose_spherical(33., -30., 4.)[:3, :4].cuda()
with torch.no_grad():
    pts = ps.sample_test(c2w)
    full = net.forward_points(pts)
    s0, s1 = E.sharding.shard_rays(H * W)
    mine = net.forward_points(pts[s0:s1])
    frame = torch.empty_like(full)
    E.sharding.gather_rays_into(mine.contiguous(), frame, H * W)
torch.cuda.synchronize()
assert torch.equal(frame, full), "sharded frame != single-GPU frame"
assert E.sharding.shard_poses(5) == list(range(rank, 5, 2))
# fused tile gather: the MLP kernel stores this rank's rows into BOTH GPUs' symmetric-memory frame buffers
# (peer-to-peer stores), a cross-GPU barrier publishes the frame: same bits as kernel + NCCL all-gather
with torch.no_grad():
    pf = [E.sharding.PeerFrame(H * W), E.sharding.PeerFrame(H * W)]
    assert (pf[0].row0, pf[0].row1) == (s0, s1) and pf[0].world_size == 2
    for k in range(4):                       # alternate the two buffers, another pose every frame
        c2 = O.pose_spherical(33. + 20. * k, -30., 4.)[:3, :4].cuda()
        want = net.forward_points(ps.sample_test(c2))
        pf[k & 1].buf.fill_(-1.)
        pf[k & 1].publish()                  # nobody starts storing before every rank has cleared its buffer
        got = net.forward_points_gather(ps.sample_test(c2)[s0:s1], pf[k & 1])
        assert got is pf[k & 1].buf
        pf[k & 1].publish()
        torch.cuda.synchronize()
        assert torch.equal(pf[k & 1].buf, want), f"fused gather frame {k} != single-GPU frame"
    # the same through the CUDA-graph replay
    g = E.GraphedR2L(net, ps, 1, frame=pf[0])
    c2 = O.pose_spherical(-75., -30., 4.)[:3, :4].cuda()
    want = net.forward_points(ps.sample_test(c2))
    for _ in range(2):
        pf[0].buf.fill_(-1.)
        pf[0].publish()                      # nobody starts storing before every rank has cleared its buffer
        g(c2)
        pf[0].publish()
        torch.cuda.synchronize()
        assert torch.equal(pf[0].buf, want), "graphed fused gather != single-GPU frame"
    # ray generation fused as well: one kernel per rank from pose to every GPU's frame buffer
    pf[1].buf.fill_(-1.)
    pf[1].publish()
    net.render_poses(ps, c2, frame=pf[1])
    pf[1].publish()
    torch.cuda.synchronize()
    assert torch.equal(pf[1].buf, want), "render_poses(frame=) != single-GPU frame"
    try:
        net.forward_points_gather(ps.sample_test(c2)[:128], pf[0])
        raise SystemExit("a wrong-sized block must be rejected")
    except ValueError:
        pass
dist.destroy_process_group()
print("ok", rank)
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_ray_sharded_frame_two_gpus_nccl(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29600 + os.getpid() % 300)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = []
    try:
        for p in procs:
            outs.append(p.communicate(timeout=240)[0])
    except subprocess.TimeoutExpired:      # one rank failed and left its peer inside a barrier
        for p in procs:
            p.kill()
        outs = [p.communicate()[0] for p in procs]
        raise AssertionError("2-GPU worker timed out:\n" + "\n".join(o[-3000:] for o in outs))
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o
