"""In-kernel counters of the R2L ping-pong kernel (r2l_resmlp_profile on a pair-mode handle with R2L_PP=1)."""
import sys, torch
sys.path.insert(0, '.')
import efficient_nerf_b200 as E
from oracle import ref_torch as O
L = E._lib
sd = O.r2l_state_dict(0)
net = E.NeRF_v3_2(O.r2l_args(), 1008, 3, precision="fp16"); net.load_state_dict(sd); net = net.cuda().eval()
ps = E.PointSampler(400, 400, O.LEGO["focal"], 16, 2., 6.)
pts = ps.sample_test_batch(torch.stack([O.pose_spherical(10. * k, -30., 4.)[:3, :4] for k in range(4)]).cuda(), lazy=False)
N = pts.shape[0]
h = net.packed_handle()
rgb = torch.empty(N, 3, device="cuda"); prof = torch.zeros(148, 8, dtype=torch.int64, device="cuda")
for _ in range(3):
    L.call("r2l_resmlp_profile", h.h, N, L.ptr(pts), pts.stride(0), L.ptr(rgb), L.ptr(prof), L.stream_ptr())
torch.cuda.synchronize()
p = prof.double().cpu()[0::2]
units = N / 256 / 74
layers = units * 91
names = ["T0 total", "T0 wait a_ready", "T0 wait weights", "T1 total", "T1 wait a_ready", "T1 wait weights", "epi wait d_full", "epi total"]
for i, n in enumerate(names):
    print(f"{n:18s} mean {p[:, i].mean():12.0f} cyc  ({100 * p[:, i].mean() / p[:, 0].mean():5.1f}% of T0 total)  per layer {p[:, i].mean() / layers:8.0f}")
