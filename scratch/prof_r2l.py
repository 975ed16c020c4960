import sys, torch
sys.path.insert(0, '.')
import efficient_nerf_b200 as E
from oracle import ref_torch as O
L = E._lib
sd = O.r2l_state_dict(0)
net = E.NeRF_v3_2(O.r2l_args(), 1008, 3, precision="fp16"); net.load_state_dict(sd); net = net.cuda().eval()
ps = E.PointSampler(400, 400, O.LEGO["focal"], 16, 2., 6.)
pts = ps.sample_test(O.pose_spherical(0., -30., 4.)[:3, :4].cuda())
h = net.packed_handle()
rgb = torch.empty(160000, 3, device="cuda"); prof = torch.zeros(148, 8, dtype=torch.int64, device="cuda")
for _ in range(3):
    L.call("r2l_resmlp_profile", h.h, 160000, L.ptr(pts), pts.stride(0), L.ptr(rgb), L.ptr(prof), L.stream_ptr())
torch.cuda.synchronize()
p = prof.double().cpu()
names = ["MMA total", "MMA wait A", "MMA wait W", "WG0 wait D", "WG0 epilogue", "WG1 wait D", "WG1 epilogue", "WG0 encode"]
tiles = 1250 / 148
layers = tiles * 90
for i, n in enumerate(names):
    print(f"{n:14s} mean {p[:, i].mean():12.0f} cyc  ({100 * p[:, i].mean() / p[:, 0].mean():5.1f}% of MMA total)  per layer-step {p[:, i].mean() / layers:8.0f}")
