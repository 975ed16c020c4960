"""Where a create_data pose goes: one stochastic teacher render (perturb=1, device RNG) under the torch profiler."""
import sys, time, torch
sys.path.insert(0, '.')
import efficient_nerf_b200 as E
from oracle import ref_torch as O
sdc, sdf = O.nerf_state_dicts(0)
nets = []
for sd in (sdc, sdf):
    n = E.NeRF(8, 256, 63, 27, 5, [4], True, precision="fp16"); n.load_state_dict(sd); nets.append(n.cuda().eval())
H = W = 400
focal = O.LEGO["focal"]
c2w = O.pose_spherical(30., -30., 4.)[:3, :4].cuda()
kw = dict(network_fn=nets[0], network_fine=nets[1], network_query_fn=None, N_samples=64, N_importance=128, perturb=1.,
          raw_noise_std=0., white_bkgd=True, use_viewdirs=True, lindisp=False, ndc=False, near=2., far=6., return_depth=True)
def pose(stochastic):
    ro, rd = E.get_rays(H, W, focal, c2w)
    k = dict(kw)
    if stochastic:
        k['t_rand'] = torch.rand((H * W, 64), device="cuda"); k['u'] = torch.rand((H * W, 128), device="cuda")
    else:
        k['perturb'] = 0.
    return E.render_image(H, W, focal, chunk=H * W, rays=(ro, rd), **k)
with torch.no_grad():
    for st in (False, True):
        for _ in range(3): pose(st)
        torch.cuda.synchronize(); t0 = time.time()
        for _ in range(10): pose(st)
        torch.cuda.synchronize(); print("stochastic" if st else "deterministic", (time.time() - t0) / 10 * 1e3, "ms/pose")
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3): pose(True)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
