"""Time the fused NeRF MLP call (view bias + MLP kernel + far fix-up) for S = 64 / 192 on a 160000-ray frame."""
import sys, torch
sys.path.insert(0, '.')
import efficient_nerf_b200 as E
def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
with torch.no_grad():
    coarse, fine = E.synthetic.seeded_nerf_pair(0, "fp16")
    c2w = E.synthetic.test_pose(0).cuda()
    ro, rd = E.get_rays(400, 400, E.synthetic.LEGO["focal"], c2w)
    ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
    vd = E.normalize_dirs(rd)
    for S in (64, 192):
        z = torch.sort(torch.rand(160000, S, device="cuda") * 4 + 2, -1)[0]
        for fix in (True, False):
            fine.set_far_fixup(fix)
            ms = timeit(lambda: fine.forward_samples(ro, rd, vd, z))
            print(f"S={S} far_fixup={fix}: {ms:.3f} ms  flagged={fine.far_flagged() if fix else '-'}")
