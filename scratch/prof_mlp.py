"""In-kernel cycle counters of the fused MLP kernels (r2l_resmlp_profile / r2l_nerf_profile) + event timing."""
import sys, torch
sys.path.insert(0, '.')
import efficient_nerf_b200 as E
from oracle import ref_torch as O
L = E._lib
which = sys.argv[1] if len(sys.argv) > 1 else "both"

def report(p, names, layers):
    for i, n in enumerate(names):
        if n:
            print(f"  {n:14s} mean {p[:, i].mean():12.0f} cyc ({100 * p[:, i].mean() / p[:, 0].mean():5.1f}% of MMA total) per layer {p[:, i].mean() / layers:8.0f}   max-CTA {p[:, i].max():12.0f}")

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

with torch.no_grad():
    if which in ("r2l", "both"):
        sd = O.r2l_state_dict(0)
        net = E.NeRF_v3_2(O.r2l_args(), 1008, 3, precision="fp16"); net.load_state_dict(sd); net = net.cuda().eval()
        ps = E.PointSampler(400, 400, O.LEGO["focal"], 16, 2., 6.)
        pts = ps.sample_test(O.pose_spherical(0., -30., 4.)[:3, :4].cuda())
        h = net.packed_handle()
        rgb = torch.empty(160000, 3, device="cuda"); prof = torch.zeros(148, 8, dtype=torch.int64, device="cuda")
        for _ in range(3):
            L.call("r2l_resmlp_profile", h.h, 160000, L.ptr(pts), pts.stride(0), L.ptr(rgb), L.ptr(prof), L.stream_ptr())
        torch.cuda.synchronize()
        ms = timeit(lambda: net.forward_points(pts))
        print(f"R2L 160000 rays: {ms:.3f} ms  -> {11789824 * 160000 / ms / 1e9:.1f} TFLOP/s")
        report(prof.double().cpu(), ["MMA total", "MMA wait A", "MMA wait W", "WG0 wait D", "WG0 epilogue", "WG1 wait D", "WG1 epilogue", "WG0 encode"], 1250 / 148 * 90)
    if which in ("nerf", "both"):
        sdc, _ = O.nerf_state_dicts(0)
        net = E.NeRF(8, 256, 63, 27, 5, [4], True, precision="fp16"); net.load_state_dict(sdc); net = net.cuda().eval()
        c2w = O.pose_spherical(0., -30., 4.)[:3, :4].cuda()
        ro, rd = E.get_rays(400, 400, O.LEGO["focal"], c2w)
        ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
        vd = E.normalize_dirs(rd)
        N, S = 160000, 192
        z = torch.sort(torch.rand(N, S, device="cuda") * 4 + 2, -1)[0]
        h = net.packed_handle()
        raw = torch.empty(N, S, 4, device="cuda"); prof = torch.zeros(148, 8, dtype=torch.int64, device="cuda")
        for _ in range(2):
            L.call("r2l_nerf_profile", h.h, N, S, L.ptr(ro), ro.stride(0), L.ptr(rd), rd.stride(0), L.ptr(vd), vd.stride(0),
                   L.ptr(z), L.ptr(raw), L.ptr(prof), L.stream_ptr())
        torch.cuda.synchronize()
        ms = timeit(lambda: net.forward_samples(ro, rd, vd, z), n=3)
        print(f"NeRF {N}x{S} samples: {ms:.3f} ms -> {1186816 * N * S / ms / 1e9:.1f} TFLOP/s")
        report(prof.double().cpu(), ["MMA total", "MMA wait A", "MMA wait W", "WG0 wait D", "WG0 epilogue", "WG1 wait D", "WG1 epilogue", "MMA wait enc"], N * S / 128 / 148 * 10)
