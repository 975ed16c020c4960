import sys, torch
sys.path.insert(0, '.')
import efficient_nerf_b200 as E
def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
N = 640000
L = E._lib
for S in (64, 128, 192):
    raw = torch.randn(N, S, 4, device="cuda"); z = torch.sort(torch.rand(N, S, device="cuda") * 4 + 2, -1)[0]; d = torch.randn(N, 3, device="cuda")
    rgb = torch.empty(N, 3, device="cuda"); disp = torch.empty(N, device="cuda"); acc = torch.empty(N, device="cuda"); w = torch.empty(N, S, device="cuda"); dep = torch.empty(N, device="cuda")
    ms = timeit(lambda: L.call("r2l_raw2outputs", N, S, L.ptr(raw), L.ptr(z), L.ptr(d), 3, None, 1, L.ptr(rgb), L.ptr(disp), L.ptr(acc), L.ptr(w), L.ptr(dep), L.stream_ptr()))
    print(f"raw2outputs S={S}: {ms:.3f} ms  {N * (24 * S + 36) / ms / 1e6:.0f} GB/s")
    del raw, w
z = torch.sort(torch.rand(N, 64, device="cuda") * 4 + 2, -1)[0]
bins = (.5 * (z[:, 1:] + z[:, :-1])).contiguous(); wts = torch.rand(N, 62, device="cuda")
u = torch.linspace(0, 1, 128).cuda(); out = torch.empty(N, 128, device="cuda")
ms = timeit(lambda: L.call("r2l_sample_pdf", N, 63, 128, L.ptr(bins), 63, L.ptr(wts), 62, L.ptr(u), 0, L.ptr(out), None, L.stream_ptr()))
print(f"sample_pdf Ni=128: {ms:.3f} ms  {N * 1012 / ms / 1e6:.0f} GB/s")
zs = torch.sort(out, -1)[0].contiguous(); zo = torch.empty(N, 192, device="cuda"); std = torch.empty(N, device="cuda")
ms = timeit(lambda: L.call("r2l_merge_sorted", N, 64, 128, L.ptr(z), L.ptr(zs), L.ptr(zo), L.ptr(std), L.stream_ptr()))
print(f"merge_sorted 64+128 (sorted fast path): {ms:.3f} ms  {N * 4 * (64 + 128 + 192) / ms / 1e6:.0f} GB/s")
ms = timeit(lambda: L.call("r2l_merge_sorted", N, 64, 128, L.ptr(z), L.ptr(out.flip(-1).contiguous()), L.ptr(zo), L.ptr(std), L.stream_ptr()), n=3)
print(f"merge_sorted 64+128 (general bitonic): {ms:.3f} ms")
