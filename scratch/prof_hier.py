"""hier_sample / sample_pdf on 640 000 rays (inputs > L2): timing, and the launch ncu captures (scratch/README.md)."""
import sys, torch
sys.path.insert(0, '.')
import efficient_nerf_b200 as E
H = E.run_nerf_raybased_helpers
def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
N = int(sys.argv[1]) if len(sys.argv) > 1 else 640000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
torch.manual_seed(0)
z = torch.sort(torch.rand(N, 64, device="cuda") * 4 + 2, -1)[0].contiguous()
w = torch.rand(N, 64, device="cuda")
w8 = (torch.rand(N, 64, device="cuda") ** 8).contiguous()      # sparse weights (a trained model's)
ut = torch.linspace(0., 1., 128)
for name, ww in (("uniform", w), ("sparse", w8)):
    ms = timeit(lambda: H.hier_sample(z, ww, 128, ut), n=n)
    print(f"hier_sample 64+128 {name}: {ms:.4f} ms  {N * 1284 / ms / 1e6:.0f} GB/s")
ut64 = torch.linspace(0., 1., 64)
ms = timeit(lambda: H.hier_sample(z, w, 64, ut64), n=n)
print(f"hier_sample 64+64: {ms:.4f} ms  {N * (512 + 4 * 129) / ms / 1e6:.0f} GB/s")
bins = (.5 * (z[:, 1:] + z[:, :-1])).contiguous(); wts = w[:, 1:-1].contiguous()
ms = timeit(lambda: E.sample_pdf(bins, wts, 128, det=True), n=n)
print(f"sample_pdf Ni=128: {ms:.4f} ms  {N * 1012 / ms / 1e6:.0f} GB/s")
