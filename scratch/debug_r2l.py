import sys, ctypes, numpy as np, torch
sys.path.insert(0, '.')
import efficient_nerf_b200 as E
from oracle import ref_torch as O
L = E._lib
torch.manual_seed(0)
n_points = 4
args = O.r2l_args(netdepth=6, use_residual=True)
net = E.NeRF_v3_2(args, n_points * 63, 3, precision="fp16").cuda().eval()
pts = ((torch.rand(128, n_points * 3) * 2 - 1) * 4).cuda()
h = net.packed_handle()
rgb = torch.empty(128, 3, device="cuda"); acc = torch.zeros(128, 256, device="cuda"); x0 = torch.zeros(128, 256, device="cuda")
a_img = torch.zeros(65536 + 256, dtype=torch.uint8, device="cuda")
for it in range(2):
    L.call("r2l_resmlp_debug_head", h.h, 128, L.ptr(pts), pts.stride(0), L.ptr(rgb), L.ptr(acc), L.ptr(x0), L.ptr(a_img), L.stream_ptr())
    torch.cuda.synchronize()
    clk = a_img[65536:65536 + 16 * 8].cpu().view(torch.int64)
    base = int(clk[clk > 0].min())
    names = {0: "WG0 encode done+arrive", 1: "WG1 encode done+arrive", 2: "MMA: d_full commit issued", 4: "WG0 passed d_full", 5: "WG1 passed d_full",
             8: "MMA passed a_ready[0][0]", 9: "MMA passed a_ready[0][1]", 10: "MMA passed a_ready[1][0]", 11: "MMA passed a_ready[1][1]"}
    print("iteration", it, "nan rows:", torch.nonzero(torch.isnan(acc).any(1)).flatten().tolist())
    for k in sorted(names):
        if int(clk[k]) > 0: print(f"  {names[k]:32s} +{int(clk[k]) - base}")
