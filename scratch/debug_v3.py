import sys, ctypes, torch
sys.path.insert(0, '.')
import efficient_nerf_b200 as E
from oracle import ref_torch as O
L = E._lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sd = O.r2l_state_dict(0)
net = E.NeRF_v3_2(O.r2l_args(), 1008, 3, precision="fp16"); net.load_state_dict(sd); net = net.cuda().eval()
torch.manual_seed(0)
pts = ((torch.rand(n, 48) * 2 - 1) * 4).cuda()
h = net.packed_handle()
rgb = torch.empty(n, 3, device="cuda")
try:
    L.call("r2l_resmlp_forward", h.h, n, L.ptr(pts), pts.stride(0), L.ptr(rgb), L.stream_ptr())
    torch.cuda.synchronize()
    ref = O.r2l_forward(sd, O.embed_r2l(pts.cpu(), 10))
    print("ok, max err", float((rgb.cpu() - ref).abs().max()))
except Exception as e:
    print("FAILED:", str(e).splitlines()[0])
    out = (ctypes.c_uint * 8)()
    rc = L.load().r2l_mlp_status(h.h, out)
    print("status rc", rc, "dbg: flag %d block %d thread %d barrier %d parity %d" % tuple(out[:5]))
