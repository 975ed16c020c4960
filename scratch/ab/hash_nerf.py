"""sha256 of one coarse + fine NeRF forward on a fixed input: bit-identity check between two builds of the library"""
import hashlib, sys, torch
sys.path.insert(0, '.')
import efficient_nerf_b200 as E
from oracle import ref_torch as O
with torch.no_grad():
    sdc, sdf = O.nerf_state_dicts(0)
    g = torch.Generator().manual_seed(5)
    n = 20011
    o = (torch.randn(n, 3, generator=g) * 0.3).cuda(); d = torch.randn(n, 3, generator=g).cuda()
    v = d / d.norm(dim=-1, keepdim=True)
    h = hashlib.sha256()
    for sd, S in ((sdc, 64), (sdf, 192)):
        net = E.NeRF(8, 256, 63, 27, 5, [4], True, precision="fp16"); net.load_state_dict(sd); net = net.cuda().eval()
        z = torch.sort(2 + 4 * torch.rand(n, S, generator=g), -1)[0].cuda()
        raw = net.forward_samples(o, d, v, z)
        h.update(raw.cpu().numpy().tobytes())
    print("sha256", h.hexdigest()[:16])
