"""sha256 of one R2L frame (fused ray generation + MLP) and the handle-creation launch count: bit-identity between builds"""
import hashlib, sys, torch
sys.path.insert(0, '.')
import efficient_nerf_b200 as E
from oracle import ref_torch as O
with torch.no_grad():
    sd = O.r2l_state_dict(0)
    net = E.NeRF_v3_2(O.r2l_args(), 1008, 3, precision="fp16"); net.load_state_dict(sd); net = net.cuda().eval()
    ps = E.PointSampler(400, 400, O.LEGO["focal"], 16, 2., 6.)
    k0 = E._lib.kernel_launches()
    net.packed_handle()
    k1 = E._lib.kernel_launches()
    rgb = net.render_poses(ps, O.pose_spherical(30., -30., 4.)[:3, :4].cuda())
    torch.cuda.synchronize()
    print("pack launches", k1 - k0, "sha256", hashlib.sha256(rgb.cpu().numpy().tobytes()).hexdigest()[:16])
