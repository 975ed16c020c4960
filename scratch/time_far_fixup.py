"""Cost of the far-sample fix-up per MLP call: forward_samples with / without it, alternating, median of 9."""
import sys, torch, statistics
sys.path.insert(0, '.')
import efficient_nerf_b200 as E
def once(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)
with torch.no_grad():
    coarse, fine = E.synthetic.seeded_nerf_pair(0, "fp16")
    c2w = E.synthetic.test_pose(0).cuda()
    ro, rd = E.get_rays(400, 400, E.synthetic.LEGO["focal"], c2w)
    ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
    vd = E.normalize_dirs(rd)
    for net, S, name in ((coarse, 64, "coarse"), (fine, 192, "fine")):
        z = (torch.linspace(0, 1, S).cuda() * 4 + 2)[None].expand(160000, S).contiguous()
        f = lambda: net.forward_samples(ro, rd, vd, z)
        for _ in range(3): f()
        on, off = [], []
        for i in range(9):
            net.set_far_fixup(True); on.append(once(f))
            net.set_far_fixup(False); off.append(once(f))
        net.set_far_fixup(True); f()
        print(f"{name} S={S}: with fix-up {statistics.median(on):.3f} ms, without {statistics.median(off):.3f} ms, "
              f"delta {statistics.median(on) - statistics.median(off):+.3f} ms, flagged {net.far_flagged()}")
