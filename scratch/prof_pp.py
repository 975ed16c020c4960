"""Ping-pong NeRF kernel: ring turnaround / turn-token waits of issuer T0 (R2L_NERF_PP=1 R2L_PROF_MODE=1)."""
import sys, torch
sys.path.insert(0, '.')
import efficient_nerf_b200 as E
from oracle import ref_torch as O
L = E._lib
with torch.no_grad():
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
    import os
    p = prof.double().cpu()[0::2]   # leader CTAs
    tot = p[:, 0].mean()
    import os
    if os.environ.get("R2L_PROF_MODE") == "5":
        prof = torch.zeros(512, 8, dtype=torch.int64, device="cuda")
        L.call("r2l_nerf_profile", h.h, N, S, L.ptr(ro), ro.stride(0), L.ptr(rd), rd.stride(0), L.ptr(vd), vd.stride(0),
               L.ptr(z), L.ptr(raw), L.ptr(prof), L.stream_ptr())
        torch.cuda.synchronize()
        flat = prof.cpu().reshape(-1)
        ev = [(int(flat[9 + 2 * i]), int(flat[8 + 2 * i])) for i in range(600) if int(flat[8 + 2 * i]) != 0]
        ev.sort(); t0 = ev[0][0]
        kinds = {1: "issue", 2: "a_done", 3: "epi-start", 4: "epi-end"}
        for tm, tag in ev:
            k, t, step, st = tag // 100000, (tag // 10000) % 10, (tag // 100) % 100, tag % 100
            print(f"{tm - t0:7d}  T{t} step {step} {kinds[k]}" + (f" stage {'P' if st == 9 else st}" if k == 1 else ""))
        sys.exit(0)
    if os.environ.get("R2L_PROF_MODE") == "4":
        pa = prof.double().cpu()
        n_tl = 160000 * 192 / 128 / 148 * 10
        print(f"per tile-layer (mean over CTAs): wait::ld WG0 {pa[:,3].mean()/n_tl:.0f} WG1 {pa[:,5].mean()/n_tl:.0f}; fences+arrive (WG0) {pa[:,4].mean()/n_tl:.0f}; bias barriers (WG1) {pa[:,6].mean()/n_tl:.0f}")
    who = "T1" if os.environ.get("R2L_PROF_MODE") == "1" else "T0"
    print(f"issuer {who} total {tot:.0f} cyc: wait a_done {100*p[:,1].mean()/tot:.1f}%  wait W+bias {100*p[:,2].mean()/tot:.1f}%  wait enc {100*p[:,7].mean()/tot:.1f}%")
    # determinism stress: a protocol race shows up as run-to-run differences
    ref = net.forward_samples(ro, rd, vd, z).clone()
    bad = 0
    for i in range(8):
        out = net.forward_samples(ro, rd, vd, z)
        bad += int((out != ref).any())
    print("determinism: %d of 8 reruns differ; finite=%s" % (bad, bool(torch.isfinite(ref).all())))
