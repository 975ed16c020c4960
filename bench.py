#!/usr/bin/env python
"""Benchmark of the per-ray rendering hot path (BASELINE.json metric: Mrays/s and ms/frame at 400x400).

    python bench.py --gpus N --steps K --warmup W [--workload r2l|nerf] [--impl reference]

Workload (config.workload):
  r2l  (default, BASELINE configs[1]): R2L lego_noview resmlp W256 D88, n_sample_per_ray=16, 400x400
       synthetic test poses; one STEP = one launch of the fused kernel (PointSampler ray generation +
       PositionalEmbedder + 88-layer ResMLP) over the step's poses.
  nerf (BASELINE configs[0]): NeRF lego W256 D8, 64 coarse + 128 fine samples, one STEP = one 400x400 frame
       through get_rays -> fused encode+MLP (coarse) -> raw2outputs -> hier_sample -> fused encode+MLP (fine) ->
       raw2outputs.  At N = 1 the default run reports it as a full record of its own under `extras.nerf`
       (roofline with event-timed MLP launches, e2e with declared copies, cpu_baseline, parity census).
Random-init weights (torch.manual_seed(0), the reference's construction order), synthetic poses
pose_spherical(theta_k, -30, 4): data = "synthetic".

Multi-GPU: one process per GPU (torchrun), poses sharded round-robin, NO data-path collective (weak
scaling: every rank renders K poses); time = max over ranks; value = all rays / that time.  For N > 1 a short
second phase renders single frames RAY-sharded over the ranks (strong scaling: the gather of the rgb tiles is the
one collective of the path; R2L fuses it into the MLP kernel over peer memory) and reports `extras.ray_sharded`.

Timing: W >= 3 warm-up steps, then K steps; every step is bracketed by CUDA events on the launching
stream (torch's current stream, which is the stream the C ABI launches on); an L2 flush (256 MiB
memset) runs between steps outside the event brackets.  `value` is the device-resident number (inputs
already in HBM); `e2e` is measured separately through the public API with the pose coming from pinned
host memory and the finished frame copied back to pinned host memory inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if "--impl" in sys.argv and sys.argv[sys.argv.index("--impl") + 1:][:1] == ["reference"] or "--impl=reference" in sys.argv:
    # the reference arm times the reference's CPU path: its modules pick `device` at import (model/nerf_raybased.py:10,
    # utils/run_nerf_raybased_helpers.py:13), so this process must not see a GPU (set before torch is imported)
    os.environ["CUDA_VISIBLE_DEVICES"] = ""

import numpy as np  # noqa: E402
import torch  # noqa: E402

H = W = 400
RAYS = H * W
FLOP_PER_RAY = {"r2l": 11789824, "nerf": 303824896}   # BASELINE.md §2 (unpadded MACs x2)


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernels, from the committed `ncu --set
    full` captures (profiles/traffic.json names the capture each figure was read from)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d.get("hbm_gbs", 6650.0), tf_burst=d.get("bf16_tflops", 1590.0),
                    tf_sust=d.get("bf16_tflops_sustained", 1400.0), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the GPU is under load.  Reported: the samples of
    the timed region plus the last 0.5 s of the (identical) warm-up loop before it — a 20-step timed region lasts
    ~0.13 s, less than nvidia-smi's own latency."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.t_mark = None
        self.t_end = None

    def start(self):
        """Start sampling (before the warm-up: nvidia-smi needs ~0.5 s to deliver its first line)."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
            t0 = time.time()
            while not self.lines and time.time() - t0 < 5.0:
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def mark(self):
        """The timed region starts now."""
        self.t_mark = time.time()

    def end(self):
        """The timed region has ended (samples that arrive within 0.15 s still describe it)."""
        self.t_end = time.time()

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        if self.t_end is not None:
            time.sleep(max(0.0, self.t_end + 0.15 - time.time()))
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lo = (self.t_mark or 0.0) - 0.5
        hi = (self.t_end or time.time()) + 0.15
        n_timed = 0
        for ts, ln in self.lines:
            if ts < lo or ts > hi:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])), mx.append(float(f[1])), pw.append(float(f[2]))
            except ValueError:
                continue
            n_timed += ts >= (self.t_mark or 0.0)
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm),
                "samples_in_timed_region": int(n_timed),
                "window": "timed region + the last 0.5 s of the identical warm-up loop, 50 ms period"}


# ------------------------------------------------------------------------------------ workloads
def make_poses(E, n, offset=0, stride=1):
    return [E.synthetic.test_pose(offset + k * stride) for k in range(n)]


class R2LWorkload:
    name = "r2l"
    kernels_per_step = 1   # r2l_mlp_kernel (ray generation fused)

    def __init__(self, E, precision, poses_per_launch=1):
        self.P = int(poses_per_launch)
        self.rays_per_step = RAYS * self.P
        self.block = None
        self.graph = None
        self.frames = None
        self.net = E.synthetic.seeded_r2l(0, precision)
        self.ps = E.PointSampler(H, W, E.synthetic.LEGO["focal"], 16, 2., 6.)
        self.E = E
        self.net.packed_handle()
        self.ev = None

    def enable_fused_gather(self):
        """--shard rays --gather fused: two symmetric-memory frame buffers (alternating), the MLP kernel stores its
        tiles into every GPU's buffer by peer-to-peer stores; a barrier replaces the NCCL all-gather."""
        self.frames = [self.E.sharding.PeerFrame(RAYS), self.E.sharding.PeerFrame(RAYS)]

    def enable_graph(self):
        """--graph: sampler + fused MLP captured once into a CUDA graph, one replay per step."""
        if self.frames is not None:
            self.graphs = [self.E.GraphedR2L(self.net, self.ps, 1, frame=f) for f in self.frames]
            self.graph = self.graphs[0]
        else:
            self.graph = self.E.GraphedR2L(self.net, self.ps, self.P, rows=self.block)

    def step(self, c2w_dev, mlp_events=None, k=0):
        # c2w_dev: [P, 3, 4] — P consecutive test poses rendered by ONE fused launch
        if self.graph is not None:
            if mlp_events is not None:      # no events inside a graph: the bracket is the whole replay
                mlp_events[0].record()
            out = (self.graphs[k & 1] if self.frames is not None else self.graph)(c2w_dev)
            if mlp_events is not None:
                mlp_events[1].record()
            return out
        if mlp_events is not None:
            mlp_events[0].record()
        # ONE kernel: rays from pixel index + pose, stratified points, encoding, 88-layer ResMLP.  --shard rays: this
        # rank's contiguous block of the frame's rays; --gather fused: tiles stored into every GPU's frame buffer k & 1
        rgb = self.net.render_poses(self.ps, c2w_dev, rows=self.block,
                                    frame=self.frames[k & 1] if self.frames is not None else None)
        if mlp_events is not None:
            mlp_events[1].record()
        return rgb

    def mlp_ms(self, mev, steps):
        return sum(a.elapsed_time(b) for a, b in mev) / steps

    def describe(self):
        return {"workload": "R2L lego_noview resmlp W256 D88, n_sample_per_ray=16, 400x400 synthetic poses "
                            f"(BASELINE configs[1]); step = one launch of {self.P} consecutive test pose(s) = "
                            f"{self.rays_per_step} rays (1250 tiles of 128 rays per frame leave a ragged 9th wave on "
                            "148 SMs; batching poses fills whole waves)",
                "poses_per_launch": self.P, "rays_per_step": self.rays_per_step, "operands": self.net.precision,
                "accumulate": "fp32", "cuda_graph": self.graph is not None}


class NerfWorkload:
    name = "nerf"
    kernels_per_step = 15

    def __init__(self, E, precision, poses_per_launch=1):
        self.P, self.rays_per_step = 1, RAYS
        self.block = None
        self.frames = None
        self.graph = None
        self.coarse, self.fine = E.synthetic.seeded_nerf_pair(0, precision)
        self.coarse.packed_handle(), self.fine.packed_handle()
        self.E, self.focal = E, E.synthetic.LEGO["focal"]
        self.kw = dict(network_query_fn=None, perturb=0., N_importance=128, network_fine=self.fine, N_samples=64,
                       network_fn=self.coarse, use_viewdirs=True, white_bkgd=True, raw_noise_std=0., ndc=False,
                       near=2., far=6.)
        self.mlp_events = []
        self.coarse._mlp_events = self.fine._mlp_events = self.mlp_events   # events around the MLP launches

    def step(self, c2w_dev, mlp_events=None, k=0):
        if c2w_dev.dim() == 3:
            c2w_dev = c2w_dev[0]
        if self.block is not None:          # --shard rays: render only this rank's block of the frame's rays
            ro, rd = self.E.get_rays(H, W, self.focal, c2w_dev)
            s0, s1 = self.block
            rays = torch.stack([ro.reshape(-1, 3)[s0:s1], rd.reshape(-1, 3)[s0:s1]], 0)
            rgb, disp, acc, _ = self.E.render_image(H, W, self.focal, chunk=32768, rays=rays, **self.kw)
            return rgb.reshape(-1, 3)
        rgb, disp, acc, _ = self.E.render_image(H, W, self.focal, chunk=32768, c2w=c2w_dev, **self.kw)
        return rgb.reshape(-1, 3)

    def mlp_ms(self, mev, steps):
        """CUDA-event time of the MLP launches (coarse + fine: view bias, fused MLP, far-sample fix-up) per step, over
        the LAST `steps` steps recorded."""
        evs = self.mlp_events[-2 * steps:]
        ms = sum(a.elapsed_time(b) for a, b in evs) / steps
        del self.mlp_events[:]
        return ms

    def describe(self):
        return {"workload": "NeRF lego W256 D8, 64 coarse + 128 fine samples, 400x400 synthetic poses "
                            "(BASELINE configs[0]); step = 1 frame = 160000 rays",
                "rays_per_step": RAYS, "operands": self.coarse.precision, "accumulate": "fp32", "cuda_graph": False}


# ------------------------------------------------------------------------------------ CPU arm
def cpu_reference(workload, steps, warmup, sample_rays):
    """The reference's OWN code on a bounded sample of the same frame, all host threads: the unmodified
    model/nerf_raybased.py, utils/run_nerf_raybased_helpers.py and main.py render functions from baseline/_ref
    (staged by baseline/stage_reference.py, hashes pinned; kind = "reference").  Only if no staged tree exists does it
    fall back to the oracle port (kind = "port": the same torch-CPU ops, pinned bit-exact to the reference).  The ONLY
    place bench.py touches oracle/ (the checker's loader is used to time the CPU baseline; the GPU arm never does)."""
    from oracle import ref_torch as O
    from oracle import ref_real as R
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    cam = O.LEGO
    c2w = O.pose_spherical(-180., -30., 4.)[:3, :4]
    idx = torch.linspace(0, RAYS - 1, sample_rays).long()
    times = []
    root = R.reference_root(required=False)
    kind = "reference" if root is not None else "port"
    with torch.no_grad():
        if root is not None:
            ref = R.load(root, device="cpu")
            if workload == "r2l":
                # main.py:297-309: model(positional_embedder(point_sampler.sample_test(c2w))), on the sampled rows
                model, pe, ps = R.build_r2l(ref, 0, H, W, cam["focal"], 16, 2., 6.)
                fn = lambda: model(pe(ps.sample_test(c2w)[idx]))
            else:
                # main.py:283-290: render(H, W, focal, chunk, rays=...) with create_nerf's render_kwargs_test
                coarse, fine, kw = R.build_nerf(ref, 0)
                ro, rd = ref.Hh.get_rays(H, W, cam["focal"], c2w)
                rays = (ro.reshape(-1, 3)[idx], rd.reshape(-1, 3)[idx])
                fn = lambda: ref.main["render"](H, W, cam["focal"], chunk=1024 * 32, rays=rays, **kw)[0]
        elif workload == "r2l":
            sd = O.r2l_state_dict(0)
            fn = lambda: O.render_r2l(sd, H, W, cam["focal"], 2., 6., c2w, rows=idx)
        else:
            sdc, sdf = O.nerf_state_dicts(0)
            ro, rd = O.get_rays(H, W, cam["focal"], c2w)
            batch = O.pack_rays(ro.reshape(-1, 3)[idx], rd.reshape(-1, 3)[idx], 2., 6.)
            fn = lambda: O.render_rays(batch, sdc, sdf, 64, 128, white_bkgd=True)["rgb_map"]
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            fn()
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    mean = float(np.mean(times))
    what = ("the unmodified reference from baseline/_ref: model/nerf_raybased.py + utils/run_nerf_raybased_helpers.py + "
            "main.py render functions" if kind == "reference" else
            "oracle/ref_torch.py = the reference's torch-CPU path (no staged reference tree found)")
    return dict(value=sample_rays / mean / 1e6, unit="Mrays/s", cores=cores, kind=kind,
                sample=f"{sample_rays} evenly spaced rays of one 400x400 frame per step, {len(times)} timed steps "
                       f"({what}); ms/frame extrapolated = {mean / sample_rays * RAYS * 1e3:.0f}"), mean


def run_reference_arm(args):
    rank, local, world = dist_env()
    if rank != 0:
        return 0
    sample = 65536 if args.workload == "r2l" else 4096
    steps = max(1, min(args.steps, 3))
    warmup = min(args.warmup, 1)
    cb, mean = cpu_reference(args.workload, steps, warmup, sample)
    line = {"impl": "reference", "metric": "render_throughput", "value": cb["value"], "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": mean * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[args.workload], "rays_per_step": sample,
                       "note": "CPU reference arm: each step is a bounded sample of the 160000-ray frame"},
            "ms_per_frame_400x400": mean / sample * RAYS * 1e3,
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0,
                                        "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


WORKLOAD_DESC = {
    "r2l": "R2L lego_noview resmlp W256 D88, n_sample_per_ray=16, 400x400 synthetic poses (BASELINE configs[1])",
    "nerf": "NeRF lego W256 D8, 64 coarse + 128 fine samples, 400x400 synthetic pose (BASELINE configs[0])",
}


# ------------------------------------------------------------------------------------ GPU arm
def roofline_block(workload, rays_per_launch, kernel_ms, peaks, P=1):
    flops = FLOP_PER_RAY[workload] * rays_per_launch
    ach = flops / (kernel_ms * 1e-3) / 1e12
    tr = ncu_traffic().get(workload, {})
    traffic = tr.get("bytes_per_frame")
    if workload == "r2l" and "bytes_weights" in tr:
        # the weights are read once per launch, the rgb rows scale with the poses of the launch
        traffic = tr.get("bytes_weights", 0) + tr.get("bytes_per_pose", 0) * P
    return {"bound": "tensor",
            "kernel": "r2l_mlp_kernel" if workload == "r2l" else
                      "nerf_mlp_pp_kernel, coarse + fine launch of one frame (+ view bias and far-sample fix-up)",
            "achieved": ach, "peak": peaks["tf_sust"], "unit": "TFLOP/s", "frac": ach / peaks["tf_sust"],
            "frac_of_burst": ach / peaks["tf_burst"],
            "peak_source": f"{peaks['src']} bf16_tflops_sustained (kernel timed inside a long step; burst "
                           f"{peaks['tf_burst']})",
            "kernel_ms": kernel_ms, "algorithmic_flop_per_launch": flops, "traffic": traffic,
            "traffic_unit": f"B per step (dram__bytes_read.sum + dram__bytes_write.sum, {tr.get('source', 'no capture')})"}


def timed_run(wl, poses_dev, warmup, steps, run_step, flush, barrier, sampler=None, on_start=None):
    """Device-resident timing: W warm-up steps (+ 0.6 s more of the identical loop), then K event-bracketed steps with
    an L2 flush before each; returns (per-step ms, MLP events, extra warm-up steps, wall seconds)."""
    def timed_pass(pose_list, with_mlp_events):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in pose_list]
        mevs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in pose_list]
        for k, pose in enumerate(pose_list):
            flush.zero_()                              # L2 flush, outside the event bracket
            evs[k][0].record()
            run_step(pose, mevs[k] if with_mlp_events else None, k)
            evs[k][1].record()
        return evs, mevs

    # Warm-up = the SAME loop (same allocation pattern, same kernels, host running ahead): W steps, then more until
    # 0.6 s have passed.  Two start-up transients otherwise land in the K timed steps: the power-cap controller
    # throttles hard ~0.3 s after a cold GPU starts a tensor-heavy kernel (one 60-70 ms NeRF step among 42 ms
    # ones), and the first pass through the un-synchronised loop contains one step with a ~90 ms host-side stall.
    timed_pass(poses_dev[:warmup], True)
    torch.cuda.synchronize()
    t_warm = time.perf_counter()
    extra_warm = 0
    while time.perf_counter() - t_warm < 0.6 and extra_warm < 400:
        n_more = max(warmup, 4)
        timed_pass([poses_dev[k % warmup] for k in range(n_more)], True)
        torch.cuda.synchronize()
        extra_warm += n_more
    barrier()
    if on_start is not None:
        on_start()
    if sampler is not None:
        sampler.mark()
    t_wall0 = time.perf_counter()
    ev, mev = timed_pass(poses_dev[warmup:warmup + steps], True)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    if sampler is not None:
        sampler.end()
    return [a.elapsed_time(b) for a, b in ev], mev, extra_warm, t_wall


def e2e_run(wl, poses, warmup, steps, run_step, barrier, rank, by_rays=False, fused=False):
    """End to end through the public API with HOST buffers: every step copies its pose(s) from pinned host memory
    (H2D) and its finished frame(s) back to pinned host memory (D2H, on a second stream, double-buffered, so the copy
    of step i overlaps the kernels of step i+1 — what a caller streaming frames does).  Returns total ms."""
    P = wl.P
    pose_host = [p.pin_memory() for p in poses]
    frame_host2 = [torch.empty((wl.rays_per_step if not by_rays else RAYS, 3), dtype=torch.float32).pin_memory()
                   for _ in range(2)]
    c2w_buf = torch.empty((P, 3, 4), dtype=torch.float32, device="cuda")
    copy_stream = torch.cuda.Stream()
    main_stream = torch.cuda.current_stream()
    copied = [None, None]

    def one(i, pose):
        c2w_buf.copy_(pose, non_blocking=True)                            # H2D of this step's input
        # fused gather: peers overwrite symmetric buffer (i + 1) & 1 once they pass publish(i); the read-back of
        # the frame it still holds (step i - 1) must be over before this rank joins that barrier
        wait_prev = (lambda: main_stream.wait_event(copied[(i + 1) & 1])) if fused and copied[(i + 1) & 1] else None
        out = run_step(c2w_buf, None, i, wait_prev)
        if not by_rays or rank == 0:
            done = torch.cuda.Event()
            done.record(main_stream)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done)
                frame_host2[i & 1].copy_(out, non_blocking=True)          # D2H of this step's result
                if not fused:      # (symmetric-memory buffers are not the caching allocator's)
                    out.record_stream(copy_stream)
                copied[i & 1] = torch.cuda.Event()
                copied[i & 1].record(copy_stream)

    # untimed warm-up of THIS loop (an even number of steps, so the double buffering and the fused gather's buffer
    # parity continue seamlessly): first use of the copy stream and of the pinned buffers — a 2-GPU run once lost
    # 330 ms of a 240 ms timed region to it
    n_w = 2 * max(1, min(warmup, 4) // 2)
    for i in range(n_w):
        one(i, pose_host[i % max(1, warmup)])
    main_stream.wait_stream(copy_stream)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    verbose = bool(os.environ.get("BENCH_VERBOSE"))
    t_host = []
    for i in range(steps):
        if verbose:
            t_host.append(time.perf_counter())
        one(n_w + i, pose_host[warmup + i])
    main_stream.wait_stream(copy_stream)
    e1.record()
    barrier()
    if verbose and len(t_host) > 1:      # host-side pacing of the loop: where an end-to-end step loses time
        d = [1e3 * (b - a) for a, b in zip(t_host, t_host[1:])]
        print(f"[e2e rank {rank}] host ms between step launches: " + " ".join(f"{x:.2f}" for x in d[:60]), file=sys.stderr)
    return e0.elapsed_time(e1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", choices=["r2l", "nerf"], default="r2l")
    ap.add_argument("--precision", choices=["fp16", "bf16"], default="fp16")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--poses-per-launch", type=int, default=0,
                    help="R2L: test poses rendered per launch (step); default 4.  NeRF always renders 1 frame per step")
    ap.add_argument("--shard", choices=["poses", "rays"], default="poses",
                    help="poses (default): rank r renders poses r, r+G, ... (weak scaling, no data-path collective); "
                         "rays: every frame is split into contiguous ray blocks over the ranks and the tiles are "
                         "all-gathered with NCCL (strong scaling: ms per frame)")
    ap.add_argument("--gather", choices=["nccl", "fused"], default="nccl",
                    help="--shard rays, R2L: nccl = one all_gather_into_tensor per frame after the MLP kernel; fused = the "
                         "MLP kernel stores its tiles into every GPU's symmetric-memory frame buffer (peer-to-peer "
                         "stores over NVLink) and a cross-GPU barrier publishes the frame")
    ap.add_argument("--graph", action="store_true",
                    help="R2L: replay the step (sampler + fused MLP) as one CUDA graph instead of launching it from Python")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    rank, local, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (sm_100a); there is no CPU fallback")
    torch.cuda.set_device(local)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    import efficient_nerf_b200 as E
    E._lib.load()
    by_rays = args.shard == "rays"
    P = (args.poses_per_launch or (1 if by_rays else 4)) if args.workload == "r2l" else 1
    wl = (R2LWorkload if args.workload == "r2l" else NerfWorkload)(E, args.precision, P)
    if by_rays:
        if P != 1:
            raise SystemExit("--shard rays renders one frame per step")
        wl.block = E.sharding.shard_rays(RAYS, rank, world)
    fused = args.gather == "fused"
    if fused:
        if not (by_rays and args.workload == "r2l" and world > 1):
            raise SystemExit("--gather fused applies to --shard rays with the R2L workload on more than one GPU")
        wl.enable_fused_gather()
    if args.graph:
        if args.workload != "r2l":
            raise SystemExit("--graph is implemented for the R2L workload")
        wl.enable_graph()
    warmup = max(args.warmup, 3)
    steps = max(args.steps, 1)
    # pose-sharded: rank r renders poses r, r+G, ...; ray-sharded: every rank works on the SAME pose sequence
    flat = make_poses(E, (steps + warmup) * P, offset=0 if by_rays else rank, stride=1 if by_rays else world)
    poses = [torch.stack(flat[i * P:(i + 1) * P], 0).contiguous() for i in range(steps + warmup)]   # [P, 3, 4] per step
    poses_dev = [p.cuda() for p in poses]
    RAYS_STEP = wl.rays_per_step
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    frame_buf = torch.empty((RAYS, 3), dtype=torch.float32, device="cuda") if by_rays else None

    def run_step(pose, mev=None, k=0, before_publish=None):
        out = wl.step(pose, mev, k)
        if fused:       # the tiles are already in every GPU's buffer k & 1: a cross-GPU barrier publishes the frame
            if before_publish is not None:
                before_publish()
            out = wl.frames[k & 1].publish()
        elif by_rays:   # NCCL all-gather of the finished tiles: the whole frame on every rank
            out = E.sharding.gather_rays_into(out.contiguous(), frame_buf, RAYS)
        return out
    peaks = measured_peaks()

    with torch.no_grad():
        sampler = ClockSampler(local)
        if rank == 0 and not os.environ.get("BENCH_NO_SAMPLER"):
            sampler.start()
        counts = {}

        def on_start():      # the library's launch counters are read around the timed pass only
            counts["abi"], counts["kernels"] = E._lib.launch_count, E._lib.kernel_launches()

        # ---------------- device-resident timing
        step_ms, mev, extra_warm, t_wall = timed_run(wl, poses_dev, warmup, steps, run_step, flush, barrier, sampler,
                                                     on_start)
        launches = E._lib.launch_count - counts["abi"]
        kernels = E._lib.kernel_launches() - counts["kernels"]
        if getattr(wl, "graph", None) is not None:     # replays do not pass through the library's launch counter
            kernels += wl.graph.kernels_per_replay * steps
        clocks = sampler.stop() if rank == 0 else None
        dev_ms = sum(step_ms)
        if os.environ.get("BENCH_VERBOSE"):
            print("per-step ms:", " ".join(f"{x:.2f}" for x in step_ms[:40]), file=sys.stderr)
        mlp_ms = wl.mlp_ms(mev, steps)

        # ---------------- end-to-end timing through the public API with host buffers
        e2e_steps = steps
        e2e_ms = e2e_run(wl, poses, warmup, e2e_steps, run_step, barrier, rank, by_rays, fused)

    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    n_jobs = 1 if by_rays else world      # ray-sharded: all ranks work on the same frames
    total_rays = RAYS_STEP * steps * n_jobs
    value = total_rays / (dev_ms * 1e-3) / 1e6
    e2e_value = RAYS_STEP * e2e_steps * n_jobs / (e2e_ms * 1e-3) / 1e6

    line = None
    if rank == 0:
        line = {"metric": "render_throughput", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": steps,
                "warmup": warmup, "warmup_extra_steps": extra_warm, "ms_per_step": dev_ms / steps,
                "ms_per_step_median": float(np.median(step_ms)), "ms_per_step_min": float(min(step_ms)),
                "ms_per_step_max": float(max(step_ms)), "higher_is_better": True,
                "scaling": "strong" if by_rays else "weak",
                "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                "config": dict(wl.describe(),
                               sharding=(("contiguous ray blocks of every frame (multiples of 128 rays); the MLP kernel "
                                          "stores its rgb tiles into every GPU's symmetric-memory frame buffer "
                                          "(peer-to-peer stores) + one cross-GPU barrier per frame, inside the timed "
                                          "region" if fused else
                                          "contiguous ray blocks of every frame (multiples of 128 rays) + one NCCL "
                                          "all_gather_into_tensor of the rgb tiles per frame, inside the timed region")
                                         if by_rays else "pose round-robin, no data-path collective"),
                               l2="flushed between steps (256 MiB memset outside the event brackets)",
                               timing="sum of per-step CUDA-event durations, max over ranks"),
                "ms_per_frame_400x400": dev_ms / steps / P,
                "wall_ms_per_step_incl_flush": t_wall / steps * 1e3,
                "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 48 * P,
                        "d2h_bytes_per_step": RAYS_STEP * 3 * 4, "ms_per_step": e2e_ms / e2e_steps},
                "gpu_launches": int(kernels), "abi_calls": int(launches),
                "clocks": clocks}
        line["roofline"] = roofline_block(args.workload, RAYS_STEP // (world if by_rays else 1), mlp_ms, peaks, P)
        if world == 1 and not args.no_cpu_baseline:
            sample = 16384 if args.workload == "r2l" else 1024
            line["cpu_baseline"], _ = cpu_reference(args.workload, 2, 1, sample)
    if world == 1 and not args.no_extras:
        line["extras"] = extras(E, peaks, args.precision, args.workload, flush, no_cpu=args.no_cpu_baseline)
    if world > 1 and not args.no_extras and not by_rays:
        # strong scaling of ONE frame, driver-visible: rays of every frame sharded over the ranks.  The headline line
        # must survive this secondary phase: an exception is recorded, and a hang (a collective that one rank never
        # enters) is cut off by an alarm that prints the line without the phase and leaves.
        import signal

        def give_up(signum, frame):
            if rank == 0:
                line.setdefault("extras", {})["ray_sharded"] = {"error": "timed out after 240 s"}
                print(json.dumps(line), flush=True)
            os._exit(0)
        signal.signal(signal.SIGALRM, give_up)
        signal.alarm(240)
        try:
            rs = ray_sharded_phase(E, dist, args.precision, rank, world, barrier)
        except Exception as e:
            rs = {"error": f"{type(e).__name__}: {e}"}
        signal.alarm(0)
        if rank == 0:
            line.setdefault("extras", {})["ray_sharded"] = rs
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def ray_sharded_phase(E, dist, precision, rank, world, barrier):
    """One 400x400 frame split into contiguous ray blocks over the ranks: R2L with the tile gather fused
    into the MLP kernel (peer-to-peer stores into every GPU's symmetric-memory frame buffer + one cross-GPU barrier) and
    the rank's kernel replayed as a CUDA graph; NeRF with one NCCL all_gather_into_tensor per frame.  Reports ms per
    frame (device, max over ranks), e2e (pose from pinned host memory, frame back to rank 0's host memory) and whether
    the gathered frame is bit-identical to the same frame rendered by ONE GPU."""
    out = {}
    with torch.no_grad():
        def measure(wl, run_step, fused, frames):
            poses1 = [p.reshape(1, 3, 4).contiguous() for p in make_poses(E, frames + 4)]
            poses_dev = [p.cuda() for p in poses1]
            for k in range(4):
                run_step(poses_dev[k], None, k)
            barrier()
            evs = []
            for k in range(frames):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                run_step(poses_dev[4 + k], None, k)
                b.record()
                evs.append((a, b))
            barrier()
            ms = sum(a.elapsed_time(b) for a, b in evs)
            e2e = e2e_run(wl, poses1, 4, frames, run_step, barrier, rank, True, fused)
            t = torch.tensor([ms, e2e], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t[0]) / frames, float(t[1]) / frames, poses_dev[0]

        def all_same(a, b):
            same = torch.tensor([int(torch.equal(a, b))], device="cuda")
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            return bool(int(same[0]))

        # ---- R2L: fused gather + graph replay
        wl = R2LWorkload(E, precision, 1)
        wl.block = E.sharding.shard_rays(RAYS, rank, world)
        wl.enable_fused_gather()
        wl.enable_graph()

        def step_r2l(pose, mev=None, k=0, before_publish=None):
            wl.step(pose, None, k)
            if before_publish is not None:
                before_publish()
            return wl.frames[k & 1].publish()
        ms, e2e, pose0 = measure(wl, step_r2l, True, 60)
        got = step_r2l(pose0, None, 0).clone()
        whole = wl.net.render_poses(wl.ps, pose0)           # the same frame on this GPU alone
        out["r2l"] = {"ms_per_frame_400x400": ms, "Mrays_per_s": RAYS / ms / 1e3,
                      "e2e": {"ms_per_frame": e2e, "Mrays_per_s": RAYS / e2e / 1e3, "h2d_bytes_per_step": 48,
                              "d2h_bytes_per_step": RAYS * 12},
                      "bit_identical_to_single_gpu": all_same(got, whole),
                      "gather": "fused into r2l_mlp_kernel: rgb tiles stored to every GPU's frame buffer over NVLink + "
                                "one symmetric-memory barrier per frame; CUDA-graph replay",
                      "tiles_per_gpu": (wl.block[1] - wl.block[0] + 127) // 128}
        barrier()
        # ---- NeRF: NCCL gather of the rgb tiles
        wn = NerfWorkload(E, precision)
        wn.block = E.sharding.shard_rays(RAYS, rank, world)
        frame_buf = torch.empty((RAYS, 3), dtype=torch.float32, device="cuda")

        def step_nerf(pose, mev=None, k=0, before_publish=None):
            return E.sharding.gather_rays_into(wn.step(pose).contiguous(), frame_buf, RAYS)
        ms, e2e, pose0 = measure(wn, step_nerf, False, 12)
        got = step_nerf(pose0).clone()
        wn.block = None
        whole = wn.step(pose0)
        out["nerf"] = {"ms_per_frame_400x400": ms, "Mrays_per_s": RAYS / ms / 1e3,
                       "e2e": {"ms_per_frame": e2e, "Mrays_per_s": RAYS / e2e / 1e3, "h2d_bytes_per_step": 48,
                               "d2h_bytes_per_step": RAYS * 12},
                       "bit_identical_to_single_gpu": all_same(got, whole),
                       "gather": "one NCCL all_gather_into_tensor of the rgb tiles per frame"}
    out["scaling"] = "strong"
    out["timing"] = "sum of per-frame CUDA-event durations (gather / barrier inside), max over ranks"
    return out


def extras(E, peaks, precision, main_workload, flush, no_cpu=False):
    """Secondary measurements reported beside the headline: the OTHER model as a full record of its own (roofline with
    event-timed MLP launches, e2e with declared copies, cpu_baseline, parity census), single-pose R2L, the other BASELINE
    configs and the HBM-bound kernels."""
    out = {}
    with torch.no_grad():
        def timeit(fn, n=10, warm=3):
            for _ in range(warm):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / n

        def full_record(wl, name, steps, warmup=3):
            """The same measurement the headline gets (timed_run + e2e_run), for another workload."""
            P = wl.P
            flat = make_poses(E, (steps + warmup) * P)
            poses = [torch.stack(flat[i * P:(i + 1) * P], 0).contiguous() for i in range(steps + warmup)]
            poses_dev = [p.cuda() for p in poses]
            run_step = lambda pose, mev=None, k=0, before_publish=None: wl.step(pose, mev, k)
            sync = torch.cuda.synchronize
            step_ms, mev, extra_warm, _ = timed_run(wl, poses_dev, warmup, steps, run_step, flush, sync)
            mlp_ms = wl.mlp_ms(mev, steps)
            e2e_ms = e2e_run(wl, poses, warmup, steps, run_step, sync, 0)
            ms = sum(step_ms) / steps
            rec = {"workload": wl.describe()["workload"], "steps": steps, "ms_per_step": ms,
                   "ms_per_step_median": float(np.median(step_ms)), "ms_per_step_max": float(max(step_ms)),
                   "ms_per_frame_400x400": ms / P, "Mrays_per_s": wl.rays_per_step / ms / 1e3,
                   "e2e": {"value": wl.rays_per_step * steps / e2e_ms / 1e3, "unit": "Mrays/s",
                           "ms_per_step": e2e_ms / steps, "h2d_bytes_per_step": 48 * P,
                           "d2h_bytes_per_step": wl.rays_per_step * 12},
                   "roofline": roofline_block(name, wl.rays_per_step, mlp_ms, peaks, P),
                   "l2": "flushed between steps", "timing": "sum of per-step CUDA-event durations"}
            rec["tensor_TFLOPs"] = FLOP_PER_RAY[name] * wl.rays_per_step / ms / 1e9
            rec["frac_of_sustained_peak"] = rec["tensor_TFLOPs"] / peaks["tf_sust"]
            return rec

        other = "nerf" if main_workload == "r2l" else "r2l"
        wl = (R2LWorkload if other == "r2l" else NerfWorkload)(E, precision, 4 if other == "r2l" else 1)
        out[other] = full_record(wl, other, 12 if other == "nerf" else 100)
        if not no_cpu:
            out[other]["cpu_baseline"], _ = cpu_reference(other, 2, 1, 1024 if other == "nerf" else 16384)
        pose = make_poses(E, 1)[0].cuda()
        if main_workload == "r2l":   # the reference's own step: ONE pose per forward (main.py:297-309)
            w1 = R2LWorkload(E, precision, 1)
            out["r2l_one_pose_per_launch"] = full_record(w1, "r2l", 100)
        nw = wl if other == "nerf" else NerfWorkload(E, precision)
        out["parity"] = parity_block(E, nw, precision)
        # the other BASELINE configs, one frame each (configs[3]: 800x800; configs[4]: LLFF fern 504x378, NDC, 64+64)
        if main_workload == "r2l":
            cam8 = E.synthetic.LEGO_800
            r2l = R2LWorkload(E, precision, 1)
            ps8 = E.PointSampler(cam8["H"], cam8["W"], cam8["focal"], 16, 2., 6.)
            n8 = cam8["H"] * cam8["W"]
            ms = timeit(lambda: r2l.net.render_poses(ps8, pose), n=20)
            out["r2l_800x800"] = {"ms_per_frame": ms, "Mrays_per_s": n8 / ms / 1e3,
                                  "tensor_TFLOPs": FLOP_PER_RAY["r2l"] * n8 / ms / 1e9}
            kw8 = dict(nw.kw)
            ms = timeit(lambda: E.render_image(cam8["H"], cam8["W"], cam8["focal"], chunk=32768, c2w=pose, **kw8), n=3, warm=1)
            out["nerf_800x800"] = {"ms_per_frame": ms, "Mrays_per_s": n8 / ms / 1e3,
                                   "tensor_TFLOPs": FLOP_PER_RAY["nerf"] * n8 / ms / 1e9}
            fern = E.synthetic.FERN
            kwf = dict(nw.kw, N_importance=64, white_bkgd=False, ndc=True, near=0., far=1.)
            nf = fern["H"] * fern["W"]
            ms = timeit(lambda: E.render_image(fern["H"], fern["W"], fern["focal"], chunk=32768, c2w=pose, **kwf), n=3, warm=1)
            out["nerf_fern_504x378_ndc_64+64"] = {"ms_per_frame": ms, "Mrays_per_s": nf / ms / 1e3,
                                                  "tensor_TFLOPs": 227868672 * nf / ms / 1e9}
        # HBM-bound kernels on 4x the frame (inputs > L2): raw2outputs S=192 / S=64, sample_pdf Ni=128, hier_sample
        N = 4 * RAYS
        raw = torch.randn(N, 192, 4, device="cuda")
        z = torch.sort(torch.rand(N, 192, device="cuda") * 4 + 2, -1)[0]
        d = torch.randn(N, 3, device="cuda")
        ms = timeit(lambda: E.raw2outputs(raw, z, d, 0, True))
        gbs = N * (24 * 192 + 36) / ms / 1e6
        out["raw2outputs_S192"] = {"ms": ms, "GB_per_s": gbs, "frac_of_hbm": gbs / peaks["hbm"], "rays": N}
        raw64, z64 = raw[:, :64].contiguous(), z[:, :64].contiguous()
        ms = timeit(lambda: E.raw2outputs(raw64, z64, d, 0, True))
        gbs = N * (24 * 64 + 36) / ms / 1e6
        out["raw2outputs_S64"] = {"ms": ms, "GB_per_s": gbs, "frac_of_hbm": gbs / peaks["hbm"], "rays": N}
        del raw, raw64
        bins = z[:, :63].contiguous()
        w = torch.rand(N, 62, device="cuda")
        ms = timeit(lambda: E.sample_pdf(bins, w, 128, det=True))
        gbs = N * 1012 / ms / 1e6
        out["sample_pdf_Ni128"] = {"ms": ms, "GB_per_s": gbs, "frac_of_hbm": gbs / peaks["hbm"], "rays": N}
        # the fused form render_rays uses (mids + sample_pdf + sorted merge + z_std): 512 B in, 772 B out per ray
        wc = torch.rand(N, 64, device="cuda")
        ut = torch.linspace(0., 1., 128)
        ms = timeit(lambda: E.run_nerf_raybased_helpers.hier_sample(z64, wc, 128, ut))
        gbs = N * 1284 / ms / 1e6
        out["hier_sample_64+128"] = {"ms": ms, "GB_per_s": gbs, "frac_of_hbm": gbs / peaks["hbm"], "rays": N,
                                     "note": "the unfused kernels move 2548 B/ray for the same result"}
    return out


def parity_block(E, nerf_wl, precision):
    """Whole-view parity of the bench pose, on the device: the fused tensor-core path against the precision='fp32'
    CUDA-core path (pinned <= 1e-5 to the torch-CPU goldens in tests/).  NeRF: rays beyond the 2e-3 gate in rgb_map /
    rgb0 with and without the far-sample fix-up, rays flagged; R2L: max |rgb - fp32| over the frame."""
    out = {}
    pose = make_poses(E, 1)[0].cuda()
    focal = E.synthetic.LEGO["focal"]
    with torch.no_grad():
        c32, f32 = E.synthetic.seeded_nerf_pair(0, "fp32")
        embed_fn, _ = E.get_embedder(10, 0)
        embeddirs_fn, _ = E.get_embedder(4, 0)
        qfn = lambda inputs, viewdirs, fn: E.run_network(inputs, viewdirs, fn, embed_fn, embeddirs_fn, 1024 * 64)
        kw32 = dict(nerf_wl.kw, network_fn=c32, network_fine=f32, network_query_fn=qfn)
        rgb32, _, _, ex32 = E.render_image(H, W, focal, chunk=8192, c2w=pose, **kw32)
        res = {}
        for name, on in (("with_far_fixup", True), ("without_far_fixup", False)):
            nerf_wl.coarse.set_far_fixup(on), nerf_wl.fine.set_far_fixup(on)
            rgb, _, _, ex = E.render_image(H, W, focal, chunk=32768, c2w=pose, **nerf_wl.kw)
            d = (rgb - rgb32).abs().reshape(-1, 3).max(-1)[0]
            d0 = (ex["rgb0"] - ex32["rgb0"]).abs().reshape(-1, 3).max(-1)[0]
            mse = float(((rgb - rgb32).double() ** 2).mean())
            res[name] = {"rays_beyond_2e-3_rgb_map": int((d > 2e-3).sum()), "rays_beyond_2e-3_rgb0": int((d0 > 2e-3).sum()),
                         "max_abs_rgb_map": float(d.max()), "max_abs_rgb0": float(d0.max()),
                         "psnr_vs_fp32_dB": float(-10. * np.log10(max(mse, 1e-30)))}
            if on:
                res[name]["rays_flagged_fine_pass"] = nerf_wl.fine.far_flagged()
        nerf_wl.coarse.set_far_fixup(True), nerf_wl.fine.set_far_fixup(True)
        out["nerf_400x400_vs_fp32_path"] = dict(res, operands=precision, rays=RAYS,
                                                gate="rgb within 2e-3 max-abs per view (north_star)")
        net = E.synthetic.seeded_r2l(0, precision)
        net32 = E.synthetic.seeded_r2l(0, "fp32")
        ps = E.PointSampler(H, W, focal, 16, 2., 6.)
        rgb = net.render_poses(ps, pose)
        pe = E.PositionalEmbedder(10)
        pts = ps.sample_test(pose)
        rgb32 = torch.cat([net32(pe(pts[i:i + 32768])) for i in range(0, RAYS, 32768)], 0)
        d = (rgb - rgb32).abs()
        out["r2l_400x400_vs_fp32_path"] = {"max_abs_rgb": float(d.max()), "rays_beyond_2e-3": int((d.max(-1)[0] > 2e-3).sum()),
                                           "operands": precision, "rays": RAYS}
    return out


if __name__ == "__main__":
    sys.exit(main())
