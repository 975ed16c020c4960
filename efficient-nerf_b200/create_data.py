"""Teacher-NeRF pseudo-data generation: the `--create_data rand` mode of the reference's utils/create_data.py
(:777-872), BASELINE config 3, with the renders going through the fused B200 path and the work sharded
over ranks.

What the reference does, and what is kept bit-for-bit in the file format:
  for i in 1..n_pose_kd:
      pose  = get_rand_pose()                          dataset/load_blender.py:359-368 (2 np.random draws)
      focal_= focal * (np.random.rand() + 1)           create_data.py:816-818 (unless --no_rand_focal)
      rays  = get_rays(H, W, focal_, pose[:3, :4])     create_data.py:819-820
      rgb, depth = render(rays=..., **train kwargs)    create_data.py:824-832 (perturb = 1 by default)
      data += [o(3), d(3), rgb(3) (, depth 1 | surface point 3)]   create_data.py:836-841
      every 100 poses: rows shuffled by two successive np.random.permutation, cut into float32 [4096, C] files
      data_{k}.npy, k counting from 1 (or from the number of files already present); the tail that does not fill
      a file is dropped                                create_data.py:854-872
The files are what BlenderDataset_v2 (dataset/load_blender.py:257-324) trains on.

Sharding: a GROUP of `i_save` = 100 poses is the unit.  Group g always produces the same
file indices  g*F+1 .. (g+1)*F  (F = i_save*H*W // split_size), so ranks write disjoint files with no
communication; rank r takes groups r, r+G, ...  Random streams:
  stream="reference": ONE global np.random / torch CPU stream consumed exactly like the reference does (every
                      rank replays the draws of the groups it skips) — world_size 1 reproduces the reference's
                      pose / focal / shuffle sequence for a given np.random.seed;
  stream="per_group": group g uses np.random.RandomState(seed + g) and torch.manual_seed(seed + g): independent
                      of the number of ranks, no replay cost (default when world_size > 1).
`fast_rng=True` draws the stratified-sampling / inverse-CDF randoms on the device instead of the CPU generator
(not parity with the reference's draws; the file format and statistics are unchanged).
"""
import os
import queue
import threading
import time

import numpy as np
import torch

from .render import render
from .run_nerf_raybased_helpers import get_rays


def pose_spherical(theta, phi, radius):
    """Camera-to-world matrix on a sphere (dataset/load_blender.py:10-28), float32 like the reference."""
    def trans_t(t):
        return torch.Tensor([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, t], [0, 0, 0, 1]]).float()

    def rot_phi(a):
        return torch.Tensor([[1, 0, 0, 0], [0, np.cos(a), -np.sin(a), 0], [0, np.sin(a), np.cos(a), 0],
                             [0, 0, 0, 1]]).float()

    def rot_theta(a):
        return torch.Tensor([[np.cos(a), 0, -np.sin(a), 0], [0, 1, 0, 0], [np.sin(a), 0, np.cos(a), 0],
                             [0, 0, 0, 1]]).float()

    c2w = trans_t(radius)
    c2w = rot_phi(phi / 180. * np.pi) @ c2w
    c2w = rot_theta(theta / 180. * np.pi) @ c2w
    c2w = torch.Tensor([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]) @ c2w
    return c2w


def get_rand_pose(rng=np.random):
    """Random pose on the upper hemisphere, radius 4 (dataset/load_blender.py:359-368): theta ~ U[-180, 180),
    phi ~ U[-90, 0), drawn in that order from `rng`."""
    theta = -180 + rng.rand() * 360
    phi = -90 + rng.rand() * 90
    return pose_spherical(theta, phi, 4)


def files_per_group(H, W, i_save=100, split_size=4096):
    return (i_save * H * W) // split_size


class _AsyncNpySaver:
    """np.save on worker threads: a group is ~3 900 files of 147 KB; without this the host I/O dominates."""

    def __init__(self, n_threads=4):
        self.q = queue.Queue(maxsize=8)
        self.err = []
        self.threads = [threading.Thread(target=self._run, daemon=True) for _ in range(max(1, n_threads))]
        for t in self.threads:
            t.start()

    def _run(self):
        while True:
            job = self.q.get()
            if job is None:
                return
            barrier = None
            try:
                datadir, first, arr, split_size, barrier = job
                for j in range(arr.shape[0] // split_size):
                    np.save(os.path.join(datadir, f"data_{first + j}.npy"), arr[j * split_size:(j + 1) * split_size])
            except Exception as e:  # surfaced by close()
                self.err.append(e)
            finally:
                if barrier is not None:
                    barrier.done()

    class Barrier:
        """Counts the jobs of one submit() call; wait() returns when all of them have been written."""

        def __init__(self, _n_hint=0):
            self.n, self.cv = 0, threading.Condition()

        def add(self):
            with self.cv:
                self.n += 1

        def done(self):
            with self.cv:
                self.n -= 1
                self.cv.notify_all()

        def wait(self):
            with self.cv:
                while self.n > 0:
                    self.cv.wait()

    def submit(self, datadir, first, arr, split_size, n_parts, barrier=None):
        """Split the group's files into n_parts contiguous jobs."""
        n_files = arr.shape[0] // split_size
        per = (n_files + n_parts - 1) // max(1, n_parts)
        for a in range(0, n_files, max(1, per)):
            b = min(n_files, a + per)
            if barrier is not None:
                barrier.add()
            self.q.put((datadir, first + a, arr[a * split_size:b * split_size], split_size, barrier))

    def close(self):
        for _ in self.threads:
            self.q.put(None)
        for t in self.threads:
            t.join()
        if self.err:
            raise self.err[0]


def create_data_rand(teacher_fn, teacher_fine, datadir, n_pose_kd, H, W, focal, near=2., far=6., chunk=1024 * 32,
                     use_rand_focal=True, learn_depth='', perturb=1., raw_noise_std=0., N_samples=64,
                     N_importance=128, white_bkgd=True, use_viewdirs=True, lindisp=False, rank=0, world_size=1,
                     stream=None, seed=0, i_save=100, split_size=4096, resume=True, fast_rng=False,
                     writer_threads=4, render_fn=None, device=None, progress=None):
    """Render `n_pose_kd` random poses with the teacher and write the shuffled [split_size, C] float32 .npy shards.
    Returns the list of file indices written by this rank.  `render_fn` (tests) replaces the renderer:
    render_fn(H, W, focal, rays_o, rays_d) -> (rgb [H,W,3], depth [H,W])."""
    if stream is None:
        stream = "reference" if world_size == 1 else "per_group"
    if stream not in ("reference", "per_group"):
        raise ValueError("stream must be 'reference' or 'per_group'")
    if n_pose_kd % i_save != 0:
        raise ValueError(f"n_pose_kd ({n_pose_kd}) must be a multiple of i_save ({i_save}): the reference only "
                         "writes complete groups (create_data.py:854)")
    os.makedirs(datadir, exist_ok=True)
    F = files_per_group(H, W, i_save, split_size)
    n_groups = n_pose_kd // i_save
    split0 = 0
    if resume and stream == "reference" and world_size == 1:
        # the reference resumes by counting the .npy files already present (create_data.py:790-796)
        split0 = len([x for x in os.listdir(datadir) if x.endswith('.npy')])
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
    C = 9 + (0 if not learn_depth else (3 if learn_depth == 'surface' else 1))
    saver = _AsyncNpySaver(writer_threads)
    written = []
    rng_global = np.random  # the reference's global stream

    query_fn = None
    if render_fn is None:
        # a teacher that is not on the fused tensor-core path (other widths / depths, precision='fp32') is queried like
        # the reference does (create_data.py:283-293): embed + batchified forward
        from .render import _fused_ok, run_network
        if not all(_fused_ok(n, True if use_viewdirs else None) for n in (teacher_fn, teacher_fine) if n is not None):
            from .run_nerf_raybased_helpers import get_embedder
            embed_fn, _ = get_embedder(10, 0)
            embeddirs_fn = get_embedder(4, 0)[0] if use_viewdirs else None

            def query_fn(inputs, viewdirs, network_fn):
                return run_network(inputs, viewdirs, network_fn, embed_fn, embeddirs_fn, netchunk=1024 * 64)

    def render_pose(focal_, c2w):
        rays_o, rays_d = get_rays(H, W, focal_, c2w)  # [H, W, 3] on the device
        kw = dict(network_fn=teacher_fn, network_fine=teacher_fine, network_query_fn=query_fn, N_samples=N_samples,
                  N_importance=N_importance, perturb=perturb, raw_noise_std=raw_noise_std, white_bkgd=white_bkgd,
                  use_viewdirs=use_viewdirs, lindisp=lindisp, ndc=False, near=near, far=far, return_depth=True)
        if fast_rng and perturb > 0.:
            N = H * W
            kw['t_rand'] = torch.rand((N, N_samples), device=rays_o.device)
            kw['u'] = torch.rand((N, N_importance), device=rays_o.device)
            chunk_ = max(chunk, N)   # one launch per pose: the injected draws cover the whole frame
        else:
            chunk_ = chunk
        rgb, _, _, extras = render(H, W, focal, chunk=chunk_, rays=(rays_o, rays_d), **kw)
        return rays_o, rays_d, rgb, extras['depth_map']

    # Host random numbers of a group (poses, focal scales, the two shuffle permutations) do not depend on the
    # renders, so a helper thread draws them — in the reference's stream order — one group ahead of the GPU
    # (two np.random.permutation(16M) cost ~1 s per group on the host).
    n_rows = i_save * H * W
    todo = queue.Queue(maxsize=2)

    def draw_groups():
        try:
            for g in range(n_groups):
                mine = (g % world_size) == rank
                first = split0 + g * F + 1
                if stream == "per_group":
                    if not mine:
                        continue
                    if resume and os.path.exists(_group_marker(datadir, first, F)):
                        continue   # the marker is written after ALL of the group's files are on disk
                    rng = np.random.RandomState(seed + g)
                else:
                    rng = rng_global
                poses, focals = [], []
                for _ in range(i_save):
                    poses.append(get_rand_pose(rng))
                    focals.append(focal * (rng.rand() + 1) if use_rand_focal else focal)
                # the group's poses are published BEFORE its two permutations are drawn (same stream order as the
                # reference: poses, focals, then the permutations), so the GPU starts rendering the group while the
                # host spends ~1 s on np.random.permutation(16M) x 2; the renderer picks the shuffle up when it is done
                ix_box = {"ready": threading.Event()}
                if mine:
                    todo.put((g, first, poses, focals, ix_box))
                try:
                    # drawn even for groups this rank skips: same global stream as the reference
                    ix1 = rng.permutation(n_rows)
                    ix2 = rng.permutation(n_rows)
                    ix_box["ix"] = torch.from_numpy(ix1[ix2]) if mine else None   # data[ix1][ix2] == data[ix1[ix2]]
                except Exception as e:
                    ix_box["err"] = e
                    raise
                finally:
                    ix_box["ready"].set()
            todo.put(None)
        except Exception as e:
            todo.put(e)

    drawer = threading.Thread(target=draw_groups, daemon=True)
    drawer.start()
    # GPU path: poses are written straight into a [rows, C] group buffer (two, alternating); the group's shuffle
    # (one device gather) and its read-back into pinned host memory run on a side stream while the next group
    # renders; a finisher thread hands each host group to the .npy writers when its copy event has fired.
    on_gpu = render_fn is None and device is not None and device.type == "cuda"
    finisher = None
    if on_gpu:
        n_buf = 3   # pinned host groups in flight: one being filled, up to two with the writers
        copy_stream = torch.cuda.Stream(device)
        group_bufs = [torch.empty((n_rows, C), dtype=torch.float32, device=device) for _ in range(2)]
        shuffled = torch.empty((n_rows, C), dtype=torch.float32, device=device)
        host_bufs = [torch.empty((n_rows, C), dtype=torch.float32).pin_memory() for _ in range(n_buf)]
        host_free = queue.Queue()
        for b in range(n_buf):
            host_free.put(b)
        buf_idle = [None, None]          # event: the side stream has finished reading group_bufs[i]
        fin_q = queue.Queue()
        fin_err = []

        def finish_groups():
            try:
                while True:
                    job = fin_q.get()
                    if job is None:
                        return
                    ev, b, first = job
                    ev.synchronize()
                    host = host_bufs[b].numpy()
                    # the writers read the pinned buffer in place, so it goes back to the free list only when this
                    # group's files are on disk
                    done = _AsyncNpySaver.Barrier(max(1, writer_threads))
                    saver.submit(datadir, first, host[:F * split_size], split_size, n_parts=max(1, writer_threads),
                                 barrier=done)
                    done.wait()
                    _mark_group_done(datadir, first, F)
                    host_free.put(b)
            except Exception as e:
                fin_err.append(e)
                host_free.put(-1)

        finisher = threading.Thread(target=finish_groups, daemon=True)
        finisher.start()
    try:
        n_done = 0
        while True:
            item = todo.get()
            if item is None:
                break
            if isinstance(item, Exception):
                raise item
            g, first, poses, focals, ix_box = item
            _t0 = time.time()
            if stream == "per_group":
                torch.manual_seed(seed + g)   # CPU-generator draws inside the renders (t_rand, u, noise)
            rows = []
            gbuf = None
            if on_gpu:
                gi = n_done & 1
                gbuf = group_bufs[gi]
                if buf_idle[gi] is not None:
                    torch.cuda.current_stream(device).wait_event(buf_idle[gi])
            for k, (pose, focal_) in enumerate(zip(poses, focals)):
                c2w = pose[:3, :4]
                if render_fn is not None:
                    o, d = _host_rays(H, W, focal_, c2w)
                    rgb, depth = render_fn(H, W, focal_, o, d)
                else:
                    o, d, rgb, depth = render_pose(focal_, c2w.to(device))
                parts = [o, d, rgb]
                if learn_depth:
                    depth = depth[..., None]
                    parts.append(o + d * depth.expand_as(d) if learn_depth == 'surface' else depth)
                parts = [p.reshape(H * W, -1).float() for p in parts]
                if gbuf is not None:
                    torch.cat(parts, -1, out=gbuf[k * H * W:(k + 1) * H * W])
                else:
                    rows.append(torch.cat(parts, -1))
                if progress is not None:
                    progress(g, k + 1)
            _t1 = time.time()
            ix_box["ready"].wait()
            _t2 = time.time()
            if "err" in ix_box:
                raise ix_box["err"]
            ix = ix_box["ix"]
            if gbuf is not None:
                b = host_free.get()
                if os.environ.get("R2L_CD_TIMING"):
                    torch.cuda.synchronize()
                    print(f"[create_data] group {g}: enqueue {_t1 - _t0:.2f} s, wait perm {_t2 - _t1:.2f} s, "
                          f"wait host buffer {time.time() - _t2:.2f} s (incl. device sync)", flush=True)
                if fin_err:
                    raise fin_err[0]
                rendered = torch.cuda.Event()
                rendered.record(torch.cuda.current_stream(device))
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(rendered)
                    torch.index_select(gbuf, 0, ix.to(device, non_blocking=True), out=shuffled)   # data[ix1][ix2]
                    host_bufs[b].copy_(shuffled, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                buf_idle[n_done & 1] = ev
                fin_q.put((ev, b, first))
            else:
                data = torch.cat(rows, 0)                       # [i_save*H*W, C]
                data = data[ix.to(data.device)]
                host = data.cpu().numpy() if data.is_cuda else data.numpy()
                assert host.dtype == np.float32 and host.shape[1] == C
                done = _AsyncNpySaver.Barrier(max(1, writer_threads))
                saver.submit(datadir, first, host[:F * split_size], split_size, n_parts=max(1, writer_threads),
                             barrier=done)
                done.wait()
                _mark_group_done(datadir, first, F)
            written.extend(range(first, first + F))
            n_done += 1
        if finisher is not None:
            fin_q.put(None)
            finisher.join()
            finisher = None
            if fin_err:
                raise fin_err[0]
    finally:
        if finisher is not None:     # error path: let the finisher drain and stop
            fin_q.put(None)
            finisher.join()
        saver.close()
    return written


def _group_marker(datadir, first, F):
    """Completion marker of the group whose files are data_{first} .. data_{first+F-1}.  A group's files are written
    by several threads over disjoint ranges, so 'the last file exists' does not mean 'the group is complete'; the
    marker is created only after every writer of the group has finished.  (Not a .npy: the reference's own resume
    counts the .npy files of the directory, create_data.py:790-796.)"""
    return os.path.join(datadir, f".group_{first}_{first + F - 1}.done")


def _mark_group_done(datadir, first, F):
    with open(_group_marker(datadir, first, F), "w") as f:
        f.write("ok\n")


def _host_rays(H, W, focal, c2w):
    """CPU get_rays for the render_fn test hook only (utils/run_nerf_raybased_helpers.py:231-247)."""
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing='ij')
    i, j = i.t(), j.t()
    dirs = torch.stack([(i - W * .5) / focal, -(j - H * .5) / focal, -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., np.newaxis, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def load_shards(datadir, indices=None):
    """Read data_{k}.npy shards back (what BlenderDataset_v2 does, dataset/load_blender.py:271-290)."""
    if indices is None:
        indices = sorted(int(x[5:-4]) for x in os.listdir(datadir) if x.startswith('data_') and x.endswith('.npy'))
    return np.concatenate([np.load(os.path.join(datadir, f"data_{k}.npy")) for k in indices], 0)
