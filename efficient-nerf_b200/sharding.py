"""Multi-GPU sharding of the rendering path (one process per GPU, torch.distributed).

Every ray is independent, so the path shards with no data-path collective:
  * many poses (R2L test set, create_data):  rank r renders poses r, r+G, r+2G, ...
  * one frame:  contiguous ray blocks, multiples of 128 rays so MMA tiles stay full
The only collective is one all_gather of the finished image tiles (NCCL on GPUs, gloo in CPU tests).
The reference's equivalent is nn.DataParallel scatter/gather per forward (main.py:37-42, 472-479).
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_poses(n_poses, rank=None, world_size=None):
    """Indices of the poses this rank renders (round-robin, like a strided test split)."""
    if rank is None:
        rank, world_size = world()
    return list(range(rank, n_poses, world_size))


def shard_rays(n_rays, rank=None, world_size=None, multiple=128):
    """[start, stop) of this rank's contiguous ray block; every block but the last is a multiple of
    `multiple` rays.  Blocks cover [0, n_rays) exactly once and differ by at most `multiple` rays."""
    if rank is None:
        rank, world_size = world()
    n_units = (n_rays + multiple - 1) // multiple
    base, rem = divmod(n_units, world_size)
    u0 = rank * base + min(rank, rem)
    u1 = u0 + base + (1 if rank < rem else 0)
    return min(u0 * multiple, n_rays), min(u1 * multiple, n_rays)


def gather_rays(local, n_rays, multiple=128, group=None):
    """All-gather per-rank ray blocks [n_local, C] into the full [n_rays, C] tensor on every rank."""
    rank, ws = world()
    if ws == 1:
        return local
    bounds = [shard_rays(n_rays, r, ws, multiple) for r in range(ws)]
    max_len = max(b[1] - b[0] for b in bounds)
    C = local.shape[1:]
    pad = torch.zeros((max_len,) + tuple(C), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:b[1] - b[0]] for o, b in zip(outs, bounds)], 0)


def gather_rays_into(local, out, n_rays, multiple=128, group=None):
    """gather_rays without temporaries when every rank holds the same number of rays (n_rays/multiple divisible by
    the world size: 160 000 rays on 2 GPUs, 640 000 on 2/4/8 ...): ONE all_gather_into_tensor straight into the
    caller's [n_rays, C] frame buffer.  Falls back to gather_rays (and copies) for ragged splits.  Returns `out`."""
    rank, ws = world()
    if ws == 1:
        out.copy_(local)
        return out
    lens = {b[1] - b[0] for b in (shard_rays(n_rays, r, ws, multiple) for r in range(ws))}
    if len(lens) == 1 and out.is_contiguous() and local.is_contiguous():
        dist.all_gather_into_tensor(out, local, group=group)
    else:
        out.copy_(gather_rays(local, n_rays, multiple, group))
    return out


class PeerFrame:
    """One frame buffer [n_rays, channels] fp32 per GPU, allocated in symmetric memory and mapped into every rank of
    the box (torch.distributed._symmetric_memory: CUDA VMM handles exchanged once at construction), for the fused tile
    gather of ray-sharded frames: every rank's MLP kernel stores its rows [row0, row1) into ALL buffers by peer-to-peer
    stores over NVLink (NeRF_v3_2.forward_points_gather), `publish()` is the cross-GPU barrier after which `buf` holds
    the whole frame on every rank.  Replaces kernel + NCCL all_gather_into_tensor (gather_rays_into) by ONE kernel +
    a ~7 us barrier.

    Reuse contract: a peer that has left THIS publish() may launch its next frame at once, and that kernel stores into
    this rank's buffer — possibly while this rank is still reading the frame just published.  A single PeerFrame
    rendered back to back is therefore a cross-GPU data race.  Either alternate TWO PeerFrames (frame k+1 goes to the
    other buffer) and consume frame k, stream-ordered, before joining publish() k+1 — a peer can only reach the
    kernel of frame k+2 (same buffer as k) after that barrier — or, with one PeerFrame, call `release()` (a second
    barrier) once the frame has been consumed and before any rank renders into it again."""

    def __init__(self, n_rays, channels=3, multiple=128, group=None):
        import ctypes
        import torch.distributed._symmetric_memory as symm_mem
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerFrame needs an initialised process group (one process per GPU)")
        self.rank, self.world_size = world()
        if self.world_size > 8:
            raise ValueError("peer-to-peer frames cover the GPUs of one box (<= 8 ranks)")
        grp = group if group is not None else dist.group.WORLD
        dev = torch.device("cuda", torch.cuda.current_device())
        self.n_rays = int(n_rays)
        self.buf = symm_mem.empty((self.n_rays, int(channels)), dtype=torch.float32, device=dev)
        self.hdl = symm_mem.rendezvous(self.buf, group=grp.group_name)
        self.ptrs = (ctypes.c_void_p * self.world_size)(*[int(p) for p in self.hdl.buffer_ptrs])
        self.row0, self.row1 = shard_rays(self.n_rays, self.rank, self.world_size, multiple)

    def publish(self):
        """Cross-GPU barrier on the current stream: every rank's stores of this frame have landed everywhere."""
        self.hdl.barrier()
        return self.buf

    def release(self):
        """Second barrier for single-buffer use: every rank has finished reading the frame (reads enqueued on the
        current stream before this call), so peers may overwrite it."""
        self.hdl.barrier()


def gather_frames(local_frames, n_poses, group=None):
    """local_frames: list of [H*W, C] tensors for shard_poses(n_poses).  Returns the list of all
    n_poses frames in pose order on every rank (one all_gather per round of G poses)."""
    rank, ws = world()
    if ws == 1:
        return list(local_frames)
    frames = [None] * n_poses
    rounds = (n_poses + ws - 1) // ws
    proto = local_frames[0]
    for r in range(rounds):
        mine = local_frames[r] if r < len(local_frames) else torch.zeros_like(proto)
        outs = [torch.empty_like(proto) for _ in range(ws)]
        dist.all_gather(outs, mine.contiguous(), group=group)
        for k in range(ws):
            idx = r * ws + k
            if idx < n_poses:
                frames[idx] = outs[k]
    return frames
