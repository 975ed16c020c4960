"""create_data `rand` sharded over the GPUs of one box (BASELINE configs[2] shape: 400x400, 64+128 samples, 4096-row
.npy shards), one process per GPU under torchrun; groups of `i_save` poses go round-robin to the ranks, every rank
writes its own disjoint data_{k}.npy range, no data-path collective.  Prints aggregate poses/s (wall clock between
two barriers, max over ranks) and checks the files.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scratch/bench_create_data_multi.py [groups_per_rank] [i_save]"""
import json, os, shutil, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '.')
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import efficient_nerf_b200 as E
from efficient_nerf_b200 import create_data as CD
gpr = int(sys.argv[1]) if len(sys.argv) > 1 else 2
i_save = int(sys.argv[2]) if len(sys.argv) > 2 else 25
n_pose = gpr * i_save * world
out = "/tmp/cd_multi"
if rank == 0:
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
coarse, fine = E.synthetic.seeded_nerf_pair(0, "fp16")
coarse.packed_handle(), fine.packed_handle()
def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
with torch.no_grad():
    # warm-up: one small group elsewhere (kernels, allocator, writer threads)
    CD.create_data_rand(coarse, fine, out + f"_warm{rank}", 4, 400, 400, E.synthetic.LEGO["focal"], perturb=1.,
                        fast_rng=True, writer_threads=4, i_save=4)
    shutil.rmtree(out + f"_warm{rank}", ignore_errors=True)
    barrier()
    t0 = time.time()
    w = CD.create_data_rand(coarse, fine, out, n_pose, 400, 400, E.synthetic.LEGO["focal"], perturb=1., fast_rng=True,
                            writer_threads=4, i_save=i_save, rank=rank, world_size=world)
    barrier()
    dt = time.time() - t0
t = torch.tensor([dt], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    files = sorted(f for f in os.listdir(out) if f.endswith(".npy"))
    F = CD.files_per_group(400, 400, i_save, 4096)
    a = np.load(os.path.join(out, files[0]))
    ok = (len(files) == F * gpr * world and a.shape == (4096, 9) and a.dtype == np.float32 and np.isfinite(a).all())
    nbytes = sum(os.path.getsize(os.path.join(out, f)) for f in files)
    print(json.dumps({"workload": "create_data rand 400x400 64+128, 4096-row .npy shards (BASELINE configs[2] shape)",
                      "n_gpus": world, "poses": n_pose, "i_save": i_save, "seconds": float(t[0]),
                      "poses_per_s": n_pose / float(t[0]), "Mrays_per_s": n_pose * 0.16 / float(t[0]),
                      "files": len(files), "GB_written": nbytes / 1e9, "files_ok": bool(ok),
                      "rank0_files": len(w)}))
    shutil.rmtree(out, ignore_errors=True)
if world > 1:
    dist.destroy_process_group()
