"""Throughput of create_data `rand` at BASELINE config 3's shape (400x400, 64+128 samples, 4096-row shards)."""
import sys, time, shutil, os, numpy as np, torch
sys.path.insert(0, '.')
import efficient_nerf_b200 as E
from efficient_nerf_b200 import create_data as CD
from oracle import ref_torch as O
sdc, sdf = O.nerf_state_dicts(0)
nets = []
for sd in (sdc, sdf):
    n = E.NeRF(8, 256, 63, 27, 5, [4], True, precision="fp16"); n.load_state_dict(sd); nets.append(n.cuda().eval())
n_pose = int(sys.argv[1]) if len(sys.argv) > 1 else 100
for mode in ("fast_rng", "reference_rng"):
    out = f"/tmp/cd_{mode}"
    shutil.rmtree(out, ignore_errors=True)
    np.random.seed(0); torch.manual_seed(0)
    t0 = time.time()
    with torch.no_grad():
        w = CD.create_data_rand(nets[0], nets[1], out, n_pose, 400, 400, O.LEGO["focal"], perturb=1.,
                                fast_rng=(mode == "fast_rng"), writer_threads=8)
    torch.cuda.synchronize()
    dt = time.time() - t0
    nbytes = sum(os.path.getsize(os.path.join(out, f)) for f in os.listdir(out))
    print(f"{mode}: {n_pose} poses in {dt:.2f} s = {n_pose / dt:.2f} poses/s = {n_pose * 0.16 / dt:.2f} Mrays/s; "
          f"{len(w)} files, {nbytes / 1e6:.0f} MB")
    shutil.rmtree(out, ignore_errors=True)
