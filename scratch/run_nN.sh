# bash scratch/run_nN.sh <N> <tag>: the driver's multi-GPU bench command on N GPUs of this box
N=$1; O=gpurun_out/${2:-r2e}; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 40 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench n$N rc=$?"
tail -3 $O/bench_n$N.err
python - <<PY
import json
d=json.loads(open("$O/bench_n$N.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value","ms_per_step","gpu_launches","n_gpus")}, d["e2e"]["value"])
rs=d.get("extras",{}).get("ray_sharded",{})
for k in ("r2l","nerf"):
    if k in rs: print(k, rs[k]["ms_per_frame_400x400"], rs[k]["e2e"]["ms_per_frame"], rs[k]["bit_identical_to_single_gpu"])
if "error" in rs: print("ERROR", rs["error"])
PY
