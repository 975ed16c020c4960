set -u
O=gpurun_out/r2c; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > $O/tests_gpu_multi_n2.log 2>&1; echo "multi tests rc=$?"; tail -5 $O/tests_gpu_multi_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"
tail -5 $O/bench_n2.err
python - <<PY
import json
d=json.loads(open("$O/bench_n2.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value","ms_per_step","gpu_launches","n_gpus")})
print(json.dumps(d.get("extras"), indent=1)[:3000])
PY
