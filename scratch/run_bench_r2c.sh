set -u
O=gpurun_out/r2c; mkdir -p $O
timeout 600 python bench.py --steps 100 > $O/bench_r2l.json 2> $O/bench_r2l.err; echo "bench r2l rc=$?"
tail -3 $O/bench_r2l.err
python - <<PY
import json
d=json.load(open("$O/bench_r2l.json"))
print({k: d[k] for k in ("value","ms_per_step","gpu_launches","clocks")})
print(d["roofline"]); print(d["e2e"]); print(d.get("cpu_baseline"))
ex=d["extras"]
for k,v in ex.items():
    print(k, json.dumps(v)[:900])
PY
