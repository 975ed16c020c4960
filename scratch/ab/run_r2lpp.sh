timeout 300 python -m pytest tests/test_gpu_mlp.py -m gpu -x -q -k "r2l or R2L or resmlp" 2>&1 | tail -5
for v in 0 1 0 1; do
  R2L_PP=$v timeout 300 python bench.py --steps 100 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('R2L_PP=$v', round(d['value'],2), 'Mrays/s', round(d['ms_per_step'],3), 'ms frac', round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"
done
