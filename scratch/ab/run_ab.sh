# A/B of two builds of the library on the same box: bash scratch/ab/run_ab.sh <workload> [steps]
W=${1:-nerf}; K=${2:-20}
for v in old new old new; do
  cp scratch/ab/lib_$v.so efficient-nerf_b200/lib/libr2l_b200.so
  python bench.py --workload $W --steps $K --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],3), 'mlp', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],4), 'launches', d['gpu_launches'], d['clocks']['sm_mhz'])"
done
