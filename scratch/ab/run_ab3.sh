# A/B of the committed tree (scratch/ab/head_tree, its own library) against the working tree on ONE box, plus the
# working tree with R2L_NERF_FUSED=0:  bash scratch/ab/run_ab3.sh [steps]
K=${1:-20}
ROOT=$PWD
line() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), 'mlp', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],4), 'launches', d['gpu_launches'], d['clocks']['sm_mhz'])"; }
for rep in 1 2; do
  (cd scratch/ab/head_tree && python bench.py --workload nerf --steps $K --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | line head)
  (cd $ROOT && python bench.py --workload nerf --steps $K --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | line fused)
  (cd $ROOT && R2L_NERF_FUSED=0 python bench.py --workload nerf --steps $K --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | line unfused)
done
