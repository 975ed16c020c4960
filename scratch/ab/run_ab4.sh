# head tree vs working tree library variants (lib_old = -DR2L_NO_COMP build, lib_new = normal build), all UNFUSED
ROOT=$PWD
line() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), 'mlp', round(d['roofline']['kernel_ms'],3), 'launches', d['gpu_launches'], d['clocks']['sm_mhz'])"; }
export R2L_NERF_FUSED=0
for rep in 1 2; do
  (cd scratch/ab/head_tree && python bench.py --workload nerf --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | line head)
  cp scratch/ab/lib_old.so efficient-nerf_b200/lib/libr2l_b200.so
  python bench.py --workload nerf --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | line nocomp_unfused
  cp scratch/ab/lib_new.so efficient-nerf_b200/lib/libr2l_b200.so
  python bench.py --workload nerf --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | line comp_unfused
  R2L_NERF_FUSED=1 python bench.py --workload nerf --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | line comp_fused
done
