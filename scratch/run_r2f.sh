# round-2 "f" evidence: full GPU test suite with the fused-compositing tests, NeRF bench both routes on one box,
# ncu traffic of the NeRF MLP kernels in fused mode
O=gpurun_out/r2f; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/tests_gpu.log 2>&1; echo "tests rc=$?"; tail -n 3 $O/tests_gpu.log
timeout 300 python bench.py --workload nerf --steps 20 --warmup 3 --no-extras --no-cpu-baseline > $O/bench_nerf_twostep.json 2>$O/b1.err; echo "rc=$?"
R2L_NERF_FUSED=1 timeout 300 python bench.py --workload nerf --steps 20 --warmup 3 --no-extras --no-cpu-baseline > $O/bench_nerf_fused.json 2>$O/b2.err; echo "rc=$?"
R2L_NERF_FUSED=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:nerf_mlp -s 6 -c 2 -o $O/prof_nerf_fused -f python bench.py --workload nerf --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu_nerf_fused.log 2>&1; echo "ncu rc=$?"
R2L_NERF_FUSED=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_nerf_fused.csv python bench.py --workload nerf --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu_nerf_list.log 2>&1; echo "ncu list rc=$?"
python - <<PY
import json
for f in ("bench_nerf_twostep","bench_nerf_fused"):
    d=json.load(open("$O/"+f+".json")); print(f, d["ms_per_step"], d["roofline"]["kernel_ms"], d["gpu_launches"], d["clocks"]["sm_mhz"])
PY
