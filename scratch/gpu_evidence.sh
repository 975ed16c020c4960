#!/bin/bash
# Full evidence run: tests, benches, launch lists, ncu captures. Outputs under gpurun_out/$1/
set -u
O=gpurun_out/${1:-r2}; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/smi.log 2>&1; nproc >> $O/smi.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/tests_gpu.log 2>&1; echo "tests rc=$?"
timeout 600 python bench.py > $O/bench_r2l.json 2> $O/bench_r2l.err; echo "bench r2l rc=$?"
timeout 600 python bench.py --workload nerf --steps 20 --warmup 3 --no-extras > $O/bench_nerf.json 2> $O/bench_nerf.err; echo "bench nerf rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_r2l.json 2> $O/bench_ref_r2l.err; echo "ref rc=$?"
if [ "${2:-full}" = "full" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r2l.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu_r2l.log 2>&1; echo "ncu list r2l rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_nerf.csv python bench.py --workload nerf --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu_nerf.log 2>&1; echo "ncu list nerf rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:r2l_mlp_kernel -s 3 -c 1 -o $O/prof_r2l -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu_r2l_full.log 2>&1; echo "ncu full r2l rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nerf_mlp -s 6 -c 2 -o $O/prof_nerf -f python bench.py --workload nerf --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu_nerf_full.log 2>&1; echo "ncu full nerf rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:r2l_mlp_kernel -s 3 -c 1 -o $O/prof_r2l_bf16 -f python bench.py --precision bf16 --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu_r2l_bf16_full.log 2>&1; echo "ncu full r2l bf16 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"raw2outputs|sample|merge|far_fixup" -c 8 -o $O/prof_hbm -f python bench.py --workload nerf --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu_hbm_full.log 2>&1; echo "ncu full hbm rc=$?"
fi
tail -n 3 $O/tests_gpu.log; cat $O/bench_r2l.json | cut -c1-300; python - <<PY
import json
for f in ("bench_r2l","bench_nerf","bench_ref_r2l"):
    try:
        d=json.load(open("$O/"+f+".json")); print(f, d["value"], d["unit"], d.get("ms_per_step"), d.get("roofline",{}).get("achieved"), d.get("roofline",{}).get("frac"), d.get("clocks"), d.get("e2e",{}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
