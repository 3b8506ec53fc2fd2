#!/bin/bash
# tests + bench (c3, c2, c5, big) + ncu launch list + one full capture of the force kernel
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -15
echo "== bench c3"
timeout 900 python bench.py --workload c3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 2500 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
for w in c2 c5 big c1; do
  echo "== bench $w"
  timeout 900 python bench.py --workload $w --e2e-steps 3 --cpu-rows -1 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$w.json"))
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches","rebuilds_in_timed_region")}, d["roofline"]["kernels_ms"], d["roofline"]["frac"], d["roofline"]["step"]["frac"], d["config"]["skin"], d["clocks"])
except Exception as e: print("ERR", e, open("gpurun_out/bench_$w.err").read()[-800:])
PY
done
echo "== ncu launch list (c3)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload c3 --steps 200 --warmup 20 --e2e-steps 0 --cpu-rows -1 > gpurun_out/ncu_bench_c3.log 2>&1; tail -2 gpurun_out/ncu_bench_c3.log | cut -c1-300
echo "== ncu full capture k_force + k_kick_drift (c3)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_force|k_kick_drift' -s 40 -c 4 -o gpurun_out/prof_c3 -f python bench.py --workload c3 --steps 100 --warmup 20 --e2e-steps 0 --cpu-rows -1 > gpurun_out/ncu_full_c3.log 2>&1; tail -2 gpurun_out/ncu_full_c3.log | cut -c1-300
echo "== ncu full capture k_force (c5)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_force' -s 20 -c 2 -o gpurun_out/prof_c5 -f python bench.py --workload c5 --steps 60 --warmup 10 --e2e-steps 0 --cpu-rows -1 > gpurun_out/ncu_full_c5.log 2>&1; tail -2 gpurun_out/ncu_full_c5.log | cut -c1-300
ls -la gpurun_out
