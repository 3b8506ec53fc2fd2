#!/bin/bash
# Round 2, call A: the persistent step loop's first contact with a B200 — parity file, then quick bench lines (loop vs chunk).
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee $O/a_smi.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/a_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/a_smoke.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not config_size and not c4_size and not full_size and not long_run" > $O/a_pytest_fast.log 2>&1; echo "pytest fast rc=$?"; tail -15 $O/a_pytest_fast.log
for w in c1 c2 c3; do
  for loop in auto chunk; do
    timeout 300 python bench.py --workload $w --loop $loop --steps 2000 --warmup 500 --e2e-steps 0 --cpu-rows -1 > $O/a_bench_${w}_${loop}.json 2> $O/a_bench_${w}_${loop}.err
    python - <<PY
import json
try:
    d=json.loads(open("$O/a_bench_${w}_${loop}.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$w $loop", "%.3e" % d["value"], "us/step %.2f" % (d["ms_per_step"]*1e3), "rebuilds", d["rebuilds_in_timed_region"], r.get("phases_us") or r.get("kernels_ms"), "rebuild", r["rebuild"], "steady", d["steady_state"] and ("%.3e" % d["steady_state"]["value"], d["steady_state"]["rebuilds"], d["steady_state"]["nbr_mean"]), d["state_check"])
except Exception as e:
    print("$w $loop FAILED", e); print(open("$O/a_bench_${w}_${loop}.err").read()[-1500:])
PY
  done
done
timeout 300 python bench.py --workload c3 --steps 20 --warmup 5 --e2e-steps 3 --cpu-rows -1 > $O/a_bench_c3_driverlike.json 2>&1; tail -c 600 $O/a_bench_c3_driverlike.json; echo
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_cli.py -x -q -m gpu -k "config_size or c4_size or full_size or long_run or cli" > $O/a_pytest_slow.log 2>&1; echo "pytest slow rc=$?"; tail -15 $O/a_pytest_slow.log
