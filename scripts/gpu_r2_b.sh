#!/bin/bash
# Round 2, call B: loop v2 (cheaper barriers, grouped gathers) — fast parity, phase traces, loop vs chunk bench lines.
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cli.py -x -q -m gpu -k "not config_size and not c4_size and not full_size and not long_run and not 100000" > $O/b_pytest_fast.log 2>&1; echo "pytest fast rc=$?"; tail -6 $O/b_pytest_fast.log
for w in c1 c2 c3; do
  MOLDYN_B200_LIBRARY=moldyn_b200/lib/libmoldyn_b200_trace.so timeout 200 python scripts/loop_trace.py $w 500 2>&1 | tail -13
  MOLDYN_B200_LIBRARY=moldyn_b200/lib/libmoldyn_b200_trace.so timeout 200 python scripts/loop_trace.py $w 8000 2>&1 | tail -13
done
for w in c1 c2 c3; do
  for loop in auto chunk; do
    timeout 300 python bench.py --workload $w --loop $loop --steps 2000 --warmup 500 --e2e-steps 0 --cpu-rows -1 > $O/b_bench_${w}_${loop}.json 2> $O/b_bench_${w}_${loop}.err
    python - <<PY
import json
try:
    d=json.loads(open("$O/b_bench_${w}_${loop}.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$w $loop", "%.3e" % d["value"], "us/step %.2f" % (d["ms_per_step"]*1e3), r.get("phases_us") or r.get("kernels_ms"), "rebuild", r["rebuild"], "steady", d["steady_state"] and ("%.3e" % d["steady_state"]["value"], d["steady_state"]["us_per_step"], d["steady_state"]["rebuilds"], d["steady_state"]["nbr_mean"]))
except Exception as e:
    print("$w $loop FAILED", e); print(open("$O/b_bench_${w}_${loop}.err").read()[-1500:])
PY
  done
done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "100000" > $O/b_pytest_100k.log 2>&1; echo "pytest 100k rc=$?"; tail -6 $O/b_pytest_100k.log
