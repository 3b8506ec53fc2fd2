#!/bin/bash
# Round artefacts: GPU tests, the default bench line (+ reference arm), the ncu launch list of the bench command and one
# ncu --set full capture of each hot kernel.  Outputs land in gpurun_out/ (copied into profiles/ by hand).
mkdir -p gpurun_out
T=${1:-r01}
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; tail -c 600 gpurun_out/${T}_bench_default.json; echo
timeout 900 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; tail -c 400 gpurun_out/${T}_bench_reference.json; echo
timeout 900 python bench.py --workload c5 --steps 2000 --warmup 500 --e2e-steps 3 > gpurun_out/${T}_bench_c5.json 2> gpurun_out/${T}_bench_c5.err; tail -c 300 gpurun_out/${T}_bench_c5.json; echo
for w in c1 c2 c4 big; do
  timeout 900 python bench.py --workload $w --steps 1000 --warmup 500 --e2e-steps 2 --cpu-rows -1 > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_bench_$w.json")); print("$w", d["value"], d["ms_per_step"], d["e2e"]["value"], d["state_check"])
except Exception as e: print("ERR $w", e, open("gpurun_out/${T}_bench_$w.err").read()[-500:])
PY
done
export MOLDYN_B200_LOOP=host
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_c3.csv python bench.py --steps 60 --warmup 3 --e2e-steps 1 --cpu-rows -1 > gpurun_out/${T}_launches_c3.log 2>&1; tail -2 gpurun_out/${T}_launches_c3.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_force|k_kick_drift' -s 12200 -c 2 -o gpurun_out/${T}_prof_c3 -f python bench.py --workload c3 --steps 300 --warmup 6000 --e2e-steps 0 --cpu-rows -1 > gpurun_out/${T}_ncu_c3.log 2>&1; tail -1 gpurun_out/${T}_ncu_c3.log | cut -c1-150
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_force' -s 320 -c 1 -o gpurun_out/${T}_prof_c5 -f python bench.py --workload c5 --steps 60 --warmup 300 --e2e-steps 0 --cpu-rows -1 > gpurun_out/${T}_ncu_c5.log 2>&1; tail -1 gpurun_out/${T}_ncu_c5.log | cut -c1-150
