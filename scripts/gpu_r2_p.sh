#!/bin/bash
# Round 2, call P: how many of a thread's pairs keep their velocities in shared memory (shared memory vs L1 for the gathers).
O=gpurun_out; mkdir -p $O
line() {
python - "$1" "$2" <<'PY'
import json, sys
tag, path = sys.argv[1], sys.argv[2]
try:
    d=json.loads(open(path).read().strip().splitlines()[-1])
    r=d["roofline"]
    print(tag, "%.3e" % d["value"], "us/step %.2f" % (d["ms_per_step"]*1e3), r.get("phases_us") or r.get("kernels_ms"), "steady", d["steady_state"] and ("%.3e" % d["steady_state"]["value"], round(d["steady_state"]["us_per_step"],2), d["steady_state"]["rebuilds"]))
except Exception as e:
    print(tag, "FAILED", e); print(open(path.replace(".json",".err")).read()[-1500:])
PY
}
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "loop_drivers or golden" 2>&1 | tail -2
for ps in -1 6 5 4 2 0; do
  if [ $ps -ge 0 ]; then export MOLDYN_B200_LOOP_PSMEM=$ps; else unset MOLDYN_B200_LOOP_PSMEM; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --e2e-steps 0 --cpu-rows -1 > $O/p_c3_$ps.json 2> $O/p_c3_$ps.err; line "c3 driver flags PS=$ps" $O/p_c3_$ps.json
done
for ps in 9 4 0; do
  export MOLDYN_B200_LOOP_PSMEM=$ps
  timeout 300 python bench.py --workload big --steps 600 --warmup 2500 --e2e-steps 0 --cpu-rows -1 --steady-steps 0 > $O/p_big_$ps.json 2> $O/p_big_$ps.err; line "big (3100 steps in) PS=$ps" $O/p_big_$ps.json
done
