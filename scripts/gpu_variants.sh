#!/bin/bash
# quick A/B of build variants: $1.. = values of MD_FORCE_MINB
mkdir -p gpurun_out
for mb in "$@"; do
  echo "=== MD_FORCE_MINB=$mb"
  MD_NVCC_EXTRA="-DMD_FORCE_MINB=$mb" python -m moldyn_b200.build --force 2>&1 | grep -v Warning | tail -1
  for w in c3 c5 big; do
    steps=3000; [ $w = big ] && steps=500; [ $w = c5 ] && steps=1000
    timeout 600 python bench.py --workload $w --steps $steps --warmup 200 --e2e-steps 0 --cpu-rows -1 > gpurun_out/v_${mb}_$w.json 2> gpurun_out/v_${mb}_$w.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/v_${mb}_$w.json"))
    print("$w", {k:d[k] for k in ("value","ms_per_step","rebuilds_in_timed_region")}, d["roofline"]["kernels_ms"])
except Exception as e: print("ERR", e, open("gpurun_out/v_${mb}_$w.err").read()[-600:])
PY
  done
done
python -m moldyn_b200.build --force > /dev/null 2>&1
echo "== pytest (default build)"
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -5
