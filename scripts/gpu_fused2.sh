#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open(f))
    print(f.split('/')[-1], {k:round(d[k],5) if isinstance(d[k],float) else d[k] for k in ("value","ms_per_step","rebuilds_in_timed_region")}, {k:(round(v,5) if v else v) for k,v in d["roofline"]["kernels_ms"].items()})
except Exception as e: print("ERR", f, e, open(f.replace('.json','.err')).read()[-1500:])
PY
}
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -3
for mb in 4 5; do
  echo "=== STEP_MINB=$mb"
  MD_NVCC_EXTRA="-DMD_STEP_MINB=$mb" python -m moldyn_b200.build --force > /dev/null 2>&1
  for w in c3 big; do
    timeout 600 python bench.py --workload $w --steps 2000 --warmup 6000 --e2e-steps 0 --cpu-rows -1 > gpurun_out/g${mb}_$w.json 2> gpurun_out/g${mb}_$w.err; show gpurun_out/g${mb}_$w.json
  done
done
python -m moldyn_b200.build --force > /dev/null 2>&1
for w in c3 big; do
timeout 600 python bench.py --workload $w --steps 2000 --warmup 6000 --e2e-steps 0 --cpu-rows -1 --split-step > gpurun_out/gs_$w.json 2> gpurun_out/gs_$w.err; show gpurun_out/gs_$w.json
done
