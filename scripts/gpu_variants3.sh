#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open(f))
    print(f.split('/')[-1], {k:round(d[k],5) if isinstance(d[k],float) else d[k] for k in ("value","ms_per_step","rebuilds_in_timed_region")}, {k:(round(v,5) if v else v) for k,v in d["roofline"]["kernels_ms"].items()})
except Exception as e: print("ERR", f, e, open(f.replace('.json','.err')).read()[-500:])
PY
}
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -3
for mb in 4 5; do
  echo "=== MINB=$mb"
  MD_NVCC_EXTRA="-DMD_FORCE_MINB=$mb -DMD_FORCE_MINB_DILUTE=$mb" python -m moldyn_b200.build --force > /dev/null 2>&1
  timeout 600 python bench.py --workload c3 --steps 3000 --warmup 12000 --e2e-steps 0 --cpu-rows -1 > gpurun_out/m${mb}_c3_late.json 2> gpurun_out/m${mb}_c3_late.err; show gpurun_out/m${mb}_c3_late.json
  timeout 600 python bench.py --workload c3 --steps 3000 --warmup 300 --e2e-steps 0 --cpu-rows -1 > gpurun_out/m${mb}_c3_early.json 2> gpurun_out/m${mb}_c3_early.err; show gpurun_out/m${mb}_c3_early.json
  timeout 600 python bench.py --workload big --steps 300 --warmup 100 --e2e-steps 0 --cpu-rows -1 > gpurun_out/m${mb}_big.json 2> gpurun_out/m${mb}_big.err; show gpurun_out/m${mb}_big.json
  timeout 600 python bench.py --workload c5 --steps 1000 --warmup 300 --cell-subdiv 2 --e2e-steps 0 --cpu-rows -1 > gpurun_out/m${mb}_c5.json 2> gpurun_out/m${mb}_c5.err; show gpurun_out/m${mb}_c5.json
done
echo "=== MINB=4 without fast rcp (c5)"
MD_NVCC_EXTRA="-DMD_FAST_RCP=0" python -m moldyn_b200.build --force > /dev/null 2>&1
timeout 600 python bench.py --workload c5 --steps 1000 --warmup 300 --cell-subdiv 2 --e2e-steps 0 --cpu-rows -1 > gpurun_out/m4n_c5.json 2> gpurun_out/m4n_c5.err; show gpurun_out/m4n_c5.json
python -m moldyn_b200.build --force > /dev/null 2>&1
