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
timeout 600 python bench.py --workload c5 --steps 1000 --warmup 300 --e2e-steps 0 --cpu-rows -1 > gpurun_out/c5b_$1.json 2> gpurun_out/c5b_$1.err; show gpurun_out/c5b_$1.json
export MOLDYN_B200_LOOP=host
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_force' -s 320 -c 1 -o gpurun_out/prof_c5_$1 -f python bench.py --workload c5 --steps 60 --warmup 300 --e2e-steps 0 --cpu-rows -1 > gpurun_out/ncu_c5_$1.log 2>&1; tail -1 gpurun_out/ncu_c5_$1.log | cut -c1-150
