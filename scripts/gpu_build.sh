#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open(f))
    print(f.split('/')[-1], {k:round(d[k],5) if isinstance(d[k],float) else d[k] for k in ("value","ms_per_step","rebuilds_in_timed_region")}, {k:(round(v,5) if v else v) for k,v in d["roofline"]["kernels_ms"].items()}, d["state_check"]["temperature"])
except Exception as e: print("ERR", f, e, open(f.replace('.json','.err')).read()[-1500:])
PY
}
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -3
timeout 600 python bench.py --workload c5 --steps 1000 --warmup 300 --e2e-steps 0 --cpu-rows -1 > gpurun_out/b_c5.json 2> gpurun_out/b_c5.err; show gpurun_out/b_c5.json
timeout 600 python bench.py --workload c3 --steps 3000 --warmup 6000 --e2e-steps 0 --cpu-rows -1 > gpurun_out/b_c3.json 2> gpurun_out/b_c3.err; show gpurun_out/b_c3.json
