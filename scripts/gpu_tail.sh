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
MD_NVCC_EXTRA="-DMD_TIMING_PROBES" python -m moldyn_b200.build --force > /dev/null 2>&1; timeout 400 python scripts/probe.py
python -m moldyn_b200.build --force > /dev/null 2>&1
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -3
for w in c3 c2 c1; do
timeout 600 python bench.py --workload $w --steps 3000 --warmup 6000 --e2e-steps 0 --cpu-rows -1 > gpurun_out/t_$w.json 2> gpurun_out/t_$w.err; show gpurun_out/t_$w.json
done
