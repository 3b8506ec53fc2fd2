#!/bin/bash
# A/B of MD_FORCE_MINB_DILUTE on late-time c3 and big; then skin / cell sweeps with the default build
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open(f))
    print(f.split('/')[-1], {k:round(d[k],5) if isinstance(d[k],float) else d[k] for k in ("value","ms_per_step","rebuilds_in_timed_region")}, {k:(round(v,5) if v else v) for k,v in d["roofline"]["kernels_ms"].items()}, "skin",round(d["config"]["skin"],3),"cells",d["config"]["cells"][0])
except Exception as e: print("ERR", f, e, open(f.replace('.json','.err')).read()[-500:])
PY
}
for mb in "$@"; do
  echo "=== MD_FORCE_MINB_DILUTE=$mb"
  MD_NVCC_EXTRA="-DMD_FORCE_MINB_DILUTE=$mb" python -m moldyn_b200.build --force 2>&1 | grep -v Warning | tail -1
  timeout 600 python bench.py --workload c3 --steps 3000 --warmup 12000 --e2e-steps 0 --cpu-rows -1 > gpurun_out/d_${mb}_c3.json 2> gpurun_out/d_${mb}_c3.err; show gpurun_out/d_${mb}_c3.json
  timeout 600 python bench.py --workload big --steps 300 --warmup 100 --e2e-steps 0 --cpu-rows -1 > gpurun_out/d_${mb}_big.json 2> gpurun_out/d_${mb}_big.err; show gpurun_out/d_${mb}_big.json
done
python -m moldyn_b200.build --force > /dev/null 2>&1
echo "=== default build: pytest"
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -4
echo "=== skin sweep c3 late"
for skin in 0.3 0.45 0.6 0.8545; do
  timeout 600 python bench.py --workload c3 --steps 3000 --warmup 12000 --skin $skin --e2e-steps 0 --cpu-rows -1 > gpurun_out/s_${skin}_c3.json 2> gpurun_out/s_${skin}_c3.err; show gpurun_out/s_${skin}_c3.json
done
echo "=== cell_atoms sweep c3 late (default skin)"
for ca in 0.5 2 3; do
  timeout 600 python bench.py --workload c3 --steps 3000 --warmup 12000 --cell-atoms $ca --e2e-steps 0 --cpu-rows -1 > gpurun_out/ca_${ca}_c3.json 2> gpurun_out/ca_${ca}_c3.err; show gpurun_out/ca_${ca}_c3.json
done
echo "=== c5: skin and subdiv"
for skin in 0.08 0.12 0.16; do for sub in 1 2; do
  timeout 600 python bench.py --workload c5 --steps 1000 --warmup 300 --skin $skin --cell-subdiv $sub --e2e-steps 0 --cpu-rows -1 > gpurun_out/c5_${skin}_${sub}.json 2> gpurun_out/c5_${skin}_${sub}.err; show gpurun_out/c5_${skin}_${sub}.json
done; done
