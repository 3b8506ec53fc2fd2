#!/bin/bash
# Round 2, call S: tile kernels v5 (lists padded with the nobody slot, one 8-byte index load per lane and trip).
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tile or dense or config_size or full_size or neighbour or update_force" 2>&1 | tail -3
line() {
python - "$1" "$2" <<'PY'
import json, sys
tag, path = sys.argv[1], sys.argv[2]
try:
    d=json.loads(open(path).read().strip().splitlines()[-1])
    r=d["roofline"]
    print(tag, "%.3e" % d["value"], "us/step %.2f" % (d["ms_per_step"]*1e3), r.get("phases_us") or r.get("kernels_ms"), "frac", r.get("frac"), "rebuild", r["rebuild"])
except Exception as e:
    print(tag, "FAILED", e); print(open(path.replace(".json",".err")).read()[-1500:])
PY
}
timeout 300 python bench.py --workload c5 --steps 1000 --warmup 300 --e2e-steps 0 --cpu-rows -1 --steady-steps 0 > $O/s_c5.json 2> $O/s_c5.err; line "c5 tile v5" $O/s_c5.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_force_tile' -s 640 -c 1 -o $O/r02_prof_c5_force_tile_v5 -f \
  python bench.py --workload c5 --steps 60 --warmup 300 --e2e-steps 0 --cpu-rows -1 --steady-steps 0 --no-time-rebuild > $O/s_ncu_force.log 2>&1
python scripts/ncu_summary.py $O/r02_prof_c5_force_tile_v5.ncu-rep > $O/r02_prof_c5_force_tile_v5.txt 2>&1; head -31 $O/r02_prof_c5_force_tile_v5.txt
