#!/bin/bash
# Round 2, call T: tile kernels v6 (chunks stored in lane order) + masked block sums in the loop's tail.
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tile or dense or config_size or neighbour or update_force or loop_drivers or golden or nve or nose or determinism" 2>&1 | tail -3
line() {
python - "$1" "$2" <<'PY'
import json, sys
tag, path = sys.argv[1], sys.argv[2]
try:
    d=json.loads(open(path).read().strip().splitlines()[-1])
    r=d["roofline"]
    print(tag, "%.3e" % d["value"], "us/step %.2f" % (d["ms_per_step"]*1e3), r.get("phases_us") or r.get("kernels_ms"), "frac", r.get("frac"), "rebuild", r["rebuild"]["ms_each"], "steady", d["steady_state"] and ("%.3e" % d["steady_state"]["value"], round(d["steady_state"]["us_per_step"],2)))
except Exception as e:
    print(tag, "FAILED", e); print(open(path.replace(".json",".err")).read()[-1500:])
PY
}
timeout 300 python bench.py --workload c5 --steps 1000 --warmup 300 --e2e-steps 0 --cpu-rows -1 --steady-steps 0 > $O/t_c5.json 2> $O/t_c5.err; line "c5 tile v6" $O/t_c5.json
timeout 600 python bench.py --steps 20 --warmup 5 --e2e-steps 0 --cpu-rows -1 > $O/t_c3.json 2> $O/t_c3.err; line "c3 driver flags" $O/t_c3.json
timeout 300 python bench.py --workload c2 --steps 2000 --warmup 500 --e2e-steps 0 --cpu-rows -1 > $O/t_c2.json 2> $O/t_c2.err; line "c2" $O/t_c2.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_force_tile' -s 640 -c 1 -o $O/r02_prof_c5_force_tile_v6 -f \
  python bench.py --workload c5 --steps 60 --warmup 300 --e2e-steps 0 --cpu-rows -1 --steady-steps 0 --no-time-rebuild > $O/t_ncu_force.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_build_tile' -s 3 -c 1 -o $O/r02_prof_c5_build_tile_v6 -f \
  python bench.py --workload c5 --steps 60 --warmup 300 --e2e-steps 0 --cpu-rows -1 --steady-steps 0 --no-time-rebuild > $O/t_ncu_build.log 2>&1
for f in r02_prof_c5_force_tile_v6 r02_prof_c5_build_tile_v6; do python scripts/ncu_summary.py $O/$f.ncu-rep > $O/$f.txt 2>&1; head -31 $O/$f.txt | grep -E "##|duration|fp64_cycles|issue_active|inst_executed.sum|stalls|lsu_wavefronts.avg"; done
MOLDYN_B200_LIBRARY=moldyn_b200/lib/libmoldyn_b200_trace.so timeout 200 python scripts/loop_trace.py c2 8000 2>&1 | tail -12
