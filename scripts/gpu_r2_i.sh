#!/bin/bash
# Round 2, call I: full GPU suite on the current tree, loop + tile bench lines, phase traces, ncu of the tile kernels and the loop.
O=gpurun_out; mkdir -p $O
timeout 2400 python -m pytest tests -x -q -m gpu > $O/i_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/i_pytest.log
line() {
python - "$1" "$2" <<'PY'
import json, sys
tag, path = sys.argv[1], sys.argv[2]
try:
    d=json.loads(open(path).read().strip().splitlines()[-1])
    r=d["roofline"]
    print(tag, "%.3e" % d["value"], "us/step %.2f" % (d["ms_per_step"]*1e3), r.get("phases_us") or r.get("kernels_ms"), "frac", r.get("frac"), "rebuild", r["rebuild"], "steady", d["steady_state"] and ("%.3e" % d["steady_state"]["value"], round(d["steady_state"]["us_per_step"],2), d["steady_state"]["rebuilds"], d["steady_state"]["nbr_mean"]), "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],3))
except Exception as e:
    print(tag, "FAILED", e); print(open(path.replace(".json",".err")).read()[-1500:])
PY
}
timeout 300 python bench.py --workload c5 --steps 1000 --warmup 300 --e2e-steps 0 --cpu-rows -1 --steady-steps 0 > $O/i_c5.json 2> $O/i_c5.err; line "c5 tile" $O/i_c5.json
for w in c1 c2 c3 big; do
  timeout 300 python bench.py --workload $w --steps 2000 --warmup 500 --e2e-steps 0 --cpu-rows -1 > $O/i_${w}.json 2> $O/i_${w}.err; line "$w auto" $O/i_${w}.json
done
timeout 300 python bench.py --workload c3 --loop chunk --steps 2000 --warmup 500 --e2e-steps 0 --cpu-rows -1 > $O/i_c3_chunk.json 2> $O/i_c3_chunk.err; line "c3 chunk" $O/i_c3_chunk.json
timeout 600 python bench.py --steps 20 --warmup 5 > $O/i_default_driver.json 2> $O/i_default_driver.err; line "default (driver flags)" $O/i_default_driver.json
for w in c2 c3; do
  MOLDYN_B200_LIBRARY=moldyn_b200/lib/libmoldyn_b200_trace.so timeout 200 python scripts/loop_trace.py $w 8000 2>&1 | tail -13
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_force_tile' -s 640 -c 1 -o $O/r02_prof_c5_force_tile_v3 -f \
  python bench.py --workload c5 --steps 60 --warmup 300 --e2e-steps 0 --cpu-rows -1 --steady-steps 0 --no-time-rebuild > $O/i_ncu_force.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_build_tile' -s 3 -c 1 -o $O/r02_prof_c5_build_tile_v3 -f \
  python bench.py --workload c5 --steps 60 --warmup 300 --e2e-steps 0 --cpu-rows -1 --steady-steps 0 --no-time-rebuild > $O/i_ncu_build.log 2>&1
MOLDYN_B200_LOOP=host timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_md_loop' -s 6200 -c 2 -o $O/r02_prof_c3_loop -f \
  python bench.py --workload c3 --steps 100 --warmup 6100 --e2e-steps 0 --cpu-rows -1 --steady-steps 0 --no-time-rebuild > $O/i_ncu_loop.log 2>&1; tail -2 $O/i_ncu_loop.log | cut -c1-200
for f in r02_prof_c5_force_tile_v3 r02_prof_c5_build_tile_v3 r02_prof_c3_loop; do python scripts/ncu_summary.py $O/$f.ncu-rep > $O/$f.txt 2>&1; head -31 $O/$f.txt; done
ls -la $O/*.ncu-rep
