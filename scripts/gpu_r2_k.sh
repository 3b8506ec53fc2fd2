#!/bin/bash
# Round 2, call K: full GPU suite after the md_step restructure (speculative enqueue), brick-height A/B on C5, two-session e2e,
# driver-flag line, ncu of the tile kernels (v4: flattened staging, four candidates per builder trip).
O=gpurun_out; mkdir -p $O
timeout 2400 python -m pytest tests -x -q -m gpu > $O/k_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/k_pytest.log
line() {
python - "$1" "$2" <<'PY'
import json, sys
tag, path = sys.argv[1], sys.argv[2]
try:
    d=json.loads(open(path).read().strip().splitlines()[-1])
    r=d["roofline"]
    e=d["e2e"]
    print(tag, "%.3e" % d["value"], "us/step %.2f" % (d["ms_per_step"]*1e3), r.get("phases_us") or r.get("kernels_ms"), "frac", r.get("frac"), "rebuild", r["rebuild"], "steady", d["steady_state"] and ("%.3e" % d["steady_state"]["value"], round(d["steady_state"]["us_per_step"],2), d["steady_state"]["rebuilds"], d["steady_state"]["nbr_mean"]), "e2e", e and (round(e["ms_per_step"],3), e.get("single_session")))
except Exception as e:
    print(tag, "FAILED", e); print(open(path.replace(".json",".err")).read()[-1500:])
PY
}
for bz in 0 4 2 3; do
  MOLDYN_B200_TILE_BZ=$bz timeout 300 python bench.py --workload c5 --steps 1000 --warmup 300 --e2e-steps 0 --cpu-rows -1 --steady-steps 0 > $O/k_c5_bz$bz.json 2> $O/k_c5_bz$bz.err; line "c5 tile bz=$bz" $O/k_c5_bz$bz.json
done
timeout 600 python bench.py --steps 20 --warmup 5 > $O/k_default_driver.json 2> $O/k_default_driver.err; line "default (driver flags)" $O/k_default_driver.json
for w in c1 c2 c3 big; do
  timeout 300 python bench.py --workload $w --steps 2000 --warmup 500 --e2e-steps 0 --cpu-rows -1 > $O/k_${w}.json 2> $O/k_${w}.err; line "$w auto" $O/k_${w}.json
done
timeout 300 python bench.py --workload c3 --loop chunk --steps 20 --warmup 5 --e2e-steps 0 --cpu-rows -1 --steady-steps 0 > $O/k_c3_chunk_drv.json 2> $O/k_c3_chunk_drv.err; line "c3 chunk driver flags" $O/k_c3_chunk_drv.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_force_tile' -s 640 -c 1 -o $O/r02_prof_c5_force_tile_v4 -f \
  python bench.py --workload c5 --steps 60 --warmup 300 --e2e-steps 0 --cpu-rows -1 --steady-steps 0 --no-time-rebuild > $O/k_ncu_force.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_build_tile' -s 3 -c 1 -o $O/r02_prof_c5_build_tile_v4 -f \
  python bench.py --workload c5 --steps 60 --warmup 300 --e2e-steps 0 --cpu-rows -1 --steady-steps 0 --no-time-rebuild > $O/k_ncu_build.log 2>&1
for f in r02_prof_c5_force_tile_v4 r02_prof_c5_build_tile_v4; do python scripts/ncu_summary.py $O/$f.ncu-rep > $O/$f.txt 2>&1; head -31 $O/$f.txt; done
ls -la $O/*_v4.ncu-rep
MOLDYN_B200_LOOP=host timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_md_loop' -s 6200 -c 2 -o $O/r02_prof_c3_loop_v4 -f \
  python bench.py --workload c3 --steps 100 --warmup 6100 --e2e-steps 0 --cpu-rows -1 --steady-steps 0 --no-time-rebuild > $O/k_ncu_loop.log 2>&1; tail -2 $O/k_ncu_loop.log | cut -c1-200
python scripts/ncu_summary.py $O/r02_prof_c3_loop_v4.ncu-rep > $O/r02_prof_c3_loop_v4.txt 2>&1; head -31 $O/r02_prof_c3_loop_v4.txt
