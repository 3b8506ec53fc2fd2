#!/bin/bash
# Round 2, call U: the final tree — full GPU suite, both arms with the driver's flags, every workload, ncu captures for
# profiles/ncu_traffic.json, smoke().
O=gpurun_out; mkdir -p $O
timeout 2400 python -m pytest tests -x -q -m gpu > $O/u_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/u_pytest.log
line() {
python - "$1" "$2" <<'PY'
import json, sys
tag, path = sys.argv[1], sys.argv[2]
try:
    d=json.loads(open(path).read().strip().splitlines()[-1])
    r=d["roofline"]; e=d.get("e2e")
    print(tag, "%.3e" % d["value"], "us/step %.2f" % (d["ms_per_step"]*1e3), r.get("phases_us") or r.get("kernels_ms"), "frac", r.get("frac"), "traffic", r.get("traffic"), "rebuild", r["rebuild"]["ms_each"], "steady", d["steady_state"] and ("%.3e" % d["steady_state"]["value"], round(d["steady_state"]["us_per_step"],2)), "e2e", e and (round(e["ms_per_step"],3), round(e["single_session"]["ms_per_step"],3)))
except Exception as e:
    print(tag, "FAILED", e); print(open(path.replace(".json",".err")).read()[-1500:])
PY
}
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/u_reference.json 2> $O/u_reference.err; cut -c1-300 $O/u_reference.json
timeout 600 python bench.py --steps 20 --warmup 5 > $O/u_default_driver.json 2> $O/u_default_driver.err; line "default (driver flags)" $O/u_default_driver.json
for w in c1 c2 c5 big; do
  timeout 300 python bench.py --workload $w --steps 2000 --warmup 500 --e2e-steps 0 --cpu-rows -1 > $O/u_${w}.json 2> $O/u_${w}.err; line "$w" $O/u_${w}.json
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_force_tile' -s 640 -c 1 -o $O/r02_prof_c5_force_tile_final -f \
  python bench.py --workload c5 --steps 60 --warmup 300 --e2e-steps 0 --cpu-rows -1 --steady-steps 0 --no-time-rebuild > $O/u_ncu_force.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_build_tile' -s 3 -c 1 -o $O/r02_prof_c5_build_tile_final -f \
  python bench.py --workload c5 --steps 60 --warmup 300 --e2e-steps 0 --cpu-rows -1 --steady-steps 0 --no-time-rebuild > $O/u_ncu_build.log 2>&1
MOLDYN_B200_LOOP=host timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_md_loop' -s 6200 -c 2 -o $O/r02_prof_c3_loop_final -f \
  python bench.py --workload c3 --steps 100 --warmup 6100 --e2e-steps 0 --cpu-rows -1 --steady-steps 0 --no-time-rebuild > $O/u_ncu_loop.log 2>&1
for f in r02_prof_c5_force_tile_final r02_prof_c5_build_tile_final r02_prof_c3_loop_final; do python scripts/ncu_summary.py $O/$f.ncu-rep > $O/$f.txt 2>&1; head -31 $O/$f.txt | grep -E "##|duration|fp64_cycles|issue_active|inst_executed.sum|stalls|dram__bytes"; done
MOLDYN_B200_LOOP=host timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_c3_final.csv python bench.py --steps 20 --warmup 5 --e2e-steps 0 --cpu-rows -1 --steady-steps 0 --no-time-rebuild > $O/u_launches.log 2>&1; wc -l $O/r02_launches_c3_final.csv
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | cut -c1-300
