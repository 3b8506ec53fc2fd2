#!/bin/bash
# Round 2, call F: tile v2 (AoS shell, 8 lanes per atom, lane-per-column builder) — parity, bench, small ncu capture; loop after
# the L1-gather change.
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_cli.py -x -q -m gpu -k "not config_size and not c4_size and not long_run and not (full_size and 100)" > $O/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $O/f_pytest.log
line() {
python - "$1" "$2" <<'PY'
import json, sys
tag, path = sys.argv[1], sys.argv[2]
try:
    d=json.loads(open(path).read().strip().splitlines()[-1])
    r=d["roofline"]
    print(tag, "%.3e" % d["value"], "us/step %.2f" % (d["ms_per_step"]*1e3), r.get("phases_us") or r.get("kernels_ms"), "frac", r.get("frac"), "rebuild", r["rebuild"], "steady", d["steady_state"] and ("%.3e" % d["steady_state"]["value"], round(d["steady_state"]["us_per_step"],2), d["steady_state"]["rebuilds"], d["steady_state"]["nbr_mean"]))
except Exception as e:
    print(tag, "FAILED", e); print(open(path.replace(".json",".err")).read()[-1500:])
PY
}
B="--workload c5 --steps 1000 --warmup 300 --e2e-steps 0 --cpu-rows -1 --steady-steps 0"
timeout 300 python bench.py $B > $O/f_c5_tile.json 2> $O/f_c5_tile.err; line "c5 tile v2" $O/f_c5_tile.json
MOLDYN_B200_TILE_BZ=2 timeout 300 python bench.py $B > $O/f_c5_bz2.json 2> $O/f_c5_bz2.err; line "c5 tile v2 bz2" $O/f_c5_bz2.json
MOLDYN_B200_TILE_BZ=3 timeout 300 python bench.py $B > $O/f_c5_bz3.json 2> $O/f_c5_bz3.err; line "c5 tile v2 bz3" $O/f_c5_bz3.json
timeout 300 python bench.py $B --skin 0.13 > $O/f_c5_skin13.json 2> $O/f_c5_skin13.err; line "c5 tile v2 skin .13" $O/f_c5_skin13.json
MOLDYN_B200_TILE=0 timeout 300 python bench.py $B > $O/f_c5_notile.json 2> $O/f_c5_notile.err; line "c5 notile" $O/f_c5_notile.json
for w in c1 c2 c3; do
  timeout 300 python bench.py --workload $w --steps 2000 --warmup 500 --e2e-steps 0 --cpu-rows -1 > $O/f_${w}.json 2> $O/f_${w}.err; line "$w auto" $O/f_${w}.json
done
timeout 300 python bench.py --workload c3 --steps 20 --warmup 5 --e2e-steps 3 --cpu-rows -1 --steady-steps 0 > $O/f_c3_drv.json 2> $O/f_c3_drv.err; line "c3 driver-like" $O/f_c3_drv.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_force_tile' -s 640 -c 2 -o $O/r02_prof_c5_force_tile -f \
  python bench.py --workload c5 --steps 60 --warmup 300 --e2e-steps 0 --cpu-rows -1 --steady-steps 0 --no-time-rebuild > $O/f_ncu_force.log 2>&1; tail -1 $O/f_ncu_force.log | cut -c1-120
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_build_tile' -s 3 -c 1 -o $O/r02_prof_c5_build_tile -f \
  python bench.py --workload c5 --steps 60 --warmup 300 --e2e-steps 0 --cpu-rows -1 --steady-steps 0 --no-time-rebuild > $O/f_ncu_build.log 2>&1; tail -1 $O/f_ncu_build.log | cut -c1-120
python scripts/ncu_summary.py $O/r02_prof_c5_force_tile.ncu-rep > $O/r02_ncu_c5_force_tile.txt 2>&1; cat $O/r02_ncu_c5_force_tile.txt | head -64
python scripts/ncu_summary.py $O/r02_prof_c5_build_tile.ncu-rep > $O/r02_ncu_c5_build_tile.txt 2>&1; cat $O/r02_ncu_c5_build_tile.txt | head -34
ls -la $O/*.ncu-rep
