#!/bin/bash
# Round 2, call D: tile kernels after the re-imaging fix — parity, staging A/B (cooperative loads vs TMA), ncu of both kernels.
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tile or nve or forces_match or neighbour_sets or determinism or rebuild_stress or (full_size and 64)" > $O/d_pytest_tile.log 2>&1; echo "pytest tile rc=$?"; tail -25 $O/d_pytest_tile.log
line() {
python - "$1" "$2" <<'PY'
import json, sys
tag, path = sys.argv[1], sys.argv[2]
try:
    d=json.loads(open(path).read().strip().splitlines()[-1])
    r=d["roofline"]
    print(tag, "%.3e" % d["value"], "us/step %.2f" % (d["ms_per_step"]*1e3), r.get("phases_us") or r.get("kernels_ms"), "frac", r.get("frac"), "rebuild", r["rebuild"], d["state_check"])
except Exception as e:
    print(tag, "FAILED", e); print(open(path.replace(".json",".err")).read()[-1500:])
PY
}
B="--workload c5 --steps 1000 --warmup 300 --e2e-steps 0 --cpu-rows -1 --steady-steps 0"
timeout 300 python bench.py $B > $O/d_c5_coop.json 2> $O/d_c5_coop.err; line "c5 tile coop-stage" $O/d_c5_coop.json
MOLDYN_B200_TILE_TMA=1 timeout 300 python bench.py $B > $O/d_c5_tma.json 2> $O/d_c5_tma.err; line "c5 tile tma-stage" $O/d_c5_tma.json
MOLDYN_B200_TILE=0 timeout 300 python bench.py $B > $O/d_c5_notile.json 2> $O/d_c5_notile.err; line "c5 notile" $O/d_c5_notile.json
MOLDYN_B200_TILE_BZ=2 timeout 300 python bench.py $B > $O/d_c5_bz2.json 2> $O/d_c5_bz2.err; line "c5 tile bz2" $O/d_c5_bz2.json
# ncu: one k_force_tile (steady, forces only), one k_build_tile
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_force_tile|k_build_tile' -s 330 -c 40 -o $O/r02_prof_c5_tile -f \
  python bench.py --workload c5 --steps 60 --warmup 300 --e2e-steps 0 --cpu-rows -1 --steady-steps 0 --no-time-rebuild > $O/d_ncu_tile.log 2>&1; tail -2 $O/d_ncu_tile.log | cut -c1-200
python scripts/ncu_summary.py $O/r02_prof_c5_tile.ncu-rep > $O/r02_ncu_c5_tile_all.txt 2>&1; grep -A30 "k_build_tile" $O/r02_ncu_c5_tile_all.txt | head -40; grep -A30 "k_force_tile" $O/r02_ncu_c5_tile_all.txt | head -36
