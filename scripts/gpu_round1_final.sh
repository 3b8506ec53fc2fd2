#!/bin/bash
# Round-1 closing run (one gpurun call, most important first; every phase under its own timeout):
#   1. the driver's own command: pytest -m gpu -x -q
#   2. default bench (C3) — plain launches vs programmatic dependent launch (MOLDYN_B200_PDL=1) vs 5 blocks/SM
#   3. C5 bench — this tree vs the previous commit's kernels (variants/libmd_r1base.so) on the same box
#   4. the graph-loop / trajectory / determinism tests again with MOLDYN_B200_PDL=1
#   5. ncu --set full of the dense force kernel, launch list of the default bench command
#   6. smoke(), reference arm
# Variant libraries are builds of this repo's sources (scripts/README.md); they are selected with MOLDYN_B200_LIBRARY.
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
lap() { echo "$(( $(date +%s) - T0 )) s  $1" | tee -a $O/final_timing.log; }
V=moldyn_b200/lib/variants
: > $O/final_timing.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader | tee -a $O/final_timing.log

timeout 600 python -m pytest tests -x -q -m gpu --durations=8 > $O/final_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/final_timing.log
tail -15 $O/final_pytest_gpu.log
lap "pytest -m gpu"

bench() {  # name, env..., -- args
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 300 python bench.py "$@" > $O/final_bench_$name.json 2> $O/final_bench_$name.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/final_bench_$name.json") if l.startswith("{")][-1])
    r = d.get("roofline") or {}
    print("$name", "value %.4g" % d["value"], "ms/step %.5f" % d["ms_per_step"], "e2e", d.get("e2e") and "%.3g" % d["e2e"]["value"],
          "k_ms", r.get("kernels_ms"), "clk", d.get("clocks", {}).get("sm_mhz"), d.get("state_check"))
except Exception as e:
    print("ERR $name", e, open("$O/final_bench_$name.err").read()[-600:])
PY
  lap "bench $name"
}

bench c3_default --
bench c3_pdl MOLDYN_B200_PDL=1 -- --e2e-steps 0 --cpu-rows -1
bench c5_default -- --workload c5 --steps 2000 --warmup 500 --e2e-steps 3
bench c5_r1base MOLDYN_B200_LIBRARY=$PWD/$V/libmd_r1base.so -- --workload c5 --steps 2000 --warmup 500 --e2e-steps 0 --cpu-rows -1
bench c5_pdl MOLDYN_B200_PDL=1 -- --workload c5 --steps 2000 --warmup 500 --e2e-steps 0 --cpu-rows -1
bench c2_default -- --workload c2 --e2e-steps 0 --cpu-rows -1
bench c2_pdl MOLDYN_B200_PDL=1 -- --workload c2 --e2e-steps 0 --cpu-rows -1
bench c3_minb5 MOLDYN_B200_LIBRARY=$PWD/$V/libmd_dilute_minb5.so -- --e2e-steps 0 --cpu-rows -1
bench c3_minb5_pdl MOLDYN_B200_PDL=1 MOLDYN_B200_LIBRARY=$PWD/$V/libmd_dilute_minb5.so -- --e2e-steps 0 --cpu-rows -1

MOLDYN_B200_PDL=1 timeout 420 python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
  -k "golden or graph_loop or determinism or thermostat_barostat or nve_trajectory or nose_hoover or rebuild_stress or per_call or macro or full_size" \
  > $O/final_pytest_pdl.log 2>&1; echo "pytest(PDL) rc=$?" | tee -a $O/final_timing.log
tail -5 $O/final_pytest_pdl.log
lap "pytest PDL subset"

export MOLDYN_B200_LOOP=host
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_force' -s 320 -c 1 -o $O/r01_prof_c5_v10 -f \
  python bench.py --workload c5 --steps 60 --warmup 300 --e2e-steps 0 --cpu-rows -1 > $O/final_ncu_c5.log 2>&1
tail -1 $O/final_ncu_c5.log | cut -c1-150
python scripts/ncu_summary.py $O/r01_prof_c5_v10.ncu-rep > $O/r01_ncu_c5_v10_k_force.txt 2>&1; head -30 $O/r01_ncu_c5_v10_k_force.txt
lap "ncu c5"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r01_launches_c3_v10.csv \
  python bench.py --steps 60 --warmup 3 --e2e-steps 1 --cpu-rows -1 > $O/final_launches_c3.log 2>&1
tail -2 $O/final_launches_c3.log | cut -c1-200
lap "launch list c3"
unset MOLDYN_B200_LOOP

timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/final_timing.log
tail -2 $O/final_smoke.log
lap "smoke"
bench reference_arm -- --impl reference
lap "done"
