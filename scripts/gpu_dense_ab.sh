#!/bin/bash
# Dense force kernel A/B on one box: per-thread lists (default) vs union lists vs warp-cooperative (k_force_coop),
# parity tests of the variants first, one ncu --set full capture of k_force_coop.
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
lap() { echo "$(( $(date +%s) - T0 )) s  $1" | tee -a $O/dense_timing.log; }
: > $O/dense_timing.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "warp_cooperative or union or forces_match_oracle or determinism" \
  --durations=5 > $O/dense_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/dense_timing.log
tail -12 $O/dense_pytest.log
lap "pytest"
bench() {
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 300 python bench.py "$@" > $O/dense_bench_$name.json 2> $O/dense_bench_$name.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/dense_bench_$name.json") if l.startswith("{")][-1])
    r = d.get("roofline") or {}
    print("$name", "value %.4g" % d["value"], "ms/step %.5f" % d["ms_per_step"], "k_ms", r.get("kernels_ms"),
          "rebuilds", d.get("rebuilds_in_timed_region"), d.get("state_check"))
except Exception as e:
    print("ERR $name", e, open("$O/dense_bench_$name.err").read()[-800:])
PY
  lap "bench $name"
}
A="--workload c5 --steps 2000 --warmup 500 --e2e-steps 0 --cpu-rows -1"
bench c5_default -- $A
bench c5_coop -- $A --dense-kernel coop
bench c5_union -- $A --dense-kernel union
bench c5_coop_pdl MOLDYN_B200_PDL=1 -- $A --dense-kernel coop
export MOLDYN_B200_LOOP=host
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_force_coop' -s 320 -c 1 -o $O/r01_prof_c5_coop -f \
  python bench.py --workload c5 --steps 60 --warmup 300 --e2e-steps 0 --cpu-rows -1 --dense-kernel coop > $O/dense_ncu_coop.log 2>&1
tail -1 $O/dense_ncu_coop.log | cut -c1-150
python scripts/ncu_summary.py $O/r01_prof_c5_coop.ncu-rep > $O/r01_ncu_c5_coop_k_force_coop.txt 2>&1; cat $O/r01_ncu_c5_coop_k_force_coop.txt | head -30
lap "ncu coop"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r01_launches_c5_coop.csv \
  python bench.py --workload c5 --steps 60 --warmup 3 --e2e-steps 0 --cpu-rows -1 --dense-kernel coop > $O/dense_launches.log 2>&1
lap "launch list"
