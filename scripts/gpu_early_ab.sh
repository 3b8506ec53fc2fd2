#!/bin/bash
# Early-start k_kick_drift (MOLDYN_B200_PDL=2) A/B on one box + parity under the flag + the cooperative-kernel test.
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
lap() { echo "$(( $(date +%s) - T0 )) s  $1" | tee -a $O/early_timing.log; }
: > $O/early_timing.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "programmatic or warp_cooperative" > $O/early_pytest1.log 2>&1; echo "pytest(new tests) rc=$?" | tee -a $O/early_timing.log
tail -4 $O/early_pytest1.log
lap "pytest new tests"
bench() {
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 300 python bench.py "$@" > $O/early_bench_$name.json 2> $O/early_bench_$name.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/early_bench_$name.json") if l.startswith("{")][-1])
    print("$name", "value %.4g" % d["value"], "ms/step %.5f" % d["ms_per_step"], "rebuilds", d.get("rebuilds_in_timed_region"), d.get("state_check"))
except Exception as e:
    print("ERR $name", e, open("$O/early_bench_$name.err").read()[-800:])
PY
  lap "bench $name"
}
Q="--e2e-steps 0 --cpu-rows -1"
for w in ${EARLY_WORKLOADS:-c3 c2 c1 c5 big}; do
  S=""; [ $w = c5 ] && S="--steps 2000 --warmup 500"
  bench ${w}_default -- --workload $w $Q $S
  bench ${w}_pdl1 MOLDYN_B200_PDL=1 -- --workload $w $Q $S
  bench ${w}_pdl2 MOLDYN_B200_PDL=2 -- --workload $w $Q $S
done
MOLDYN_B200_PDL=2 timeout 420 python -m pytest tests/test_gpu_parity.py tests/test_cli.py -x -q -m gpu > $O/early_pytest2.log 2>&1; echo "pytest(PDL=2, parity+cli) rc=$?" | tee -a $O/early_timing.log
tail -4 $O/early_pytest2.log
lap "pytest PDL=2"
