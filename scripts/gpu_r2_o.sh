#!/bin/bash
# Round 2, call O: list heads of the loop's pairs stashed in shared memory — parity subset, bench lines, phase trace.
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "loop or nve or nvt or npt or golden or trajectory or determinism or rebuild" > $O/o_parity.log 2>&1; echo "parity rc=$?"; tail -4 $O/o_parity.log
line() {
python - "$1" "$2" <<'PY'
import json, sys
tag, path = sys.argv[1], sys.argv[2]
try:
    d=json.loads(open(path).read().strip().splitlines()[-1])
    r=d["roofline"]
    print(tag, "%.3e" % d["value"], "us/step %.2f" % (d["ms_per_step"]*1e3), r.get("phases_us") or r.get("kernels_ms"), "frac", r.get("frac"), "steady", d["steady_state"] and ("%.3e" % d["steady_state"]["value"], round(d["steady_state"]["us_per_step"],2), d["steady_state"]["rebuilds"], d["steady_state"]["nbr_mean"]))
except Exception as e:
    print(tag, "FAILED", e); print(open(path.replace(".json",".err")).read()[-1500:])
PY
}
timeout 600 python bench.py --steps 20 --warmup 5 --e2e-steps 0 --cpu-rows -1 > $O/o_default.json 2> $O/o_default.err; line "c3 driver flags" $O/o_default.json
for w in c2 c3 big; do
  timeout 300 python bench.py --workload $w --steps 2000 --warmup 500 --e2e-steps 0 --cpu-rows -1 > $O/o_${w}.json 2> $O/o_${w}.err; line "$w auto" $O/o_${w}.json
done
timeout 300 python bench.py --workload c3 --loop chunk --steps 2000 --warmup 500 --e2e-steps 0 --cpu-rows -1 > $O/o_c3_chunk.json 2> $O/o_c3_chunk.err; line "c3 chunk" $O/o_c3_chunk.json
for w in c3; do
  MOLDYN_B200_LIBRARY=moldyn_b200/lib/libmoldyn_b200_trace.so timeout 200 python scripts/loop_trace.py $w 8000 2>&1 | tail -13
done
