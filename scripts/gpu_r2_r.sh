#!/bin/bash
# Round 2, call R: next pair's list head prefetched in the loop's force phase — parity subset, bench lines, phase trace.
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "loop_drivers or golden or nve or determinism" 2>&1 | tail -2
line() {
python - "$1" "$2" <<'PY'
import json, sys
tag, path = sys.argv[1], sys.argv[2]
try:
    d=json.loads(open(path).read().strip().splitlines()[-1])
    r=d["roofline"]
    print(tag, "%.3e" % d["value"], "us/step %.2f" % (d["ms_per_step"]*1e3), r.get("phases_us") or r.get("kernels_ms"), "steady", d["steady_state"] and ("%.3e" % d["steady_state"]["value"], round(d["steady_state"]["us_per_step"],2), d["steady_state"]["rebuilds"]))
except Exception as e:
    print(tag, "FAILED", e); print(open(path.replace(".json",".err")).read()[-1500:])
PY
}
timeout 600 python bench.py --steps 20 --warmup 5 --e2e-steps 0 --cpu-rows -1 > $O/r_c3.json 2> $O/r_c3.err; line "c3 driver flags" $O/r_c3.json
timeout 300 python bench.py --workload c2 --steps 2000 --warmup 500 --e2e-steps 0 --cpu-rows -1 > $O/r_c2.json 2> $O/r_c2.err; line "c2" $O/r_c2.json
timeout 300 python bench.py --workload c3 --loop chunk --steps 20 --warmup 5 --e2e-steps 0 --cpu-rows -1 > $O/r_c3_chunk.json 2> $O/r_c3_chunk.err; line "c3 chunk" $O/r_c3_chunk.json
MOLDYN_B200_LIBRARY=moldyn_b200/lib/libmoldyn_b200_trace.so timeout 200 python scripts/loop_trace.py c3 8000 2>&1 | tail -13
