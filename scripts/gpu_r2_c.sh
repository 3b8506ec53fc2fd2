#!/bin/bash
# Round 2, call C: loop v3 (velocities in shared memory, warp folds) + tile kernels (dense) — parity, traces, A/B bench lines.
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_cli.py -x -q -m gpu -k "not config_size and not c4_size and not full_size and not long_run" > $O/c_pytest_fast.log 2>&1; echo "pytest fast rc=$?"; tail -25 $O/c_pytest_fast.log
for w in c2 c3; do
  MOLDYN_B200_LIBRARY=moldyn_b200/lib/libmoldyn_b200_trace.so timeout 200 python scripts/loop_trace.py $w 500 2>&1 | tail -13
  MOLDYN_B200_LIBRARY=moldyn_b200/lib/libmoldyn_b200_trace.so timeout 200 python scripts/loop_trace.py $w 8000 2>&1 | tail -13
done
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
for w in c1 c2 c3; do
  for loop in auto chunk; do
    timeout 300 python bench.py --workload $w --loop $loop --steps 2000 --warmup 500 --e2e-steps 0 --cpu-rows -1 > $O/c_bench_${w}_${loop}.json 2> $O/c_bench_${w}_${loop}.err
    line "$w $loop" $O/c_bench_${w}_${loop}.json
  done
done
timeout 300 python bench.py --workload c5 --steps 1000 --warmup 300 --e2e-steps 0 --cpu-rows -1 > $O/c_bench_c5_tile.json 2> $O/c_bench_c5_tile.err; line "c5 tile" $O/c_bench_c5_tile.json
MOLDYN_B200_TILE=0 timeout 300 python bench.py --workload c5 --steps 1000 --warmup 300 --e2e-steps 0 --cpu-rows -1 > $O/c_bench_c5_notile.json 2> $O/c_bench_c5_notile.err; line "c5 notile" $O/c_bench_c5_notile.json
for bz in 2 3 6; do
MOLDYN_B200_TILE_BZ=$bz timeout 300 python bench.py --workload c5 --steps 1000 --warmup 300 --e2e-steps 0 --cpu-rows -1 --steady-steps 0 > $O/c_bench_c5_bz$bz.json 2> $O/c_bench_c5_bz$bz.err; line "c5 bz$bz" $O/c_bench_c5_bz$bz.json
done
timeout 300 python bench.py --workload big --steps 500 --warmup 100 --e2e-steps 0 --cpu-rows -1 --steady-steps 0 > $O/c_bench_big.json 2> $O/c_bench_big.err; line "big auto" $O/c_bench_big.json
timeout 300 python bench.py --workload big --loop chunk --steps 500 --warmup 100 --e2e-steps 0 --cpu-rows -1 --steady-steps 0 > $O/c_bench_big_chunk.json 2> $O/c_bench_big_chunk.err; line "big chunk" $O/c_bench_big_chunk.json
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "config_size or full_size" > $O/c_pytest_slow.log 2>&1; echo "pytest slow rc=$?"; tail -15 $O/c_pytest_slow.log
