#!/bin/bash
# Round 2, call E ($1 GPUs): the persistent loop across GPUs — slab parity worker, then strong-scaling lines of C3 (loop vs the
# two-kernel chunk path) in the driver's configuration (20 steps from the lattice) and in the collisional steady state.
N=${1:-2}
O=gpurun_out; mkdir -p $O
nvidia-smi -L | head -8
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tests/dist_worker.py > $O/e_worker_$N.log 2>&1; echo "worker rc=$?"; tail -14 $O/e_worker_$N.log
show() {
python - "$1" "$2" <<'PY'
import json, sys
tag, path = sys.argv[1], sys.argv[2]
try:
    d=json.loads([l for l in open(path) if l.startswith("{")][-1])
    print(tag, "%.3e" % d["value"], "us/step %.2f" % (d["ms_per_step"]*1e3), "rebuilds", d["rebuilds_in_timed_region"], "steady", d["steady_state"] and ("%.3e" % d["steady_state"]["value"], round(d["steady_state"]["us_per_step"],2), d["steady_state"]["rebuilds"]), d["state_check"]["temperature"])
    for r, pr in enumerate(d.get("per_rank") or []):
        print("   rank", r, {k: (round(v,2) if isinstance(v,float) else v) for k,v in pr.items() if k not in ("loop_us_per_step",)}, {k: round(v,2) for k,v in pr["loop_us_per_step"].items()})
except Exception as e:
    print(tag, "FAILED", e); print(open(path.replace(".json",".err")).read()[-2500:])
PY
}
run() {  # tag, gpus, extra env, bench args
  tag=$1; g=$2; envs=$3; shift 3
  if [ $g = 1 ]; then
    env $envs timeout 600 python bench.py "$@" > $O/e_$tag.json 2> $O/e_$tag.err
  else
    env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $g "$@" > $O/e_$tag.json 2> $O/e_$tag.err
  fi
  show "$tag" $O/e_$tag.json
}
D="--workload c3 --steps 20 --warmup 5 --e2e-steps 0 --cpu-rows -1"
S="--workload c3 --steps 2000 --warmup 6000 --e2e-steps 0 --cpu-rows -1 --steady-steps 0"
for g in 1 $N; do
  run c3_drv_loop_$g $g "A=1" $D
  run c3_drv_chunk_$g $g "MOLDYN_B200_LOOP=chunk" $D
  run c3_steady_loop_$g $g "A=1" $S
  run c3_steady_chunk_$g $g "MOLDYN_B200_LOOP=chunk" $S
done
run c4_loop_$N $N "A=1" --workload c4 --steps 300 --warmup 100 --e2e-steps 0 --cpu-rows -1 --steady-steps 0
run c4_chunk_$N $N "MOLDYN_B200_LOOP=chunk" --workload c4 --steps 300 --warmup 100 --e2e-steps 0 --cpu-rows -1 --steady-steps 0
if [ "${2:-}" = "extra" ]; then
  timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "loop_drivers or tile or determinism" > $O/e_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/e_pytest.log
  run c5_tile_1 1 "A=1" --workload c5 --steps 1000 --warmup 300 --e2e-steps 0 --cpu-rows -1 --steady-steps 0
  python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/e_c5_tile_1.json") if l.startswith("{")][-1]); print(d["roofline"]["kernels_ms"], d["roofline"]["rebuild"], d["roofline"].get("frac"))
PY
fi
