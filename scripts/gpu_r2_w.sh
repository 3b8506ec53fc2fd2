#!/bin/bash
# Round 2, call W (8 GPUs), final tree: slab worker at 8 ranks, the driver's scaling series of C3 (20 steps from the lattice) at
# N = 1, 2, 4, 8, the steady state at N = 8, C4 on 8 GPUs.
O=gpurun_out; mkdir -p $O
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tests/dist_worker.py > $O/w_worker_8.log 2>&1; echo "worker rc=$?"; grep -v "^\*\|OMP_NUM\|NCCL version" $O/w_worker_8.log | tail -9
show() {
python - "$1" "$2" <<'PY'
import json, sys
tag, path = sys.argv[1], sys.argv[2]
try:
    d=json.loads([l for l in open(path) if l.startswith("{")][-1])
    print(tag, "%.3e" % d["value"], "us/step %.2f" % (d["ms_per_step"]*1e3), "rebuilds", d["rebuilds_in_timed_region"], "T", d["state_check"]["temperature"], "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    pr = d.get("per_rank") or []
    if pr:
        import statistics as st
        keys = ("wait_halo_us_per_step", "wait_sums_us_per_step", "rebuild_ms_each")
        print("   per rank:", {k: [round(p[k], 2) for p in pr] for k in keys})
        print("   loop phases:", {k: [round(p["loop_us_per_step"][k], 2) for p in pr] for k in pr[0]["loop_us_per_step"]})
        print("   owned:", [p["n_owned"] for p in pr], "ghosts:", [p["n_ghost"] for p in pr])
except Exception as e:
    print(tag, "FAILED", e); print(open(path.replace(".json",".err")).read()[-2500:])
PY
}
run() {  # tag, gpus, extra env, bench args
  tag=$1; g=$2; envs=$3; shift 3
  if [ $g = 1 ]; then
    env $envs timeout 400 python bench.py "$@" > $O/w_$tag.json 2> $O/w_$tag.err
  else
    env $envs timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $g "$@" > $O/w_$tag.json 2> $O/w_$tag.err
  fi
  show "$tag" $O/w_$tag.json
}
D="--workload c3 --steps 20 --warmup 5 --e2e-steps 0 --cpu-rows -1 --steady-steps 0"
for g in 1 2 4 8; do run c3_drv_$g $g "A=1" $D; done
run c3_steady_8 8 "A=1" --workload c3 --steps 2000 --warmup 6000 --e2e-steps 0 --cpu-rows -1 --steady-steps 0
run c4_8 8 "A=1" --workload c4 --steps 300 --warmup 100 --e2e-steps 0 --cpu-rows -1 --steady-steps 0
