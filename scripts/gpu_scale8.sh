#!/bin/bash
# 8-GPU box: parity worker at 4 and 8 ranks, then the scaling series of the bench
mkdir -p gpurun_out
nvidia-smi -L | head -8
for N in 4 8; do
  echo "== dist_worker N=$N"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N tests/dist_worker.py 2>&1 | grep -E "ok|DIST_OK|Error|error|assert" | tail -8
done
run() { # N workload steps warm
  N=$1; W=$2
  if [ $N = 1 ]; then
    timeout 600 python bench.py --workload $W --steps $3 --warmup $4 --e2e-steps 0 --cpu-rows -1 > gpurun_out/s8_${W}_$N.json 2> gpurun_out/s8_${W}_$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --workload $W --steps $3 --warmup $4 --e2e-steps 0 --cpu-rows -1 > gpurun_out/s8_${W}_$N.json 2> gpurun_out/s8_${W}_$N.err
  fi
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/s8_${W}_$N.json") if l.startswith("{")][-1])
    pr=d.get("per_rank") or []
    print("$W N=$N", round(d["value"]/1e9,2), "e9", round(d["ms_per_step"]*1e3,2), "us/step rebuilds", d["rebuilds_in_timed_region"], [(r["n_ghost"], round(r["wait_halo_us_per_step"],1), round(r["wait_sums_us_per_step"],1), round(r["force_atoms_us_per_step"],1), round(r["force_tail_us_per_step"],1)) for r in pr][:3])
except Exception as e: print("ERR $W $N", e, open("gpurun_out/s8_${W}_$N.err").read()[-600:])
PY
}
for N in 1 2 4 8; do run $N c3 3000 3000; done
for N in 1 8; do run $N big 500 500; done
run 8 c4 300 300
