#!/bin/bash
# $1 = number of GPUs on the box. single-GPU regression + multi-GPU parity worker + multi-GPU bench
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest gpu (incl. 2-GPU worker)"
timeout 2400 python -m pytest tests -m gpu -q --timeout 1500 -x 2>&1 | tail -15
for w in c3 big; do
  for g in 1 $N; do
    steps=2000; [ $w = big ] && steps=300
    echo "== bench $w gpus=$g"
    if [ $g = 1 ]; then
      timeout 900 python bench.py --workload $w --steps $steps --warmup 200 --e2e-steps 0 --cpu-rows -1 > gpurun_out/dist_${w}_$g.json 2> gpurun_out/dist_${w}_$g.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $g --workload $w --steps $steps --warmup 200 --e2e-steps 3 --cpu-rows -1 > gpurun_out/dist_${w}_$g.json 2> gpurun_out/dist_${w}_$g.err
    fi
    python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/dist_${w}_$g.json") if l.startswith("{")][-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus","rebuilds_in_timed_region")}, d.get("per_rank"), d.get("e2e") and d["e2e"]["ms_per_step"])
except Exception as e: print("ERR", e, open("gpurun_out/dist_${w}_$g.err").read()[-1500:])
PY
  done
done
