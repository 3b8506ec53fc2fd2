#!/bin/bash
# $1 = GPUs.  A/B of the multi-GPU step variants on c3
N=${1:-2}
mkdir -p gpurun_out
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload ${W:-c3} --steps ${STEPS:-2000} --warmup ${WARM:-200} --e2e-steps 0 --cpu-rows -1 > gpurun_out/ab_${tag}_$N.json 2> gpurun_out/ab_${tag}_$N.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/ab_${tag}_$N.json") if l.startswith("{")][-1])
    print("$tag", round(d["value"]/1e9,2), "e9", round(d["ms_per_step"]*1e3,2), "us/step", [(r["n_ghost"], round(r["wait_halo_us_per_step"],2), round(r["wait_sums_us_per_step"],2), round(r["force_atoms_us_per_step"],2), round(r["force_tail_us_per_step"],2), round(r["rebuild_ms_each"],3)) for r in d["per_rank"]][:2])
except Exception as e: print("ERR $tag", e, open("gpurun_out/ab_${tag}_$N.err").read()[-800:])
PY
}
timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 1100 -x 2>&1 | tail -3
run default X=1
W=c3 STEPS=2000 WARM=6000 run late X=1
