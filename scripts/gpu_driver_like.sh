#!/bin/bash
# the exact launches the driver uses at N=2: our arm and the reference arm
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --impl reference --steps 3 --warmup 1 > gpurun_out/drv_ref_$N.json 2> gpurun_out/drv_ref_$N.err; tail -c 500 gpurun_out/drv_ref_$N.json; echo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N > gpurun_out/drv_ours_$N.json 2> gpurun_out/drv_ours_$N.err; grep '^{' gpurun_out/drv_ours_$N.json | tail -c 1500; echo; tail -3 gpurun_out/drv_ours_$N.err
