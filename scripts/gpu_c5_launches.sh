#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -3
export MOLDYN_B200_LOOP=host
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r01f_launches_c5.csv python bench.py --workload c5 --steps 60 --warmup 200 --e2e-steps 0 --cpu-rows -1 > gpurun_out/r01f_launches_c5.log 2>&1; tail -1 gpurun_out/r01f_launches_c5.log | cut -c1-100
