#!/bin/bash
mkdir -p gpurun_out
export MOLDYN_B200_LOOP=host
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_force' -s 6100 -c 1 -o gpurun_out/prof_c3_force_$1 -f python bench.py --workload c3 --steps 300 --warmup 6000 --e2e-steps 0 --cpu-rows -1 > gpurun_out/ncu_c3_force_$1.log 2>&1; tail -2 gpurun_out/ncu_c3_force_$1.log | cut -c1-200
