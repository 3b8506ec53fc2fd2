#!/bin/bash
# ncu --set full capture of one k_step_dilute launch in the late (mixed) state of C3
mkdir -p gpurun_out
export MOLDYN_B200_LOOP=host
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_step_dilute' -s 3200 -c 1 -o gpurun_out/prof_c3_step_$1 -f python bench.py --workload c3 --steps 300 --warmup 3000 --e2e-steps 0 --cpu-rows -1 > gpurun_out/ncu_c3_step_$1.log 2>&1; tail -2 gpurun_out/ncu_c3_step_$1.log | cut -c1-200
