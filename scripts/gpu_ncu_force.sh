#!/bin/bash
# full ncu capture of k_force late in a c3 run (lists non-empty) and in c5; host-stepped so ncu can see the kernels
mkdir -p gpurun_out
export MOLDYN_B200_LOOP=host
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_force' -s 6100 -c 2 -o gpurun_out/prof_c3_late -f python bench.py --workload c3 --steps 100 --warmup 6000 --e2e-steps 0 --cpu-rows -1 > gpurun_out/ncu_c3_late.log 2>&1; tail -2 gpurun_out/ncu_c3_late.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_force' -s 320 -c 1 -o gpurun_out/prof_c5_v3 -f python bench.py --workload c5 --steps 60 --warmup 300 --e2e-steps 0 --cpu-rows -1 --cell-subdiv 2 > gpurun_out/ncu_c5_v3.log 2>&1; tail -2 gpurun_out/ncu_c5_v3.log | cut -c1-200
