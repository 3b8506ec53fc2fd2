#!/bin/bash
mkdir -p gpurun_out
export MOLDYN_B200_LOOP=host
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_build_list' -s 3 -c 1 -o gpurun_out/prof_c5_build -f python bench.py --workload c5 --steps 40 --warmup 200 --e2e-steps 0 --cpu-rows -1 > gpurun_out/ncu_c5_build.log 2>&1; tail -1 gpurun_out/ncu_c5_build.log | cut -c1-100
