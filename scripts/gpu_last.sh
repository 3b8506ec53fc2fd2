#!/bin/bash
# last call of the round: the driver's pytest command + default bench (+ C2, C5 lines) on the final tree
mkdir -p gpurun_out
O=gpurun_out
timeout 150 python -m pytest tests -x -q -m gpu > $O/last_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/last_pytest.log
timeout 60 python bench.py --e2e-steps 3 > $O/last_bench_c3.json 2> $O/last_bench_c3.err; cut -c1-330 $O/last_bench_c3.json; tail -2 $O/last_bench_c3.err
timeout 40 python bench.py --workload c2 --e2e-steps 0 --cpu-rows -1 > $O/last_bench_c2.json 2>&1; cut -c1-230 $O/last_bench_c2.json
timeout 40 python bench.py --workload c5 --steps 2000 --warmup 500 --e2e-steps 0 --cpu-rows -1 > $O/last_bench_c5.json 2>&1; cut -c1-230 $O/last_bench_c5.json
