#!/bin/bash
# Driver-like verification of the committed tree on one box: pytest -m gpu -x -q, smoke(), both bench arms,
# the launch list of the default bench command; with N > 1 GPUs also the multi-GPU tests and the torchrun launches.
mkdir -p gpurun_out
O=gpurun_out
N=${1:-1}
T0=$(date +%s)
lap() { echo "$(( $(date +%s) - T0 )) s  $1" | tee -a $O/verify_timing_$N.log; }
: > $O/verify_timing_$N.log
timeout 600 python -m pytest tests -x -q -m gpu --durations=5 > $O/verify_pytest_$N.log 2>&1; echo "pytest rc=$?" | tee -a $O/verify_timing_$N.log
tail -9 $O/verify_pytest_$N.log
lap "pytest -m gpu"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/verify_smoke_$N.log 2>&1; echo "smoke rc=$?" | tee -a $O/verify_timing_$N.log
lap "smoke"
if [ "$N" = 1 ]; then
  timeout 600 python bench.py --impl reference > $O/verify_bench_reference.json 2> $O/verify_bench_reference.err; cut -c1-300 $O/verify_bench_reference.json
  lap "reference arm"
  timeout 600 python bench.py > $O/verify_bench_default.json 2> $O/verify_bench_default.err; cut -c1-420 $O/verify_bench_default.json; tail -2 $O/verify_bench_default.err
  lap "bench default"
  MOLDYN_B200_PDL=1 timeout 300 python bench.py --e2e-steps 0 --cpu-rows -1 > $O/verify_bench_pdl1.json 2>&1; cut -c1-200 $O/verify_bench_pdl1.json
  MOLDYN_B200_PDL=2 timeout 300 python bench.py --e2e-steps 0 --cpu-rows -1 > $O/verify_bench_pdl2.json 2>&1; cut -c1-200 $O/verify_bench_pdl2.json
  lap "bench pdl"
  MOLDYN_B200_LOOP=host timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r01_launches_c3_v11.csv \
    python bench.py --steps 60 --warmup 3 --e2e-steps 1 --cpu-rows -1 > $O/verify_launches.log 2>&1
  lap "launch list"
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --impl reference --steps 3 --warmup 1 > $O/verify_ref_$N.json 2> $O/verify_ref_$N.err; cut -c1-300 $O/verify_ref_$N.json
  lap "reference arm x$N"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N > $O/verify_ours_$N.json 2> $O/verify_ours_$N.err; grep '^{' $O/verify_ours_$N.json | cut -c1-600; tail -2 $O/verify_ours_$N.err
  lap "bench x$N"
fi
