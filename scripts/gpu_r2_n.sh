#!/bin/bash
# Round 2, call N: CLI tests (two particle types through moldyn_cli) and the multi-type parity tests.
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_cli.py tests/test_gpu_multi_type.py -x -q -m gpu > $O/n_cli_multi.log 2>&1; echo "rc=$?"; tail -25 $O/n_cli_multi.log
