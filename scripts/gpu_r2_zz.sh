#!/bin/bash
# Round 2, last call: multi-type parity (incl. the list-radius test) and the fast single-type subset on the shipped library.
timeout 120 python -m pytest tests/test_gpu_multi_type.py tests/test_cli.py -x -q -m gpu 2>&1 | tail -3
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or loop_drivers or update_force or determinism" 2>&1 | tail -2
