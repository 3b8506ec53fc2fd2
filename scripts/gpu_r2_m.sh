#!/bin/bash
# Round 2, call M: the multi-type path (parity tests), then the fast part of the single-type suite as a regression check.
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi_type.py -x -q -m gpu > $O/m_multi.log 2>&1; echo "multi rc=$?"; tail -25 $O/m_multi.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or update_force or loop_drivers or nve or initialize" > $O/m_parity.log 2>&1; echo "parity rc=$?"; tail -5 $O/m_parity.log
