#!/bin/bash
O=gpurun_out/r2c32
mkdir -p $O
BELLMAN_B200_WATCHDOG_S=2 timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "extreme or pruned_scan_is_exact" > $O/pytest_sub.log 2>&1; echo "pytest rc=$?" >> $O/pytest_sub.log
tail -15 $O/pytest_sub.log
