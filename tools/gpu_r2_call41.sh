#!/bin/bash
O=gpurun_out/r2c41
mkdir -p $O
BELLMAN_B200_WATCHDOG_S=2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shortest_horizons" > $O/pytest_short.log 2>&1; echo "short: pytest rc=$? $(tail -1 $O/pytest_short.log)"
grep -E "^E " $O/pytest_short.log | head -20
