#!/bin/bash
O=gpurun_out/r2c45
mkdir -p $O
BELLMAN_B200_WATCHDOG_S=2 timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wide_level or uint16" > $O/pytest_wide.log 2>&1; echo "wide: pytest rc=$? $(tail -1 $O/pytest_wide.log)"
timeout 600 python tools/wide_k_probe.py > $O/wide_k_probe.txt 2>&1; tail -5 $O/wide_k_probe.txt
