#!/bin/bash
O=gpurun_out/r2c38
mkdir -p $O
BELLMAN_B200_WATCHDOG_S=2 timeout 2400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "all: pytest rc=$? $(tail -1 $O/pytest_gpu.log)"
