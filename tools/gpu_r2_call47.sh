#!/bin/bash
# compile-time level count for the heat-shaped direct tile: parity (example shapes, heat n = 8192) + latencies
O=gpurun_out/r2c48
mkdir -p $O
BELLMAN_B200_WATCHDOG_S=2 timeout 1200 python -m pytest tests -m gpu -x -q -k "example_shapes or heat or trm or geometries" > $O/pytest_sub.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest_sub.log)"
timeout 300 python tools/example_latency.py > $O/example_latency.txt 2>&1; grep heat $O/example_latency.txt
