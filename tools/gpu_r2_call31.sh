#!/bin/bash
# compute-sanitizer memcheck of the production geometry on a small config-4-shaped instance (plain run first)
O=gpurun_out/r2c31
mkdir -p $O
BELLMAN_B200_WATCHDOG_S=0 timeout 300 python tools/sanitize_case.py 40 > $O/plain.log 2>&1 && \
BELLMAN_B200_WATCHDOG_S=0 timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_case.py 40 > $O/memcheck.log 2>&1
echo "rc=$?"; tail -5 $O/plain.log; grep -c "Invalid\|Error" $O/memcheck.log; tail -12 $O/memcheck.log
