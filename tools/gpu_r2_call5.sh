#!/bin/bash
O=gpurun_out/r2c5
mkdir -p $O
timeout 600 python tools/prune_check.py 120 25 > $O/prune_check_25.txt 2>&1
timeout 600 python tools/prune_check.py 120 26 > $O/prune_check_26.txt 2>&1
timeout 300 python tools/phase_profile.py 20000 0 0 25 > $O/phase_profile_v25.txt 2>&1
timeout 300 python tools/phase_profile.py 20000 0 0 26 > $O/phase_profile_v26.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
cat $O/prune_check_25.txt; tail -2 $O/prune_check_26.txt; head -4 $O/phase_profile_v25.txt; tail -1 $O/phase_profile_v25.txt; head -1 $O/phase_profile_v26.txt; tail -1 $O/phase_profile_v26.txt; tail -5 $O/pytest_gpu.log
