#!/bin/bash
O=gpurun_out/r2c6
mkdir -p $O
timeout 300 python tools/phase_profile.py 20000 0 0 25 > $O/phase_profile_v25.txt 2>&1
timeout 900 python tools/prune_check.py 500 25 > $O/prune_check_25_n500.txt 2>&1
head -4 $O/phase_profile_v25.txt | cut -c1-400; tail -1 $O/phase_profile_v25.txt; grep -v "stage" $O/prune_check_25_n500.txt
