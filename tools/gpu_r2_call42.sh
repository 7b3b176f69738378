#!/bin/bash
# heat-shaped hand-over wait: split into halo / ring, all slices; poll-period A/B
O=gpurun_out/r2c42
mkdir -p $O
PROFILE_DUMP=$O/prof_heat8192.npy timeout 300 python tools/phase_profile.py 8192 0 0 0 heat > $O/phase_heat8192.txt 2>&1
PROFILE_DUMP=$O/prof_heat1024.npy timeout 300 python tools/phase_profile.py 1024 0 0 0 heat > $O/phase_heat1024.txt 2>&1
head -1 $O/phase_heat8192.txt | cut -c1-200; tail -1 $O/phase_heat8192.txt | cut -c1-250
head -1 $O/phase_heat1024.txt | cut -c1-200; tail -1 $O/phase_heat1024.txt | cut -c1-250
timeout 300 python tools/example_latency.py > $O/lat_default.txt 2>&1; grep heat $O/lat_default.txt
BELLMAN_B200_LIB=$PWD/build/libbb_fastpoll.so timeout 300 python tools/example_latency.py > $O/lat_fastpoll.txt 2>&1; grep heat $O/lat_fastpoll.txt
