#!/bin/bash
# per-CTA phase profile dump (who is the slowest slice?) + ncu --set full of the FP32-bound kernel at n = 20000
O=gpurun_out/r2c19
mkdir -p $O
PROFILE_DUMP=$O/prof_all.npy timeout 300 python tools/phase_profile.py 100000 > $O/phase_profile.txt 2>&1
head -1 $O/phase_profile.txt | cut -c1-200
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --no-batched --no-verify --n 20000 > $O/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:wavefront -c 1 -o $O/wave_fp32 python bench.py --steps 1 --warmup 1 --no-cpu --no-batched --no-verify --n 20000 > $O/ncu.log 2>&1
tail -3 $O/ncu.log
ls -la $O
