#!/bin/bash
# round 2, call 2: issue-slot probe, lane-split tiles (parity + speed), ncu of the current production kernel
O=gpurun_out/r2c2
mkdir -p $O
./tools/issue_probe.bin > $O/issue_probe.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python tools/tune_sweep.py 3000 synthetic lane > $O/tune_sweep_lane.txt 2>&1
for v in 25 29 30 26 28; do timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --variant $v 2> $O/bench_v$v.err | cut -c1-260 > $O/bench_v$v.json; done
timeout 300 python tools/phase_profile.py 100000 0 0 25 > $O/phase_profile_v25.txt 2>&1
timeout 300 python tools/phase_profile.py 100000 0 0 29 > $O/phase_profile_v29.txt 2>&1
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --n 20000 > $O/plain_n20000.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:wavefront -c 1 -o $O/wavefront_predmoves python bench.py --steps 1 --warmup 1 --no-cpu --n 20000 > $O/ncu.log 2>&1
cat $O/issue_probe.txt; tail -3 $O/pytest_gpu.log; cat $O/tune_sweep_lane.txt | tail -14; for v in 25 29 30 26 28; do cat $O/bench_v$v.json; echo; done; head -2 $O/phase_profile_v25.txt; tail -1 $O/phase_profile_v25.txt; head -2 $O/phase_profile_v29.txt; tail -1 $O/phase_profile_v29.txt; tail -3 $O/ncu.log
