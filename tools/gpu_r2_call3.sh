#!/bin/bash
# round 2, call 3: pruned (branch-and-bound) scan: parity + speed
O=gpurun_out/r2c3
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
for v in 25 26 27; do timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --variant $v > $O/bench_v$v.json 2> $O/bench_v$v.err; done
timeout 300 python tools/phase_profile.py 100000 0 0 25 > $O/phase_profile_v25.txt 2>&1
timeout 300 python tools/phase_profile.py 100000 0 0 27 > $O/phase_profile_v27.txt 2>&1
tail -4 $O/pytest_gpu.log; for v in 25 26 27; do cut -c1-330 $O/bench_v$v.json; tail -2 $O/bench_v$v.err; done; head -3 $O/phase_profile_v25.txt; tail -1 $O/phase_profile_v25.txt; head -1 $O/phase_profile_v27.txt; tail -1 $O/phase_profile_v27.txt
