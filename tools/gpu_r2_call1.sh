#!/bin/bash
# round 2, call 1: julia probe, scan-update probes, parity tests, short bench, in-kernel profile
O=gpurun_out/r2c1
mkdir -p $O
{ which julia; julia --version; ls /opt /usr/local | head -50; nproc; free -g | head -2; } > $O/env.txt 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,driver_version --format=csv > $O/gpu.txt 2>&1
./tools/scan_probe.bin > $O/scan_probe.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > $O/bench.json 2> $O/bench.err
timeout 300 python tools/phase_profile.py 100000 > $O/phase_profile.txt 2>&1
timeout 300 python tools/tune_sweep.py 3000 synthetic > $O/tune_sweep.txt 2>&1
cat $O/scan_probe.txt; tail -3 $O/pytest_gpu.log; cut -c1-600 $O/bench.json; head -3 $O/phase_profile.txt; tail -2 $O/phase_profile.txt; cat $O/env.txt | head -5
