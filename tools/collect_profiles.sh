#!/bin/bash
# One GPU-box session that regenerates the non-ncu evidence under profiles/ (copy from gpurun_out/ afterwards).
# Usage on the box: bash tools/collect_profiles.sh
set -u
O=gpurun_out/collect
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,driver_version --format=csv > $O/gpu.txt 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --steps 3 --warmup 3 > $O/bench.json 2> $O/bench.err
{ for n in 3000 20000 100000; do python tools/phase_profile.py $n; done
  python tools/phase_profile.py 8192 0 0 0 heat; python tools/phase_profile.py 1024 0 0 0 heat; } > $O/phase_profile.txt 2>&1
BELLMAN_B200_DECOUPLE=1 python tools/phase_profile.py 20000 > $O/phase_profile_decoupled.txt 2>&1
python tools/tune_sweep.py 3000 synthetic > $O/tune_sweep.txt 2>&1
python tools/tune_sweep.py 8192 heat > $O/tune_sweep_heat8192.txt 2>&1
python tools/example_latency.py > $O/example_latency.txt 2>&1
./tools/pipe_probe.bin > $O/pipe_probe.txt 2>&1
{ ./tools/sm_speed_probe.bin; ./tools/sm_speed_probe.bin lds; } > $O/sm_speed_probe.txt 2>&1
./tools/blocked_probe.bin > $O/blocked_probe.txt 2>&1
tail -n 2 $O/bench.json | cut -c1-300
