#!/bin/bash
O=gpurun_out/r2c8
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --no-batched --no-verify --n 20000 --variant 25 > $O/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:wavefront -c 1 -o $O/wavefront_pruned python bench.py --steps 1 --warmup 1 --no-cpu --no-batched --no-verify --n 20000 --variant 25 > $O/ncu.log 2>&1
tail -5 $O/pytest_gpu.log; tail -3 $O/ncu.log
