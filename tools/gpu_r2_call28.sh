#!/bin/bash
O=gpurun_out/r2c28
mkdir -p $O
timeout 600 python tools/wide_k_probe.py > $O/wide_k_probe.txt 2>&1
cat $O/wide_k_probe.txt | tail -8
