#!/bin/bash
# compile-time level count for the exhaustive production tile: parity of the exhaustive geometries + bench --variant -1
O=gpurun_out/r2c50
mkdir -p $O
BELLMAN_B200_WATCHDOG_S=2 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -x -q -k "config4 or geometries or full_size or partially or pruned_scan_is_exact" > $O/pytest_sub.log 2>&1; echo "pytest rc=$?" >> $O/pytest_sub.log
tail -3 $O/pytest_sub.log
for var in -1 0; do
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-batched --variant $var > $O/bench_$var.json 2> $O/bench_$var.err
python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$var.json").read().strip().splitlines()[-1])
    print("variant $var: value %.4e ms/step %.1f" % (d["value"], d["ms_per_step"]), "verified", d.get("verified"), "kernel_ms %.1f frac %.3f" % (d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
except Exception as e:
    print("failed", e); print(open("$O/bench_$var.err").read()[-1500:])
PY
done
