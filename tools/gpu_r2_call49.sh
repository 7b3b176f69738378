#!/bin/bash
# no "scanned" hand-over for the pruned tiles (three buffers), straight-line level test with paired block minima: parity + bench
O=gpurun_out/r2c49
mkdir -p $O
BELLMAN_B200_WATCHDOG_S=2 timeout 2400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-batched > $O/bench.json 2> $O/bench.err
python - <<PY
import json
try:
    d=json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
    print("bench value %.4e ms/step %.1f" % (d["value"], d["ms_per_step"]), "e2e %.4e" % d["e2e"]["value"], "verified", d.get("verified"))
    print("   roofline", {k: d["roofline"].get(k) for k in ("achieved","peak","frac","executed_frac","kernel_ms","traffic")})
except Exception as e:
    print("failed", e); print(open("$O/bench.err").read()[-1500:])
PY
