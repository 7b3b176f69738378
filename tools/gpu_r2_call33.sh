#!/bin/bash
# branch-free fast path of the backtrack chase: full parity suite + bench (backtrack time from the plan's stats)
O=gpurun_out/r2c33
mkdir -p $O
BELLMAN_B200_WATCHDOG_S=2 timeout 2400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-batched > $O/bench.json 2> $O/bench.err
python - <<PY
import json
try:
    d=json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
    print("bench value %.4e ms/step %.1f" % (d["value"], d["ms_per_step"]), "e2e %.4e" % d["e2e"]["value"], "verified", d.get("verified"), "kernel_ms %.1f" % d["roofline"]["kernel_ms"])
except Exception as e:
    print("failed", e); print(open("$O/bench.err").read()[-1500:])
PY
timeout 120 python tools/component_latency.py > $O/component_latency.txt 2>&1; tail -8 $O/component_latency.txt
