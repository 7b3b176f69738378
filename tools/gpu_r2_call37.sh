#!/bin/bash
# per-stage pruned kernels for wide level sets + shared pruned_scan.cuh: new tests first, then the whole suite, bench, wide-K probe
O=gpurun_out/r2c37
mkdir -p $O
BELLMAN_B200_WATCHDOG_S=2 timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wide_level or uint16" > $O/pytest_wide.log 2>&1; echo "wide: pytest rc=$? $(tail -1 $O/pytest_wide.log)"
grep -E "^E " $O/pytest_wide.log | head -20
BELLMAN_B200_WATCHDOG_S=2 timeout 2400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "all: pytest rc=$? $(tail -1 $O/pytest_gpu.log)"
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-batched > $O/bench.json 2> $O/bench.err
python - <<PY
import json
try:
    d=json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
    print("bench value %.4e ms/step %.1f" % (d["value"], d["ms_per_step"]), "verified", d.get("verified"), "kernel_ms %.1f" % d["roofline"]["kernel_ms"])
except Exception as e:
    print("failed", e); print(open("$O/bench.err").read()[-1500:])
PY
timeout 600 python tools/wide_k_probe.py > $O/wide_k_probe.txt 2>&1; tail -6 $O/wide_k_probe.txt
