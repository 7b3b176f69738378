#!/bin/bash
# soak: 60 timed steps of the headline DP (+ verify), the config-5 batch, then the whole GPU suite on the final library
O=gpurun_out/r2c46
mkdir -p $O
timeout 600 python bench.py --steps 60 --warmup 3 --no-cpu > $O/bench_soak.json 2> $O/bench_soak.err
python - <<PY
import json
try:
    d=json.loads(open("$O/bench_soak.json").read().strip().splitlines()[-1])
    print("soak: steps", d["steps"], "value %.4e ms/step %.2f" % (d["value"], d["ms_per_step"]), "verified", d.get("verified"), "clocks", d.get("clocks"))
    print("   batched %.4e" % d["batched"]["value"], d["batched"]["wall_ms"])
except Exception as e:
    print("failed", e); print(open("$O/bench_soak.err").read()[-1500:])
PY
timeout 2400 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest_gpu.log)"
