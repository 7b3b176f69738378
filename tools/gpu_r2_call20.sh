#!/bin/bash
# UB from the seed successor + compile-time level count: parity of the pruned path, A/B against the runtime-Kp instantiation
O=gpurun_out/r2c20
mkdir -p $O
BELLMAN_B200_WATCHDOG_S=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -x -q -k "pruned or config4 or full_size or batched or resident or partially" > $O/pytest_sub.log 2>&1; echo "pytest rc=$?" >> $O/pytest_sub.log
tail -3 $O/pytest_sub.log
for v in default nokpc; do
  if [ $v = default ]; then unset BELLMAN_B200_LIB; else export BELLMAN_B200_LIB=$PWD/build/libbb_$v.so; fi
  BELLMAN_B200_WATCHDOG_S=1 timeout 120 python bench.py --steps 3 --warmup 2 --no-cpu --no-batched > $O/bench_$v.json 2> $O/bench_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$v.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("$v: value %.3e ms %.1f kernel_ms %.1f frac %.3f executed_frac %.3f verified %s threads %s" % (d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], r["executed_frac"], d["verified"], d["config"].get("threads_per_cta")))
except Exception as e:
    print("$v: failed", e); print(open("$O/bench_$v.err").read()[-800:])
PY
done
