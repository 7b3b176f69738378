#!/bin/bash
# row-group pipeline with an initial stagger: A/B
O=gpurun_out/r2c21
mkdir -p $O
for v in default pipe stag600 stag1200; do
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
