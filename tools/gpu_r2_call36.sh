#!/bin/bash
# stress: the GPU parity suite three times, the headline bench with --verify at several horizons, smoke()
O=gpurun_out/r2c36
mkdir -p $O
for k in 1 2 3; do
  timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/pytest_$k.log 2>&1; echo "run $k: pytest rc=$? $(tail -1 $O/pytest_$k.log)"
done
for n in 100000 99991 50021 33333 7919; do
  timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-batched --n $n > $O/bench_$n.json 2> $O/bench_$n.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$n.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("n=$n: kernel_ms %.2f frac %.3f executed_frac %.3f verified %s variant-threads %s" % (r["kernel_ms"], r["frac"], r["executed_frac"], d["verified"], d["config"]["threads_per_cta"]))
except Exception as e:
    print("n=$n: failed", e); print(open("$O/bench_$n.err").read()[-600:])
PY
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
