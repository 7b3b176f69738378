#!/bin/bash
# break-even of the pruned tiles against the exhaustive tiles as a function of the executed fraction (horizon sweep)
O=gpurun_out/r2c24
mkdir -p $O
for n in 2500 5000 10000 20000; do
 for var in 28 -1; do
  timeout 120 python bench.py --steps 3 --warmup 2 --no-cpu --no-batched --no-verify --n $n --variant $var > $O/b_${n}_$var.json 2> $O/b_${n}_$var.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/b_${n}_$var.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("n=$n variant=$var: kernel_ms %.2f us/stage %.3f executed_frac %.3f frac %.3f ctas %s" % (r["kernel_ms"], r["kernel_ms"]*1e3/($n-1), r["executed_frac"], r["frac"], d["config"]["ctas"]))
except Exception as e:
    print("n=$n variant=$var: failed", e); print(open("$O/b_${n}_$var.err").read()[-600:])
PY
 done
done
