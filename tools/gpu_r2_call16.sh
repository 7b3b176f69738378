#!/bin/bash
# FP32 bound tests + three value-row buffers + barrier hand-over: parity of the pruned path, then A/B of the four builds
O=gpurun_out/r2c16
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -x -q -k "pruned or config4 or full_size or batched or resident" > $O/pytest_sub.log 2>&1; echo "pytest rc=$?" >> $O/pytest_sub.log
tail -4 $O/pytest_sub.log
for v in default b2bar b3mb b2mb; do
  if [ $v = default ]; then unset BELLMAN_B200_LIB; else export BELLMAN_B200_LIB=$PWD/build/libbb_$v.so; fi
  timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-batched > $O/bench_$v.json 2> $O/bench_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$v.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("$v: value %.3e ms %.1f kernel_ms %.1f frac %.3f executed_frac %.3f verified %s variant-cfg %s" % (d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], r["executed_frac"], d["verified"], d["config"].get("threads_per_cta")))
except Exception as e:
    print("$v: failed", e); print(open("$O/bench_$v.err").read()[-800:])
PY
done
unset BELLMAN_B200_LIB
for var in 24 26 27; do
  timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu --no-batched --no-verify --variant $((var+1)) > $O/bench_var$var.json 2> $O/bench_var$var.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/bench_var$var.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("variant $var: value %.3e ms %.1f kernel_ms %.1f frac %.3f executed_frac %.3f threads %s" % (d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], r["executed_frac"], d["config"].get("threads_per_cta")))
except Exception as e:
    print("variant $var: failed", e); print(open("$O/bench_var$var.err").read()[-800:])
PY
done
timeout 300 python tools/phase_profile.py 100000 > $O/phase_profile.txt 2>&1
head -1 $O/phase_profile.txt | cut -c1-200; tail -1 $O/phase_profile.txt | cut -c1-330
