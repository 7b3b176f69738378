#!/bin/bash
# two-zone slices (upper CTAs own one row group less, all 148 SMs): full GPU parity suite, then A/B against uniform slices
O=gpurun_out/r2c22
mkdir -p $O
BELLMAN_B200_WATCHDOG_S=2 timeout 2400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
for v in default uniform; do
  if [ $v = default ]; then unset BELLMAN_B200_LIB; else export BELLMAN_B200_LIB=$PWD/build/libbb_$v.so; fi
  BELLMAN_B200_WATCHDOG_S=1 timeout 120 python bench.py --steps 3 --warmup 2 --no-cpu --no-batched > $O/bench_$v.json 2> $O/bench_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$v.json").read().strip().splitlines()[-1]); r=d["roofline"]; c=d["config"]
    print("$v: value %.3e ms %.1f kernel_ms %.1f frac %.3f executed_frac %.3f verified %s ctas %s rows %s full %s top %s" % (d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], r["executed_frac"], d["verified"], c["ctas"], c["rows_per_cta"], c.get("ctas_with_full_rows"), c.get("rows_per_cta_upper_zone")))
except Exception as e:
    print("$v: failed", e); print(open("$O/bench_$v.err").read()[-800:])
PY
done
unset BELLMAN_B200_LIB
PROFILE_DUMP=$O/prof_all.npy timeout 300 python tools/phase_profile.py 100000 > $O/phase_profile.txt 2>&1
head -1 $O/phase_profile.txt | cut -c1-200
