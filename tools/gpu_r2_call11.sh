#!/bin/bash
O=gpurun_out/r2c11
mkdir -p $O
timeout 300 python tools/phase_profile.py 100000 > $O/phase_profile.txt 2>&1
timeout 300 python tools/phase_profile.py 100000 0 0 25 > $O/phase_profile_v25.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
for v in 0 25; do timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-batched --variant $v > $O/bench_v$v.json 2> $O/bench_v$v.err; done
for f in phase_profile phase_profile_v25; do head -1 $O/$f.txt | cut -c1-200; tail -1 $O/$f.txt | cut -c1-330; done
tail -4 $O/pytest_gpu.log
for v in 0 25; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench_v$v.json").read()); r=d["roofline"]
    print("variant $v: value %.3e ms %.1f kernel_ms %.1f frac %.3f executed_frac %.3f verified %s e2e %.3e" % (d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], r["executed_frac"], d["verified"], d["e2e"]["value"]))
except Exception as e:
    print("variant $v: failed", e); print(open("$O/bench_v$v.err").read()[-800:])
PY
done
