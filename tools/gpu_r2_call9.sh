#!/bin/bash
O=gpurun_out/r2c9
mkdir -p $O
timeout 300 python tools/phase_profile.py 20000 0 0 25 > $O/phase_profile_v25.txt 2>&1
timeout 300 python tools/phase_profile.py 20000 0 0 26 > $O/phase_profile_v26.txt 2>&1
timeout 300 python tools/phase_profile.py 20000 0 0 28 > $O/phase_profile_v28.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
for v in 25 26; do timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-batched --variant $v > $O/bench_v$v.json 2> $O/bench_v$v.err; done
for v in 25 26 28; do head -2 $O/phase_profile_v$v.txt | cut -c1-200 | grep -v "^pruned"; grep "executed fraction" $O/phase_profile_v$v.txt | sed 's/.*executed fraction/executed fraction/'; tail -1 $O/phase_profile_v$v.txt | cut -c1-330; done
tail -4 $O/pytest_gpu.log
for v in 25 26; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench_v$v.json").read()); r=d["roofline"]
    print("variant $v: value %.3e ms %.1f kernel_ms %.1f frac %.3f executed_frac %.3f verified %s e2e %.3e" % (d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], r["executed_frac"], d["verified"], d["e2e"]["value"]))
except Exception as e:
    print("variant $v: failed", e); print(open("$O/bench_v$v.err").read()[-800:])
PY
done
