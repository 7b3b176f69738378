#!/bin/bash
# round 2, call 4: two-pass pruned scan (parity + speed), new bench legs (verify, batched), batched pipeline tests
O=gpurun_out/r2c4
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
for v in 25 26 27; do timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-batched --no-verify --variant $v > $O/bench_v$v.json 2> $O/bench_v$v.err; done
timeout 300 python tools/phase_profile.py 100000 0 0 25 > $O/phase_profile_v25.txt 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 --variant 25 > $O/bench_full_v25.json 2> $O/bench_full_v25.err
timeout 900 python bench.py --steps 3 --warmup 3 > $O/bench_full_auto.json 2> $O/bench_full_auto.err
tail -4 $O/pytest_gpu.log; for v in 25 26 27; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench_v$v.json").read()); r=d["roofline"]
    print("variant $v: value %.3e ms %.1f kernel_ms %.1f frac %.3f executed_frac %.3f" % (d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], r["executed_frac"]))
except Exception as e:
    print("variant $v: failed", e); print(open("$O/bench_v$v.err").read()[-800:])
PY
done
head -3 $O/phase_profile_v25.txt; tail -1 $O/phase_profile_v25.txt
for f in bench_full_v25 bench_full_auto; do python - <<PY
import json
try:
    d=json.loads(open("$O/$f.json").read())
    print("$f: value %.3e e2e %.3e verified %s frac %.3f exec %.3f" % (d["value"], d["e2e"]["value"], d["verified"], d["roofline"]["frac"], d["roofline"]["executed_frac"]))
    print("   batched:", {k: d["batched"][k] for k in ("value","wall_ms","device_ms_max","waves_per_gpu","host_waits_per_wave","best")})
except Exception as e:
    print("$f failed:", e); print(open("$O/$f.err").read()[-1500:])
PY
done
