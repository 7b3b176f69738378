#!/bin/bash
# round 2, call 10: full validation of the default configuration (pruned tiles by default) + launch list
O=gpurun_out/r2c10
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,driver_version --format=csv > $O/gpu.txt 2>&1
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 2400 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 300 python tools/phase_profile.py 100000 > $O/phase_profile.txt 2>&1
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-batched --no-verify > $O/plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-batched --no-verify > $O/ncu_launches.log 2>&1
tail -3 $O/smoke.log; tail -6 $O/pytest_gpu.log
python - <<PY
import json
for f in ("bench", "bench_reference"):
    try:
        d=json.loads(open("$O/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.4e ms/step %.1f" % (d["value"], d["ms_per_step"]), "e2e %.4e" % d["e2e"]["value"], "verified", d.get("verified"))
        if "roofline" in d and d["roofline"]: print("   roofline", {k: d["roofline"][k] for k in ("achieved","peak","frac","executed_frac","kernel_ms","traffic")})
        if d.get("batched"): print("   batched", {k: d["batched"][k] for k in ("value","wall_ms","device_ms_max","waves_per_gpu","host_waits_per_wave")})
        if d.get("cpu_baseline"): print("   cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
        print("   clocks", d.get("clocks"))
    except Exception as e:
        print(f, "failed", e); print(open("$O/%s.err" % f).read()[-1500:])
PY
head -3 $O/phase_profile.txt | cut -c1-300; tail -1 $O/phase_profile.txt | cut -c1-330
grep -c wavefront $O/launches.csv
