#!/bin/bash
# round-2 final evidence: bench lines (ours + reference arm), launch list, ncu --set full, DRAM traffic, phase profiles, example latencies
O=gpurun_out/r2c51
mkdir -p $O
BELLMAN_B200_WATCHDOG_S=2 timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2c51/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c51/pytest_gpu.log; tail -3 gpurun_out/r2c51/pytest_gpu.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,driver_version --format=csv > $O/gpu.txt 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-batched --variant -1 > $O/bench_exhaustive.json 2> $O/bench_exhaustive.err
{ for n in 3000 20000 100000; do timeout 300 python tools/phase_profile.py $n; done; } > $O/phase_profile.txt 2>&1
timeout 300 python tools/example_latency.py > $O/example_latency.txt 2>&1
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-batched --no-verify > $O/plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-batched --no-verify > $O/ncu_launches.log 2>&1
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --no-batched --no-verify --n 20000 > $O/plain_n20000.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:wavefront -c 1 -o $O/wavefront_final python bench.py --steps 1 --warmup 1 --no-cpu --no-batched --no-verify --n 20000 > $O/ncu_full.log 2>&1
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --no-batched --no-verify > $O/plain_full.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:wavefront -c 1 --csv --log-file $O/wavefront_dram_n100000.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-batched --no-verify > $O/ncu_dram.log 2>&1
python - <<PY
import json
for f in ("bench", "bench_reference", "bench_exhaustive"):
    try:
        d=json.loads(open("$O/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.4e ms/step %.1f" % (d["value"], d["ms_per_step"]), "e2e %.4e" % d["e2e"]["value"], "verified", d.get("verified"))
        if d.get("roofline"): print("   roofline", {k: d["roofline"].get(k) for k in ("achieved","peak","frac","executed_frac","kernel_ms","traffic")})
        if d.get("batched"): print("   batched", {k: d["batched"][k] for k in ("value","wall_ms","device_ms_max","waves_per_gpu","host_waits_per_wave")})
        if d.get("cpu_baseline"): print("   cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "clocks", d.get("clocks"))
    except Exception as e:
        print(f, "failed", e); print(open("$O/%s.err" % f).read()[-1500:])
PY
grep -v "^==" $O/wavefront_dram_n100000.csv | tail -4 | cut -c1-300
grep "us/stage" $O/phase_profile.txt | cut -c1-200
cat $O/example_latency.txt | tail -12
