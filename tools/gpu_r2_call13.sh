#!/bin/bash
O=gpurun_out/r2c13
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-batched > $O/bench_auto.json 2> $O/bench_auto.err
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-batched --variant -1 > $O/bench_exhaustive.json 2> $O/bench_exhaustive.err
timeout 300 python tools/phase_profile.py 100000 > $O/phase_profile.txt 2>&1
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --no-batched --no-verify --n 20000 > $O/plain_n20000.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:wavefront -c 1 -o $O/wavefront_pruned_final python bench.py --steps 1 --warmup 1 --no-cpu --no-batched --no-verify --n 20000 > $O/ncu_full.log 2>&1
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --no-batched --no-verify > $O/plain_full.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:wavefront -c 1 --csv --log-file $O/wavefront_dram_n100000.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-batched --no-verify > $O/ncu_dram.log 2>&1
tail -4 $O/pytest_gpu.log
for v in auto exhaustive; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$v.json").read()); r=d["roofline"]
    print("$v: value %.3e ms %.1f kernel_ms %.1f frac %.3f executed_frac %.3f verified %s e2e %.3e" % (d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], r["executed_frac"], d["verified"], d["e2e"]["value"]))
except Exception as e:
    print("$v: failed", e); print(open("$O/bench_$v.err").read()[-800:])
PY
done
head -1 $O/phase_profile.txt | cut -c1-200; tail -1 $O/phase_profile.txt | cut -c1-330
grep -v "^==" $O/wavefront_dram_n100000.csv | tail -4 | cut -c1-400
