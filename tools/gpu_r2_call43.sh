#!/bin/bash
# final check as the driver runs it: GPU suite, smoke(), default bench, reference arm
O=gpurun_out/r2c43
mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
S0=$SECONDS; timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "default bench: $((SECONDS-S0)) s wall"
S0=$SECONDS; timeout 900 python bench.py --impl reference > $O/bench_ref_default.json 2> $O/bench_ref_default.err; echo "reference arm: $((SECONDS-S0)) s wall"
python - <<PY
import json
for f in ("bench_default", "bench_ref_default"):
    d=json.loads(open("$O/%s.json" % f).read().strip().splitlines()[-1])
    print(f, {k: d.get(k) for k in ("metric","value","unit","n_gpus","steps","warmup","ms_per_step","higher_is_better","scaling","vs_baseline","dtype","data","gpu_launches","verified","impl")})
    print("   keys:", sorted(d.keys()))
    if "e2e" in d: print("   e2e", d["e2e"])
    if d.get("roofline"): print("   roofline frac", d["roofline"]["frac"], "bound", d["roofline"]["bound"], "traffic", d["roofline"]["traffic"])
PY
