#!/bin/bash
# 8-GPU check of the torchrun bench line (one subproblem per rank + the config-5 batch sharded s mod 8)
O=gpurun_out/r2c52
mkdir -p $O
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 2 --warmup 1 --no-cpu > $O/bench_8gpu.json 2> $O/bench_8gpu.err
python - <<PY
import json
try:
    d=json.loads(open("$O/bench_8gpu.json").read().strip().splitlines()[-1])
    print("n_gpus", d["n_gpus"], "value %.4e ms/step %.1f" % (d["value"], d["ms_per_step"]), "e2e %.4e" % d["e2e"]["value"], "verified", d.get("verified"))
    print("   batched", {k: d["batched"][k] for k in ("value","wall_ms","device_ms_max","waves_per_gpu","host_waits_per_wave","collective")})
except Exception as e:
    print("failed", e); print(open("$O/bench_8gpu.err").read()[-2500:])
PY
