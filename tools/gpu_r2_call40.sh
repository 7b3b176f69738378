#!/bin/bash
# 2-GPU call: multi-GPU behind the C ABI (NCCL inside the library) + the torchrun bench line
O=gpurun_out/r2c40
mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_gpu or batched" > $O/pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > $O/bench_2gpu.json 2> $O/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_ref_2gpu.json 2> $O/bench_ref_2gpu.err
cat $O/gpus.txt; tail -4 $O/pytest_multi.log
python - <<PY
import json
for f in ("bench_2gpu", "bench_ref_2gpu"):
    try:
        d=json.loads(open("$O/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "n_gpus", d["n_gpus"], "value %.4e ms/step %.1f" % (d["value"], d["ms_per_step"]), "e2e %.4e" % d["e2e"]["value"], "verified", d.get("verified"))
        if d.get("batched"): print("   batched", {k: d["batched"][k] for k in ("value","wall_ms","device_ms_max","waves_per_gpu","host_waits_per_wave","best","collective")})
    except Exception as e:
        print(f, "failed", e); print(open("$O/%s.err" % f).read()[-2500:])
PY
