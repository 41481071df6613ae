#!/bin/bash
# 8 GPUs: scaling bench at N = 8 and 4 (driver-style invocation), multirank parity tests
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $O/r2c10_smi.txt 2>&1
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2971$n bench.py --gpus $n --steps 20 --warmup 5 > $O/r2c10_bench_${n}gpu.json 2> $O/r2c10_bench_${n}gpu.err
python - <<PY
import json
try:
    j=json.loads(open("$O/r2c10_bench_${n}gpu.json").read().strip().splitlines()[-1])
    print("N=$n", round(j["value"]), "e2e", round(j["e2e"]["value"]), j["config"]["workload"], "per_rank", j.get("per_rank"))
    for k,v in j["extra"].items(): print("   extra", k, json.dumps(v)[:500])
except Exception as e:
    print("N=$n failed", e); print(open("$O/r2c10_bench_${n}gpu.err").read()[-1500:])
PY
done
timeout 900 python -m pytest tests/test_gpu_multirank.py -q > $O/r2c10_multirank.log 2>&1; tail -3 $O/r2c10_multirank.log
