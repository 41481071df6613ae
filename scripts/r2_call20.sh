#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29751 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2c20_bench_2gpu.json 2> $O/r2c20_bench_2gpu.err
python - <<PY
import json
try:
    j=json.loads(open("$O/r2c20_bench_2gpu.json").read().strip().splitlines()[-1])
    print("N=2", round(j["value"]), "e2e", round(j["e2e"]["value"]), j["config"]["workload"], "per_rank", j.get("per_rank"))
    for k,v in j["extra"].items(): print("   extra", k, json.dumps(v)[:260])
except Exception as e:
    print("N=2 failed", e); print(open("$O/r2c20_bench_2gpu.err").read()[-1500:])
PY
