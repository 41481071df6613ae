#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29748 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2c18_bench_8gpu.json 2> $O/r2c18_bench_8gpu.err
python - <<PY
import json
try:
    j=json.loads(open("$O/r2c18_bench_8gpu.json").read().strip().splitlines()[-1])
    print("N=8", round(j["value"]), "e2e", round(j["e2e"]["value"]), j["config"]["workload"], "per_rank", j.get("per_rank"))
    for k,v in j["extra"].items(): print("   extra", k, json.dumps(v)[:300])
except Exception as e:
    print("N=8 failed", e); print(open("$O/r2c18_bench_8gpu.err").read()[-1500:])
PY
