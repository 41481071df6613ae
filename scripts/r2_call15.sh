#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/r2c15_pytest.log 2>&1; echo "rc=$?" >> $O/r2c15_pytest.log
tail -12 $O/r2c15_pytest.log
timeout 300 python bench.py --workload periodic_bloch --steps 200 --warmup 10 --no-cpu --no-extra > $O/r2c15_bloch.json 2> $O/r2c15_bloch.err
python - <<PY
import json
try:
    j=json.loads(open("$O/r2c15_bloch.json").read().strip().splitlines()[-1])
    print("bloch", round(j["value"]), "e2e", round(j["e2e"]["value"]), j["config"]["grid"], [(k["name"][:40], k["ctas"], round(k["total_ms"]/max(1,k["launches"]),4)) for k in j["details"]["kernels"] if k["launches"]])
except Exception as e:
    print("failed", e); print(open("$O/r2c15_bloch.err").read()[-800:])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 2 --workload periodic_bloch --steps 100 --warmup 5 --no-extra > $O/r2c15_bloch_2gpu.json 2> $O/r2c15_bloch_2gpu.err
python - <<PY
import json
try:
    j=json.loads(open("$O/r2c15_bloch_2gpu.json").read().strip().splitlines()[-1]); print("bloch 2gpu", round(j["value"]), "e2e", round(j["e2e"]["value"]), j["config"]["grid"], j.get("per_rank"))
except Exception as e:
    print("failed", e); print(open("$O/r2c15_bloch_2gpu.err").read()[-1500:])
PY
