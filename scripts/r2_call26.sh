#!/bin/bash
# 8 GPUs: the full bench line with every named shape (uled on 8 ranks = 9-plane slabs inside the z PML)
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
t0=$(date +%s)
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29703 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2c26_bench8.json 2> gpurun_out/r2c26_bench8.err
echo "bench8 rc=$? wall=$(( $(date +%s) - t0 ))s"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2c26_bench8.json").read().strip().splitlines()[-1])
    print("sphere x8", round(d["value"]), "e2e", round(d["e2e"]["value"]))
    for k,v in d["extra"].items():
        if k=="parity": print("parity", v.get("ok"), v.get("field_rel_l2")); continue
        print(k, v.get("error") or (round(v["value"]), "e2e", v["e2e"] and round(v["e2e"]), v["scaling"], v["slabs"]))
except Exception as e:
    print("parse failed", e)
PY
tail -3 gpurun_out/r2c26_bench8.err
