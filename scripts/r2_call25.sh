#!/bin/bash
# 4 GPUs: thin-slab parity (every slab inside / next to the z PML), then the full bench line with every named shape
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29701 scripts/mgpu_parity.py --thin > gpurun_out/r2c25_thin4.log 2>&1
echo "thin4 rc=$?"; grep "mgpu parity" gpurun_out/r2c25_thin4.log | cut -c1-400
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29702 scripts/mgpu_parity.py --thin --tma --periodic > gpurun_out/r2c25_thin4_tma.log 2>&1
echo "thin4 tma periodic rc=$?"; grep "mgpu parity" gpurun_out/r2c25_thin4_tma.log | cut -c1-400
t0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29703 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2c25_bench4.json 2> gpurun_out/r2c25_bench4.err
echo "bench4 rc=$? wall=$(( $(date +%s) - t0 ))s"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2c25_bench4.json").read().strip().splitlines()[-1])
    print("sphere x4", round(d["value"]), "e2e", round(d["e2e"]["value"]))
    for k,v in d["extra"].items():
        if k=="parity": print("parity", v.get("ok"), v.get("field_rel_l2")); continue
        print(k, v.get("error") or (round(v["value"]), "e2e", v["e2e"] and round(v["e2e"]), v["scaling"], v["slabs"]))
except Exception as e:
    print("parse failed", e)
PY
tail -3 gpurun_out/r2c25_bench4.err
