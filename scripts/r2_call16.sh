#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 240 > $O/r2c16_pytest.log 2>&1; echo "rc=$?" >> $O/r2c16_pytest.log
tail -8 $O/r2c16_pytest.log | cut -c1-300
timeout 150 python bench.py --workload periodic_bloch --steps 200 --warmup 10 --no-cpu --no-extra > $O/r2c16_bloch.json 2> $O/r2c16_bloch.err
python - <<PY
import json
try:
    j=json.loads(open("$O/r2c16_bloch.json").read().strip().splitlines()[-1])
    print("bloch", round(j["value"]), "e2e", round(j["e2e"]["value"]), j["config"]["grid"], "prep", round(j["details"]["prepare_s"],2), [(k["name"][:40], k["ctas"], round(k["total_ms"]/max(1,k["launches"]),4)) for k in j["details"]["kernels"] if k["launches"]])
except Exception as e:
    print("failed", e); print(open("$O/r2c16_bloch.err").read()[-800:])
PY
