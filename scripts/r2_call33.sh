#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_modes.py tests/test_gpu_fullsize.py -q --timeout 240 -x 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c33_bench.json 2> gpurun_out/r2c33_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c33_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("sphere", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(r["frac"],3), "traffic_frac", r.get("traffic_frac"), "whole", round(r["whole_step"]["frac"],3), d["clocks"]["samples"], d["clocks"]["reasons"])
for k,v in d["extra"].items():
    print(k, v.get("error") or (round(v["value"]), "e2e", v["e2e"] and round(v["e2e"]), round(v["roofline"]["frac"],3), round(v["roofline"]["whole_step"]["frac"],3)))
PY
