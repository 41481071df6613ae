#!/bin/bash
# final confirmation of the round-2 defaults on one B200: GPU test suite, smoke(), the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2c31_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r2c31_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c31_bench.json 2> gpurun_out/r2c31_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c31_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("sphere", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(r["frac"],3), "traffic_frac", r.get("traffic_frac"), "whole", round(r["whole_step"]["frac"],3), "launches", d["gpu_launches"], "clocks", d["clocks"]["sm_mhz"], d["clocks"]["samples"], d["clocks"]["reasons"])
print("cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"]))
for k,v in d["extra"].items():
    print(k, v.get("error") or (round(v["value"]), "e2e", v["e2e"] and round(v["e2e"])))
PY
