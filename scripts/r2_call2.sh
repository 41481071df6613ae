#!/bin/bash
# GPU call 2: TMA kernel correctness (modes test) + full-size parity + A/B bench
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_modes.py -q -x > $O/r2c2_modes.log 2>&1; echo "rc=$?" >> $O/r2c2_modes.log
tail -15 $O/r2c2_modes.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -s > $O/r2c2_fullsize.log 2>&1; echo "rc=$?" >> $O/r2c2_fullsize.log
tail -25 $O/r2c2_fullsize.log
for w in waveguide_mode sphere; do for t in 0 1; do
KHR_TMA=$t timeout 300 python bench.py --workload $w --steps 200 --warmup 10 --no-cpu --no-extra > $O/r2c2_bench_${w}_tma$t.json 2> $O/r2c2_bench_${w}_tma$t.err
python - <<PY
import json
try:
    j=json.loads(open("$O/r2c2_bench_${w}_tma$t.json").read().strip().splitlines()[-1])
    print("$w tma=$t", round(j["value"]), "e2e", round(j["e2e"]["value"]), [(k["name"], round(k["total_ms"]/k["launches"],4)) for k in j["details"]["kernels"]])
except Exception as e:
    print("$w tma=$t failed", e); print(open("$O/r2c2_bench_${w}_tma$t.err").read()[-1500:])
PY
done; done
