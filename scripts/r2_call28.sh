#!/bin/bash
mkdir -p gpurun_out
for W in uled waveguide_mode; do for T in 0 1; do
KHR_PLAN_DUMP=1 KHR_TMA=$T timeout 120 python bench.py --workload $W --steps 100 --warmup 10 --no-extra --no-cpu > gpurun_out/r2c28_${W}_t$T.json 2> gpurun_out/r2c28_${W}_t$T.err
echo "== $W tma=$T"; grep "^\[plan\]" gpurun_out/r2c28_${W}_t$T.err
python -c "
import json;d=json.loads(open('gpurun_out/r2c28_${W}_t$T.json').read().strip().splitlines()[-1]);print(round(d['value']), [(k['name'][-28:],k['ctas'],round(k['total_ms']/k['launches'],4)) for k in d['details']['kernels']])"
done; done
