#!/bin/bash
mkdir -p gpurun_out
for W in metalens_full uled waveguide_mode; do
ST=200; if [ $W = metalens_full ]; then ST=20; fi
KHR_PLAN_DUMP=1 timeout 150 python bench.py --workload $W --steps $ST --warmup 10 --no-extra --no-cpu > gpurun_out/r2c32_${W}.json 2> gpurun_out/r2c32_${W}.err
echo "== $W"; grep "^\[plan\]" gpurun_out/r2c32_${W}.err | cut -c1-100
python -c "
import json;d=json.loads(open('gpurun_out/r2c32_${W}.json').read().strip().splitlines()[-1]);print(round(d['value']), round(d['e2e']['value']))"
done
