#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_modes.py tests/test_gpu_parity.py tests/test_gpu_workloads.py -q --timeout 240 -x > gpurun_out/r2c29_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r2c29_pytest.log
for W in uled waveguide_mode sphere dipole500; do for F in 0 1; do for T in 0 1; do
if [ $W = sphere -o $W = dipole500 ] && [ $T = 0 ]; then continue; fi
ST=200; if [ $W = sphere -o $W = dipole500 ]; then ST=40; fi
KHR_PLAN_DUMP=1 KHR_PLAN_FILL=$F KHR_TMA=$T timeout 120 python bench.py --workload $W --steps $ST --warmup 10 --no-extra --no-cpu > gpurun_out/r2c29_${W}_f${F}_t$T.json 2> gpurun_out/r2c29_${W}_f${F}_t$T.err
echo "== $W fill=$F tma=$T"; grep "^\[plan\]" gpurun_out/r2c29_${W}_f${F}_t$T.err | cut -c1-110
python -c "
import json;d=json.loads(open('gpurun_out/r2c29_${W}_f${F}_t$T.json').read().strip().splitlines()[-1]);print(round(d['value']), round(d['e2e']['value']), [(k['name'][-28:],k['ctas'],round(k['total_ms']/k['launches'],4)) for k in d['details']['kernels']])"
done; done; done
