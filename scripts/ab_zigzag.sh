#!/bin/bash
# A/B of the zigzag tile order (KHR_ZIGZAG)
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_modes.py -m gpu -q -x 2>&1 | tail -3
summ() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%s %.0f Mcells/s %.4f ms/step (serialised %.4f) e2e %.0f | '%(d['config']['workload'], d['value'], d['ms_per_step'], d['ms_per_step_serialised_with_kernel_events'], d['e2e']['value']) + ' '.join('%s=%.4f'%(k['name'][12:], k['total_ms']/max(k['launches'],1)) for k in d['kernels'] if k['launches']))"; }
run() { echo "== $*"; env "$@" timeout 200 python bench.py --steps 1000 --warmup 10 --no-cpu | summ; env "$@" timeout 200 python bench.py --workload sphere --steps 200 --warmup 10 --no-cpu | summ; env "$@" timeout 200 python bench.py --workload uled --steps 1000 --warmup 10 --no-cpu | summ; }
run3() { echo "== $*"; env "$@" timeout 200 python bench.py --steps 1000 --warmup 10 --no-cpu | summ; env "$@" timeout 200 python bench.py --workload sphere --steps 200 --warmup 10 --no-cpu | summ; env "$@" timeout 200 python bench.py --workload uled --steps 1000 --warmup 10 --no-cpu | summ; }
{
run3 KHR_ZIGZAG=0
run3 KHR_ZIGZAG=1
run3 KHR_ZIGZAG=0
run3 KHR_ZIGZAG=1
} 2>&1 | tee $O/ab_zigzag.txt
