#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_modes.py tests/test_gpu_parity.py -q --timeout 240 -x > $O/r2c24_pytest.log 2>&1; echo "rc=$?" >> $O/r2c24_pytest.log
tail -6 $O/r2c24_pytest.log | cut -c1-300
run() { tag=$1; shift
  env "$@" timeout 200 python bench.py --workload $W --steps 400 --warmup 10 --no-cpu --no-extra > $O/r2c24_${W}_$tag.json 2> $O/r2c24_${W}_$tag.err
  python - <<PY
import json
try:
    j=json.loads(open("$O/r2c24_${W}_$tag.json").read().strip().splitlines()[-1])
    print("$W $tag", round(j["value"]), "e2e", round(j["e2e"]["value"]), "step", round(j["roofline"]["whole_step"]["frac"],3), [(k["name"].replace("step_kernel<f32,","")[:26], k["ctas"], round(k["total_ms"]/max(1,k["launches"]),4)) for k in j["details"]["kernels"] if k["launches"]])
except Exception as e:
    print("$W $tag failed", e); print(open("$O/r2c24_${W}_$tag.err").read()[-400:])
PY
}
for W in uled waveguide_mode; do
run old KHR_FIXUP=0
run fixup
done
