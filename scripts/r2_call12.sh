#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_modes.py tests/test_gpu_parity.py tests/test_gpu_workloads.py -q -x > $O/r2c12_pytest.log 2>&1; echo "rc=$?" >> $O/r2c12_pytest.log
tail -4 $O/r2c12_pytest.log
run() { tag=$1; shift
  env "$@" timeout 300 python bench.py --workload $W --steps 400 --warmup 10 --no-cpu --no-extra > $O/r2c12_${W}_$tag.json 2> $O/r2c12_${W}_$tag.err
  python - <<PY
import json
try:
    j=json.loads(open("$O/r2c12_${W}_$tag.json").read().strip().splitlines()[-1])
    print("$W $tag", round(j["value"]), "e2e", round(j["e2e"]["value"]), "frac", round(j["roofline"]["frac"],3), "step", round(j["roofline"]["whole_step"]["frac"],3), [(k["name"].replace("step_kernel<f32,","")[:30], k["ctas"], round(k["total_ms"]/max(1,k["launches"]),4)) for k in j["details"]["kernels"] if k["launches"]])
except Exception as e:
    print("$W $tag failed", e); print(open("$O/r2c12_${W}_$tag.err").read()[-600:])
PY
}
for W in uled waveguide_mode; do
run split1 KHR_FULL_SPLIT=1
run split2
run split2_zf4 KHR_ZSEG_FULL=4
run split2_zf3 KHR_ZSEG_FULL=3
done
