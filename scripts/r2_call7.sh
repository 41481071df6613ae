#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2c7_pytest.log 2>&1; echo "rc=$?" >> $O/r2c7_pytest.log
tail -6 $O/r2c7_pytest.log
run() { tag=$1; shift
  env "$@" timeout 300 python bench.py --workload $W --steps 200 --warmup 10 --no-cpu --no-extra > $O/r2c7_${W}_$tag.json 2> $O/r2c7_${W}_$tag.err
  python - <<PY
import json
try:
    j=json.loads(open("$O/r2c7_${W}_$tag.json").read().strip().splitlines()[-1])
    print("$W $tag", round(j["value"]), "e2e", round(j["e2e"]["value"]), "frac", round(j["roofline"]["frac"],3), "step", round(j["roofline"]["whole_step"]["frac"],3), [(k["name"].replace("step_kernel<f32,","")[:34], k["ctas"], round(k["total_ms"]/max(1,k["launches"]),4)) for k in j["details"]["kernels"] if k["launches"]])
except Exception as e:
    print("$W $tag failed", e); print(open("$O/r2c7_${W}_$tag.err").read()[-600:])
PY
}
for W in waveguide_mode sphere uled dipole500; do
run oldcuts KHR_TMA=0 KHR_LOCAL_CUTS=0
run base KHR_TMA=0
run tma KHR_TMA=1
done
