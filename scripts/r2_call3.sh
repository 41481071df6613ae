#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_modes.py -q -x -k "TMA or tma" > $O/r2c3_modes.log 2>&1; echo "rc=$?" >> $O/r2c3_modes.log
tail -8 $O/r2c3_modes.log
run() { # name env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --workload $W --steps 200 --warmup 10 --no-cpu --no-extra > $O/r2c3_${W}_$tag.json 2> $O/r2c3_${W}_$tag.err
  python - <<PY
import json
try:
    j=json.loads(open("$O/r2c3_${W}_$tag.json").read().strip().splitlines()[-1])
    print("$W $tag", round(j["value"]), "e2e", round(j["e2e"]["value"]), [(k["name"].replace("step_kernel<f32,",""), round(k["total_ms"]/k["launches"],4)) for k in j["details"]["kernels"]])
except Exception as e:
    print("$W $tag failed", e); print(open("$O/r2c3_${W}_$tag.err").read()[-800:])
PY
}
for W in waveguide_mode sphere; do
run tma0 KHR_TMA=0
run tma1 KHR_TMA=1
run tma1s3 KHR_TMA=1 KHR_TMA_STAGES=3
run tma1s2 KHR_TMA=1 KHR_TMA_STAGES=2
done
